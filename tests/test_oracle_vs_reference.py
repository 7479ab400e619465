"""CPU: the oracle against the LIVE reference modules, imported unmodified from /root/reference.

Runs in the build container only (the GPU box has no /root/reference: every test here skips there).
The committed fixtures under tests/golden/ pin a handful of seeds; these tests draw fresh seeds and
shapes every run configuration listed below, so the oracle cannot have been fitted to the fixtures.
"""

import numpy as np
import pytest
import torch

from oracle import head as OH
from oracle import loop as OL
from oracle import scheduler as OS
from oracle._reference_import import import_reference, reference_available, reference_denoiser

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference is not mounted")


@pytest.fixture(scope="module")
def ref():
    return import_reference()


def _case(ref, depth, D, Dc, patch, chan, B, H, W, n_pred, seed):
    torch.manual_seed(seed)
    head = ref.DiffusionMLP(depth, D, Dc, patch_size=patch, image_dim=chan).eval()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, chan, H * patch, W * patch, generator=g)
    z = torch.randn(B, H * W, Dc, generator=g)
    t = torch.rand(B, generator=g) * 1000
    ids = None
    if n_pred is not None:
        ids = torch.rand(B, H * W, generator=g).argsort(dim=1)[:, :n_pred].unsqueeze(-1).contiguous()
    sd = {k: v.detach().clone() for k, v in head.state_dict().items()}
    return head, sd, x, z, t, ids


@pytest.mark.parametrize("depth,D,Dc,patch,chan,B,H,W,n_pred,seed", [
    (6, 128, 128, 1, 3, 2, 40, 1, None, 11),    # xyz point tokens (the point-cloud mapping)
    (2, 64, 96, 2, 4, 3, 4, 6, None, 12),       # registry default token dim 16, cond width != width
    (3, 96, 96, 1, 3, 2, 33, 1, 7, 13),         # pred_ids: gather + scatter into the patchified input
    (1, 64, 64, 2, 4, 1, 3, 3, 9, 14),          # every token listed in pred_ids
    (0, 64, 64, 1, 3, 2, 5, 1, None, 15),       # no blocks: embed -> final AdaLN -> head
])
def test_head_forward_fresh_seeds(ref, depth, D, Dc, patch, chan, B, H, W, n_pred, seed):
    head, sd, x, z, t, ids = _case(ref, depth, D, Dc, patch, chan, B, H, W, n_pred, seed)
    with torch.no_grad():
        want = head(x, t, z, ids)
    got = OH.head_forward(sd, x, t, z, ids)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-6 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("depth,per_token", [(2, False), (3, True), (0, False)])
def test_head_forward_pre_embedded_rows(ref, depth, per_token):
    """A 3-D x passes through PatchEmbed.forward unchanged (embeddings.py:160-166): the blocks start from the caller's rows."""
    head, sd, _, z, t, _ = _case(ref, depth, 64, 96, 1, 3, 2, 12, 1, None, 31 + depth)
    g = torch.Generator().manual_seed(77)
    x_emb = torch.randn(2, 12, 64, generator=g)
    t = torch.rand(2, 12, generator=g) * 1000 if per_token else t
    with torch.no_grad():
        want = head(x_emb, t, z)
    got = OH.head_embedded(sd, x_emb, t, z)
    assert got.shape == want.shape == (2, 12, 3)
    assert float((got - want).abs().max()) <= 2e-6 * max(1.0, float(want.abs().max()))


def test_head_forward_per_token_timesteps(ref):
    head, sd, x, z, _, _ = _case(ref, 2, 64, 64, 1, 3, 2, 12, 1, None, 21)
    t = torch.rand(2, 12) * 1000  # training-mode call: one timestep per token (transformer_3d.py:85-90)
    with torch.no_grad():
        want = head(x, t, z)
    got = OH.head_forward(sd, x, t, z)
    assert float((got - want).abs().max()) <= 2e-6 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("steps,shift", [(25, 1.0), (7, 2.0), (50, 3.0), (2, 1.0)])
def test_schedule_and_step_fresh(ref, steps, shift):
    s = ref.FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=shift)
    s.set_timesteps(steps)
    ts, sig = OS.schedule(steps, shift=shift)
    assert np.array_equal(np.asarray(s.timesteps), ts) and list(s.sigmas) == list(sig)
    g = torch.Generator().manual_seed(steps)
    for dtype in (torch.float32, torch.bfloat16):
        v, x = torch.randn(3, 5, generator=g).to(dtype), torch.randn(3, 5, generator=g).to(dtype)
        s._step_index = None
        want = s.step(v, s.timesteps[0], x).prev_sample
        assert torch.equal(OS.euler_step(v, x, sig[1] - sig[0]), want)


@pytest.mark.parametrize("mode", ["plain", "pred", "cfg", "cfg_renorm", "cfg_trunc"])
def test_denoise_loop_fresh(ref, mode):
    B, N, D = 2, 24, 64
    # the reference's maybe_disable chunks pred_ids unconditionally (guidance_scaler.py:64): truncation needs pred_ids
    head, sd, x, z, _, ids = _case(ref, 2, D, D, 1, 3, B, N, 1, 9 if mode in ("pred", "cfg_trunc") else None, 31)
    model, sched = reference_denoiser(ref, head, num_steps=6, shift=1.5)
    kw = {}
    if mode.startswith("cfg"):
        kw = dict(guidance_scale=3.0)
        if mode == "cfg_renorm":
            kw["guidance_renorm"] = 0.4
        if mode == "cfg_trunc":
            kw["guidance_trunc"] = 500.0
        z = torch.cat([z, torch.zeros_like(z)])
    gs = ref.GuidanceScaler(**kw)
    pid = None if ids is None else gs.expand(ids)
    with torch.no_grad():
        want = model.denoise(z, x.clone(), gs, None, pid)
    got = OL.denoise(sd, z, x, num_steps=6, shift=1.5, pred_ids=pid, guidance_scale=kw.get("guidance_scale", 1.0),
                     guidance_trunc=kw.get("guidance_trunc", 0.0), guidance_renorm=kw.get("guidance_renorm", 1.0))
    assert float((got - want).abs().max()) <= 5e-6 * max(1.0, float(want.abs().max()))


def test_product_module_initialises_like_the_reference(ref):
    """Same seed => the product's parameter holders draw the reference's own random init, key for key."""
    import nova_pointcloud_b200 as nb

    torch.manual_seed(77)
    want = ref.DiffusionMLP(3, 64, 32, patch_size=2, image_dim=4).state_dict()
    torch.manual_seed(77)
    got = nb.DiffusionMLP(3, 64, 32, patch_size=2, image_dim=4).state_dict()
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert torch.equal(got[k], want[k]), k


@pytest.mark.parametrize("B,N,k,seed", [(2, 20, 8, 1), (1, 25, 4, 2), (3, 150, 8, 3), (1, 600, 8, 4)])
def test_geometry_live(B, N, k, seed):
    """oracle.geometry against the reference's functions as they stand (transformer_pointcloud_nova.py:81-152),
    including the degenerate farthest_point_sampling the product leaves out."""
    from oracle import geometry as OG
    from oracle._reference_import import reference_geometry

    geo = reference_geometry()
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(B, N, 3, generator=g) * 2 - 1
    tol = 1e-6 if N <= 25 else 1e-4  # torch.cdist: exact differences up to 25 points, mm form above
    assert np.abs(OG.local_density(pts.numpy(), k) - geo.compute_local_density(pts, k).numpy()).max() < tol
    size = max(1, N // 3)
    torch.manual_seed(seed + 100)
    out = geo.feature_aware_interpolation(pts, size).numpy()
    torch.manual_seed(seed + 100)
    idx = torch.randperm(N)[:size].numpy()
    assert np.abs(OG.interpolate(pts.numpy(), size, idx) - out).max() < tol
    if N <= 25:  # exact distances: the zero diagonal wins every min, so every pick after the start is index 0
        torch.manual_seed(seed + 200)
        picked = geo.farthest_point_sampling(pts, 4)
        assert torch.equal(picked[:, 1:], pts[:, :1].expand(-1, 3, -1))
