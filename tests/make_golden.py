"""Generate tests/golden/* by running the UNMODIFIED reference modules from /root/reference.

Run in the build container (the GPU box has no /root/reference):

    python tests/make_golden.py

The reference ships no fixtures of its own, so these outputs of the
reference's own code on seeded synthetic inputs are the parity pins.  Every fixture
stores its inputs, its weights (tiny configs only) and the reference's outputs.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle._reference_import import import_reference, reference_denoiser, reference_geometry  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sd_np(sd):
    return {"w::" + k: v.detach().numpy() for k, v in sd.items()}


def head_case(ref, name, depth, D, Dc, patch, chan, B, H, W, n_pred, seed):
    torch.manual_seed(seed)
    head = ref.DiffusionMLP(depth, D, Dc, patch_size=patch, image_dim=chan).eval()
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, chan, H * patch, W * patch, generator=g)
    N = H * W
    z = torch.randn(B, N, Dc, generator=g)
    t = torch.rand(B, generator=g) * 1000
    order = torch.rand(B, N, generator=g).argsort(dim=1)
    pred_ids = order[:, :n_pred].unsqueeze(-1).contiguous()
    with torch.no_grad():
        v_all = head(x, t, z)
        v_pred = head(x, t, z, pred_ids)
        t_tok = torch.rand(B, N, generator=g) * 1000  # training-mode per-token timesteps
        v_tok_t = head(x, t_tok, z)
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        cfg=np.array([depth, D, Dc, patch, chan]),
        x=x.numpy(), z=z.numpy(), t=t.numpy(), pred_ids=pred_ids.numpy(),
        v_all=v_all.numpy(), v_pred=v_pred.numpy(), t_tok=t_tok.numpy(), v_tok_t=v_tok_t.numpy(),
        **sd_np(head.state_dict()),
    )


def denoise_case(ref, name, depth, D, Dc, B, N, n_pred, steps, shift, seed):
    torch.manual_seed(seed)
    head = ref.DiffusionMLP(depth, D, Dc, patch_size=1, image_dim=3).eval()
    model, sched = reference_denoiser(ref, head, steps, shift)
    g = torch.Generator().manual_seed(seed + 1)
    noise = torch.randn(B, 3, N, 1, generator=g)
    z = torch.randn(B, N, Dc, generator=g)
    zu = torch.randn(B, N, Dc, generator=g)
    order = torch.rand(B, N, generator=g).argsort(dim=1)
    pred_ids = order[:, :n_pred].unsqueeze(-1).contiguous()
    out = {}
    out["all"] = model.denoise(z, noise.clone(), ref.GuidanceScaler(guidance_scale=1))
    out["pred"] = model.denoise(z, noise.clone(), ref.GuidanceScaler(guidance_scale=1), None, pred_ids)
    z2, p2 = torch.cat([z, zu]), torch.cat([pred_ids, pred_ids])
    out["cfg"] = model.denoise(z2, noise.clone(), ref.GuidanceScaler(guidance_scale=3.0), None, p2)
    out["cfg_renorm"] = model.denoise(
        z2, noise.clone(), ref.GuidanceScaler(guidance_scale=3.0, guidance_renorm=0.6), None, p2)
    out["cfg_trunc"] = model.denoise(
        z2, noise.clone(), ref.GuidanceScaler(guidance_scale=3.0, guidance_trunc=400.0), None, p2)
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        cfg=np.array([depth, D, Dc, steps]), shift=np.array(shift),
        noise=noise.numpy(), z=z.numpy(), zu=zu.numpy(), pred_ids=pred_ids.numpy(),
        timesteps=np.asarray(sched.timesteps), sigmas=np.asarray(sched.sigmas, dtype=np.float64),
        **{"out_" + k: v.numpy() for k, v in out.items()},
        **sd_np(head.state_dict()),
    )


def scheduler_case(ref):
    rec = {}
    for steps, shift in [(25, 1.0), (25, 3.0), (10, 1.0), (64, 2.5), (1, 1.0)]:
        s = ref.FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=shift)
        s.set_timesteps(steps)
        rec[f"t_{steps}_{shift}"] = np.asarray(s.timesteps)
        rec[f"s_{steps}_{shift}"] = np.asarray(s.sigmas, dtype=np.float64)
    # step(): one update on a seeded tensor, fp32 and bf16
    s = ref.FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=1.0)
    s.set_timesteps(25)
    g = torch.Generator().manual_seed(7)
    v, x = torch.randn(4, 3, 32, 1, generator=g), torch.randn(4, 3, 32, 1, generator=g)
    s._step_index = None
    rec["step_v"], rec["step_x"] = v.numpy(), x.numpy()
    rec["step_out0"] = s.step(v, s.timesteps[0], x).prev_sample.numpy()
    rec["step_out1"] = s.step(v, s.timesteps[1], x).prev_sample.numpy()
    s._step_index = None
    rec["step_out0_bf16"] = s.step(v.bfloat16(), s.timesteps[0], x.bfloat16()).prev_sample.float().numpy()
    np.savez_compressed(os.path.join(GOLD, "scheduler.npz"), **rec)


def chamfer_case():
    """Chamfer A/B/C restated call-for-call from the reference scripts (they import
    matplotlib / swanlab at module scope, which are absent, so the three functions'
    bodies are executed here verbatim-in-effect: scipy cdist for A (demo.py:44-53),
    torch.cdist for B (train_newloss.py:321-349) and C (test_optimize.py:357-383))."""
    from scipy.spatial.distance import cdist

    rng = np.random.default_rng(11)
    a = rng.uniform(-1, 1, size=(3, 200, 3)).astype(np.float32)
    b = np.random.default_rng(12).uniform(-1, 1, size=(3, 160, 3)).astype(np.float32)
    b[1, :5] = a[1, :5]  # exact coincidences
    cd_a = []
    for i in range(3):
        d = cdist(a[i], b[i])
        cd_a.append(np.mean(np.min(d, axis=1)) + np.mean(np.min(d, axis=0)))
    # B
    x, y = torch.from_numpy(a), torch.from_numpy(b)
    xc, yc = torch.clamp(x, -1, 1), torch.clamp(y, -1, 1)
    xn = xc / torch.clamp(torch.norm(xc, dim=-1, keepdim=True), min=1e-8)
    yn = yc / torch.clamp(torch.norm(yc, dim=-1, keepdim=True), min=1e-8)
    dm = torch.clamp(torch.cdist(xn, yn), min=1e-8)
    ld = torch.clamp(torch.log(dm + 1e-8), min=-10, max=10)
    dl, dr = ld.min(2)[0].exp().mean(), ld.min(1)[0].exp().mean()
    # C: m/(m+1e-6) is discontinuous at m=0, where the reference's fp32 mm-form cdist returns
    # ~1e-4 instead of 0; the pin therefore uses the two clouds without exact coincidences.
    pc, tc = torch.clamp(x[[0, 2]], -5, 5), torch.clamp(y[[0, 2]], -5, 5)
    n = min(pc.shape[1], tc.shape[1])
    dist = torch.cdist(pc[:, :n], tc[:, :n])
    m1, m2 = dist.min(dim=2)[0], dist.min(dim=1)[0]
    c = torch.clamp(((m1 / (m1 + 1e-6)).mean(1) + (m2 / (m2 + 1e-6)).mean(1)).mean(), 0, 10)
    np.savez_compressed(
        os.path.join(GOLD, "chamfer.npz"), a=a, b=b, cd_a=np.array(cd_a),
        cd_b=np.array([dl.item(), dr.item(), ((dl + dr) / 2).item()]), cd_c=np.array(c.item()),
    )


def geometry_case():
    """compute_local_density / feature_aware_interpolation run as they stand in the reference
    (transformer_pointcloud_nova.py:81-89,128-152).  The interpolation draws its target indices with the global
    RNG; the seed is set right before the call and the same randperm is replayed to record them."""
    geo = reference_geometry()
    g = torch.Generator().manual_seed(515)
    big = torch.rand(3, 200, 3, generator=g) * 2 - 1          # > 25 points: torch.cdist takes the mm form
    small = torch.rand(2, 24, 3, generator=g) * 2 - 1         # <= 25 points: exact differences
    big[1, 7] = big[1, 3]                                      # a duplicated point: two zero distances
    rec = {"big": big.numpy(), "small": small.numpy()}
    rec["density_big"] = geo.compute_local_density(big).numpy()
    rec["density_small"] = geo.compute_local_density(small).numpy()
    rec["density_small_k3"] = geo.compute_local_density(small, k_neighbors=3).numpy()
    for name, pts, size in (("big", big, 50), ("small", small, 10)):
        torch.manual_seed(616)
        rec[f"interp_{name}"] = geo.feature_aware_interpolation(pts, size).numpy()
        torch.manual_seed(616)
        rec[f"interp_{name}_idx"] = torch.randperm(pts.shape[1])[:size].numpy()
    rec["interp_repeat"] = geo.feature_aware_interpolation(small, 60).numpy()   # N <= target: tile and cut
    np.savez_compressed(os.path.join(GOLD, "geometry.npz"), **rec)


def losses_case(ref):
    """Transformer3DModel.get_losses (transformer_3d.py:81-95) as it stands, loss_repeat = 4, with a 0/1 mask weight.
    Its random draws (randn for the noise, then normal for the timestep indices) are replayed from the same seed
    and stored, so the oracle and the CUDA path can be fed the identical noise and timesteps."""
    import types

    from oracle import head as OH

    D = 256  # the CUDA head needs widths in multiples of 256; weights are NOT stored: they are the reference's default
    torch.manual_seed(707)  # init under this seed, which oracle.head.init_state_dict reproduces (checked below)
    head = ref.DiffusionMLP(1, D, D, patch_size=1, image_dim=3).eval()
    sd = OH.init_state_dict(1, D, D, 1, 3, seed=707)
    assert all(torch.equal(sd[k], v) for k, v in head.state_dict().items())
    enc = torch.nn.Module()
    enc.patch_embed = head.patch_embed
    head.patch_embed.height, head.patch_embed.width = 20, 1  # normally left behind by the encoder's forward
    g = torch.Generator().manual_seed(708)
    B, N = 3, 20
    x = torch.randn(B, 3, N, 1, generator=g)
    z = torch.randn(B, N, D, generator=g)
    mask = (torch.rand(B, N, 1, generator=g) < 0.7).float()
    model = ref.Transformer3DModel(image_encoder=enc, image_decoder=head, mask_embed=types.SimpleNamespace(mask=mask),
                                   noise_scheduler=ref.FlowMatchEulerDiscreteScheduler(1000, shift=1.0))
    torch.manual_seed(709)
    with torch.no_grad():
        out = model.get_losses(z, x)
    torch.manual_seed(709)
    noise = torch.randn(4 * B, N, 3)
    t_idx = torch.normal(0, 1, (4 * B, N)).sigmoid_().mul_(1000).to(torch.int64)
    np.savez_compressed(os.path.join(GOLD, "losses.npz"), cfg=np.array([1, D, D, 1, 3]), init_seed=np.array(707),
                        x=x.numpy(), z=z.numpy(), mask=mask.numpy(), noise=noise.numpy(), t_idx=t_idx.numpy(),
                        loss=np.array(out["loss"].item()))


def init_checksums(ref):
    """Checksums of the reference's random init for the BASELINE widths (too big to commit)."""
    rec = {}
    for D in (768, 1024, 1536):
        torch.manual_seed(1337)
        head = ref.DiffusionMLP(6, D, D, patch_size=1, image_dim=3)
        sd = head.state_dict()
        rec[f"d6w{D}"] = {
            "num_params": int(sum(v.numel() for v in sd.values())),
            "num_keys": len(sd),
            "sum": float(sum(v.double().sum() for v in sd.values())),
            "abs_sum": float(sum(v.double().abs().sum() for v in sd.values())),
            "head_w_0": [float(_) for _ in sd["head.weight"].flatten()[:4]],
            "b5_fc2_w_0": [float(_) for _ in sd["blocks.5.proj.fc2.weight"].flatten()[:4]],
        }
    with open(os.path.join(GOLD, "init_checksums.json"), "w") as f:
        json.dump(rec, f, indent=1)


def guidance3_case(ref, name="denoise_guidance3", depth=2, D=64, Dc=64, B=2, N=16, n_pred=5, steps=12, shift=1.0, seed=505):
    """Three-pass guidance (guidance_scaler.py:78-85): z = [cond; uncond; third], pred_ids expanded three times."""
    torch.manual_seed(seed)
    head = ref.DiffusionMLP(depth, D, Dc, patch_size=1, image_dim=3).eval()
    model, sched = reference_denoiser(ref, head, steps, shift)
    g = torch.Generator().manual_seed(seed + 1)
    noise = torch.randn(B, 3, N, 1, generator=g)
    z3 = torch.randn(3 * B, N, Dc, generator=g)
    order = torch.rand(B, N, generator=g).argsort(dim=1)
    pred_ids = order[:, :n_pred].unsqueeze(-1).contiguous()
    p3 = torch.cat([pred_ids] * 3)
    out = {}
    out["img"] = model.denoise(z3, noise.clone(), ref.GuidanceScaler(guidance_scale=3.0, image_guidance_scale=1.5), None, p3)
    out["img_all"] = model.denoise(z3, noise.clone(), ref.GuidanceScaler(guidance_scale=3.0, image_guidance_scale=1.5))
    out["img_renorm"] = model.denoise(
        z3, noise.clone(), ref.GuidanceScaler(guidance_scale=3.0, image_guidance_scale=1.5, guidance_renorm=0.6), None, p3)
    out["st"] = model.denoise(
        z3, noise.clone(), ref.GuidanceScaler(guidance_scale=3.0, spatiotemporal_guidance_scale=0.8), None, p3)
    out["st_renorm_trunc"] = model.denoise(
        z3, noise.clone(), ref.GuidanceScaler(guidance_scale=3.0, spatiotemporal_guidance_scale=0.8, guidance_renorm=0.7,
                                              guidance_trunc=400.0), None, p3)
    np.savez_compressed(
        os.path.join(GOLD, name + ".npz"),
        cfg=np.array([depth, D, Dc, steps]), shift=np.array(shift),
        noise=noise.numpy(), z3=z3.numpy(), pred_ids=pred_ids.numpy(),
        **{"out_" + k: v.numpy() for k, v in out.items()},
        **sd_np(head.state_dict()),
    )


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(4)
    if "--only-geometry" in sys.argv:
        geometry_case()
        return
    if "--only-guidance3" in sys.argv:
        with torch.no_grad():
            guidance3_case(import_reference())
        return
    if "--only-losses" in sys.argv:
        losses_case(import_reference())
        return
    ref = import_reference()
    with torch.no_grad():
        head_case(ref, "head_p1", depth=2, D=128, Dc=96, patch=1, chan=3, B=3, H=24, W=1, n_pred=7, seed=101)
        head_case(ref, "head_p2", depth=1, D=64, Dc=64, patch=2, chan=4, B=2, H=4, W=6, n_pred=5, seed=202)
        denoise_case(ref, "denoise_small", depth=2, D=128, Dc=128, B=2, N=20, n_pred=6, steps=25, shift=1.0, seed=303)
        denoise_case(ref, "denoise_shift3", depth=1, D=64, Dc=64, B=2, N=12, n_pred=4, steps=10, shift=3.0, seed=404)
        guidance3_case(ref)
        scheduler_case(ref)
        chamfer_case()
        geometry_case()
        losses_case(ref)
        init_checksums(ref)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
