"""Pin the CPU oracle to outputs of the reference's own modules (tests/golden, made by
tests/make_golden.py from /root/reference).  CPU only."""

import json
import os

import numpy as np
import pytest
import torch

from oracle import chamfer as OC
from oracle import geometry as OG
from oracle import head as OH
from oracle import loop as OL
from oracle import partition as OP
from oracle import scheduler as OS
from oracle import training as OT


def load(golden_dir, name):
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = {k[3:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("w::")}
    return d, sd


def relmax(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("name", ["head_p1", "head_p2"])
def test_head_forward_matches_reference(golden_dir, name):
    d, sd = load(golden_dir, name)
    x, z, t = (torch.from_numpy(d[k]) for k in ("x", "z", "t"))
    pred_ids = torch.from_numpy(d["pred_ids"])
    assert relmax(OH.head_forward(sd, x, t, z), d["v_all"]) < 2e-6
    assert relmax(OH.head_forward(sd, x, t, z, pred_ids), d["v_pred"]) < 2e-6
    # training-mode per-token timesteps (diffusion_mlp.py:70,75)
    assert relmax(OH.head_forward(sd, x, torch.from_numpy(d["t_tok"]), z), d["v_tok_t"]) < 2e-6


def test_head_dims_and_key_contract(golden_dir):
    d, sd = load(golden_dir, "head_p1")
    depth, D, Dc, T, p, C = OH.head_dims(sd)
    assert (depth, D, Dc, T, p, C) == (2, 128, 96, 3, 1, 3)
    assert len(sd) == 14 + 8 * depth
    mine = OH.init_state_dict(depth, D, Dc, p, C, seed=101)
    assert list(mine.keys()) == list(sd.keys())
    for k in sd:  # same construction order => same draws as the reference
        assert torch.equal(mine[k], sd[k]), k


def test_init_checksums_full_size(golden_dir):
    with open(os.path.join(golden_dir, "init_checksums.json")) as f:
        rec = json.load(f)
    D = 768
    sd = OH.init_state_dict(6, D, D, 1, 3, seed=1337)
    r = rec[f"d6w{D}"]
    assert sum(v.numel() for v in sd.values()) == r["num_params"]
    assert len(sd) == r["num_keys"] == 62
    assert abs(float(sum(v.double().sum() for v in sd.values())) - r["sum"]) < 1e-6 * r["abs_sum"]
    assert [float(_) for _ in sd["head.weight"].flatten()[:4]] == r["head_w_0"]
    assert [float(_) for _ in sd["blocks.5.proj.fc2.weight"].flatten()[:4]] == r["b5_fc2_w_0"]


def test_patchify_roundtrip_and_order():
    x = torch.arange(2 * 4 * 4 * 6, dtype=torch.float32).reshape(2, 4, 4, 6)
    tok = OH.patchify(x, 2)
    assert tok.shape == (2, 6, 16)
    # token (h=0,w=1), inner order (ph, pw, c) with c fastest
    assert tok[0, 1, 0] == x[0, 0, 0, 2] and tok[0, 1, 1] == x[0, 1, 0, 2] and tok[0, 1, 4] == x[0, 0, 0, 3]
    assert torch.equal(OH.unpatchify(tok, 2, 4, 2, 3), x)


@pytest.mark.parametrize("steps,shift", [(25, 1.0), (25, 3.0), (10, 1.0), (64, 2.5), (1, 1.0)])
def test_schedule_matches_reference(golden_dir, steps, shift):
    d = np.load(os.path.join(golden_dir, "scheduler.npz"))
    ts, sig = OS.schedule(steps, shift=shift)
    assert ts.dtype == np.float32
    assert np.array_equal(ts, d[f"t_{steps}_{shift}"])
    assert np.array_equal(np.asarray(sig, dtype=np.float64), d[f"s_{steps}_{shift}"])


def test_schedule_known_answers():
    """SURVEY 8(c) survey-time known answers."""
    ts, sig = OS.schedule(25)
    assert ts[:4].tolist() == [1000.0, 958.375, 916.75, 875.125] and ts[-1] == 1.0
    assert sig[:3] == [1.0, 0.9583749771118164, 0.9167500138282776]
    assert sig[-3:] == [0.04262499883770943, 0.0010000000474974513, 0]
    dt = OS.dts(sig)
    assert dt[0] == -0.041625022888183594
    assert abs(sum(dt) + 1.0) < 1e-12
    assert abs(np.prod([1 + v for v in dt]) - 0.36009485691634535) < 1e-12
    ts3, sig3 = OS.schedule(25, shift=3.0)
    assert np.allclose(ts3[:3], [1000.0, 985.7583, 970.68146], rtol=1e-6)
    assert sig3[-3:-1] == [0.12268040329217911, 0.008928571827709675]


def test_euler_step_matches_reference(golden_dir):
    d = np.load(os.path.join(golden_dir, "scheduler.npz"))
    _, sig = OS.schedule(25)
    v, x = torch.from_numpy(d["step_v"]), torch.from_numpy(d["step_x"])
    assert np.array_equal(OS.euler_step(v, x, sig[1] - sig[0]).numpy(), d["step_out0"])
    assert np.array_equal(OS.euler_step(v, x, sig[2] - sig[1]).numpy(), d["step_out1"])
    out = OS.euler_step(v.bfloat16(), x.bfloat16(), sig[1] - sig[0]).float().numpy()
    assert np.array_equal(out, d["step_out0_bf16"])


@pytest.mark.parametrize("name", ["denoise_small", "denoise_shift3"])
def test_denoise_loop_matches_reference(golden_dir, name):
    d, sd = load(golden_dir, name)
    steps, shift = int(d["cfg"][3]), float(d["shift"])
    noise, z, zu = (torch.from_numpy(d[k]) for k in ("noise", "z", "zu"))
    pred_ids = torch.from_numpy(d["pred_ids"])
    kw = dict(num_steps=steps, shift=shift)
    assert relmax(OL.denoise(sd, z, noise, **kw), d["out_all"]) < 5e-6
    assert relmax(OL.denoise(sd, z, noise, pred_ids=pred_ids, **kw), d["out_pred"]) < 5e-6
    z2, p2 = torch.cat([z, zu]), torch.cat([pred_ids, pred_ids])
    out = OL.denoise(sd, z2, noise, pred_ids=p2, guidance_scale=3.0, **kw)
    assert relmax(out, d["out_cfg"]) < 5e-6
    out = OL.denoise(sd, z2, noise, pred_ids=p2, guidance_scale=3.0, guidance_renorm=0.6, **kw)
    assert relmax(out, d["out_cfg_renorm"]) < 5e-6
    out = OL.denoise(sd, z2, noise, pred_ids=p2, guidance_scale=3.0, guidance_trunc=400.0, **kw)
    assert relmax(out, d["out_cfg_trunc"]) < 5e-6


GUIDANCE3 = {
    "img": dict(image_guidance_scale=1.5),
    "img_all": dict(image_guidance_scale=1.5),
    "img_renorm": dict(image_guidance_scale=1.5, guidance_renorm=0.6),
    "st": dict(spatiotemporal_guidance_scale=0.8),
    "st_renorm_trunc": dict(spatiotemporal_guidance_scale=0.8, guidance_renorm=0.7, guidance_trunc=400.0),
}


@pytest.mark.parametrize("mode", sorted(GUIDANCE3))
def test_three_pass_guidance_matches_reference(golden_dir, mode):
    """image_guidance_scale / spatiotemporal_guidance_scale (guidance_scaler.py:78-85) against the reference's own denoise."""
    d, sd = load(golden_dir, "denoise_guidance3")
    noise, z3 = torch.from_numpy(d["noise"]), torch.from_numpy(d["z3"])
    p3 = None if mode == "img_all" else torch.cat([torch.from_numpy(d["pred_ids"])] * 3)
    out = OL.denoise(sd, z3, noise, num_steps=int(d["cfg"][3]), shift=float(d["shift"]), pred_ids=p3, guidance_scale=3.0,
                     **GUIDANCE3[mode])
    assert relmax(out, d["out_" + mode]) < 5e-6


def test_denoise_unpredicted_rows_closed_form(golden_dir):
    """Rows outside pred_ids follow x <- x + dt x: noise * prod(1 + dt_i)."""
    d, sd = load(golden_dir, "denoise_small")
    noise = torch.from_numpy(d["noise"])
    ids = d["pred_ids"][..., 0]
    tok = OH.patchify(noise, 1).numpy()
    out = d["out_pred"]
    for b in range(out.shape[0]):
        rest = np.setdiff1d(np.arange(out.shape[1]), ids[b])
        assert np.allclose(out[b, rest], tok[b, rest] * 0.36009485691634535, rtol=2e-6)


def test_hoisted_loop_equals_plain(golden_dir):
    d, sd = load(golden_dir, "denoise_small")
    noise, z = torch.from_numpy(d["noise"]), torch.from_numpy(d["z"])
    out = OL.denoise_tokens_fast(sd, z, OH.patchify(noise, 1))
    assert relmax(out, d["out_all"]) < 5e-6


def test_chamfer_variants_match_reference(golden_dir):
    d = np.load(os.path.join(golden_dir, "chamfer.npz"))
    a, b = d["a"], d["b"]
    for i in range(a.shape[0]):
        assert abs(OC.chamfer_a(a[i], b[i]) - d["cd_a"][i]) < 1e-12
    dl, dr, mean = OC.chamfer_b(a, b)
    assert np.allclose([dl, dr, mean], d["cd_b"], atol=1e-4)  # reference uses fp32 mm-form cdist
    sel = [0, 2]  # variant C is discontinuous at coincident points (see tests/make_golden.py)
    assert abs(OC.chamfer_c(a[sel], b[sel]) - float(d["cd_c"])) < 1e-4
    assert abs(OC.chamfer_c(a[sel], b[sel]) - 2.0) < 1e-3  # SURVEY: variant C ~ 2.0 on separated clouds


def test_chamfer_edge_cases():
    p = np.random.default_rng(0).normal(size=(50, 3)).astype(np.float32)
    assert OC.chamfer_a(p, p) == 0.0
    assert abs(OC.chamfer_a(p, p + np.float32(0.5)) - OC.chamfer_a(p + np.float32(0.5), p)) < 1e-12
    m1, m2, i1, i2 = OC.nn_dist(p[:1], p)  # single point vs many
    assert m1.shape == (1,) and m2.shape == (50,) and m1[0] == 0.0 and i1[0] == 0


def test_partition_shapes():
    """SURVEY a15 known shapes."""
    n1024 = OP.cosine_num_preds(1024)
    assert len(n1024) == 64 and sum(n1024) == 1024 and n1024[:5] == [0, 1, 2, 2, 3] and max(n1024) == 25
    n2048 = OP.cosine_num_preds(2048)
    assert sum(n2048) == 2048 and n2048[:5] == [1, 1, 4, 4, 5] and max(n2048) in (50, 51)
    assert OP.equal_subset_sizes(1024) == [51] * 19 + [55]
    assert OP.equal_subset_sizes(2048) == [102] * 19 + [110]
    order = np.stack([np.random.default_rng(i).permutation(1024) for i in range(2)])
    parts = OP.split_order(order, n1024)
    assert len(parts) == 63 and np.array_equal(np.concatenate(parts, axis=1), order)


def test_geometry_matches_reference(golden_dir):
    """oracle.geometry against the reference's own compute_local_density / feature_aware_interpolation outputs
    (<= 25 points the reference's torch.cdist is exact: 1e-6; above it takes the mm form: 1e-4)."""
    g = np.load(os.path.join(golden_dir, "geometry.npz"))
    assert np.abs(OG.local_density(g["small"]) - g["density_small"]).max() < 1e-6
    assert np.abs(OG.local_density(g["small"], 3) - g["density_small_k3"]).max() < 1e-6
    assert np.abs(OG.local_density(g["big"]) - g["density_big"]).max() < 1e-4
    for name, tol in (("small", 1e-6), ("big", 1e-4)):
        idx = g[f"interp_{name}_idx"]
        assert np.abs(OG.interpolate(g[name], len(idx), idx) - g[f"interp_{name}"]).max() < tol
    assert np.array_equal(OG.interpolate(g["small"], 60, None).astype(np.float32), g["interp_repeat"])


def test_geometry_oracle_edges():
    p = np.random.default_rng(3).uniform(-1, 1, (1, 40, 3)).astype(np.float32)
    d, i = OG.knn(p[0], p[0], 5)
    assert (d[:, 0] == 0).all() and (i[:, 0] == np.arange(40)).all() and (np.diff(d, axis=1) >= 0).all()
    t = np.array([[0.0, 0, 0], [1, 0, 0], [1, 0, 0], [0, 0, 0], [1, 0, 0]])
    assert OG.knn(np.array([[1.0, 0, 0]]), t, 4)[1].tolist() == [[1, 2, 4, 0]]  # ties: lowest index first
    with pytest.raises(ValueError):
        OG.knn(t, t, 6)
    far = OG.softmax_interp(p * 300, p * 300)  # weights collapse onto the point itself, no underflow to 0/0
    assert np.isfinite(far).all() and np.abs(far - p.astype(np.float64) * 300).max() < 1e-3
    assert OG.target_size(0.5, 15000, 20) == 750 and OG.target_size(-10.0, 15000, 20) == 100
    assert OG.target_size(10.0, 15000, 20) == 1500


def test_training_loss_matches_reference(golden_dir):
    """oracle.training.get_losses against Transformer3DModel.get_losses run as it stands (tests/make_golden.py
    losses_case): same noise and timestep indices (replayed from the reference's seed) -> the same loss."""
    d = np.load(os.path.join(golden_dir, "losses.npz"))
    depth, D, Dc, patch, chan = [int(v) for v in d["cfg"]]
    sd = OH.init_state_dict(depth, D, Dc, patch, chan, seed=int(d["init_seed"]))  # the reference's default init
    out = OT.get_losses(sd, torch.from_numpy(d["z"]), torch.from_numpy(d["x"]), torch.from_numpy(d["noise"]),
                        torch.from_numpy(d["t_idx"]), torch.from_numpy(d["mask"]))
    assert abs(float(out["loss"]) - float(d["loss"])) <= 1e-6 * abs(float(d["loss"]))
    # masked-out tokens carry no loss; the per-token losses add up to the scalar
    w = torch.from_numpy(d["mask"]).repeat(4, 1, 1).squeeze(-1)
    assert float(out["loss_per_token"][w == 0].abs().max()) == 0.0
    assert abs(float(out["loss_per_token"].sum()) - float(out["loss"])) < 1e-6


def test_training_tables_and_timestep_sampling():
    sig, tt = OT.training_tables(1000, 1.0)
    assert sig.shape == (1000,) and float(sig[0]) == 1.0 and abs(float(sig[-1]) - 1e-3) < 1e-9 and float(tt[0]) == 1000.0
    sig3, _ = OT.training_tables(1000, 3.0)
    assert abs(float(sig3[500]) - 3 * 0.5 / (1 + 2 * 0.5)) < 1e-6
    idx = OT.sample_timesteps((64, 50), generator=torch.Generator().manual_seed(0))
    assert idx.dtype == torch.int64 and int(idx.min()) >= 0 and int(idx.max()) <= 999
    x, n = torch.randn(2, 5, 3), torch.randn(2, 5, 3)
    zero = torch.zeros(2, 5, dtype=torch.int64)
    assert torch.equal(OT.add_noise(x, n, zero, sig), n)  # sigma = 1 at index 0: pure noise


def test_emd_oracle_is_the_exact_assignment_optimum():
    """oracle/chamfer.py::emd restates demo.py:57-74 (scipy linear_sum_assignment on cdist): on clouds small enough to
    enumerate, it equals the minimum over all permutations; emd_approx clamps to [-2, 2] first (train_newloss.py:363-364)."""
    import itertools

    from oracle import chamfer as OC

    g = np.random.default_rng(3)
    a, b = g.normal(size=(6, 3)) * 1.5, g.normal(size=(6, 3)) * 1.5
    best = min(np.linalg.norm(a - b[list(p)], axis=1).mean() for p in itertools.permutations(range(6)))
    assert abs(OC.emd(a, b) - best) < 1e-12
    ac, bc = np.clip(a, -2, 2), np.clip(b, -2, 2)
    assert abs(OC.emd_approx(a[None], b[None])[0] - OC.emd(ac, bc)) < 1e-12
    assert OC.emd(a, a) == 0.0
