"""GPU: Chamfer kernel against scipy float64 (small), the committed golden values, and
size-independent properties at BASELINE's full size (256 x 2048 vs 2048)."""

import os

import numpy as np
import pytest
import torch

from gpu_util import record
from oracle import chamfer as OC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,N,M", [(1, 1, 1), (2, 37, 129), (3, 1000, 777), (1, 2048, 2048), (2, 5, 3000)])
def test_nn_matches_scipy(B, N, M):
    import nova_pointcloud_b200 as nb

    rng = np.random.default_rng(N * 7 + M)
    a = rng.uniform(-1, 1, (B, N, 3)).astype(np.float32)
    b = rng.uniform(-1, 1, (B, M, 3)).astype(np.float32)
    d1, d2, i1, i2 = nb.chamfer_nn(torch.from_numpy(a), torch.from_numpy(b))
    e1, e2, n1, n2 = nb.chamfer_nn(torch.from_numpy(a), torch.from_numpy(b), with_indices=False)
    assert torch.equal(e1, d1) and torch.equal(e2, d2) and n1.numel() == 0 and n2.numel() == 0  # distance-only kernel
    for k in range(B):
        m1, m2, j1, j2 = OC.nn_dist(a[k], b[k])
        assert np.abs(d1[k].cpu().numpy() - m1).max() < 1e-6
        assert np.abs(d2[k].cpu().numpy() - m2).max() < 1e-6
        # indices: the chosen neighbour realises the minimum (ties may differ in fp32 vs fp64)
        g1 = np.linalg.norm(a[k].astype(np.float64) - b[k][i1[k].cpu().numpy()].astype(np.float64), axis=1)
        assert np.abs(g1 - m1).max() < 1e-6
        g2 = np.linalg.norm(b[k].astype(np.float64) - a[k][i2[k].cpu().numpy()].astype(np.float64), axis=1)
        assert np.abs(g2 - m2).max() < 1e-6


def test_variants_match_golden_and_oracle(golden_dir):
    import nova_pointcloud_b200 as nb

    d = np.load(os.path.join(golden_dir, "chamfer.npz"))
    a, b = d["a"], d["b"]
    for i in range(a.shape[0]):
        assert abs(nb.chamfer_distance(a[i], b[i]) - d["cd_a"][i]) < 1e-6
    cd = nb.chamfer_distance(a, b)
    assert np.abs(cd.cpu().numpy() - d["cd_a"]).max() < 1e-6
    dl, dr = nb.dist_chamfer(a, b)
    assert abs(float(dl) - d["cd_b"][0]) < 1e-4 and abs(float(dr) - d["cd_b"][1]) < 1e-4
    assert abs(float(nb.robust_chamfer_distance(a, b)) - OC.chamfer_b(a, b)[2]) < 1e-6
    sel = [0, 2]
    assert abs(float(nb.compute_chamfer_distance(a[sel], b[sel])) - float(d["cd_c"])) < 1e-4
    assert abs(float(nb.compute_chamfer_distance(a[sel], b[sel])) - OC.chamfer_c(a[sel], b[sel])) < 1e-5


def test_edge_cases():
    import nova_pointcloud_b200 as nb

    p = torch.randn(50, 3)
    assert nb.chamfer_distance(p, p) == 0.0
    d1, d2, i1, i2 = nb.chamfer_nn(p, p)
    assert torch.equal(i1.cpu()[0], torch.arange(50, dtype=torch.int32))
    with pytest.raises(nb.NovaError):
        nb.chamfer_nn(torch.zeros(1, 0, 3), torch.zeros(1, 4, 3))  # empty cloud: the reference raises too
    with pytest.raises(nb.NovaError):
        nb.chamfer_nn(torch.zeros(1, 4, 2), torch.zeros(1, 4, 2))
    # duplicated points: ties resolve to the lowest index, like argmin
    q = torch.tensor([[0.0, 0, 0], [1, 0, 0], [1, 0, 0], [0, 0, 0]])
    _, _, i1, _ = nb.chamfer_nn(torch.tensor([[1.0, 0, 0]]), q)
    assert int(i1[0, 0]) == 1


def test_full_size_properties():
    """BASELINE cfg5: 256 x (2048 vs 2048).  Symmetry, self-distance, permutation invariance,
    translation covariance, plus a scipy spot check on two pairs."""
    import nova_pointcloud_b200 as nb

    a = nb.synth.make_clouds(256, 2048, 11)
    b = nb.synth.make_clouds(256, 2048, 12)
    d1, d2, i1, i2 = nb.chamfer_nn(a, b)
    e1, e2, j1, j2 = nb.chamfer_nn(b, a)
    assert torch.equal(d1, e2) and torch.equal(d2, e1) and torch.equal(i1, j2) and torch.equal(i2, j1)
    z1, z2, _, _ = nb.chamfer_nn(a, a)
    assert float(z1.abs().max()) == 0.0 and float(z2.abs().max()) == 0.0
    perm = torch.randperm(2048, device=a.device)
    p1, p2, _, _ = nb.chamfer_nn(a[:, perm].contiguous(), b)
    assert torch.equal(p1, d1[:, perm]) and torch.equal(p2, d2)
    # the reported neighbour realises the reported distance
    nb_pts = torch.gather(b, 1, i1.long().unsqueeze(-1).expand(-1, -1, 3))
    assert float(((a - nb_pts).norm(dim=-1) - d1).abs().max()) < 1e-6
    for k in (0, 255):
        m1, m2, _, _ = OC.nn_dist(a[k].cpu().numpy(), b[k].cpu().numpy())
        assert np.abs(d1[k].cpu().numpy() - m1).max() < 1e-6 and np.abs(d2[k].cpu().numpy() - m2).max() < 1e-6
    cd = nb.chamfer_distance(a, b)
    assert cd.shape == (256,)
    assert abs(float(cd[255]) - (m1.mean() + m2.mean())) < 1e-6


@pytest.mark.parametrize("B,N,M", [(256, 2048, 2048), (3, 1000, 777), (2, 33, 4097), (5, 1, 1), (1, 2049, 31)])
def test_packed_sweep_is_bit_identical_to_the_scalar_sweep(monkeypatch, B, N, M):
    """The one-sweep distance-only kernel on packed fp32 pairs (FADD2 / FMUL2 / FFMA2 + three-input minima, the default;
    4 or 8 queries per lane) evaluates (t - q)^2 where the scalar kernel evaluates (q - t)^2: the same bits."""
    import nova_pointcloud_b200 as nb

    g = torch.Generator().manual_seed(B * 31 + N + M)
    a = torch.randn(B, N, 3, generator=g).cuda()
    b = torch.randn(B, M, 3, generator=g).cuda()
    out = {}
    idx = {}
    for variant in ("0", "1", "8"):
        monkeypatch.setenv("NOVA_B200_CHAMFER_PACKED", variant)
        d1, d2, _, _ = nb.chamfer_nn(a, b, with_indices=False)
        out[variant] = (d1.clone(), d2.clone())
        idx[variant] = [t.clone() for t in nb.chamfer_nn(a, b)]  # the indexed two-sweep kernel: scalar (0) / packed pairs
    monkeypatch.delenv("NOVA_B200_CHAMFER_PACKED")
    for j in range(4):
        assert torch.equal(idx["1"][j], idx["0"][j]), j  # same distances, same arg-min (ties included)
    for variant in ("1", "8"):
        assert torch.equal(out[variant][0], out["0"][0]) and torch.equal(out[variant][1], out["0"][1]), variant
    d1, d2, _, _ = nb.chamfer_nn(a, b)  # the two-sweep kernel with indices
    assert torch.equal(d1, out["1"][0]) and torch.equal(d2, out["1"][1])


# ---------------------------------------------------------------- earth mover's distance
def _emd_clouds(B, N, seed, spread=1.0):
    g = np.random.default_rng(seed)
    a = g.uniform(-spread, spread, size=(B, N, 3)).astype(np.float32)
    b = (a[:, g.permutation(N)] * 0.9 + g.normal(0, 0.15, size=(B, N, 3))).astype(np.float32)  # a shuffled, perturbed copy
    return a, b


@pytest.mark.parametrize("N", [1, 2, 7, 33, 256, 700])
def test_emd_matches_hungarian(N):
    """nova_emd (auction algorithm) against scipy's linear_sum_assignment on float64 distances (demo.py:57-74): the mean
    matched distance is within eps = 1e-5 of the optimum (N * eps on the matching cost), the assignment is a permutation
    and its own cost is what the kernel reports."""
    import nova_pointcloud_b200 as nb

    a, b = _emd_clouds(3, N, 100 + N)
    got, assign = nb.emd_matching(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    for i in range(3):
        want = OC.emd(a[i], b[i])
        assert abs(float(got[i]) - want) <= 1.5e-5 + 1e-6 * want, (N, i, float(got[i]), want)
        perm = assign[i].cpu().numpy()
        assert sorted(perm.tolist()) == list(range(N))
        own = np.linalg.norm(a[i].astype(np.float64) - b[i][perm].astype(np.float64), axis=1).mean()
        assert abs(own - float(got[i])) <= 1e-6 * max(own, 1.0) + 1e-7
        assert float(got[i]) >= want - 1e-6  # never below the optimum (beyond fp32 summation noise)
    record(f"emd vs Hungarian (N={N}): worst abs error of the mean distance",
           max(abs(float(got[i]) - OC.emd(a[i], b[i])) for i in range(3)))


def test_emd_variants_and_properties():
    """The three reference surfaces (demo.py:57-74, train_newloss.py:352-377, test_optimize.py:395-414), identical
    clouds -> 0, a pure permutation -> 0, symmetry, determinism, unequal sizes rejected (the reference asserts)."""
    import nova_pointcloud_b200 as nb

    a, b = _emd_clouds(4, 512, 9, spread=2.5)  # beyond +-2: emd_approx's clamp matters
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    got = nb.emd_approx(ta, tb).cpu().numpy()
    want = OC.emd_approx(a, b)
    assert np.abs(got - want).max() <= 2e-5
    assert abs(float(nb.robust_emd(ta, tb)) - want.mean()) <= 2e-5
    full = np.array([OC.emd(a[i], b[i]) for i in range(4)])
    assert abs(float(nb.compute_emd(ta, tb)) - min(max(full.mean(), 0.0), 10.0)) <= 2e-5
    assert abs(nb.earth_mover_distance(a[0], b[0]) - full[0]) <= 2e-5  # (N,3) numpy in, float out
    assert float(nb.emd_matching(ta, ta)[0].abs().max()) == 0.0
    perm = torch.randperm(512, device="cuda")
    assert float(nb.emd_matching(ta, ta[:, perm].contiguous())[0].abs().max()) <= 1e-5
    fwd, bwd = nb.emd_matching(ta, tb)[0], nb.emd_matching(tb, ta)[0]
    assert float((fwd - bwd).abs().max()) <= 2e-5
    again = nb.emd_matching(ta, tb)
    assert torch.equal(again[0], fwd) and torch.equal(again[1], nb.emd_matching(ta, tb)[1])
    with pytest.raises(nb.NovaError):
        nb.emd_matching(ta, tb[:, :100].contiguous())


def test_emd_full_size_cfg5():
    """BASELINE configs[4] shapes: 2048 x 2048 points.  One pair against the Hungarian oracle (~1 s of scipy), 16 pairs
    for the size-independent properties: a valid permutation, cost >= the Chamfer lower bound (every point's matched
    distance is at least its nearest-neighbour distance), all pairs converged."""
    import nova_pointcloud_b200 as nb

    a, b = _emd_clouds(16, 2048, 77)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out, assign, status = torch.ops.nova_b200.emd(ta, tb, 1e-5)
    assert bool((status > 0).all())
    want = OC.emd(a[0], b[0])
    record("emd vs Hungarian (N=2048): abs error of the mean distance", abs(float(out[0]) - want))
    assert abs(float(out[0]) - want) <= 1.5e-5 + 1e-6 * want
    d1, d2, _, _ = nb.chamfer_nn(ta, tb, with_indices=False)
    assert bool((out >= torch.maximum(d1.mean(1), d2.mean(1)) - 1e-6).all())
    srt = assign.sort(dim=1).values.cpu()
    assert torch.equal(srt, torch.arange(2048, dtype=torch.int32).expand(16, -1))


def test_emd_thread_count_variants_are_bit_identical(monkeypatch):
    """nova_emd runs a pair with 1024 threads (one CTA per SM) or, when there are more pairs than SMs, with 512 (two
    CTAs per SM).  Bids, prices and the compaction order do not depend on the thread count and the closing sum keeps the
    1024-thread order: same matching, same mean, same round count.  The bidding loop steers by an approximate square root
    (the default; NOVA_B200_EMD_FAST_SQRT=0 restores the correctly rounded one), which may settle near-ties differently:
    same mean within the auction's own tolerance."""
    a, b = _emd_clouds(6, 700, 5)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    res = {}
    for threads in ("1024", "512"):
        monkeypatch.setenv("NOVA_B200_EMD_THREADS", threads)
        out, assign, status = torch.ops.nova_b200.emd(ta, tb, 1e-5)
        res[threads] = (out.clone(), assign.clone(), status.clone())
    for k in range(3):
        assert torch.equal(res["512"][k], res["1024"][k]), k
    monkeypatch.setenv("NOVA_B200_EMD_FAST_SQRT", "0")
    out, assign, status = torch.ops.nova_b200.emd(ta, tb, 1e-5)
    monkeypatch.delenv("NOVA_B200_EMD_FAST_SQRT")
    monkeypatch.delenv("NOVA_B200_EMD_THREADS")
    assert bool((status > 0).all())
    assert float((out - res["1024"][0]).abs().max()) <= 2e-5
    srt = assign.sort(dim=1).values.cpu()
    assert torch.equal(srt, torch.arange(700, dtype=torch.int32).expand(6, -1))
