"""GPU: kNN / local density / soft interpolation kernels against the float64 oracle (small), the committed
golden outputs of the reference's own functions, and size-independent properties at 256 x 2048 points."""

import os

import numpy as np
import pytest
import torch

from oracle import geometry as OG

pytestmark = pytest.mark.gpu


def clouds(B, N, seed):
    return np.random.default_rng(seed).uniform(-1, 1, (B, N, 3)).astype(np.float32)


@pytest.mark.parametrize("B,Nq,Nt,k", [(1, 1, 1, 1), (2, 37, 129, 4), (3, 300, 1500, 9), (1, 130, 2048, 16),
                                       (2, 64, 1025, 32), (1, 5, 7, 7)])
def test_knn_matches_oracle(B, Nq, Nt, k):
    import nova_pointcloud_b200 as nb

    q, t = clouds(B, Nq, Nq + k), clouds(B, Nt, Nt + 1)
    d, i = nb.knn(torch.from_numpy(q), torch.from_numpy(t), k)
    assert d.shape == (B, Nq, k) and i.dtype == torch.int32
    d, i = d.cpu().numpy(), i.cpu().numpy()
    for b in range(B):
        od, _ = OG.knn(q[b], t[b], k)
        assert np.abs(d[b] - od).max() < 1e-6  # bar: 1e-6 abs vs scipy float64, as for Chamfer
        # the reported neighbours realise the reported distances, are distinct, and come out ascending
        real = np.linalg.norm(q[b][:, None, :].astype(np.float64) - t[b][i[b]].astype(np.float64), axis=-1)
        assert np.abs(real - d[b]).max() < 1e-6
        assert all(len(set(row)) == k for row in i[b])
        assert (np.diff(d[b], axis=1) >= 0).all()


@pytest.mark.parametrize("B,N,k", [(160, 1024, 4), (150, 1030, 9), (150, 1000, 16), (256, 2048, 9)])
def test_packed_pair_arithmetic_is_bit_identical_to_the_scalar_form(monkeypatch, B, N, k):
    """Two queries per thread (grids that still cover the chip) take their distances on packed fp32 pairs (FADD2 / FMUL2 /
    FFMA2, (t - q)^2 instead of (q - t)^2): same distances, same indices, same densities as NOVA_B200_KNN_PACKED=0."""
    import nova_pointcloud_b200 as nb

    a = torch.from_numpy(clouds(B, N, 7 * N + k)).cuda()
    out = {}
    for packed in ("0", "1"):
        monkeypatch.setenv("NOVA_B200_KNN_PACKED", packed)
        d, i = nb.knn(a, a, k)
        out[packed] = (d.clone(), i.clone(), nb.compute_local_density(a, k_neighbors=k - 1).clone())
    monkeypatch.delenv("NOVA_B200_KNN_PACKED")
    for j in range(3):
        assert torch.equal(out["0"][j], out["1"][j]), j
    od, _ = OG.knn(a[0].cpu().numpy(), a[0].cpu().numpy(), k)
    assert np.abs(out["1"][0][0].cpu().numpy() - od).max() < 1e-6


def test_knn_ties_and_errors():
    import nova_pointcloud_b200 as nb

    t = torch.tensor([[0.0, 0, 0], [1, 0, 0], [1, 0, 0], [0, 0, 0], [1, 0, 0]])
    d, i = nb.knn(torch.tensor([[1.0, 0, 0]]), t, 4)
    assert i.cpu().tolist() == [[1, 2, 4, 0]] and d.cpu().tolist() == [[0.0, 0.0, 0.0, 1.0]]  # ties: lowest index first
    with pytest.raises(nb.NovaError):
        nb.knn(t, t, 6)  # k > targets: torch.topk raises as well
    with pytest.raises(nb.NovaError):
        nb.knn(t, t, 0)
    with pytest.raises(nb.NovaError):
        nb.knn(t, t, 33)
    with pytest.raises(nb.NovaError):
        nb.knn(torch.zeros(1, 0, 3), torch.zeros(1, 4, 3), 1)
    with pytest.raises(nb.NovaError):
        nb.compute_local_density(torch.zeros(1, 8, 3), k_neighbors=8)  # needs k_neighbors + 1 points


def test_geometry_matches_reference_golden(golden_dir):
    """Outputs of the reference's own compute_local_density / feature_aware_interpolation (tests/make_golden.py)."""
    import nova_pointcloud_b200 as nb

    g = np.load(os.path.join(golden_dir, "geometry.npz"))
    # <= 25 points the reference's cdist is exact differences: tight; above it is the mm form: 1e-4
    assert np.abs(nb.compute_local_density(g["small"]).cpu().numpy() - g["density_small"]).max() < 1e-6
    assert np.abs(nb.compute_local_density(g["small"], 3).cpu().numpy() - g["density_small_k3"]).max() < 1e-6
    assert np.abs(nb.compute_local_density(g["big"]).cpu().numpy() - g["density_big"]).max() < 1e-4
    assert np.abs(nb.compute_local_density(g["big"]).cpu().numpy() - OG.local_density(g["big"])).max() < 1e-6
    for name, tol in (("small", 1e-6), ("big", 1e-4)):
        idx = g[f"interp_{name}_idx"]
        out = nb.feature_aware_interpolation(g[name], len(idx), indices=torch.from_numpy(idx)).cpu().numpy()
        assert np.abs(out - g[f"interp_{name}"]).max() < tol
        assert np.abs(out - OG.interpolate(g[name], len(idx), idx)).max() < 2e-6
    rep = nb.feature_aware_interpolation(g["small"], 60).cpu().numpy()
    assert np.array_equal(rep, g["interp_repeat"])
    one = nb.compute_local_density(g["small"][0])  # single cloud (N,3) -> (N,)
    assert one.shape == (24,)


def test_softmax_interp_far_from_origin():
    """exp(-d) of the nearest source underflows in fp32 once d > 88: the running-minimum form must not."""
    import nova_pointcloud_b200 as nb

    p = clouds(2, 400, 5) * 300.0
    tgt = clouds(2, 33, 6) * 300.0
    out = torch.ops.nova_b200.softmax_interp(torch.from_numpy(tgt).cuda(), torch.from_numpy(p).cuda()).cpu().numpy()
    ref = OG.softmax_interp(tgt, p)
    assert np.isfinite(out).all()
    assert np.abs(out - ref).max() / np.abs(ref).max() < 1e-5


def test_full_size_properties():
    """256 clouds x 2048 points (the cfg5 shape): self is the nearest neighbour at distance 0, density is
    permutation-equivariant bit for bit, scales with the cloud, and agrees with kNN; scipy spot check."""
    import nova_pointcloud_b200 as nb

    a = nb.synth.make_clouds(256, 2048, 11)
    d, i = nb.knn(a, a, 9)
    assert float(d[..., 0].abs().max()) == 0.0
    assert torch.equal(i[..., 0].long(), torch.arange(2048, device=a.device).expand(256, -1))
    dens = nb.compute_local_density(a)
    assert dens.shape == (256, 2048)
    # density == mean of kNN distances 1..8, summed in the same (ascending) order
    acc = torch.zeros_like(dens)
    for r in range(1, 9):
        acc = acc + d[..., r]
    assert torch.equal(dens, acc / 8.0)
    perm = torch.randperm(2048, device=a.device)
    assert torch.equal(nb.compute_local_density(a[:, perm].contiguous()), dens[:, perm])
    assert float((nb.compute_local_density(a * 2.0) - 2.0 * dens).abs().max()) < 1e-6  # power-of-two scale: exact up to sqrt
    # Chamfer's nearest neighbour of a against b equals kNN with k = 1
    b = nb.synth.make_clouds(256, 2048, 12)
    d1, _, i1, _ = nb.chamfer_nn(a, b)
    k1, j1 = nb.knn(a, b, 1)
    assert torch.equal(k1[..., 0], d1) and torch.equal(j1[..., 0], i1)
    for c in (0, 255):
        assert np.abs(dens[c].cpu().numpy() - OG.local_density(a[c:c + 1].cpu().numpy())[0]).max() < 1e-6
    # soft interpolation: a convex combination stays inside the bounding box; far-apart copies do not mix
    idx = torch.randperm(2048, device=a.device)[:512]
    out = nb.feature_aware_interpolation(a, 512, indices=idx)
    assert out.shape == (256, 512, 3)
    assert bool((out <= a.amax(dim=1, keepdim=True) + 1e-6).all()) and bool((out >= a.amin(dim=1, keepdim=True) - 1e-6).all())
    ref = OG.interpolate(a[:2].cpu().numpy(), 512, idx.cpu().numpy())
    assert np.abs(out[:2].cpu().numpy() - ref).max() < 2e-6


def test_dynamic_partition_and_target_size():
    import nova_pointcloud_b200 as nb

    pts = nb.synth.make_clouds(1, 1024, 3)
    g = torch.Generator().manual_seed(9)
    order, subsets = nb.dynamic_partition(pts, k=20, generator=g)
    assert sorted(order.cpu().tolist()) == list(range(20))
    assert [s.shape[1] for s in subsets] == [51] * 19 + [55]
    allpts = torch.cat(subsets, dim=1)[0]
    assert torch.equal(allpts.sort(dim=0).values, pts[0].sort(dim=0).values)  # a partition: every point exactly once
    dens = nb.compute_local_density(subsets[int(order[0])])
    size = nb.density_target_size(dens, 15000, 20)
    assert size == OG.target_size(float(dens.mean()), 15000, 20) and 100 <= size <= 1500


@pytest.mark.parametrize("B,N,S", [(3, 200, 50), (2, 2048, 512), (1, 5, 5), (4, 1000, 1)])
def test_farthest_point_sampling_matches_oracle_bit_for_bit(B, N, S):
    """Textbook FPS (the algorithm transformer_pointcloud_nova.py:100-125 is named after): same picks as the numpy
    float32 oracle, index for index; duplicated points exercise the lowest-index tie-break."""
    import nova_pointcloud_b200 as nb
    from oracle import geometry as OG

    g = torch.Generator().manual_seed(B * 1000 + N)
    pts = torch.rand(B, N, 3, generator=g) * 2 - 1
    if N >= 200:
        pts[:, 150] = pts[:, 7]  # exact duplicates: equal distances everywhere
        pts[:, 151] = pts[:, 7]
    start = torch.randint(0, N, (B,), generator=g)
    out, idx = nb.farthest_point_sampling(pts.cuda(), S, start_indices=start, return_indices=True)
    assert out.shape == (B, S, 3) and idx.shape == (B, S)
    for b in range(B):
        want = OG.farthest_point_sampling(pts[b].numpy(), S, int(start[b]))
        assert np.array_equal(idx[b].cpu().numpy(), want), b
        assert torch.equal(out[b].cpu(), pts[b][torch.from_numpy(want)])
    if S > 1 and N > S:  # property: picks are distinct points and the running minimum distance never grows
        for b in range(B):
            sel = pts[b][idx[b].cpu()]
            assert len(set(idx[b].tolist())) == S or N >= 200  # duplicates may be picked last only
            d = torch.cdist(sel, sel)
            gaps = [float(d[i, :i].min()) for i in range(1, S)]
            assert all(gaps[i] >= gaps[i + 1] - 1e-6 for i in range(len(gaps) - 1))


def test_farthest_point_sampling_reference_mode_is_the_exact_arithmetic_result():
    """mode="reference": the reference's loop in exact arithmetic picks [start, 0, 0, ...] (the zero diagonal wins every
    min; tests/test_oracle_vs_reference.py::test_geometry_live shows it on the live function for exact distances)."""
    import nova_pointcloud_b200 as nb

    pts = torch.rand(2, 30, 3, generator=torch.Generator().manual_seed(1)).cuda()
    start = torch.tensor([5, 11])
    out, idx = nb.farthest_point_sampling(pts, 4, start_indices=start, mode="reference", return_indices=True)
    assert idx.cpu().tolist() == [[5, 0, 0, 0], [11, 0, 0, 0]]
    assert torch.equal(out[:, 1:], pts[:, :1].expand(-1, 3, -1))
