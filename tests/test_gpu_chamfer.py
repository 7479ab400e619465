"""GPU: Chamfer kernel against scipy float64 (small), the committed golden values, and
size-independent properties at BASELINE's full size (256 x 2048 vs 2048)."""

import os

import numpy as np
import pytest
import torch

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
