"""GPU: training-mode forward (add_noise, per-token-timestep head, masked flow-matching loss) against the oracle and
the committed output of the reference's own get_losses."""

import os

import numpy as np
import pytest
import torch

from gpu_util import cpu_sd
from oracle import training as OT

pytestmark = pytest.mark.gpu


def _golden(golden_dir):
    from oracle import head as OH

    d = np.load(os.path.join(golden_dir, "losses.npz"))
    depth, D, Dc, patch, chan = [int(v) for v in d["cfg"]]
    return d, OH.init_state_dict(depth, D, Dc, patch, chan, seed=int(d["init_seed"]))  # the reference's default init


def _head(sd, cfg, dtype=torch.float32):
    import nova_pointcloud_b200 as nb

    depth, D, Dc, patch, chan = [int(v) for v in cfg]
    head = nb.DiffusionMLP(depth, D, Dc, patch_size=patch, image_dim=chan).eval()
    head.load_state_dict(sd)
    return head.to("cuda", dtype)


def test_get_losses_matches_reference_golden(golden_dir):
    import nova_pointcloud_b200 as nb

    d, sd = _golden(golden_dir)
    head = _head(sd, d["cfg"])
    sched = nb.FlowMatchEulerDiscreteScheduler(1000, shift=1.0)
    out = nb.get_losses(head, sched, torch.from_numpy(d["z"]).cuda(), torch.from_numpy(d["x"]).cuda(),
                        mask=torch.from_numpy(d["mask"]).cuda(), noise=torch.from_numpy(d["noise"]).cuda(),
                        timesteps=torch.from_numpy(d["t_idx"]).cuda())
    want = float(d["loss"])
    assert abs(float(out["loss"]) - want) <= 1e-5 * abs(want)  # fp32 bar: 1e-5 relative
    ref = OT.get_losses(sd, torch.from_numpy(d["z"]), torch.from_numpy(d["x"]), torch.from_numpy(d["noise"]),
                        torch.from_numpy(d["t_idx"]), torch.from_numpy(d["mask"]))
    got = out["loss_per_token"].cpu()
    assert float((got - ref["loss_per_token"]).abs().max()) <= 1e-5 * float(ref["loss_per_token"].abs().max())
    assert abs(float(out["weight_sum"]) - float(torch.from_numpy(d["mask"]).sum()) * 4) < 1e-3
    # bf16 head: same loss within the bf16 bar
    out16 = nb.get_losses(_head(sd, d["cfg"], torch.bfloat16), sched, torch.from_numpy(d["z"]).cuda().bfloat16(),
                          torch.from_numpy(d["x"]).cuda(), mask=torch.from_numpy(d["mask"]).cuda(),
                          noise=torch.from_numpy(d["noise"]).cuda(), timesteps=torch.from_numpy(d["t_idx"]).cuda())
    assert abs(float(out16["loss"]) - want) <= 2e-2 * abs(want)


def test_add_noise_bit_exact_and_scheduler_state():
    import nova_pointcloud_b200 as nb

    sched = nb.FlowMatchEulerDiscreteScheduler(1000, shift=3.0)
    g = torch.Generator().manual_seed(5)
    x, n = torch.randn(4, 33, 16, generator=g), torch.randn(4, 33, 16, generator=g)
    idx = OT.sample_timesteps((4, 33), generator=g)
    sig, tt = OT.training_tables(1000, 3.0)
    got = sched.add_noise(x.cuda(), n.cuda(), idx.cuda())
    assert torch.equal(got.cpu(), OT.add_noise(x, n, idx, sig))  # same three roundings as the reference expression
    assert torch.equal(sched.timestep.cpu(), tt[idx]) and sched.sigma.shape == (4, 33, 1)
    # image-layout samples with per-sample timesteps (B,): sigma broadcasts over every trailing dim
    xi, ni = torch.randn(3, 4, 6, 6, generator=g), torch.randn(3, 4, 6, 6, generator=g)
    ib = torch.tensor([0, 500, 999])
    assert torch.equal(sched.add_noise(xi.cuda(), ni.cuda(), ib.cuda()).cpu(), OT.add_noise(xi, ni, ib, sig))
    with pytest.raises(nb.NovaError):
        sched.add_noise(x.cuda(), n.cuda(), torch.full((4, 33), 1000).cuda())  # index outside the table
    sched.set_timesteps(25)
    with pytest.raises(nb.NovaError):
        sched.add_noise(x.cuda(), n.cuda(), idx.cuda())  # the inference schedule replaced the training tables
    draws = nb.FlowMatchEulerDiscreteScheduler().sample_timesteps((1000, 64), device="cuda")
    assert draws.dtype == torch.int64 and int(draws.min()) >= 0 and int(draws.max()) <= 999
    assert abs(float(draws.float().mean()) - 499.5) < 5.0  # sigmoid of a standard normal is symmetric about 1/2


def test_flow_loss_properties_full_size():
    """4 x 32 clouds x 2048 tokens (loss_repeat x cfg2 batch): a perfect prediction has zero loss, the loss is
    quadratic in the error, invariant to a token permutation up to summation order, and deterministic."""
    pred = torch.randn(128, 2048, 3, device="cuda")
    noise = torch.randn(128, 2048, 3, device="cuda")
    x = torch.randn(128, 2048, 3, device="cuda")
    w = (torch.rand(128, 2048, device="cuda") < 0.6).float()
    tok, sums = torch.ops.nova_b200.flow_loss(noise - x, noise, x, w)
    assert float(sums[0]) == 0.0 and float(sums[1]) == float(w.sum())
    tok1, s1 = torch.ops.nova_b200.flow_loss(pred, noise, x, w)
    tok2, s2 = torch.ops.nova_b200.flow_loss(pred, noise, x, w)
    assert torch.equal(tok1, tok2) and torch.equal(s1, s2)
    want = (((pred - (noise - x)).double() ** 2).mean(-1) * w.double()).sum() / (w.double().sum() + 1e-5)
    assert abs(float(s1[0]) - float(want)) <= 1e-5 * float(want)
    err = pred - (noise - x)
    _, s4 = torch.ops.nova_b200.flow_loss(noise - x + 2 * err, noise, x, w)
    assert abs(float(s4[0]) - 4 * float(s1[0])) <= 1e-5 * float(s4[0])
    perm = torch.randperm(2048, device="cuda")
    _, sp = torch.ops.nova_b200.flow_loss(pred[:, perm].contiguous(), noise[:, perm].contiguous(),
                                          x[:, perm].contiguous(), w[:, perm].contiguous())
    assert abs(float(sp[0]) - float(s1[0])) <= 1e-5 * float(s1[0])
    _, s_none = torch.ops.nova_b200.flow_loss(pred, noise, x, None)
    assert abs(float(s_none[1]) - 128 * 2048) < 1.0
