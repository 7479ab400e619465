"""GPU: training mode of the head -- add_noise, the per-token-timestep forward, the masked flow-matching loss and the
backward pass (nova_head_train_forward / nova_head_backward) -- against the oracle (torch autograd through
oracle/head.py) and the committed output of the reference's own get_losses."""

import os

import numpy as np
import pytest
import torch

from gpu_util import cpu_sd, record, relmax
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
    with torch.no_grad():  # evaluation: nova_head_forward + nova_flow_loss, nothing kept for a backward
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
    with torch.no_grad():
        out16 = nb.get_losses(_head(sd, d["cfg"], torch.bfloat16), sched, torch.from_numpy(d["z"]).cuda().bfloat16(),
                              torch.from_numpy(d["x"]).cuda(), mask=torch.from_numpy(d["mask"]).cuda(),
                              noise=torch.from_numpy(d["noise"]).cuda(), timesteps=torch.from_numpy(d["t_idx"]).cuda())
    assert abs(float(out16["loss"]) - want) <= 2e-2 * abs(want)


def _oracle_loss_and_grads(sd, z, x, noise, t_idx, mask, R):
    """torch autograd through the oracle's restatement of get_losses (transformer_3d.py:81-95) on the CPU, fp32."""
    leaves = {k: v.detach().clone().float().requires_grad_(True) for k, v in sd.items()}
    zl = z.detach().clone().float().requires_grad_(True)
    out = OT.get_losses(leaves, zl, x.float(), noise.float(), t_idx, mask, loss_repeat=R)
    out["loss"].backward()
    return float(out["loss"].detach()), {k: v.grad for k, v in leaves.items()}, zl.grad


def _train_case(depth, D, Dc, patch, chan, B, H, W, R, seed, dtype):
    import nova_pointcloud_b200 as nb
    from oracle import head as OH

    sd = OH.init_state_dict(depth, D, Dc, patch, chan, seed=seed)
    if dtype == torch.bfloat16:
        sd = {k: v.bfloat16().float() for k, v in sd.items()}  # the oracle differentiates the weights the GPU head holds
    g = torch.Generator().manual_seed(seed + 1)
    N, T = H * W, patch * patch * chan
    x = torch.randn(B, chan, H * patch, W * patch, generator=g)
    z = torch.randn(B, N, Dc, generator=g)
    if dtype == torch.bfloat16:
        z = z.bfloat16().float()
    mask = (torch.rand(B, N, 1, generator=g) < 0.7).float()
    noise = torch.randn(R * B, N, T, generator=g)
    t_idx = OT.sample_timesteps((R * B, N), generator=g)
    head = nb.DiffusionMLP(depth, D, Dc, patch_size=patch, image_dim=chan).train()
    head.load_state_dict(sd)
    return sd, head.to("cuda", dtype), x, z, mask, noise, t_idx


def _gpu_loss_and_grads(head, x, z, mask, noise, t_idx, R):
    import nova_pointcloud_b200 as nb

    sched = nb.FlowMatchEulerDiscreteScheduler(1000, shift=1.0)
    head.zero_grad(set_to_none=True)
    zc = z.to("cuda", head.dtype).requires_grad_(True)
    out = nb.get_losses(head, sched, zc, x.cuda(), mask=mask.cuda(), loss_repeat=R, noise=noise.cuda(), timesteps=t_idx.cuda())
    out["loss"].backward()
    return float(out["loss"].detach()), {k: p.grad.detach().float().cpu() for k, p in head.named_parameters()}, zc.grad.float().cpu()


def test_backward_matches_autograd_fp32_on_the_reference_golden(golden_dir):
    """The reference's own get_losses case (tests/golden/losses.npz): loss through the training-mode forward, and
    every gradient of loss.backward() against autograd through the oracle -- fp32 handle, <= 5e-5 of each tensor's max
    (measured values recorded under profiles/)."""
    d, sd = _golden(golden_dir)
    head = _head(sd, d["cfg"]).train()
    z, x, mask = torch.from_numpy(d["z"]), torch.from_numpy(d["x"]), torch.from_numpy(d["mask"])
    noise, t_idx = torch.from_numpy(d["noise"]), torch.from_numpy(d["t_idx"])
    loss, grads, dz = _gpu_loss_and_grads(head, x, z, mask, noise, t_idx, 4)
    want = float(d["loss"])
    assert abs(loss - want) <= 1e-5 * abs(want)  # the reference's own number
    _, ref, ref_dz = _oracle_loss_and_grads(sd, z, x, noise, t_idx, mask, 4)
    assert set(grads) == set(ref)
    worst = max(relmax(grads[k], ref[k]) for k in ref)
    record("training backward fp32 (reference golden case): worst parameter gradient", worst)
    for k in ref:
        assert grads[k].shape == ref[k].shape and relmax(grads[k], ref[k]) <= 5e-5, k  # measured 1.2e-5
    assert relmax(dz, ref_dz) <= 5e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-5), (torch.bfloat16, 2e-2)])  # measured 2.2e-5 / 6.5e-3
def test_backward_matches_autograd_two_blocks_patch2(dtype, tol):
    """depth 2, patch 2 x 2 x 4 channels (T = 16: the Conv2d-layout weight gradient), Dc = 64, 150 rows (not a multiple
    of anything), masked loss, loss_repeat 2.  bf16: against the fp32 oracle on the bf16-rounded weights."""
    sd, head, x, z, mask, noise, t_idx = _train_case(2, 256, 64, 2, 4, 3, 5, 5, 2, 11, dtype)
    loss, grads, dz = _gpu_loss_and_grads(head, x, z, mask, noise, t_idx, 2)
    want, ref, ref_dz = _oracle_loss_and_grads(sd, z, x, noise, t_idx, mask, 2)
    assert abs(loss - want) <= (1e-5 if dtype == torch.float32 else 2e-2) * abs(want)
    errs = {k: relmax(grads[k], ref[k]) for k in ref}
    record(f"training backward {str(dtype).split('.')[-1]} (depth 2, T 16, 150 rows): worst parameter gradient", max(errs.values()))
    record(f"training backward {str(dtype).split('.')[-1]} (depth 2, T 16, 150 rows): dz", relmax(dz, ref_dz))
    for k, e in errs.items():
        assert e <= tol, (k, e)
    assert relmax(dz, ref_dz) <= tol


@pytest.mark.parametrize("depth,D,B,N", [
    (6, 768, 3, 111),     # NOVA-0.3B head as BASELINE's cfg2 has it, 333 rows (ragged against every tile size)
    (2, 1024, 2, 331),    # NOVA-0.6B width, 662 rows: three 256-row blocks, the last ragged
    (2, 1536, 1, 257),    # NOVA-1.4B width, one row into the second 256-row block
])
def test_backward_bf16_at_the_registry_widths(depth, D, B, N):
    """The bf16 training step (tcgen05 dgrad / wgrad GEMMs, MN-major weight-gradient operands) at the widths of the
    registry heads against torch autograd through the fp32 oracle on the bf16-rounded weights: <= 2e-2 of each tensor's
    maximum (north_star's bf16 tolerance), loss included."""
    sd, head, x, z, mask, noise, t_idx = _train_case(depth, D, D, 1, 3, B, N, 1, 1, 23, torch.bfloat16)
    loss, grads, dz = _gpu_loss_and_grads(head, x, z, mask, noise, t_idx, 1)
    want, ref, ref_dz = _oracle_loss_and_grads(sd, z, x, noise, t_idx, mask, 1)
    assert abs(loss - want) <= 2e-2 * abs(want)
    errs = {k: relmax(grads[k], ref[k]) for k in ref}
    record(f"training backward bf16 (depth {depth}, D {D}, {B * N} rows): worst parameter gradient", max(errs.values()))
    record(f"training backward bf16 (depth {depth}, D {D}, {B * N} rows): dz", relmax(dz, ref_dz))
    for k, e in errs.items():
        assert e <= 2e-2, (k, e)
    assert relmax(dz, ref_dz) <= 2e-2


def test_backward_split_reduction_on_tensor_cores_matches_fp32_handle():
    """4096 rows: the bf16 handle's weight gradients run as ONE batched tcgen05 launch over a split M-reduction
    (tc::launch_batched); compared with the fp32 handle (SIMT GEMMs, itself pinned to autograd above) on the same
    bf16-rounded weights.  Also: a repeat is bit-identical (deterministic reductions, no atomics)."""
    sd, head16, x, z, mask, noise, t_idx = _train_case(2, 256, 64, 1, 3, 4, 32, 32, 1, 5, torch.bfloat16)
    import nova_pointcloud_b200 as nb

    head32 = nb.DiffusionMLP(2, 256, 64, patch_size=1, image_dim=3).train()
    head32.load_state_dict(sd)
    head32 = head32.cuda()
    l16, g16, dz16 = _gpu_loss_and_grads(head16, x, z, mask, noise, t_idx, 1)
    l16b, g16b, dz16b = _gpu_loss_and_grads(head16, x, z, mask, noise, t_idx, 1)
    assert l16 == l16b and all(torch.equal(g16[k], g16b[k]) for k in g16) and torch.equal(dz16, dz16b)
    l32, g32, dz32 = _gpu_loss_and_grads(head32, x, z, mask, noise, t_idx, 1)
    assert abs(l16 - l32) <= 2e-2 * abs(l32)
    errs = {k: relmax(g16[k], g32[k]) for k in g32}
    record("training backward bf16 vs fp32 handle (4096 rows, split reduction): worst parameter gradient", max(errs.values()))
    for k, e in errs.items():
        assert e <= 2e-2, (k, e)  # measured 5e-3
    assert relmax(dz16, dz32) <= 2e-2


def test_backward_only_requested_gradients_and_frozen_parameters():
    """Frozen parameters get no gradient (and cost no GEMM); z without requires_grad gets none either."""
    import nova_pointcloud_b200 as nb

    sd, head, x, z, mask, noise, t_idx = _train_case(1, 256, 64, 1, 3, 2, 4, 4, 1, 3, torch.float32)
    for k, p in head.named_parameters():
        p.requires_grad_(k.startswith("blocks.0.proj"))
    sched = nb.FlowMatchEulerDiscreteScheduler(1000, shift=1.0)
    out = nb.get_losses(head, sched, z.cuda(), x.cuda(), loss_repeat=1, noise=noise.cuda(), timesteps=t_idx.cuda())
    out["loss"].backward()
    got = {k for k, p in head.named_parameters() if p.grad is not None}
    assert got == {k for k, _ in head.named_parameters() if k.startswith("blocks.0.proj")}
    _, ref, _ = _oracle_loss_and_grads(sd, z, x, noise, t_idx, None, 1)
    for k in got:
        assert relmax(dict(head.named_parameters())[k].grad, ref[k]) <= 1e-4


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
