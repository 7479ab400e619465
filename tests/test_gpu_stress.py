"""GPU: randomised shapes through every tcgen05 kernel class of the fused large-M step (modulation GEMM with TMA-staged x,
fc1, fc2 with partial statistics, gate GEMM with the block tail) and through the batched weight-gradient GEMM.

compute-sanitizer is closed on this GPU pool, so the hand-rolled mbarrier / TMEM / TMA protocols are exercised the other
way round: row counts that are NOT multiples of the 256-row tile pair or even of 8 (clipped stores, zero-filled loads,
epilogue warps without rows), tile counts from one per CTA pair to several, every bounded wait armed (a protocol error
traps with a debug word instead of hanging), results against the fp32 handle on the same weights."""

import numpy as np
import pytest
import torch

from gpu_util import record, relmax

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def no_barrier_timeouts():
    from nova_pointcloud_b200 import _lib

    _lib.lib().nova_debug_words_clear()
    yield
    words = _lib.debug_words()
    assert (words[0] >> 16) != 0xDEAD, "tcgen05 barrier timeout words: %s" % [hex(w) for w in words]


def _heads(D, depth, seed):
    import nova_pointcloud_b200 as nb

    h32 = nb.synth.make_head(D, depth, seed=seed, dtype=torch.float32)
    sd = {k: v.detach().bfloat16().float() for k, v in h32.state_dict().items()}
    h32.load_state_dict(sd)
    h16 = nb.synth.make_head(D, depth, seed=seed, dtype=torch.bfloat16)
    h16.load_state_dict(sd)
    return h32, h16


@pytest.mark.parametrize("D,rows", [(768, 3571), (768, 4099), (768, 7777), (768, 16385), (1024, 2053), (1024, 9001),
                                    (1536, 1031), (1536, 5555)])
def test_fused_step_ragged_rows(D, rows):
    """One velocity prediction over a single cloud of `rows` tokens (all above the width's switch point, so the
    fused-AdaLN dataflow with the tail epilogue runs): bf16 handle against the fp32 handle, the bf16 bar of 2e-2."""
    h32, h16 = _heads(D, 2, seed=rows)
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(1, 3, rows, 1, generator=g).cuda()
    z = torch.randn(1, rows, D, generator=g).bfloat16().cuda()
    t = torch.full((1,), 371.0).cuda()
    v16 = h16(x.bfloat16(), t, z).float()
    v32 = h32(x.bfloat16().float(), t, z.float())
    err = relmax(v16, v32)
    record(f"fused step, ragged rows (D={D}, rows={rows}): bf16 vs fp32 handle", err)
    assert bool(torch.isfinite(v16).all()) and err < 2e-2


def test_fused_sampling_random_batches_are_reproducible_and_batch_invariant():
    """Random cloud counts / token counts above the switch point: two runs of the same call are bit-identical and a
    cloud sampled inside the batch equals the same cloud sampled alone (rows are independent) -- the test that caught the
    all-dependent-launch chain in round 2."""
    import nova_pointcloud_b200 as nb

    head = nb.synth.make_head(768, 3, dtype=torch.bfloat16)
    sched = nb.FlowMatchEulerDiscreteScheduler()
    sched.set_timesteps(5)
    rng = np.random.default_rng(5)
    for _ in range(4):
        B, N = int(rng.integers(3, 9)), int(rng.integers(1800, 2300))
        noise, z = nb.synth.make_inputs(B, N, 768, seed=B * 10000 + N, dtype=torch.bfloat16)
        a = nb.denoise(head, sched, z, noise)
        b = nb.denoise(head, sched, z, noise)
        assert torch.equal(a, b), (B, N)
        pair = nb.denoise(head, sched, z[1:3], noise[1:3])  # 2 clouds: still above the switch point (> 3560 rows)
        assert torch.equal(pair, a[1:3]), (B, N)


@pytest.mark.parametrize("rows", [513, 1000, 4097, 12345])
def test_weight_gradient_split_reduction_ragged_rows(rows):
    """nova_head_backward at row counts that do not divide into the split's 64-row chunks (zero-padded transposes,
    1 to 16 batches in the batched GEMM): bf16 handle against the fp32 handle."""
    import nova_pointcloud_b200 as nb
    from nova_pointcloud_b200 import ops

    h32, h16 = _heads(256, 2, seed=rows)
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, 3, generator=g).cuda()
    t = (torch.rand(rows, generator=g) * 1000).cuda()
    z = torch.randn(rows, 256, generator=g).bfloat16().cuda()
    dv = (torch.randn(rows, 3, generator=g) / rows).cuda()
    shapes = {k: tuple(p.shape) for k, p in h32.named_parameters()}
    out = {}
    for name, head, zz in (("f32", h32, z.float()), ("bf16", h16, z)):
        v, ws = ops.head_train_forward(head.handle(), x, t, zz)
        out[name] = ops.head_backward(head.handle(), dv, x, zz, ws, shapes)
    worst = max(relmax(out["bf16"][0][k], out["f32"][0][k]) for k in shapes)
    record(f"backward, ragged rows ({rows}): bf16 vs fp32 handle, worst parameter gradient", worst)
    assert worst < 2e-2 and relmax(out["bf16"][1], out["f32"][1]) < 2e-2
