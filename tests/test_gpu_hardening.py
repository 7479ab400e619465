"""Robustness of the C ABI and its Python mirror: out-of-range pred_ids, several devices in one process, the
empty-set call without a workspace, the single-rank all-gather."""

import ctypes as C

import pytest
import torch

from gpu_util import make_case

pytestmark = pytest.mark.gpu


def test_out_of_range_pred_ids_never_index_and_are_flagged():
    """An id < 0 or >= N must not read or write out of bounds (the reference's gather / scatter raises): the kernels
    skip it and raise the library's bad-ids flag; strict mode raises IndexError before anything is launched."""
    import nova_pointcloud_b200 as nb
    from nova_pointcloud_b200 import _lib, ops

    head, x, z, _, ids = make_case(2, 256, 64, 3, 40, 1, n_pred=6)
    head = head.cuda()
    sched = nb.FlowMatchEulerDiscreteScheduler()
    sched.set_timesteps(3)
    good = nb.denoise(head, sched, z.cuda(), x.cuda(), None, None, ids.cuda())
    torch.cuda.synchronize()
    _lib.bad_pred_ids_seen(clear=True)
    guard = torch.full((4096,), 7.0, device="cuda")  # neighbours of the output in the caching allocator
    bad = ids.clone()
    bad[1, 2, 0] = 40          # == N
    bad[2, 5, 0] = -3
    out = nb.denoise(head, sched, z.cuda(), x.cuda(), None, None, bad.cuda())
    torch.cuda.synchronize()
    assert _lib.bad_pred_ids_seen(clear=True)
    assert bool(torch.isfinite(out).all()) and bool((guard == 7.0).all())
    assert torch.equal(out[0], good[0])  # the cloud whose ids are all valid is untouched by its neighbours' bad ids
    assert not _lib.bad_pred_ids_seen()
    ops.set_strict_ids(True)
    try:
        with pytest.raises(IndexError):
            nb.denoise(head, sched, z.cuda(), x.cuda(), None, None, bad.cuda())
        nb.denoise(head, sched, z.cuda(), x.cuda(), None, None, ids.cuda())  # valid ids pass
    finally:
        ops.set_strict_ids(False)


def test_empty_set_needs_no_workspace_through_the_c_abi():
    """nova_head_sample with pred_ids given and n == 0: every token follows the x <- x + dt x recurrence; the call
    must not touch a workspace (NULL, 0 bytes is what nova_head_workspace_bytes asks for)."""
    import nova_pointcloud_b200 as nb
    from nova_pointcloud_b200 import _lib

    head, x, z, _, _ = make_case(1, 256, 64, 2, 8, 1)
    head = head.cuda()
    h = head.handle()
    tok = x.squeeze(-1).transpose(1, 2).float().cuda().contiguous()  # (B,3,N,1) -> tokens (B,N,3), patch size 1
    zz = z.cuda().contiguous()
    ids = torch.zeros(1, dtype=torch.int64, device="cuda")
    out = torch.empty_like(tok)
    ts = (C.c_float * 3)(900.0, 500.0, 100.0)
    sg = (C.c_double * 4)(0.9, 0.5, 0.1, 0.0)
    g = _lib.Guidance(1.0, 0.0, 1.0)
    rc = _lib.lib().nova_head_sample(h._h, C.c_void_p(tok.data_ptr()), C.c_void_p(zz.data_ptr()), C.c_void_p(ids.data_ptr()),
                                     2, 2, 8, 0, ts, sg, 3, C.byref(g), C.c_void_p(out.data_ptr()), None, 0,
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "nova_head_sample")
    torch.cuda.synchronize()
    want = tok.clone()
    for i in range(3):
        dt = torch.tensor(sg[i + 1] - sg[i], dtype=torch.float32)
        want = want * dt + want
    assert torch.allclose(out, want, rtol=1e-6, atol=1e-7)


def test_single_rank_allgather_through_the_c_abi():
    """nova_comm_* / nova_allgather on a one-rank communicator: the gather is a copy (the 2-rank run is
    scripts/c_abi_allgather_2gpu.py under gpurun --gpus 2)."""
    from nova_pointcloud_b200 import _lib

    lib = _lib.lib()
    uid = C.create_string_buffer(128)
    _lib.check(lib.nova_comm_unique_id(uid), "nova_comm_unique_id")
    comm = C.c_void_p()
    _lib.check(lib.nova_comm_init_rank(uid, 1, 0, C.byref(comm)), "nova_comm_init_rank")
    send = torch.arange(3 * 64, dtype=torch.float32, device="cuda")
    recv = torch.zeros_like(send)
    _lib.check(lib.nova_allgather(comm, C.c_void_p(send.data_ptr()), C.c_void_p(recv.data_ptr()), send.numel() * 4,
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)), "nova_allgather")
    torch.cuda.synchronize()
    assert torch.equal(send, recv)
    _lib.check(lib.nova_comm_destroy(comm), "nova_comm_destroy")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_one_process_two_devices():
    """Function attributes (dynamic shared memory) and the SM count are per device: a process that drives two GPUs
    must get the same bits from both."""
    import nova_pointcloud_b200 as nb

    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        head = nb.synth.make_head(768, 2, dtype=torch.bfloat16, device=dev)
        sched = nb.FlowMatchEulerDiscreteScheduler()
        sched.set_timesteps(3)
        noise, z = nb.synth.make_inputs(3, 700, 768, dtype=torch.bfloat16, device=dev)  # fused dataflow (2100 rows) ...
        big = nb.denoise(head, sched, z.repeat(2, 1, 1), noise.repeat(2, 1, 1, 1))
        small = nb.denoise(head, sched, z[:1, :100].contiguous(), noise[:1, :, :100].contiguous())  # ... and the chain kernel
        outs.append((big.cpu(), small.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
