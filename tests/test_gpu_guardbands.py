"""Guard bands around every caller-owned buffer of a C-ABI call (compute-sanitizer is closed on the GPU pool, so
out-of-bounds WRITES are looked for the plain way): inputs, output and an EXACTLY sized workspace are carved out
of one painted arena with 4 KB gaps, the call runs at ragged shapes through every dataflow (fp32 SIMT, wide + chain
kernel, fused large-M with the tail epilogue, guidance, pred_ids), and afterwards every gap byte and every input
byte must still hold what was put there."""

import ctypes as C

import pytest
import torch

from gpu_util import make_case

pytestmark = pytest.mark.gpu

GAP = 4096
PAINT = 0xA5


class Arena:
    def __init__(self, nbytes):
        self.buf = torch.full((nbytes,), PAINT, dtype=torch.uint8, device="cuda")
        self.off = GAP
        self.spans = []  # (offset, nbytes, snapshot or None)

    def put(self, t):
        """copy tensor t into the arena (an input: must come back unchanged); returns the device pointer"""
        raw = t.contiguous().view(torch.uint8).reshape(-1)
        o = self._take(raw.numel())
        self.buf[o:o + raw.numel()] = raw.cuda()
        self.spans.append((o, raw.numel(), raw.cuda().clone()))
        return C.c_void_p(self.buf.data_ptr() + o)

    def out(self, nbytes):
        """an output / scratch span of exactly nbytes; returns (pointer, offset)"""
        o = self._take(nbytes)
        self.spans.append((o, nbytes, None))
        return C.c_void_p(self.buf.data_ptr() + o), o

    def _take(self, nbytes):
        o = self.off
        self.off = (o + nbytes + GAP + 1023) // 1024 * 1024  # gap of >= 4 KB, spans 1 KB aligned (TMA needs 16 B)
        assert self.off <= self.buf.numel(), "arena too small"
        return o

    def view(self, o, shape, dtype):
        n = 1
        for d in shape:
            n *= d
        return self.buf[o:o + n * torch.empty(0, dtype=dtype).element_size()].view(dtype).view(*shape)

    def check(self):
        torch.cuda.synchronize()
        mask = torch.ones(self.buf.numel(), dtype=torch.bool, device="cuda")
        for o, n, snap in self.spans:
            mask[o:o + n] = False
            if snap is not None:
                assert torch.equal(self.buf[o:o + n], snap), f"input span at {o} (+{n}) was written"
        dirty = (self.buf != PAINT) & mask
        assert not bool(dirty.any()), f"guard bytes written at offsets {dirty.nonzero()[:8].flatten().tolist()}"


@pytest.mark.parametrize("case", ["fp32", "bf16_wide_chain", "bf16_chain_ids", "bf16_fused_tail", "bf16_fused_cfg", "bf16_wide_cfg_renorm"])
def test_sample_and_forward_write_only_what_they_own(monkeypatch, case):
    from nova_pointcloud_b200 import _lib

    bf16 = case != "fp32"
    if case.startswith("bf16_fused"):
        monkeypatch.setenv("NOVA_B200_WIDE_ADA_ROWS", "0")  # the large-M dataflow at a test-sized row count
    guided = "cfg" in case
    D = 768 if bf16 else 256
    Bx, N = 3, 77                                   # 231 rows: ragged against 64 / 128 / 256-row tiles
    n = 13 if case == "bf16_chain_ids" else N
    head, x, z, _, ids = make_case(2, D, D, Bx, N, 1, n_pred=None if n == N else n)
    head = (head.to(torch.bfloat16) if bf16 else head).cuda()
    h = head.handle()
    B = 2 * Bx if guided else Bx
    zz = torch.cat([z, z.flip(0)]) if guided else z
    zz = zz.to(torch.bfloat16 if bf16 else torch.float32)
    tok = x.squeeze(-1).transpose(1, 2).float().contiguous()  # (B,3,N,1) -> tokens (B,N,3)
    S = 4
    ts = (C.c_float * S)(1000.0, 700.0, 400.0, 100.0)
    sg = (C.c_double * (S + 1))(1.0, 0.7, 0.4, 0.1, 0.0)
    g = _lib.Guidance(3.0 if guided else 1.0, 0.0, 0.6 if case.endswith("renorm") else 1.0)
    lib = _lib.lib()
    ws_bytes = int(lib.nova_head_workspace_bytes(h._h, B * n, S))
    arena = Arena(ws_bytes + 64 * GAP + zz.numel() * zz.element_size() + 4 * tok.numel() * 4 + (1 << 20))
    p_tok, p_z = arena.put(tok), arena.put(zz)
    p_ids = arena.put(torch.cat([ids.squeeze(-1)] * (2 if guided else 1))) if ids is not None else None
    p_out, o_out = arena.out(Bx * N * 3 * 4)
    p_ws, _ = arena.out(ws_bytes)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):  # eager, graph capture, replay
        _lib.check(lib.nova_head_sample(h._h, p_tok, p_z, p_ids, B, Bx, N, n, ts, sg, S, C.byref(g), p_out, p_ws, ws_bytes,
                                        stream), "nova_head_sample")
    arena.check()
    out = arena.view(o_out, (Bx, N, 3), torch.float32)
    assert bool(torch.isfinite(out).all())
    # one forward call (per-cloud timesteps) through the same arena: v_out [B, N, T] is exactly sized too
    if not guided:
        t = torch.tensor([900.0, 500.0, 100.0])
        p_t = arena.put(t)
        fw_bytes = int(lib.nova_head_workspace_bytes(h._h, Bx * n, 0))
        assert fw_bytes <= ws_bytes
        p_v, o_v = arena.out(Bx * N * 3 * 4)
        _lib.check(lib.nova_head_forward(h._h, p_tok, p_t, 0, p_z, p_ids, Bx, Bx, N, n, p_v, p_ws, fw_bytes, stream),
                   "nova_head_forward")
        arena.check()
        assert bool(torch.isfinite(arena.view(o_v, (Bx, N, 3), torch.float32)).all())


def test_chamfer_and_knn_write_only_what_they_own():
    """Ragged cloud sizes through nova_chamfer_nn (distances + indices) with exactly sized outputs."""
    from nova_pointcloud_b200 import _lib

    lib = _lib.lib()
    B, N, M = 3, 301, 77
    gen = torch.Generator().manual_seed(3)
    a, b = torch.rand(B, N, 3, generator=gen), torch.rand(B, M, 3, generator=gen)
    arena = Arena(1 << 20)
    p_a, p_b = arena.put(a), arena.put(b)
    p_d1, o_d1 = arena.out(B * N * 4)
    p_d2, o_d2 = arena.out(B * M * 4)
    p_i1, o_i1 = arena.out(B * N * 4)
    p_i2, o_i2 = arena.out(B * M * 4)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.nova_chamfer_nn(p_a, p_b, B, N, M, p_d1, p_d2, p_i1, p_i2, stream), "nova_chamfer_nn")
    arena.check()
    d1 = arena.view(o_d1, (B, N), torch.float32)
    i1 = arena.view(o_i1, (B, N), torch.int32)
    ref = torch.cdist(a.double(), b.double())
    want, idx = ref.min(dim=2)
    assert float((d1.cpu().double() - want).abs().max()) < 1e-5
    assert bool((i1.cpu().long() == idx).float().mean() > 0.999)


@pytest.mark.parametrize("guided", [False, True])
def test_generate_sets_writes_only_what_it_owns(guided):
    """The whole set-by-set pass (gathers by order window, chain kernel, scatter) with exactly sized buffers."""
    from nova_pointcloud_b200 import _lib, partition

    D, Bx, N, S = 768, 3, 96, 3
    head, x, z, _, _ = make_case(2, D, D, Bx, N, 1)
    head = head.to(torch.bfloat16).cuda()
    h = head.handle()
    B = 2 * Bx if guided else Bx
    zz = (torch.cat([z, z.flip(0)]) if guided else z).bfloat16()
    tok = x.squeeze(-1).transpose(1, 2).float().contiguous()
    order = torch.rand(Bx, N, generator=torch.Generator().manual_seed(5)).argsort(dim=1)
    sizes = [int(v) for v in partition.cosine_num_preds(N, 9)]
    live = [v for v in sizes if v > 0]
    assert sum(sizes) == N
    c_sizes = (C.c_int32 * len(sizes))(*sizes)
    ts = (C.c_float * S)(1000.0, 600.0, 200.0)
    sg = (C.c_double * (S + 1))(1.0, 0.6, 0.2, 0.0)
    g = _lib.Guidance(3.0 if guided else 1.0, 0.0, 1.0)
    c_g = (C.c_float * len(live))(*[3.0 - 1.5 * i / len(live) for i in range(len(live))]) if guided else None
    lib = _lib.lib()
    ws_bytes = int(lib.nova_head_workspace_bytes(h._h, B * max(live), S))
    arena = Arena(ws_bytes + 64 * GAP + zz.numel() * 2 + 8 * tok.numel() * 4 + (1 << 20))
    p_tok, p_z, p_order = arena.put(tok), arena.put(zz), arena.put(order)
    p_out, o_out = arena.out(Bx * N * 3 * 4)
    p_ws, _ = arena.out(ws_bytes)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):  # eager, pass-graph capture, replay
        _lib.check(lib.nova_head_generate_sets(h._h, p_tok, p_z, p_order, B, Bx, N, c_sizes, len(sizes), ts, sg, S, C.byref(g),
                                               c_g, p_out, p_ws, ws_bytes, stream), "nova_head_generate_sets")
    arena.check()
    assert bool(torch.isfinite(arena.view(o_out, (Bx, N, 3), torch.float32)).all())  # every token belongs to one set


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_training_step_writes_only_what_it_owns(dtype):
    """nova_head_train_forward + nova_head_backward (every parameter gradient and dz) at a ragged row count."""
    from nova_pointcloud_b200 import _lib

    D, rows = (256, 203) if dtype == torch.float32 else (768, 331)
    head, _, _, _, _ = make_case(2, D, D, 1, 8, 1)
    head = head.to(dtype).cuda()
    h = head.handle()
    gen = torch.Generator().manual_seed(9)
    tok, t = torch.randn(rows, 3, generator=gen), torch.rand(rows, generator=gen) * 1000
    z, dv = torch.randn(rows, D, generator=gen).to(dtype), torch.randn(rows, 3, generator=gen) / rows
    lib = _lib.lib()
    lib.nova_head_train_bytes.restype = C.c_size_t
    tr_bytes = int(lib.nova_head_train_bytes(h._h, rows))
    sd = head.state_dict()
    names = list(sd.keys())
    total = sum(v.numel() for v in sd.values())
    arena = Arena(tr_bytes + (len(names) + 64) * 2 * GAP + 4 * total + 16 * rows * D + (1 << 20))
    p_tok, p_t, p_z, p_dv = arena.put(tok), arena.put(t), arena.put(z), arena.put(dv)
    p_v, o_v = arena.out(rows * 3 * 4)
    p_tr, _ = arena.out(tr_bytes)
    grads = [arena.out(sd[k].numel() * 4) for k in names]
    p_dz, o_dz = arena.out(rows * D * z.element_size())
    c_names = (C.c_char_p * len(names))(*[k.encode() for k in names])
    c_grads = (C.c_void_p * len(names))(*[p.value for p, _ in grads])
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.nova_head_train_forward(h._h, p_tok, p_t, p_z, rows, p_v, p_tr, tr_bytes, stream), "nova_head_train_forward")
    # the saved activations live in the workspace: only the spans around it are checked between the two calls
    _lib.check(lib.nova_head_backward(h._h, p_dv, p_tok, p_z, rows, len(names), c_names, c_grads, p_dz, p_tr, tr_bytes,
                                      stream), "nova_head_backward")
    arena.check()
    assert bool(torch.isfinite(arena.view(o_v, (rows, 3), torch.float32)).all())
    assert bool(torch.isfinite(arena.view(o_dz, (rows, D), dtype).float()).all())
    for (p, o), k in zip(grads, names):
        gk = arena.view(o, tuple(sd[k].shape), torch.float32)
        assert bool(torch.isfinite(gk).all()), k
