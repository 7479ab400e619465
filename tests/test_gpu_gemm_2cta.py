"""GPU: the CTA-pair (cta_group::2) tcgen05 GEMM against torch matmul -- in its own process so that
a protocol bug here cannot poison the single-CTA results."""

import pytest
import torch

from gpu_util import relmax

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def report_debug_words():
    yield
    from nova_pointcloud_b200 import _lib

    words = _lib.debug_words()
    if words[0]:
        print("tcgen05 barrier timeout words:", [hex(w) for w in words])


def ref_gemm(A, W, bias, epi):
    if A.dtype != torch.float64:
        A, W = A.float(), W.float()
    out = A @ W.t()
    if bias is not None:
        out = out + bias.to(out.dtype)
    return torch.nn.functional.silu(out) if epi == "bias_silu" else out


@pytest.mark.parametrize("impl", ["tcgen05_2cta"])
@pytest.mark.parametrize("M,N,K", [
    (128, 256, 64),       # exactly one tile, one k-block
    (128, 256, 256),      # 4 k-blocks: the full smem ring once
    (128, 256, 1024),     # ring wraps 4 times (phase bits)
    (256, 512, 512),      # 4 tiles
    (1000, 768, 768),     # ragged M
    (300, 200, 104),      # ragged everything, K not a multiple of 64
    (20000, 1024, 256),   # more tiles than SMs: persistent loop + TMEM double buffering
    (64, 15360, 768),     # N = 20 D of the AdaLN GEMM
])
@pytest.mark.parametrize("epi", ["bias", "bias_silu"])
def test_bf16_gemm(impl, M, N, K, epi):
    from nova_pointcloud_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    out = ops.debug_gemm(A, W, b, impl, epi)
    torch.cuda.synchronize()
    ref = ref_gemm(A, W, b, epi)
    # bf16 output rounding: 2^-9 relative per element
    err = (out.float() - ref).abs()
    tol = 2.0**-8 * ref.abs() + 1e-2
    assert bool((err <= tol).all()), f"max err {float(err.max())} at {int(err.argmax())}"


@pytest.mark.parametrize("tile_n", ["64", "128"])
@pytest.mark.parametrize("impl", ["tcgen05_1cta", "tcgen05_2cta"])
@pytest.mark.parametrize("M,N,K", [
    (128, 64, 64),        # one narrow tile
    (64, 768, 768),       # the small-M fc shape: 12 / 6 tiles instead of 3
    (300, 200, 104),      # ragged everything
    (1000, 768, 1536),    # ring wraps with the deeper (8-stage) ring of the narrow tiles
    (5000, 1024, 256),    # more tiles than SMs
])
def test_narrow_tiles(monkeypatch, tile_n, impl, M, N, K):
    """64- and 128-column tiles (chosen automatically when 256-column tiles would leave most SMs idle)."""
    from nova_pointcloud_b200 import ops

    monkeypatch.setenv("NOVA_B200_TILE_N", tile_n)
    g = torch.Generator(device="cuda").manual_seed(M + N + K + int(tile_n))
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    for epi in ("bias", "bias_silu"):
        out = ops.debug_gemm(A, W, b, impl, epi)
        torch.cuda.synchronize()
        ref = ref_gemm(A, W, b, epi)
        err = (out.float() - ref).abs()
        tol = 2.0**-8 * ref.abs() + 1e-2
        assert bool((err <= tol).all()), f"{epi}: max err {float(err.max())} at {int(err.argmax())}"
    monkeypatch.setenv("NOVA_B200_TILE_N", "256")
    wide = ops.debug_gemm(A, W, b, impl, "bias")
    monkeypatch.setenv("NOVA_B200_TILE_N", tile_n)
    assert torch.equal(wide, ops.debug_gemm(A, W, b, impl, "bias"))  # same k order per output element: same bits


@pytest.mark.parametrize("impl", ["tcgen05_2cta", "tcgen05"])
def test_tcgen05_no_bias_and_repeatability(impl):
    from nova_pointcloud_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(512, 512, device="cuda", generator=g).bfloat16()
    W = (torch.randn(768, 512, device="cuda", generator=g) / 512**0.5).bfloat16()
    o1 = ops.debug_gemm(A, W, None, impl, "bias")
    o2 = ops.debug_gemm(A, W, None, impl, "bias")
    assert torch.equal(o1, o2)
    assert relmax(o1.float(), A.float() @ W.float().t()) < 1e-2


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("M,D,n_stats", [(256, 256, 3), (1000, 768, 3), (300, 512, 2), (4096, 1024, 3)])
def test_fused_adaln_gemm(M, D, n_stats, cta_group):
    """EPI_ADALN: h = LN(x)(1+scale)+shift and gate straight from the statistics GEMM's epilogue."""
    from nova_pointcloud_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(M + D)
    A = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    W = (torch.randn(n_stats * D, D, device="cuda", generator=g) / D**0.5).bfloat16()
    b = torch.randn(n_stats * D, device="cuda", generator=g) * 0.5
    x = (torch.randn(M, D, device="cuda", generator=g) * 3 + 1).bfloat16()
    h, gate = ops.debug_adaln_gemm(A, W, b, x, n_stats, cta_group)
    st = (A.float() @ W.float().t() + b).chunk(n_stats, dim=-1)
    ref = torch.nn.functional.layer_norm(x.float(), (D,), eps=1e-6) * (1 + st[0]) + st[1]
    err = (h.float() - ref).abs()
    assert bool((err <= 2.0**-7 * ref.abs() + 3e-2).all()), float(err.max())
    if n_stats == 3:
        err = (gate.float() - st[2]).abs()
        assert bool((err <= 2.0**-8 * st[2].abs() + 1e-2).all()), float(err.max())
