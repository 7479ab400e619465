"""Is a GEMM paced by its MMAs or by its epilogue?  Cycle breakdown of epilogue warp 4 of CTA 0 (diagnostic build with
-DNOVA_TAIL_TIMELINE): time waiting for accumulators vs everything else, for the plain / SiLU GEMMs at the head's shapes."""
import json
import sys

import torch

sys.path.insert(0, ".")
from nova_pointcloud_b200 import _lib, ops  # noqa: E402

M = 65536
out = []
for N, K, epi in ((768, 768, "bias"), (768, 768, "bias_silu"), (2304, 768, "bias"), (768, 3072, "bias")):
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = (torch.randn(N, K, device="cuda") / K**0.5).bfloat16()
    b = torch.zeros(N, device="cuda")
    for _ in range(3):
        ops.debug_gemm(A, W, b, "tcgen05", epi)
    torch.cuda.synchronize()
    w = _lib.debug_words()
    tiles = (M // 256) * (N // 256) / 74.0
    out.append({"N": N, "K": K, "epi": epi, "epilogue_loop_cycles": w[0], "wait_accumulator": w[2], "wait_store": (w[3] & 0xffff) * 64,
                "wait_barriers": (w[3] >> 16) * 64, "tiles_per_cta": round(tiles, 2), "cycles_per_tile": round(w[0] / tiles),
                "mma_cycles_per_tile_at_peak": K // 64 * 4 * 128})
print(json.dumps(out, indent=1))
