"""Where does the tcgen05 GEMM lose its last quarter?  K sweep at the head's shapes, with the library GEMM
(torch.matmul -> cuBLAS) timed beside it on the same box: if both climb with K the loss is per-tile (epilogue /
tile-boundary) work that K = 768 cannot amortise; if this kernel stays flat it is in the mainloop."""
import json
import sys

import torch

sys.path.insert(0, ".")
from nova_pointcloud_b200 import ops  # noqa: E402

M = 65536
out = []
g = torch.Generator(device="cuda").manual_seed(0)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for N in (768, 2304):
    for K in (768, 1536, 3072, 6144):
        A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        W = (torch.randn(N, K, device="cuda", generator=g) / K**0.5).bfloat16()
        b = torch.zeros(N, device="cuda")
        C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ms_nova = timed(lambda: ops.debug_gemm(A, W, b, "tcgen05", "bias"))
        ms_lib = timed(lambda: torch.matmul(A, W.t(), out=C))
        fl = 2.0 * M * N * K
        rec = {"M": M, "N": N, "K": K, "nova_us": round(ms_nova * 1e3, 1), "nova_tflops": round(fl / ms_nova / 1e9, 1),
               "cublas_us": round(ms_lib * 1e3, 1), "cublas_tflops": round(fl / ms_lib / 1e9, 1)}
        print(rec, flush=True)
        out.append(rec)
        del A, W, C
json.dump(out, open("gpurun_out/probe_gemm_k.json", "w"), indent=1)
