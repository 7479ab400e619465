import time, torch, ctypes as C, sys
sys.path.insert(0, '.')
import nova_pointcloud_b200 as nb
from nova_pointcloud_b200 import _lib
dev = torch.device('cuda')
B, N = 256, 2048
g = torch.Generator(device=dev).manual_seed(11)
pa = torch.rand(B, N, 3, device=dev, generator=g) * 2 - 1
pb = torch.rand(B, N, 3, device=dev, generator=g) * 2 - 1
d1 = torch.empty(B, N, device=dev); d2 = torch.empty(B, N, device=dev)
lib = _lib.lib()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def raw():
    lib.nova_chamfer_nn(C.c_void_p(pa.data_ptr()), C.c_void_p(pb.data_ptr()), B, N, N, C.c_void_p(d1.data_ptr()), C.c_void_p(d2.data_ptr()), None, None, st)
def op():
    return torch.ops.nova_b200.chamfer_nn(pa, pb, False)
def new():
    return nb.chamfer_distance(pa, pb)
def old():
    a1, a2, _, _ = torch.ops.nova_b200.chamfer_nn(pa, pb, False)
    return a1.double().mean(dim=1) + a2.double().mean(dim=1)
def mean_only():
    return torch.ops.nova_b200.chamfer_pair_mean(d1, d2)
for name, fn in [('raw', raw), ('op', op), ('new', new), ('old', old), ('mean_only', mean_only)]:
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(50): fn()
    e1.record(); th = (time.perf_counter() - t0) / 50
    torch.cuda.synchronize()
    print(f'{name:10s} device {e0.elapsed_time(e1)/50*1e3:8.1f} us/call   host issue {th*1e6:8.1f} us/call')
