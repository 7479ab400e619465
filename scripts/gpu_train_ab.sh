#!/bin/bash
# training step with and without the dual-output SiLU epilogue (NOVA_B200_TRAIN_DUAL_SILU), after the training parity tests
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_training.py tests/test_gpu_gemm.py tests/test_gpu_gemm_2cta.py tests/test_gpu_guardbands.py -x -q -m gpu 2>&1 | tail -3
for rep in 1 2; do
for v in 1 0; do
  NOVA_B200_TRAIN_DUAL_SILU=$v python - <<'PY'
import os, sys, torch
sys.path.insert(0, '.')
import nova_pointcloud_b200 as nb
from nova_pointcloud_b200 import training
D, M = 768, 65536
head = nb.synth.make_head(D, 6, dtype=torch.bfloat16, device='cuda').train()
g = torch.Generator(device='cuda').manual_seed(3)
x = torch.randn(32, 3, 2048, 1, device='cuda', generator=g)
z = torch.randn(32, 2048, D, device='cuda', generator=g).bfloat16()
sched = nb.FlowMatchEulerDiscreteScheduler(1000, shift=1.0)
def step():
    head.zero_grad(set_to_none=True)
    out = nb.get_losses(head, sched, z, x, loss_repeat=1)
    out['loss'].backward()
    return out['loss']
for _ in range(3): l = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): l = step()
e1.record(); torch.cuda.synchronize()
print('dual', os.environ['NOVA_B200_TRAIN_DUAL_SILU'], 'ms/step', round(e0.elapsed_time(e1) / 10, 3), 'loss', float(l))
PY
done
done
