"""Where an EPI_TAIL epilogue warp spends its cycles (diagnostic build with -DNOVA_TAIL_TIMELINE, CTA 0, warp 4):
    NOVA_B200_NVCC_FLAGS=-DNOVA_TAIL_TIMELINE python -m nova_pointcloud_b200.build --force   (kept as lib/libnova_b200_tl.so)
    NOVA_B200_LIB=$PWD/nova_pointcloud_b200/lib/libnova_b200_tl.so NOVA_B200_GRAPH=0 python scripts/profile_tail_timeline.py"""
import json
import sys

import torch

sys.path.insert(0, ".")
import nova_pointcloud_b200 as nb  # noqa: E402
from nova_pointcloud_b200 import _lib  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 768
head = nb.synth.make_head(D, 6, dtype=torch.bfloat16)
sched = nb.FlowMatchEulerDiscreteScheduler()
sched.set_timesteps(2)
noise, z = nb.synth.make_inputs(32, 2048, D, dtype=torch.bfloat16)
for _ in range(2):
    nb.denoise(head, sched, z, noise)
torch.cuda.synchronize()
w = _lib.debug_words()
total, full, acc = w[0], w[1], w[2]
store, bar = (w[3] & 0xffff) * 64, (w[3] >> 16) * 64
print(json.dumps({"D": D, "epilogue_loop_cycles": total, "wait_staged_chunks": full, "wait_accumulator": acc, "wait_own_store": store,
                  "wait_epilogue_barriers": bar, "everything_else": total - full - acc - store - bar,
                  "tiles_per_cta": (65536 // 256) * (D // 256) / 74.0}))
