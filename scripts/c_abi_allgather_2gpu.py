"""Two ranks through the C ABI's multi-GPU entry points (no torch.distributed): rank 0 makes the NCCL id, ships it
through a file, both ranks sample their shard with nova_head_sample and exchange the points with nova_allgather.

    gpurun --gpus 2 -- python scripts/c_abi_allgather_2gpu.py          (spawns its two ranks itself)
"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rank_main(rank, world, id_path):
    import torch

    import nova_pointcloud_b200 as nb
    from nova_pointcloud_b200 import _lib

    torch.cuda.set_device(rank)
    lib = _lib.lib()
    uid = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(lib.nova_comm_unique_id(uid), "nova_comm_unique_id")
        with open(id_path + ".tmp", "wb") as f:
            f.write(uid.raw)
        os.replace(id_path + ".tmp", id_path)
    else:
        while not os.path.exists(id_path):
            time.sleep(0.05)
        with open(id_path, "rb") as f:
            uid = C.create_string_buffer(f.read(), 128)
    comm = C.c_void_p()
    _lib.check(lib.nova_comm_init_rank(uid, world, rank, C.byref(comm)), "nova_comm_init_rank")
    head = nb.synth.make_head(768, 6, dtype=torch.bfloat16)
    sched = nb.FlowMatchEulerDiscreteScheduler()
    sched.set_timesteps(25)
    noise, z = nb.synth.make_inputs(8, 512, 768, seed=2024, dtype=torch.bfloat16)  # the same 8 clouds on both ranks
    lo, hi = nb.shard_range(8, rank, world)
    local = nb.denoise(head, sched, z[lo:hi], noise[lo:hi]).contiguous()
    full = torch.empty(8, 512, 3, device="cuda")
    _lib.check(lib.nova_allgather(comm, C.c_void_p(local.data_ptr()), C.c_void_p(full.data_ptr()), local.numel() * 4,
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)), "nova_allgather")
    torch.cuda.synchronize()
    single = nb.denoise(head, sched, z, noise)  # every rank also samples all 8 clouds: the gather must reproduce it
    ok = bool(torch.equal(full[lo:hi], local)) and float((full - single).abs().max() / single.abs().max()) < 2e-2
    _lib.check(lib.nova_comm_destroy(comm), "nova_comm_destroy")
    print(f"rank {rank}: allgather {'ok' if ok else 'MISMATCH'}", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(rank_main(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]))
    id_path = f"/tmp/nova_nccl_id_{os.getpid()}"
    procs = [subprocess.Popen([sys.executable, __file__, str(r), "2", id_path]) for r in range(2)]
    rc = max(p.wait(timeout=600) for p in procs)
    print("c_abi_allgather_2gpu", "PASS" if rc == 0 else "FAIL")
    sys.exit(rc)
