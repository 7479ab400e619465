#!/usr/bin/env python
"""bench.py -- point clouds/s of the fused diffusion-head + flow-match Euler sampling path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4] [--impl reference]

One "step" = one full sampling pass (S = 25 Euler steps of the depth-6 head) over one batch of
B clouds x N point tokens of synthetic input.  Default workload = BASELINE.json configs[1]
(NOVA-0.3B head d6w768, 2048 points, bf16; B = 32 clouds per GPU => 65 536 rows per head call).
N > 1: one process per GPU (torchrun), the batch of clouds is sharded data-parallel (weak scaling:
B clouds per GPU), no collective inside the denoise loop, ONE all-gather of the generated points
inside the timed region.

Printed JSON (one line, rank 0):
  value / ms_per_step : device-resident inputs, CUDA events on the launching stream, max over ranks
  e2e                 : same metric through the public API with pinned HOST buffers; H2D of the
                        noise + condition and D2H of the points inside the timed region
  roofline            : the dominant kernel (tcgen05 AdaLN-statistics GEMM): algorithmic FLOP per launch /
                        average launch duration measured in situ with CUDA events the library records on
                        the launching stream around every launch; peak = MEASURED_PEAKS.json
                        bf16_tflops_sustained (kernel timed inside a long step)
  step_roofline       : whole step: algorithmic FLOP (F_min*B*N*S + 4*D^2*B*N,
                        hoisted work not credited) / step time
  kernel_shares       : in-situ share of the step per kernel class
  gemm_alone          : the plain tcgen05 GEMM at M x 20D x D timed alone (burst regime)
  cpu_baseline        : the reference's own Transformer3DModel.denoise (baseline/_ref, copied by build() from
                        /root/reference; the oracle port when that copy is absent) on a bounded sample, host cores
  target_shape / strong_cfg3 / cfg4 / chamfer_sharded : the other BASELINE.json configs, at every --gpus N:
                        NOVA-0.6B 2048-point shape (the >= 60 % target), cfg3 = 64 clouds TOTAL sharded over the
                        ranks (strong scaling), cfg4 = NOVA-1.4B 32 x 2048 per GPU with its all-gather, cfg5 = 256
                        Chamfer pairs sharded over the ranks
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Chamfer: 6 fp32 lane-operations per undirected pair, 148 SMs x 128 fp32 lanes per clock at 1.965 GHz -> directed pair-evals/s
CHAMFER_FP32_PEAK = 148 * 128 * 1.965e9 / 6.0 * 2.0

WORKLOADS = {
    # name: (width, points, clouds per GPU, dtype)
    "cfg2": dict(width=768, points=2048, batch=32, desc="NOVA-0.3B d48w768 head mlp_d6w768, 2048 points, bf16"),
    "cfg3": dict(width=1024, points=1024, batch=64, desc="NOVA-0.6B d48w1024 head mlp_d6w1024, 1024 points, bf16"),
    "cfg3-2048": dict(width=1024, points=2048, batch=32, desc="NOVA-0.6B head mlp_d6w1024, 2048 points, bf16"),
    "cfg4": dict(width=1536, points=2048, batch=32, desc="NOVA-1.4B d48w1536 head mlp_d6w1536, 2048 points, bf16"),
}
S_STEPS = 25
DEPTH = 6
T = 3


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(burst=float(p["bf16_tflops"]), sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    hbm=float(p["hbm_gbs"]), source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


def algorithmic_flops(D, rows, steps=S_STEPS):
    """per token-step 2*(32 D^2 + 2 T D), plus 4 D^2 once per token (hoisted cond)."""
    return rows * (steps * 2.0 * (32 * D * D + 2 * T * D) + 4.0 * D * D)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def _run_nvml(self) -> bool:
        """Sample through NVML in-process (about 1 ms per sample) so that even a sub-second timed region gets tens
        of samples; returns False when NVML is unavailable and the nvidia-smi loop should be used instead."""
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception:
            return False
        bits = ((0x8, 3), (0x40, 4), (0x20, 5), (0x4, 6))  # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap -> row slot
        while not self._stop_evt.is_set():
            try:
                mask = int(reasons_fn(h))
                row = [str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), "", "", "", "", ""]
                for bit, slot in bits:
                    row[slot] = "Active" if mask & bit else "Not Active"
                self.rows.append(row)
            except Exception:
                pass
            self._stop_evt.wait(0.01)
        return True

    def run(self):
        if self._run_nvml():
            return
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        mx = max((float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(self.rows)}


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def _reference_runner(width, threads):
    """(kind, fn): fn(z, noise) runs one 25-step denoise of the reference algorithm on the host cores, fp32.

    kind == "reference": the reference's OWN modules (Transformer3DModel.denoise driving DiffusionMLP and
    FlowMatchEulerDiscreteScheduler, transformer_3d.py:102-113), imported unmodified from baseline/_ref -- the copy
    __graft_entry__.build() makes from /root/reference -- or from /root/reference itself where it is mounted.
    kind == "port": oracle/loop.py, when neither exists."""
    torch.set_num_threads(threads)
    for root in (REF_DIR, "/root/reference"):
        if os.path.isdir(os.path.join(root, "diffnext")):
            os.environ["NOVA_REFERENCE_ROOT"] = root
            break
    from oracle import _reference_import as RI

    RI.REFERENCE_ROOT = os.environ.get("NOVA_REFERENCE_ROOT", RI.REFERENCE_ROOT)
    if RI.reference_available():
        ref = RI.import_reference()
        torch.manual_seed(1337)
        head = ref.DiffusionMLP(DEPTH, width, width, patch_size=1, image_dim=3).eval()
        model, _ = RI.reference_denoiser(ref, head, num_steps=S_STEPS)
        gs = ref.GuidanceScaler(guidance_scale=1)

        def run(z, noise):
            with torch.no_grad():
                return model.denoise(z, noise, gs)

        return "reference", run
    from oracle import head as OH
    from oracle import loop as OL

    sd = OH.init_state_dict(DEPTH, width, width, 1, 3, seed=1337)

    def run_port(z, noise):
        with torch.no_grad():
            return OL.denoise(sd, z, noise, num_steps=S_STEPS)

    return "port", run_port


def _cpu_inputs(width, points, clouds):
    g = torch.Generator().manual_seed(2024)
    noise = torch.randn(clouds, 3, points, 1, generator=g)
    z = torch.randn(clouds, points, width, generator=g)
    return z, noise


def cpu_baseline_run(width, points, sample_clouds, threads, repeats=1):
    """Time the reference's denoise loop (fp32, host cores) on a bounded sample of `sample_clouds` clouds."""
    kind, run = _reference_runner(width, threads)
    z, noise = _cpu_inputs(width, points, sample_clouds)
    run(z[:1, :64], noise[:1, :, :64])  # warm the thread pool
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        run(z, noise)
        best = min(best, time.perf_counter() - t0)
    return sample_clouds / best, best, kind


def cpu_cfg1(threads):
    """BASELINE.json configs[0] exactly: D = 1024, 4 clouds x 1024 points, 25 steps, fp32; 1 warm-up + best of 3."""
    kind, run = _reference_runner(1024, threads)
    z, noise = _cpu_inputs(1024, 1024, 4)
    run(z, noise)
    best = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        run(z, noise)
        best = min(best, time.perf_counter() - t0)
    return {"value": 4 / best, "unit": "clouds/s", "seconds_best_of_3": best, "kind": kind, "cores": threads,
            "token_steps_per_s": 4 * 1024 * S_STEPS / best,
            "what": "BASELINE configs[0]: DiffusionMLP(6,1024,1024) + flow-match Euler, 4 x 1024 points, 25 steps, fp32"}


def cpu_chamfer(threads):
    """Chamfer on the host: variant A per pair with scipy cdist float64 (demo.py:44-53, 1 core) and batched
    torch.cdist fp32 on all cores (train_newloss.py:337, test_optimize.py:366), 16 pairs of 2048 x 2048."""
    import numpy as np
    from oracle import chamfer as OC

    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(11)
    a = torch.rand(16, 2048, 3, generator=g) * 2 - 1
    b = torch.rand(16, 2048, 3, generator=g) * 2 - 1
    a_np, b_np = a.numpy(), b.numpy()
    t0 = time.perf_counter()
    for i in range(4):
        OC.chamfer_a(a_np[i], b_np[i])
    scipy_s = (time.perf_counter() - t0) / 4
    torch.cdist(a[:2], b[:2])
    t0 = time.perf_counter()
    d = torch.cdist(a, b)
    _ = d.min(dim=2).values.mean(dim=1) + d.min(dim=1).values.mean(dim=1)
    cdist_s = (time.perf_counter() - t0) / 16
    return {"scipy_float64_1core_pairs_per_s": 1.0 / scipy_s, "torch_cdist_fp32_allcores_pairs_per_s": 1.0 / cdist_s,
            "cores": threads, "sample": "4 pairs (scipy) / 16 pairs (torch.cdist) of 2048 x 2048 points"}


def bench_config(args, wl, world):
    D, N, B = wl["width"], wl["points"], wl["batch"]
    return {"workload": args.workload + ": " + wl["desc"], "points": N, "width": D, "depth": DEPTH,
            "diffusion_steps": S_STEPS, "clouds_per_gpu": B, "rows_per_head_call": B * N,
            "parallelism": f"dp{world} (clouds sharded, one all-gather of outputs)",
            "l2": "inputs larger than L2 (126 MB): every [M, D] bf16 activation is "
                  f"{B * N * D * 2 / 1e6:.0f} MB and a diffusion step streams ~85 of them "
                  f"({85 * B * N * D * 2 / 1e9:.1f} GB) through HBM; weights {32 * D * D * 2 / 1e6:.0f} MB stay L2-resident"}


def run_reference_arm(args, wl):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, on this arm's
    config / metric / unit; every step is a bounded sample (2 clouds) of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_clouds = 2
    vals, kind = [], "port"
    for i in range(args.warmup + args.steps):
        v, secs, kind = cpu_baseline_run(wl["width"], wl["points"], sample_clouds, threads)
        if i >= args.warmup:
            vals.append((v, secs))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(s for _, s in vals) / len(vals)
    sample = (f"{sample_clouds} clouds x {wl['points']} tokens x {S_STEPS} steps per step, fp32, "
              + ("the reference's own Transformer3DModel.denoise + DiffusionMLP + FlowMatchEulerDiscreteScheduler "
                 "(baseline/_ref, unmodified)" if kind == "reference" else "oracle/loop.py denoise (port: no reference copy present)")
              + "; the CPU is throughput-flat in the number of clouds")
    line = {
        "impl": "reference", "metric": "point_clouds_per_sec", "value": value, "unit": "clouds/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, wl, args.gpus),
        "head_tokens_per_s": value * wl["points"],
        "cpu_baseline": {"value": value, "unit": "clouds/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        line["cfg1"] = cpu_cfg1(threads)
    except Exception as e:  # report, never hide
        line["cfg1"] = {"error": str(e)[:300]}
    try:
        line["chamfer_cpu"] = cpu_chamfer(threads)
    except Exception as e:
        line["chamfer_cpu"] = {"error": str(e)[:300]}
    print(json.dumps(line), flush=True)


def eager_library_sampler(head, sched, z, noise, compiled=False):
    """The reference's algorithm as eager PyTorch ON THE GPU (cuBLAS GEMMs + ATen element-wise kernels): the
    "library kernel" bar of SURVEY 8(d).  Written here with torch.nn.functional from the module's state_dict; it
    is a reported comparator only -- neither the product path nor the oracle.  Like the reference it recomputes
    the condition projection every step and carries the latent in the model dtype
    (transformer_3d.py:102-113, diffusion_mlp.py:89-99, scheduling_cfm.py:125-140).  ``compiled=True`` runs one
    diffusion step through torch.compile (inductor), the second bar SURVEY names."""
    import math

    import torch.nn.functional as F

    sd = dict(head.state_dict())
    dt_model = z.dtype
    depth = sum(1 for k in sd if k.startswith("blocks.") and k.endswith(".norm1.proj.weight"))
    wp = sd["patch_embed.proj.weight"]
    w_tok = wp.permute(0, 2, 3, 1).reshape(wp.shape[0], -1).contiguous()
    freq = torch.exp(torch.arange(128, device=z.device, dtype=torch.float32) * (-math.log(10000.0) / 128))

    def mlp2(prefix, x):
        return F.linear(F.silu(F.linear(x, sd[prefix + ".fc1.weight"], sd[prefix + ".fc1.bias"])),
                        sd[prefix + ".fc2.weight"], sd[prefix + ".fc2.bias"])

    def adaln(prefix, x, zt, k):
        st = F.linear(F.silu(zt), sd[prefix + ".proj.weight"], sd[prefix + ".proj.bias"]).chunk(k, dim=-1)
        return F.layer_norm(x, (x.shape[-1],), None, None, 1e-6) * (1 + st[0]) + st[1], st[2:]

    def one_step(x, t, dt):
        emb = t.expand(x.shape[0], 1) * freq
        temb = mlp2("time_cond_embed.timestep_proj", torch.cat([emb.cos(), emb.sin()], dim=-1).to(dt_model))
        zt = mlp2("time_cond_embed.condition_proj", z) + temb.unsqueeze(1)
        h = F.linear(x, w_tok, sd["patch_embed.proj.bias"])
        for b in range(depth):
            y, (gate,) = adaln(f"blocks.{b}.norm1", h, zt, 3)
            u = mlp2(f"blocks.{b}.proj", y)
            u = F.layer_norm(u, (u.shape[-1],), sd[f"blocks.{b}.norm2.weight"], sd[f"blocks.{b}.norm2.bias"], 1e-5)
            h = u * gate + h
        y, _ = adaln("norm", h, zt, 2)
        v = F.linear(y, sd["head.weight"], sd["head.bias"])
        return v * dt + x

    step_fn = torch.compile(one_step) if compiled else one_step
    x = noise.squeeze(-1).transpose(1, 2).to(dt_model)  # (B,3,N,1) -> tokens (B,N,3)
    sig = sched.sigmas
    for i, t in enumerate(sched.timesteps):
        x = step_fn(x, torch.full((1, 1), float(t), device=z.device), torch.tensor(sig[i + 1] - sig[i], device=z.device, dtype=dt_model))
    return x.float()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="clouds per GPU (default: workload's)")
    ap.add_argument("--impl", default="nova", choices=["nova", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the set-by-set and Chamfer legs")
    ap.add_argument("--no-compile-bar", action="store_true",
                    help="skip the torch.compile comparator (the inductor build of one diffusion step takes ~1 min)")
    ap.add_argument("--no-north-star", action="store_true",
                    help="skip the other BASELINE.json configs (target_shape, strong_cfg3, cfg4, chamfer_sharded)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch
    if args.impl == "reference":
        return run_reference_arm(args, wl)
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist

    import nova_pointcloud_b200 as nb
    from nova_pointcloud_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    D, N, B = wl["width"], wl["points"], wl["batch"]
    total = B * world
    head = nb.synth.make_head(D, DEPTH, dtype=torch.bfloat16, device=dev)
    sched = nb.FlowMatchEulerDiscreteScheduler()
    sched.set_timesteps(S_STEPS)
    # per-rank shard of the synthetic batch (seed offset by rank), pinned host copies for the e2e leg
    noise_h, z_h = nb.synth.make_inputs(B, N, D, seed=2024 + rank, dtype=torch.bfloat16, pin=True)
    noise_d, z_d = noise_h.to(dev), z_h.to(dev)
    out_h = torch.empty(total, N, T, dtype=torch.float32).pin_memory()

    def step_resident():
        local = nb.denoise(head, sched, z_d, noise_d)
        return nb.gather_shards(local, total)

    def step_e2e():
        n_d = noise_h.to(dev, non_blocking=True)
        zz = z_h.to(dev, non_blocking=True)
        local = nb.denoise(head, sched, zz, n_d)
        full = nb.gather_shards(local, total)
        out_h.copy_(full, non_blocking=True)
        return full

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ops.launch_count_reset()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1) / steps
        launches = ops.launch_count()
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks

    ms, launches, clocks = timed(step_resident, args.steps, args.warmup)
    # e2e: HOST buffers through the public API.  Every step's H2D (pinned noise + z) and D2H (the points) is inside the
    # timed region; nb.HostSampler issues the copy of step k+1 on a copy stream under the denoise of step k, so only the
    # first copy of the region is exposed.  (step_e2e above is the same path without the overlap, kept for warm-up.)
    def timed_e2e(steps):
        pipe = nb.HostSampler(head, sched, total)
        pipe.submit(z_h, noise_h)
        pipe.collect(out_h)  # warm-up: staging slots allocated, loop graph captured for slot 0
        pipe.submit(z_h, noise_h)
        pipe.collect(out_h)  # ... and for slot 1
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.submit(z_h, noise_h)
        for k in range(steps):
            if k + 1 < steps:
                pipe.submit(z_h, noise_h)
            pipe.collect(out_h)
        e1.record()
        barrier()
        t_ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([t_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms

    timed(step_e2e, 1, 1)
    ms_e2e = timed_e2e(args.steps)
    value = total / (ms * 1e-3)
    e2e_value = total / (ms_e2e * 1e-3)

    # in-situ per-kernel-class durations: CUDA events recorded by the library around every launch, on the
    # launching stream, inside real sampling steps (a separate pass, so the timed region above is clean)
    prof_steps = 2
    ops.profile_enable(True)
    for _ in range(prof_steps):
        step_resident()
    torch.cuda.synchronize()
    prof = ops.profile_read()
    ops.profile_enable(False)

    # dominant kernel alone: the AdaLN GEMM (M x 20D x D) on the tcgen05 kernel
    pk = peaks()
    gemm = None
    try:
        M = B * N
        A = torch.randn(M, D, device=dev).bfloat16()
        W = (torch.randn(20 * D, D, device=dev) / D**0.5).bfloat16()
        bias = torch.zeros(20 * D, device=dev)
        for _ in range(2):
            ops.debug_gemm(A, W, bias, "tcgen05", "bias")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            ops.debug_gemm(A, W, bias, "tcgen05", "bias")
        e1.record()
        torch.cuda.synchronize()
        gms = e0.elapsed_time(e1) / reps
        tf = 2.0 * M * 20 * D * D / (gms * 1e-3) / 1e12
        gemm = {"kernel": "nova::tc::gemm_kernel (tcgen05, M x 20D x D, bias epilogue)", "M": M, "N": 20 * D, "K": D,
                "ms": gms, "achieved": tf, "peak": pk["burst"], "unit": "TFLOP/s", "frac": tf / pk["burst"],
                "peak_source": pk["source"] + " burst (kernel timed alone)"}
        del A, W, bias
    except Exception as e:  # report, never hide
        gemm = {"error": str(e)[:300]}

    # ---- the other BASELINE.json configs, on every rank (sharded), at every --gpus N
    north = {}
    if not args.no_north_star:
        def sharded_leg(width, points, clouds_local, clouds_total, steps=3, warmup=3):
            """One data-parallel sampling leg: this rank's shard + the all-gather, CUDA events, max over ranks."""
            hd = nb.synth.make_head(width, DEPTH, dtype=torch.bfloat16, device=dev)
            n_d, zz = nb.synth.make_inputs(max(clouds_local, 1), points, width, seed=4242 + rank, dtype=torch.bfloat16, device=dev)
            n_d, zz = n_d[:clouds_local], zz[:clouds_local]

            def one():
                if clouds_local > 0:
                    local = nb.denoise(hd, sched, zz, n_d)
                else:  # more ranks than clouds: this rank only takes part in the collective
                    local = torch.empty(0, points, T, device=dev)
                return nb.gather_shards(local, clouds_total)

            ms_leg, launches_leg, clk = timed(one, steps, warmup)
            ops.profile_enable(True)
            one()
            torch.cuda.synchronize()
            pr = ops.profile_read()
            ops.profile_enable(False)
            del hd, n_d, zz
            return ms_leg, launches_leg, clk, pr

        def leg_record(width, points, clouds_local, clouds_total, ms_leg, pr, clk, what):
            fl = algorithmic_flops(width, max(clouds_local, 0) * points)
            tf_leg = fl / (ms_leg * 1e-3) / 1e12
            ada_ms_, ada_n_ = pr["gemm_ada"]
            # the modulation GEMMs: N = 2 * width for each of the DEPTH + 1 launches of a step (the gate columns are the
            # separate gemm_tail class)
            ada_tf_ = (2.0 * clouds_local * points * 2 * width * width) / (ada_ms_ / max(ada_n_, 1) * 1e-3) / 1e12 if ada_n_ else 0.0
            tot = sum(v[0] for v in pr.values()) or 1.0
            return {"value": clouds_total / (ms_leg * 1e-3), "unit": "clouds/s", "ms_per_pass": ms_leg, "n_gpus": world,
                    "clouds_total": clouds_total, "clouds_this_gpu": clouds_local, "points": points, "width": width,
                    "rows_per_head_call": clouds_local * points, "what": what,
                    "step_roofline": {"bound": "tensor", "achieved": tf_leg, "peak": pk["sustained"], "unit": "TFLOP/s",
                                      "frac": tf_leg / pk["sustained"], "frac_of_burst": tf_leg / pk["burst"]},
                    "roofline": {"bound": "tensor", "achieved": ada_tf_, "peak": pk["sustained"], "unit": "TFLOP/s",
                                 "frac": ada_tf_ / pk["sustained"], "frac_of_burst": ada_tf_ / pk["burst"],
                                 "kernel": "modulation GEMM (M x 2D x D, EPI_ADALN), in situ (rank 0)", "launches_timed": ada_n_},
                    "kernel_shares": {k: round(v[0] / tot, 4) for k, v in pr.items() if v[1]},
                    "clocks": clk}

        try:  # the shape the >= 60 % target is stated on: NOVA-0.6B head, 2048 points, 32 clouds per GPU (weak)
            m_, _, c_, pr_ = sharded_leg(1024, 2048, 32, 32 * world)
            north["target_shape"] = leg_record(1024, 2048, 32, 32 * world, m_, pr_, c_,
                                               "NOVA-0.6B (mlp_d6w1024) 2048-point sampling, 32 clouds per GPU, bf16, one all-gather")
        except Exception as e:  # report, never hide
            north["target_shape"] = {"error": str(e)[:300]}
        try:  # BASELINE configs[2]: 64 clouds TOTAL, 1024 points, D = 1024, sharded over the ranks (strong scaling)
            lo, hi = nb.shard_range(64, rank, world)
            m_, _, c_, pr_ = sharded_leg(1024, 1024, hi - lo, 64)
            north["strong_cfg3"] = leg_record(1024, 1024, hi - lo, 64, m_, pr_, c_,
                                              "BASELINE configs[2]: NOVA-0.6B 1024-point sampling, batch 64 TOTAL sharded over "
                                              "the ranks (strong scaling: compare value across --gpus 1/2/4/8)")
            north["strong_cfg3"]["scaling"] = "strong"
        except Exception as e:
            north["strong_cfg3"] = {"error": str(e)[:300]}
        try:  # BASELINE configs[3]: NOVA-1.4B, 2048 points, 32 clouds per GPU (= 256 on 8 GPUs), all-gather of (B,2048,3) fp32
            m_, _, c_, pr_ = sharded_leg(1536, 2048, 32, 32 * world, steps=2)
            north["cfg4"] = leg_record(1536, 2048, 32, 32 * world, m_, pr_, c_,
                                       "BASELINE configs[3]: NOVA-1.4B (mlp_d6w1536) 2048-point sampling, 32 clouds per GPU "
                                       "(batch 256 on 8 GPUs), NCCL all-gather of the outputs inside the timed region")
            north["cfg4"]["allgather_bytes"] = 32 * world * 2048 * 3 * 4
        except Exception as e:
            north["cfg4"] = {"error": str(e)[:300]}
        try:  # BASELINE configs[4]: Chamfer, 256 pairs of 2048 x 2048 in total, sharded over the ranks
            Bc, Nc = 256, 2048
            lo, hi = nb.shard_range(Bc, rank, world)
            gc = torch.Generator(device=dev).manual_seed(11 + rank)
            pa = torch.rand(hi - lo, Nc, 3, device=dev, generator=gc) * 2 - 1
            pb = torch.rand(hi - lo, Nc, 3, device=dev, generator=gc) * 2 - 1

            def chamfer_pass():
                cd = nb.chamfer_distance(pa, pb).float().unsqueeze(-1).unsqueeze(-1)  # (b,1,1): per-pair Chamfer A
                return nb.gather_shards(cd, Bc)

            cms, _, _ = timed(chamfer_pass, 10, 3)
            kms, _, _ = timed(lambda: torch.ops.nova_b200.chamfer_nn(pa, pb, False), 10, 3)  # this rank's kernel alone
            pair_evals = 2.0 * Bc * Nc * Nc  # directed
            # fp32 pipe bound: 6 fp32 lane-operations per undirected pair in the exact difference form (3 subtractions,
            # 1 multiply, 2 FMAs; the minima run on the ALU pipe), 148 SMs x 128 fp32 lanes per clock at the maximum SM
            # clock.  (The one-sweep kernel issues them as packed FADD2 / FMUL2 / FFMA2, which halves the issue slots, not
            # the lane-operations.)
            issue_peak = CHAMFER_FP32_PEAK  # directed pair evaluations per second per GPU
            north["chamfer_sharded"] = {
                "value": Bc / (cms * 1e-3), "unit": "cloud pairs/s", "ms": cms, "pairs_total": Bc, "pairs_this_gpu": hi - lo,
                "n_gpus": world, "points": Nc, "pair_evals_per_s": pair_evals / (cms * 1e-3),
                "kernel_ms_this_gpu": kms, "kernel_pair_evals_per_s_per_gpu": 2.0 * (hi - lo) * Nc * Nc / (kms * 1e-3),
                "what": "BASELINE configs[4]: Chamfer A of 256 x (2048 vs 2048), pairs sharded over the ranks, one all-gather "
                        "of the per-pair distances; timed through the public chamfer_distance call (a 0.4 ms job on one GPU: "
                        "with 32 pairs per GPU the kernel is kernel_ms_this_gpu and the rest of `ms` is launch, reduction and "
                        "NCCL latency, which is what bounds its scaling)",
                "kernel_roofline_frac": 2.0 * (hi - lo) * Nc * Nc / (kms * 1e-3) / issue_peak,
                "roofline": {"bound": "fp32_pipe", "achieved": pair_evals / (cms * 1e-3) / 1e12, "peak": issue_peak * world / 1e12,
                             "unit": "T directed pair-evals/s", "frac": pair_evals / (cms * 1e-3) / (issue_peak * world),
                             "note": "algorithmically HBM-trivial (16.8 MB per 2.1 G pair evaluations); the bound is the fp32 "
                                     "pipe: 6 lane-operations per undirected pair (exact difference form), 148 x 128 lanes per "
                                     "clock at 1.965 GHz; `frac` includes the Python-side reductions and the all-gather of this "
                                     "call, kernel_roofline_frac is this rank's kernel alone"}}
            del pa, pb
        except Exception as e:
            north["chamfer_sharded"] = {"error": str(e)[:300]}
        barrier()

    # ---- secondary legs (rank 0, reported beside the headline; bounded to ~2 s)
    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:  # single-GPU only: other ranks must not wait on rank 0's extras
        try:  # set-by-set autoregressive generation (the reference's generate_frame pattern): 64 cosine-schedule sets
            sizes = nb.partition.cosine_num_preds(N, 64)
            gen = torch.Generator(device=dev).manual_seed(7)
            shape = (B, 3, N, 1)

            def ar_pass():
                return nb.generate_sets(head, sched, z_d, shape, sizes, None, gen)

            for _ in range(3):  # eager, graph capture, first replay: every set size has its own loop graph
                ar_pass()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                ar_pass()
            e1.record()
            torch.cuda.synchronize()
            ar_ms = e0.elapsed_time(e1) / 2
            extras["set_by_set"] = {
                "value": B / (ar_ms * 1e-3), "unit": "clouds/s", "ms_per_pass": ar_ms, "sets": len([k for k in sizes if k]),
                "rows_per_head_call": [int(B * min(k for k in sizes if k)), int(B * max(sizes))],
                "what": "64-set cosine schedule, every set = one fused 25-step call over B*n rows (CUDA-graph replay)"}
        except Exception as e:  # report, never hide
            extras["set_by_set"] = {"error": str(e)[:300]}
        try:  # the same 64-set pass over 4x the clouds: the chain of dependent launches is the cost, rows are nearly free
            Bw = 4 * B
            _, z_w = nb.synth.make_inputs(Bw, N, D, seed=77, dtype=torch.bfloat16, device=dev)
            shape_w = (Bw, 3, N, 1)

            def ar_wide_pass():
                return nb.generate_sets(head, sched, z_w, shape_w, sizes, None, gen)

            for _ in range(3):
                ar_wide_pass()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                ar_wide_pass()
            e1.record()
            torch.cuda.synchronize()
            arw_ms = e0.elapsed_time(e1) / 2
            extras["set_by_set_4x_batch"] = {
                "value": Bw / (arw_ms * 1e-3), "unit": "clouds/s", "ms_per_pass": arw_ms, "clouds": Bw, "sets": 64,
                "rows_per_head_call": [int(Bw * min(k for k in sizes if k)), int(Bw * max(sizes))],
                "what": "64-set cosine schedule over 4x the clouds per GPU (throughput serving: HBM holds hundreds of clouds)"}
            del z_w
        except Exception as e:  # report, never hide
            extras["set_by_set_4x_batch"] = {"error": str(e)[:300]}
        try:  # the point-cloud pipeline's own partition: 20 equal random subsets (transformer_pointcloud_nova.py:63-78)
            sizes20 = nb.partition.equal_subset_sizes(N, 20)

            def ar20_pass():
                return nb.generate_sets(head, sched, z_d, shape, sizes20, None, gen)

            for _ in range(3):
                ar20_pass()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                ar20_pass()
            e1.record()
            torch.cuda.synchronize()
            ar20_ms = e0.elapsed_time(e1) / 2
            extras["set_by_set_20"] = {
                "value": B / (ar20_ms * 1e-3), "unit": "clouds/s", "ms_per_pass": ar20_ms, "sets": 20,
                "rows_per_head_call": [int(B * min(sizes20)), int(B * max(sizes20))],
                "what": "20 equal random subsets (dynamic_partition), every set = one fused 25-step call"}
        except Exception as e:  # report, never hide
            extras["set_by_set_20"] = {"error": str(e)[:300]}
        try:  # classifier-free guidance (the default of the reference's non-point-cloud pipelines): 2x rows per cloud
            Bg = max(B // 2, 1)
            gs = nb.GuidanceScaler(guidance_scale=5.0)
            zg = torch.cat([z_d[:Bg], torch.zeros_like(z_d[:Bg])])

            def guided_pass():
                return nb.denoise(head, sched, zg, noise_d[:Bg], gs)

            for _ in range(3):
                guided_pass()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                guided_pass()
            e1.record()
            torch.cuda.synchronize()
            g_ms = e0.elapsed_time(e1) / 2
            extras["guided"] = {"value": Bg / (g_ms * 1e-3), "unit": "clouds/s", "ms_per_pass": g_ms, "clouds": Bg,
                                "rows_per_head_call": 2 * Bg * N, "guidance_scale": 5.0,
                                "what": "two-pass classifier-free guidance fused into the loop ([cond; uncond] rows, "
                                        "v = vu + (vc - vu) s, Euler update in one kernel per step)"}
        except Exception as e:
            extras["guided"] = {"error": str(e)[:300]}
        try:  # Chamfer scorer, BASELINE configs[4]: 256 pairs of 2048 x 2048 points
            Bc, Nc = 256, 2048
            gc = torch.Generator(device=dev).manual_seed(11)
            pa = torch.rand(Bc, Nc, 3, device=dev, generator=gc) * 2 - 1
            pb = torch.rand(Bc, Nc, 3, device=dev, generator=gc) * 2 - 1
            for _ in range(3):
                nb.chamfer_nn(pa, pb, with_indices=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                nb.chamfer_nn(pa, pb, with_indices=False)  # the distance-only kernel the three Chamfer variants use
            e1.record()
            torch.cuda.synchronize()
            cms = e0.elapsed_time(e1) / reps
            alg_bytes = 2 * Bc * Nc * 3 * 4 + 2 * Bc * Nc * 4  # points in, distances out (16.78 MB)
            extras["chamfer"] = {
                "value": Bc / (cms * 1e-3), "unit": "cloud pairs/s", "ms": cms, "pairs": Bc, "points": Nc,
                "pair_evals_per_s": 2.0 * Bc * Nc * Nc / (cms * 1e-3),
                "roofline_fp32": {"bound": "fp32_pipe", "achieved": 2.0 * Bc * Nc * Nc / (cms * 1e-3) / 1e12,
                                  "peak": CHAMFER_FP32_PEAK / 1e12, "unit": "T directed pair-evals/s",
                                  "frac": 2.0 * Bc * Nc * Nc / (cms * 1e-3) / CHAMFER_FP32_PEAK,
                                  "note": "6 fp32 lane-operations per undirected pair, 148 SMs x 128 lanes at 1.965 GHz"},
                "roofline": {"bound": "hbm", "achieved": alg_bytes / (cms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                             "frac": alg_bytes / (cms * 1e-3) / 1e9 / pk["hbm"], "traffic": None,
                             "note": "algorithmically HBM-trivial (%.1f MB for 2.1 G pair evaluations): the limiter is "
                                     "the fp32 pipe, see roofline_fp32" % (alg_bytes / 1e6)}}
            if not args.no_cpu_baseline:
                import numpy as np
                from oracle import chamfer as OC

                a_np, b_np = pa[:4].cpu().numpy(), pb[:4].cpu().numpy()
                t0 = time.perf_counter()
                for i in range(4):
                    OC.chamfer_a(a_np[i], b_np[i])
                cpu_s = (time.perf_counter() - t0) / 4
                extras["chamfer"]["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": "cloud pairs/s", "cores": 1, "kind": "port",
                                                     "sample": "4 pairs, scipy cdist float64 + min (oracle/chamfer.py, demo.py:38-55)"}
        except Exception as e:
            extras["chamfer"] = {"error": str(e)[:300]}
        try:  # neighbourhood ops on the same tiling: local density of 256 clouds x 2048 points
            def timed_ms(fn, reps=10):
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

            sel = torch.randperm(Nc, device=dev, generator=gc)[: Nc // 4]
            dms = timed_ms(lambda: nb.compute_local_density(pa, 8))
            kms = timed_ms(lambda: nb.knn(pa, pa, 9))
            ims = timed_ms(lambda: nb.feature_aware_interpolation(pa, Nc // 4, indices=sel))
            extras["geometry"] = {
                "value": Bc / (dms * 1e-3), "unit": "clouds/s", "what": "compute_local_density(k_neighbors=8), %d x %d points" % (Bc, Nc),
                "ms": dms, "pair_evals_per_s": Bc * Nc * Nc / (dms * 1e-3),
                "knn_k9_with_indices_ms": kms, "softmax_interp_quarter_ms": ims,
                "note": "HBM-trivial like Chamfer (%.1f MB in + out per call); bound by fp32 issue + the k-best insertion" % (Bc * Nc * 16 / 1e6)}
            if not args.no_cpu_baseline:
                from oracle import geometry as OG

                p_np = pa[:4].cpu().numpy()
                t0 = time.perf_counter()
                OG.local_density(p_np, 8)
                cpu_s = (time.perf_counter() - t0) / 4
                extras["geometry"]["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": "clouds/s", "cores": 1, "kind": "port",
                                                      "sample": "4 clouds, scipy cdist float64 + stable argsort (oracle/geometry.py, "
                                                                "transformer_pointcloud_nova.py:81-89)"}
        except Exception as e:
            extras["geometry"] = {"error": str(e)[:300]}
        try:  # earth mover's distance: the assignment problem of cfg5's pairs, auction algorithm
            emd_b = pb[:, torch.randperm(Nc, device=dev, generator=gc)] * 0.9 + 0.1 * pa  # a shuffled, perturbed copy
            emd_b = emd_b.contiguous()
            torch.ops.nova_b200.emd(pa, emd_b, 1e-5)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            emd_v, _, emd_st = torch.ops.nova_b200.emd(pa, emd_b, 1e-5)
            e1.record()
            torch.cuda.synchronize()
            ems = e0.elapsed_time(e1)
            extras["emd"] = {"value": Bc / (ems * 1e-3), "unit": "cloud pairs/s", "ms": ems, "pairs": Bc, "points": Nc,
                             "bidding_rounds_mean": float(emd_st.float().abs().mean()), "converged": bool((emd_st > 0).all()),
                             "what": "minimum-cost perfect matching of %d x (%d vs %d) points (emd_approx, train_newloss.py:352-377): "
                                     "auction algorithm with epsilon scaling to 1e-5, one CTA per pair" % (Bc, Nc, Nc)}
            if not args.no_cpu_baseline:
                from oracle import chamfer as OCe

                t0 = time.perf_counter()
                ref_e = OCe.emd(pa[0].cpu().numpy(), emd_b[0].cpu().numpy())
                cpu_s = time.perf_counter() - t0
                extras["emd"]["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": "cloud pairs/s", "cores": 1, "kind": "port",
                                                 "sample": "1 pair, scipy cdist float64 + linear_sum_assignment (oracle/chamfer.py, demo.py:57-74)",
                                                 "abs_diff_of_mean_distance": abs(float(emd_v[0]) - ref_e)}
        except Exception as e:
            extras["emd"] = {"error": str(e)[:300]}
        try:  # training step: forward with saved activations + the full backward, cfg2's rows
            from nova_pointcloud_b200 import ops as _ops

            hh = head.handle()
            Mt, Tt, Dct = B * N, hh.cfg.token_dim, hh.cfg.cond_width
            xt_tr = torch.randn(Mt, Tt, device=dev, generator=gc)
            tt_tr = torch.rand(Mt, device=dev, generator=gc) * 1000
            zt_tr = z_d.reshape(Mt, Dct)
            shapes_tr = {k: tuple(p.shape) for k, p in head.named_parameters()}

            def train_step():
                v_tr, ws_tr = _ops.head_train_forward(hh, xt_tr, tt_tr, zt_tr)
                _ops.head_backward(hh, v_tr * 1e-3, xt_tr, zt_tr, ws_tr, shapes_tr, want_dz=True)

            for _ in range(2):
                train_step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                train_step()
            e1.record()
            torch.cuda.synchronize()
            tms = e0.elapsed_time(e1) / 3
            fwd_flop = 2.0 * Mt * (256 * D + D * D + Dct * D + D * D + (3 * 6 + 2) * D * D + 12 * D * D + 2 * Tt * D)
            extras["train_step"] = {
                "value": Mt / (tms * 1e-3), "unit": "tokens/s", "ms": tms, "rows": Mt,
                "tflops": 3.0 * fwd_flop / (tms * 1e-3) / 1e12,
                "frac_of_sustained_bf16": 3.0 * fwd_flop / (tms * 1e-3) / 1e12 / peaks()["sustained"],
                "what": "one training step of the head over %d tokens with per-token timesteps: nova_head_train_forward (saved "
                        "activations) + nova_head_backward (all 14 + 8*depth parameter gradients and dz), bf16 tcgen05 GEMMs; "
                        "FLOP = 3 x the forward's GEMM FLOP" % Mt}
        except Exception as e:
            extras["train_step"] = {"error": str(e)[:300]}
        try:  # the reference's algorithm as eager PyTorch on this GPU: the library-kernel bar
            with torch.no_grad():
                ref_out = eager_library_sampler(head, sched, z_d, noise_d)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(2):
                    eager_library_sampler(head, sched, z_d, noise_d)
                e1.record()
                torch.cuda.synchronize()
                lms = e0.elapsed_time(e1) / 2
                ours = nb.denoise(head, sched, z_d, noise_d)
            extras["library_bar"] = {
                "value": B / (lms * 1e-3), "unit": "clouds/s", "ms_per_step": lms,
                "what": "the reference algorithm as eager PyTorch bf16 on the same GPU (cuBLAS + ATen kernels, ~130 "
                        "launches per diffusion step, condition projection recomputed every step, bf16 latent); a "
                        "comparator, not the product path",
                "speedup_of_this_build": lms / ms,
                "rel_max_diff_of_outputs": float((ours - ref_out).abs().max() / ref_out.abs().max())}
        except Exception as e:
            extras["library_bar"] = {"error": str(e)[:300]}
        if not args.no_compile_bar:
            try:
                with torch.no_grad():
                    for _ in range(2):  # compile + settle
                        eager_library_sampler(head, sched, z_d, noise_d, compiled=True)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(2):
                        eager_library_sampler(head, sched, z_d, noise_d, compiled=True)
                    e1.record()
                    torch.cuda.synchronize()
                    cms2 = e0.elapsed_time(e1) / 2
                extras["library_bar_compiled"] = {"value": B / (cms2 * 1e-3), "unit": "clouds/s", "ms_per_step": cms2,
                                                  "what": "same algorithm, one diffusion step under torch.compile (inductor), "
                                                          "bf16, same GPU; a comparator, not the product path",
                                                  "speedup_of_this_build": cms2 / ms}
            except Exception as e:
                extras["library_bar_compiled"] = {"error": str(e)[:300]}

    if rank == 0:
        flops = algorithmic_flops(D, B * N)  # per GPU per step
        achieved = flops / (ms * 1e-3) / 1e12
        prof_total = sum(v[0] for v in prof.values()) or 1.0
        shares = {k: {"ms_per_step": v[0] / prof_steps, "launches_per_step": v[1] // prof_steps,
                      "share": v[0] / prof_total} for k, v in prof.items() if v[1]}
        ada_ms, ada_n = prof["gemm_ada"]
        ada_avg_ms = ada_ms / max(ada_n, 1)
        # the dominant kernel: the modulation GEMM, M x 2D x D, DEPTH + 1 launches per diffusion step (every launch has
        # the same shape; the gate columns of the AdaLN projection are the separate gemm_tail class)
        ada_flops = 2.0 * B * N * 2 * D * D
        ada_tf = ada_flops / (ada_avg_ms * 1e-3) / 1e12 if ada_n else 0.0
        tail_ms, tail_n = prof.get("gemm_tail", (0.0, 0))
        tail_avg_ms = tail_ms / max(tail_n, 1)
        tail_tf = 2.0 * B * N * D * D / (tail_avg_ms * 1e-3) / 1e12 if tail_n else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):  # dram bytes per launch from the committed ncu --set full capture
            with open(tpath) as f:
                traffic = json.load(f).get(f"gemm_adaln_M{B * N}_D{D}")
        line = {
            "metric": "point_clouds_per_sec", "value": value, "unit": "clouds/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": bench_config(args, wl, world),
            "head_tokens_per_s": value * N,
            "token_steps_per_s": value * N * S_STEPS,
            "e2e": {"value": e2e_value, "unit": "clouds/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(noise_h.numel() * 4 + z_h.numel() * 2) * world,
                    "d2h_bytes_per_step": int(out_h.numel() * 4)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": ada_tf, "peak": pk["sustained"], "unit": "TFLOP/s",
                         "frac": ada_tf / pk["sustained"], "traffic": traffic,
                         "kernel": "nova::tc::gemm_kernel<EPI_ADALN, cta_group 2>: modulation GEMM M x 2D x D of the AdaLN projection "
                                   "with h = LN(x)(1+scale)+shift in its epilogue (x staged by TMA); the dominant kernel of the step "
                                   "by time (kernel_shares.gemm_ada)",
                         "flop_per_launch": ada_flops, "avg_launch_ms": ada_avg_ms, "launches_timed": ada_n,
                         "how": "CUDA events recorded by the library around every launch on the launching stream, "
                                "inside real sampling steps (nova_profile_*)",
                         "peak_source": pk["source"] + " sustained bf16 (kernel timed inside a long step)",
                         "frac_of_burst": ada_tf / pk["burst"]},
            "roofline_tail": {"bound": "tensor", "achieved": tail_tf, "peak": pk["sustained"], "unit": "TFLOP/s",
                              "frac": tail_tf / pk["sustained"], "flop_per_launch": 2.0 * B * N * D * D,
                              "avg_launch_ms": tail_avg_ms, "launches_timed": tail_n,
                              "kernel": "nova::tc::gemm_kernel<EPI_TAIL, cta_group 2>: gate GEMM M x D x D with the block tail "
                                        "x += (LN(u) gamma + beta) gate in its epilogue (in place on TMA-staged chunks); "
                                        "epilogue-paced, the step's second kernel by time"},
            "step_roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s",
                              "frac": achieved / pk["sustained"], "frac_of_burst": achieved / pk["burst"],
                              "what": "whole sampling step per GPU: algorithmic FLOP (BASELINE.md section 3: "
                                      "F_min*B*N*S + 4*D^2*B*N, hoisted work not credited) / device time"},
            "kernel_shares": shares,
            "gemm_alone": gemm,
        }
        line.update(north)
        line.update(extras)
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            sample_clouds = 2
            v, secs, kind = cpu_baseline_run(D, N, sample_clouds, threads)
            line["cpu_baseline"] = {
                "value": v, "unit": "clouds/s", "cores": threads, "kind": kind, "seconds": secs,
                "sample": f"{sample_clouds} clouds x {N} tokens x {S_STEPS} steps, fp32, "
                          + ("the reference's own Transformer3DModel.denoise (baseline/_ref, unmodified)" if kind == "reference"
                             else "oracle/loop.py denoise (port: no reference copy present)")
                          + ", condition projection not hoisted"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
