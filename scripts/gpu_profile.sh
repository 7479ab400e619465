#!/bin/bash
# ncu passes (one gpurun call): launch list of the bench command, full capture of one diffusion step,
# full capture of the Chamfer kernel.  Each ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
export NOVA_B200_GRAPH=0   # profile the individual launches, not graph replays
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 715 -c 720 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit $?"
python scripts/profile_step.py > gpurun_out/plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|resid_kernel|embed3_kernel|headout3_kernel|prep_kernel' -s 60 -c 28 \
    -o gpurun_out/prof_step python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1
echo "full capture exit $?"
python scripts/profile_chamfer.py > gpurun_out/plain_chamfer.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:nn_kernel -s 2 -c 1 \
    -o gpurun_out/prof_chamfer python scripts/profile_chamfer.py > gpurun_out/ncu_chamfer.log 2>&1
echo "chamfer capture exit $?"
