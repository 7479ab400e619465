#!/bin/bash
# ncu passes (one gpurun call): launch list of the bench command + full capture of one diffusion step.
set -u
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 715 -c 720 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list exit $?"
python scripts/profile_step.py > gpurun_out/plain_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|resid_kernel|embed_kernel|headout_kernel|prep_kernel' -s 60 -c 28 \
    -o gpurun_out/prof_step python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1
echo "full capture exit $?"
