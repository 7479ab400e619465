#!/bin/bash
# One gpurun call: what the round-end driver runs (GPU test suite in one process, smoke, both bench arms), then an ncu
# full capture of the packed one-sweep Chamfer kernel (after the plain run of the same command exited 0).
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
bash scripts/gpu_round.sh
python scripts/profile_chamfer.py > gpurun_out/plain_chamfer.log 2>&1 &&
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'nn_sym2_kernel' -s 2 -c 1 \
    -o gpurun_out/r2_prof_chamfer_sym2 -f python scripts/profile_chamfer.py > gpurun_out/ncu_chamfer.log 2>&1
echo "chamfer capture exit $?"
ncu -i gpurun_out/r2_prof_chamfer_sym2.ncu-rep --page raw --csv > gpurun_out/r2_chamfer_sym2_raw.csv 2>/dev/null
