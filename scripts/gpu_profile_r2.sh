#!/bin/bash
# round 2: whole GPU suite, then ncu --set full of one block of the fused step (mod, fc1, fc2, tail GEMMs) with source
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu exit $?"; tail -3 gpurun_out/pytest_gpu.log
export NOVA_B200_GRAPH=0
python scripts/profile_step.py > gpurun_out/plain_step.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel' -s 62 -c 8 \
    -o gpurun_out/r2_prof_step -f python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1
echo "full capture exit $?"; tail -2 gpurun_out/ncu_step.log
