#!/bin/bash
# One short gpurun call: EMD tests of the current build, then the launch-form probe.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests/test_gpu_chamfer.py -x -q -m gpu -k "emd" --tb=short > gpurun_out/emd_test.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/emd_test.log
timeout 200 python scripts/probe_emd_variants.py > gpurun_out/probe_emd_variants2.json 2> gpurun_out/probe_emd_variants2.err; echo "emd probe exit $?"; cat gpurun_out/probe_emd_variants2.json
