#!/bin/bash
# One short gpurun call: Chamfer tests of the current build, then the variant probe.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests/test_gpu_chamfer.py -x -q -m gpu -k "not emd" --tb=short > gpurun_out/chamfer_test.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/chamfer_test.log
timeout 200 python scripts/probe_chamfer_packed.py > gpurun_out/probe_chamfer_packed2.json 2> gpurun_out/probe_chamfer_packed2.err; echo "chamfer probe exit $?"; cat gpurun_out/probe_chamfer_packed2.json
