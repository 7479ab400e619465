#!/bin/bash
# One short gpurun call: geometry tests of the current build, then the packed / scalar probe.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests/test_gpu_geometry.py -x -q -m gpu --tb=short > gpurun_out/geometry_test.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/geometry_test.log
timeout 200 python scripts/probe_knn_packed.py > gpurun_out/probe_knn_packed.json 2> gpurun_out/probe_knn_packed.err; echo "knn probe exit $?"; cat gpurun_out/probe_knn_packed.json
