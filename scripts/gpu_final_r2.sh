#!/bin/bash
# One gpurun call: EMD variant probe, then what the round-end driver runs (GPU test suite in one process, smoke, both
# bench arms).  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python scripts/probe_emd_variants.py > gpurun_out/probe_emd_variants.json 2> gpurun_out/probe_emd_variants.err; echo "emd probe exit $?"; cat gpurun_out/probe_emd_variants.json
bash scripts/gpu_round.sh
