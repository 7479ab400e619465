#!/bin/bash
# ncu full capture of the neighbourhood kernels at 256 x 2048 points, after a plain run of the same command.
set -u
mkdir -p gpurun_out
python scripts/profile_geometry.py > gpurun_out/plain_geometry.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'knn_kernel|softmax_interp_kernel' -s 12 -c 4 \
    -o gpurun_out/prof_geometry python scripts/profile_geometry.py > gpurun_out/ncu_geometry.log 2>&1
echo "geometry capture exit $?"
ncu -i gpurun_out/prof_geometry.ncu-rep --page raw --csv > gpurun_out/geometry_raw.csv 2>/dev/null
