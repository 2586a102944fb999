#!/bin/bash
# one `ncu --set full` capture of the named kernels (regex) in a steady-state frame; raw + source pages as CSV
# usage: tools/gpu_ncu.sh <tag> <kernel-regex> [launch-skip] [launch-count]
tag=$1; rx=$2; skip=${3:-30}; cnt=${4:-2}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip --launch-count $cnt -o gpurun_out/${tag} -f \
    python bench.py --steps 4 --warmup 34 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv > gpurun_out/${tag}_source.csv 2>/dev/null
ls -la gpurun_out/${tag}*
