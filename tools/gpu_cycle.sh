#!/bin/bash
# One GPU iteration: parity tests (subset or all), a short bench, and the ncu launch list of a few steady-state frames.
# usage: tools/gpu_cycle.sh <tag> [pytest-args...]
tag=$1; shift
mkdir -p gpurun_out
python -m pytest "$@" -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${tag}_pytest.log
python bench.py --steps 60 --warmup 30 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cat gpurun_out/${tag}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 4 --warmup 30 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${tag}_launches.csv 110 | head -34
