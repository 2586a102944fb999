#!/bin/bash
# Round-2 (second half) evidence: bench line, reference arm, launch list of a few steady-state frames, one `ncu --set full` capture of
# the predict+update kernels and the occupancy kernels in a steady-state frame (raw page as CSV).
mkdir -p gpurun_out
python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02b_bench_reference_arm.json 2> gpurun_out/r02b_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 4 --warmup 30 --device-arm-only > gpurun_out/r02b_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02b_launches.csv 120 > gpurun_out/r02b_launch_summary.txt; head -40 gpurun_out/r02b_launch_summary.txt
ncu --set full --clock-control none --import-source on -k "regex:k_sweep|k_ck|k_weight_update|k_occ|k_birth|k_move" --launch-skip 400 --launch-count 16 -o gpurun_out/r02b_top -f \
    python bench.py --steps 4 --warmup 34 --device-arm-only > gpurun_out/r02b_top_ncu.log 2>&1
ncu -i gpurun_out/r02b_top.ncu-rep --page raw --csv > gpurun_out/r02b_top_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r02b_top_raw.csv | grep -A4 "^[a-z]\|^void" | grep -v "^--" | head -120
ls -la gpurun_out/r02b_*
