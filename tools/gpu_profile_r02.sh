#!/bin/bash
# Round-2 evidence: bench line, reference arm, launch list, one `ncu --set full` capture of every kernel of the predict+update
# stages (+ the header-streaming kernel) in a steady-state frame, raw + source pages as CSV.
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 4 --warmup 30 --no-cpu-baseline > gpurun_out/r02_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches.csv 120 > gpurun_out/r02_launch_summary.txt; head -40 gpurun_out/r02_launch_summary.txt
ncu --set full --clock-control none --import-source on -k "regex:k_vertex_mask|k_sweep|k_scan|k_fill$|k_sort_lists|k_gather_visible|k_ck|k_weight_update|k_occ_headers" --launch-skip 420 --launch-count 12 -o gpurun_out/r02_pu -f \
    python bench.py --steps 4 --warmup 34 --no-cpu-baseline > gpurun_out/r02_pu_ncu.log 2>&1
ncu -i gpurun_out/r02_pu.ncu-rep --page raw --csv > gpurun_out/r02_pu_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/r02_pu_raw.csv | grep -A3 "^[a-z]\|^void" | grep -v "^--" | head -80
ls -la gpurun_out/r02_*
