#!/bin/bash
# GPU box: (1) ncu launch list of the bench command (base1080 legs only), (2) full captures of the GEMM at the M = 778 and M = 6224
# encoder shapes and of the warp kernel at 1080p / 4K -> raw CSV (dram bytes per launch for profiles/traffic.json)
tag=${1:-rX}
mkdir -p gpurun_out /tmp/prof
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 4 --warmup 3 --slots 1 --no-cpu-baseline --no-reference-cuda --no-large4k > gpurun_out/launches_${tag}.log 2>&1
ncu --set full --clock-control none -k regex:gemm_tc -c 40 -f -o /tmp/prof/gemm python tools/bench_gemm.py > gpurun_out/gemm_${tag}.log 2>&1
ncu -i /tmp/prof/gemm.ncu-rep --page raw --csv > gpurun_out/gemm_${tag}_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:warp_sbs_fast -c 12 -f -o /tmp/prof/warp python tools/bench_warp.py > gpurun_out/warpall_${tag}.log 2>&1
ncu -i /tmp/prof/warp.ncu-rep --page raw --csv > gpurun_out/warpall_${tag}_raw.csv 2>/dev/null
ls -la gpurun_out | grep ${tag}
