#!/bin/bash
# GPU box: targeted full captures (per-launch DRAM traffic, tensor-pipe activity) of the ViT-L batch-8 GEMMs and of the 4K warp kernel.
# Launch indices follow tools/bench_gemm.py (43 launches per shape: 4 base shapes, then qkv / proj / fc1 / fc2 at M = 6224) and
# tools/bench_warp.py (the 4K fp32 six-frame-set run starts at launch 184).
tag=${1:-rX}
mkdir -p gpurun_out /tmp/prof
for spec in qkv:175 fc1:261 fc2:304; do
  name=${spec%%:*}; skip=${spec##*:}
  ncu --set full --clock-control none -k regex:gemm_tc -s $skip -c 2 -f -o /tmp/prof/g_$name python tools/bench_gemm.py > gpurun_out/gemm6224_${name}_${tag}.log 2>&1
  ncu -i /tmp/prof/g_$name.ncu-rep --page raw --csv > gpurun_out/gemm6224_${name}_${tag}_raw.csv 2>/dev/null
done
ncu --set full --clock-control none -k regex:warp_sbs_fast -s 190 -c 6 -f -o /tmp/prof/w4k python tools/bench_warp.py > gpurun_out/warp4k_${tag}.log 2>&1
ncu -i /tmp/prof/w4k.ncu-rep --page raw --csv > gpurun_out/warp4k_${tag}_raw.csv 2>/dev/null
ls -la gpurun_out | grep ${tag}
