#!/bin/bash
# GPU box: full ncu capture of the 4K warp kernel (fp16 rgb -> fp32 HWC Full-SBS), exported to CSV (reports are too big to bring back)
tag=${1:-rX}
mkdir -p gpurun_out /tmp/prof
ncu --set full --clock-control none --import-source on -k regex:warp_sbs_fast -s 30 -c 1 -f -o /tmp/prof/warp python tools/bench_warp.py > gpurun_out/warp_${tag}.log 2>&1
ncu -i /tmp/prof/warp.ncu-rep --page raw --csv > gpurun_out/warp_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof/warp.ncu-rep --page source --csv --print-source sass > gpurun_out/warp_${tag}_source_sass.csv 2>/dev/null
gzip -f gpurun_out/warp_${tag}_source_sass.csv
ls -la gpurun_out | grep warp_${tag}
