#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full captures of the warp kernel and the GEMMs.  The .ncu-rep files are
# exported to CSV on the box and deleted (gpurun_out/ is capped at 64 MiB).
# Usage: tools/gpu_profile.sh <tag> [what...]      what: launches warp gemm
tag=${1:-rX}; shift
what=${@:-launches warp gemm}
mkdir -p gpurun_out /tmp/prof
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --slots 1"
export_rep() {  # <rep> <out prefix>
    ncu -i $1 --page raw --csv > $2_raw.csv 2>/dev/null
    ncu -i $1 --page source --csv --print-source sass > $2_source_sass.csv 2>/dev/null || true
    ncu -i $1 --page details --csv > $2_details.csv 2>/dev/null || true
    gzip -f $2_source_sass.csv 2>/dev/null || true
}
for w in $what; do
case $w in
launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${tag}.csv $B > gpurun_out/launches_${tag}.log 2>&1 ;;
warp) ncu --set full --clock-control none --import-source on -k regex:warp_sbs -s 8 -c 1 -f -o /tmp/prof/warp $B > gpurun_out/warp_${tag}.log 2>&1
      export_rep /tmp/prof/warp.ncu-rep gpurun_out/warp_${tag} ;;
gemm) ncu --set full --clock-control none -k regex:gemm_tc -s 800 -c 79 -f -o /tmp/prof/gemm $B > gpurun_out/gemm_${tag}.log 2>&1
      ncu -i /tmp/prof/gemm.ncu-rep --page raw --csv > gpurun_out/gemm_${tag}_raw.csv 2>/dev/null ;;
esac
done
du -sh gpurun_out; ls -la gpurun_out
