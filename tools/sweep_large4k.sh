#!/bin/bash
# GPU box: large4k engine time under the GEMM tile-shape knobs
for cfg in "0 0" "1 0" "1 1" "1 2" "2 2"; do set -- $cfg
  echo -n "BN256=$1 PAIR=$2: "
  D2S_GEMM_BN256=$1 D2S_GEMM_PAIR=$2 timeout 200 python bench.py --workload large4k --steps 12 --warmup 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('fps %.0f engine %.2f ms  %.0f TF/s' % (d['value'], d['stage_ms']['engine (batch 8)'], d['roofline']['achieved']))"
done
