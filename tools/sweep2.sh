#!/bin/bash
# GPU box: base1080 throughput under GEMM tile knobs
for cfg in "1 1" "2 0" "2 1"; do set -- $cfg
  echo -n "BN256=$1 PERSIST=$2: "
  D2S_GEMM_BN256=$1 D2S_GEMM_PERSIST=$2 timeout 200 python bench.py --steps 150 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f  e2e %.0f  u8 %.0f  serial net %.3f ms' % (d['value'], d['e2e']['value'], d['e2e_u8']['value'], d['serial']['stage_ms']['predict_depth']))"
done
