#!/bin/bash
for sl in 12 16; do
  echo -n "slots=$sl: "
  timeout 200 python bench.py --steps 192 --warmup 5 --no-cpu-baseline --slots $sl 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f  e2e %.0f  u8 %.0f' % (d['value'], d['e2e']['value'], d['e2e_u8']['value']))"
done
