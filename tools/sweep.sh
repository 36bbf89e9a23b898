#!/bin/bash
# GPU box: throughput sweep over GEMM ring depth cap and frames in flight.
for st in 0 3 4 6; do for sl in 4 8; do
  echo -n "max_stages=$st slots=$sl: "
  D2S_GEMM_MAX_STAGES=$st python bench.py --steps 150 --warmup 5 --no-cpu-baseline --slots $sl 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f  e2e %.0f  serial_dev %.0f fps  net(serial) %.3f ms' % (d['value'], d['e2e']['value'], d['serial']['fps_device'], d['serial']['stage_ms']['predict_depth']))"
done; done
