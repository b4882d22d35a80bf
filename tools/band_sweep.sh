#!/bin/bash
# development tool: fused-kernel time vs band height (run under gpurun); prints value, ms/step, fused kernel ms
for b in "$@"; do
  echo -n "band $b: "
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --band-rows $b 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))"
done
