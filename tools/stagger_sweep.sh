#!/bin/bash
# development tool: fused-kernel time vs CTA start stagger and band height (run under gpurun)
for st in 0 700 1400 2800; do for b in 64 62; do
  echo -n "stagger $st band $b: "
  RIP_V2_STAGGER_NS=$st python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --band-rows $b 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))"
done; done
