#!/bin/bash
# device-resident timings only (variant list in $2), optional extra bench args in $3
O=gpurun_out; T=${1:-ab}; VARS=${2:-"0 0"}
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
for v in $VARS; do
  RIP_FUSED_VARIANT=$v $B $3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $v', round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))" | tee -a $O/ab_$T.log
done
