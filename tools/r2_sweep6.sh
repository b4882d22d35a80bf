#!/bin/bash
O=gpurun_out; T=${1:-sw6}
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
for br in 0 49 56 62 66 74 83 98 111; do
  RIP_FUSED_VARIANT=4 $B --band-rows $br 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v6 band_rows $br', round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))" | tee -a $O/sweep_$T.log
done
