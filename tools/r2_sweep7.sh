#!/bin/bash
# fine band-height sweep of the v6 kernel (full-wave heights 89 / 178 are anomalously slow: what about their neighbours,
# and shorter bands than the default 49?)
O=gpurun_out; mkdir -p $O
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-refpix-lookahead"
for br in 25 30 33 36 40 44 47 48 49 50 52 88 90 100 177 180 200; do
  $B --band-rows $br 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v6 band_rows $br', round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))" | tee -a $O/sweep7.log
done
