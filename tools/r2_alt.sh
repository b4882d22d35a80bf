#!/bin/bash
# A/B of two builds of the library (romanimpreprocess_b200/_alt/<name>.so), alternating; kernel time of the device-resident leg
O=gpurun_out; mkdir -p $O
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-refpix-lookahead"
for rep in 1 2; do
  for name in "$@"; do
    cp romanimpreprocess_b200/_alt/$name.so romanimpreprocess_b200/librip_b200.so
    $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$name', round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))" | tee -a $O/alt_ab.log
  done
done
