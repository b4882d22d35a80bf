#!/bin/bash
# look-ahead of the reference-pixel statistics: parity test, A/B of the device-resident leg, tall-band check
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -m gpu -k "lookahead or pipeline_matches or medium" 2>&1 | tail -5 | tee $O/la_tests.log
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e"
for rep in 1 2; do
  for fl in "" "--no-refpix-lookahead"; do
    $B $fl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lookahead[$fl]', round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))" | tee -a $O/la_ab.log
  done
done
for br in 89 178; do
  $B --band-rows $br 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('v6 band_rows $br', round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))" | tee -a $O/la_ab.log
done
