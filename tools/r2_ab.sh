#!/bin/bash
# quick A/B of the fused kernel: parity subset, then device-resident timings (variant list in $2, default "0 0")
O=gpurun_out; T=${1:-ab}; VARS=${2:-"0 0"}
timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -x -q > $O/tests_fused_$T.log 2>&1; echo "fused tests rc=$?"; tail -3 $O/tests_fused_$T.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
for v in $VARS; do
  RIP_FUSED_VARIANT=$v $B $3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $v', round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))" | tee -a $O/ab_$T.log
done
