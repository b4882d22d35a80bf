#!/bin/bash
O=gpurun_out; T=${1:-r2c}
for v in 1 2; do
  RIP_FUSED_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -x -q > $O/tests_fused_v${v}_$T.log 2>&1
  echo "variant $v fused tests rc=$?"; tail -3 $O/tests_fused_v${v}_$T.log
done
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for v in 0 1 2; do
  RIP_FUSED_VARIANT=$v $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $v', round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))"
done
for v in 2; do for b in 66 82 108 164; do
  echo -n "variant $v band $b: "
  RIP_FUSED_VARIANT=$v $B --band-rows $b 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))"
done; done
for v in 0 1 2; do
  RIP_FUSED_VARIANT=$v $B --groups 16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('G16 variant $v', round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))"
done
bash tools/r2_ncu.sh $T 2
