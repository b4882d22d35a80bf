#!/bin/bash
# round 2, first GPU contact of the role-split TMA kernel (v3): parity of both role layouts, then A/B timing vs v2
O=gpurun_out; T=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi_$T.log 2>&1
for v in 1 2; do
  RIP_FUSED_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -x -q > $O/tests_fused_v${v}_$T.log 2>&1
  echo "variant $v fused tests rc=$?"; tail -3 $O/tests_fused_v${v}_$T.log
done
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for v in 0 1 2; do
  RIP_FUSED_VARIANT=$v $B > $O/bench_v${v}_$T.json 2> $O/bench_v${v}_$T.err; echo "variant $v rc=$?"
  python -c "import sys,json; d=json.loads(open('$O/bench_v${v}_$T.json').read()); print('variant $v', round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4), d['roofline']['frac'])"
done
for v in 1 2; do for b in 50 66 82 108 128 164; do
  echo -n "variant $v band $b: "
  RIP_FUSED_VARIANT=$v $B --band-rows $b 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))"
done; done
for v in 0 1; do
  RIP_FUSED_VARIANT=$v $B --groups 16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('G16 variant $v', round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))"
done
./tools/ubench_f32x2 > $O/ubench_$T.log 2>&1; tail -12 $O/ubench_$T.log
