#!/bin/bash
O=gpurun_out; T=${1:-r2d}
(lscpu | grep -iE "numa|model name|^cpu\(s\)|socket|thread"; nvidia-smi topo -m; free -g | head -2) > $O/sysinfo_$T.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/tests_gpu_$T.log 2>&1; echo "gpu tests rc=$?"; tail -5 $O/tests_gpu_$T.log
RIP_FUSED_VARIANT=3 timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -x -q -k "small or medium" > $O/tests_fused_v3_$T.log 2>&1; echo "v2t tests rc=$?"; tail -2 $O/tests_fused_v3_$T.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for v in 0 3 0 3; do
  RIP_FUSED_VARIANT=$v $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $v', round(d['value'],1), round(d['ms_per_step'],4), round(d['fused_ms'],4))"
done
timeout 600 python bench.py > $O/bench_$T.json 2> $O/bench_$T.err; echo "bench rc=$?"; cut -c1-1500 $O/bench_$T.json; tail -3 $O/bench_$T.err
