#!/bin/bash
# end-of-round evidence: full GPU suite, smoke, default bench, variant A/B, ncu of the default kernel, launch list
O=gpurun_out; T=${1:-r2fin}
timeout 1500 python -m pytest tests -m gpu -q > $O/tests_gpu_$T.log 2>&1; echo "gpu tests rc=$?"; tail -4 $O/tests_gpu_$T.log
python __graft_entry__.py smoke > $O/smoke_$T.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$T.log
timeout 600 python bench.py > $O/bench_$T.json 2> $O/bench_$T.err; echo "bench rc=$?"; cut -c1-300 $O/bench_$T.json
bash tools/r2_ab2.sh $T "0 4 0 4"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$T.csv $CMD > $O/ncu_l_$T.log 2>&1; echo "launch list rc=$?"
bash tools/r2_ncu.sh $T 4
