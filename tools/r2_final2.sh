#!/bin/bash
O=gpurun_out; T=${1:-r2fin2}
timeout 1500 python -m pytest tests -m gpu -q > $O/tests_gpu_$T.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/tests_gpu_$T.log
python __graft_entry__.py smoke > $O/smoke_$T.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$T.log
timeout 600 python bench.py > $O/bench_$T.json 2> $O/bench_$T.err; echo "bench rc=$?"; cut -c1-200 $O/bench_$T.json
bash tools/r2_lines.sh $T
timeout 900 python bench.py --workload exposure18 --steps 3 --warmup 3 > $O/bench_exp18_$T.json 2> $O/bench_exp18_$T.err; echo "exp18 rc=$?"; cut -c1-300 $O/bench_exp18_$T.json
timeout 600 python bench.py --workload files --steps 4 > $O/bench_files_$T.json 2>/dev/null; cut -c1-300 $O/bench_files_$T.json
