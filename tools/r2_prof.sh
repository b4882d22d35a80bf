#!/bin/bash
# round-2 evidence: smoke, launch list, ncu --set full of the default fused kernel (after the plain run exited 0)
O=gpurun_out; T=${1:-r2r}
python __graft_entry__.py smoke > $O/smoke_$T.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_$T.log
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 $CMD > $O/plain_$T.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain_$T.log; exit 1; }
cat $O/plain_$T.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$T.csv $CMD > $O/ncu_l_$T.log 2>&1; echo "launch list rc=$?"
bash tools/r2_ncu.sh $T 0
