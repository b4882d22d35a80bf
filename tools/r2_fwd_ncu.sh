#!/bin/bash
O=gpurun_out; T=${1:-r2fw}
CMD="python bench.py --workload forward --n 2048 --steps 1 --warmup 1"
timeout 300 $CMD > $O/plain_fwd_$T.log 2>&1 || { echo plain failed; tail -3 $O/plain_fwd_$T.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fwd_ramp -c 1 -f -o $O/fwd_$T $CMD > $O/ncu_fwd_$T.log 2>&1; echo rc=$?
ncu -i $O/fwd_$T.ncu-rep --page raw --csv > $O/fwd_${T}_raw.csv 2>/dev/null
ncu -i $O/fwd_$T.ncu-rep --page source --csv > $O/fwd_${T}_source.csv 2>/dev/null
ls -la $O/fwd_$T*
