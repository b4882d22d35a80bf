#!/bin/bash
# ncu capture of the fused kernel for a given variant: bash tools/r2_ncu.sh TAG VARIANT [extra bench args]
O=gpurun_out; T=$1; V=$2; shift 2
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $*"
RIP_FUSED_VARIANT=$V timeout 300 $CMD > $O/plain_$T.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain_$T.log; exit 1; }
RIP_FUSED_VARIANT=$V timeout 900 ncu --set full --clock-control none --import-source on -k regex:cal_fused_v -s 3 -c 1 -f -o $O/fused_$T $CMD > $O/ncu_$T.log 2>&1
ncu -i $O/fused_$T.ncu-rep --page raw --csv > $O/fused_${T}_raw.csv 2>/dev/null
ncu -i $O/fused_$T.ncu-rep --page source --csv > $O/fused_${T}_source.csv 2>/dev/null
ls -la $O/fused_$T*
