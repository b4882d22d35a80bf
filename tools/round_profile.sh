#!/bin/bash
# Round-end evidence run (under gpurun, one GPU): GPU tests, every bench line, ncu launch list + full capture of the fused
# kernel and of the forward ramp kernel.  Everything lands in gpurun_out/ with the tag given as $1.
T=${1:-r1z}
O=gpurun_out
python -m pytest tests -m gpu -q > $O/tests_$T.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/tests_$T.log
python bench.py > $O/bench_$T.json 2> $O/bench_$T.err; echo "bench rc=$?"; cut -c1-600 $O/bench_$T.json
python bench.py --groups 16 --no-cpu-baseline > $O/bench_g16_$T.json 2>> $O/bench_$T.err
python bench.py --workload forward --steps 5 > $O/bench_forward_$T.json 2>> $O/bench_$T.err; cut -c1-300 $O/bench_forward_$T.json
python bench.py --workload realizations --realizations 64 > $O/bench_realizations_$T.json 2>> $O/bench_$T.err; cat $O/bench_realizations_$T.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_$T.json 2>> $O/bench_$T.err; cut -c1-300 $O/bench_reference_$T.json
python bench.py --workload noiselayers --steps 3 > $O/bench_noiselayers_$T.json 2>> $O/bench_$T.err; cut -c1-400 $O/bench_noiselayers_$T.json
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$T.log 2>&1; tail -1 $O/smoke_$T.log
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $O/plain_$T.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$T.csv $CMD > $O/ncu1_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cal_fused_v2 -s 3 -c 1 -f -o $O/fused_$T $CMD > $O/ncu2_$T.log 2>&1
ncu -i $O/fused_$T.ncu-rep --page raw --csv > $O/fused_${T}_raw.csv 2>/dev/null
ncu -i $O/fused_$T.ncu-rep --page source --csv > $O/fused_${T}_source.csv 2>/dev/null
FCMD="python bench.py --workload forward --steps 1 --warmup 1"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/fwd_launches_$T.csv $FCMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fwd_ramp -s 1 -c 1 -f -o $O/fwd_$T $FCMD > $O/ncu3_$T.log 2>&1
ncu -i $O/fwd_$T.ncu-rep --page raw --csv > $O/fwd_${T}_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:fwd_apportion -s 1 -c 1 -f -o $O/app_$T $FCMD > $O/ncu4_$T.log 2>&1
ncu -i $O/app_$T.ncu-rep --page raw --csv > $O/app_${T}_raw.csv 2>/dev/null
rm -f $O/app_$T.ncu-rep
echo done
