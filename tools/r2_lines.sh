#!/bin/bash
# secondary bench lines of the round: G=16 table, float64 ipc4d, forward ramp, noise layers, realisations
O=gpurun_out; T=${1:-r2t}
timeout 600 python bench.py --groups 16 --no-cpu-baseline > $O/bench_g16_$T.json 2> $O/bench_g16_$T.err; echo "g16 rc=$?"; cut -c1-400 $O/bench_g16_$T.json
timeout 600 python bench.py --ipc-dtype f64 --no-cpu-baseline > $O/bench_k64_$T.json 2> $O/bench_k64_$T.err; echo "k64 rc=$?"; cut -c1-400 $O/bench_k64_$T.json
timeout 600 python bench.py --workload forward --steps 5 > $O/bench_forward_$T.json 2> $O/bench_forward_$T.err; echo "fwd rc=$?"; cut -c1-600 $O/bench_forward_$T.json
timeout 900 python bench.py --workload noiselayers --steps 3 > $O/bench_noiselayers_$T.json 2> $O/bench_noiselayers_$T.err; echo "nl rc=$?"; cut -c1-800 $O/bench_noiselayers_$T.json; tail -3 $O/bench_noiselayers_$T.err
timeout 900 python bench.py --workload realizations --realizations 16 > $O/bench_realizations_$T.json 2> $O/bench_realizations_$T.err; echo "mr rc=$?"; cut -c1-800 $O/bench_realizations_$T.json; tail -3 $O/bench_realizations_$T.err
