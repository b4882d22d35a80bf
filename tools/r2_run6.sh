#!/bin/bash
# default bench line, float64-ipc4d line, G=16 line, then an ncu --set full capture of the default fused kernel
O=gpurun_out; T=${1:-r2g}
timeout 600 python bench.py > $O/bench_$T.json 2> $O/bench_$T.err; echo "bench rc=$?"; cut -c1-900 $O/bench_$T.json
timeout 600 python bench.py --ipc-dtype f64 --no-cpu-baseline > $O/bench_k64_$T.json 2> $O/bench_k64_$T.err; echo "bench k64 rc=$?"; cut -c1-1500 $O/bench_k64_$T.json
bash tools/r2_ncu.sh $T 0
