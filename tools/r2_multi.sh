#!/bin/bash
# multi-GPU lines: default bench and exposure18 at N ranks
N=${1:-2}; O=gpurun_out; T=${2:-multi}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $O/bench_${N}gpu_$T.json 2> $O/bench_${N}gpu_$T.err; echo "bench rc=$?"; cut -c1-2600 $O/bench_${N}gpu_$T.json; tail -3 $O/bench_${N}gpu_$T.err
timeout 900 $TR --master-port 29512 bench.py --gpus $N --workload exposure18 --steps 3 --warmup 3 > $O/bench_exp18_${N}gpu_$T.json 2> $O/bench_exp18_${N}gpu_$T.err; echo "exp18 rc=$?"; cut -c1-2600 $O/bench_exp18_${N}gpu_$T.json; tail -3 $O/bench_exp18_${N}gpu_$T.err
(lscpu | grep -iE "numa|model name|^cpu\(s\)|socket"; nvidia-smi topo -m; free -g | head -2) > $O/sysinfo_${N}gpu_$T.log 2>&1
