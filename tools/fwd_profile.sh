#!/bin/bash
# development tool (run under gpurun): forward/sim GPU tests, forward bench, then ncu captures of the forward kernels at n=2048
python -m pytest tests/test_gpu_forward.py tests/test_gpu_sim.py -q > gpurun_out/tests_fwd_r1j.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/tests_fwd_r1j.log
python bench.py --workload forward --steps 3 --warmup 1 > gpurun_out/fwd_plain_r1j.log 2>&1 || exit 1
cat gpurun_out/fwd_plain_r1j.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fwd_launches_r1j.csv \
  python bench.py --workload forward --steps 1 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:fwd_ramp -c 1 -f -o gpurun_out/fwd_r1j \
  python bench.py --workload forward --n 2048 --steps 1 --warmup 1 > gpurun_out/ncu_fwd_r1j.log 2>&1
ncu -i gpurun_out/fwd_r1j.ncu-rep --page raw --csv > gpurun_out/fwd_r1j_raw.csv 2>/dev/null
ncu -i gpurun_out/fwd_r1j.ncu-rep --page source --csv > gpurun_out/fwd_r1j_source.csv 2>/dev/null
grep -v "^==" gpurun_out/fwd_launches_r1j.csv | tail -8
