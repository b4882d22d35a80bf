#!/bin/bash
# development tool (run under gpurun): GPU test suite, then a full ncu capture of the forward ramp kernel at n=2048
python -m pytest tests -m gpu -x -q > gpurun_out/tests_r1h.log 2>&1; echo "tests rc=$?"
python bench.py --workload forward --n 2048 --steps 2 --warmup 1 > gpurun_out/fwd_plain_r1h.log 2>&1 || exit 1
cat gpurun_out/fwd_plain_r1h.log
ncu --set full --clock-control none --import-source on -k regex:fwd_ramp -c 1 -f -o gpurun_out/fwd_r1h \
  python bench.py --workload forward --n 2048 --steps 1 --warmup 1 > gpurun_out/ncu_fwd_r1h.log 2>&1
ncu -i gpurun_out/fwd_r1h.ncu-rep --page raw --csv > gpurun_out/fwd_r1h_raw.csv 2>/dev/null
ncu -i gpurun_out/fwd_r1h.ncu-rep --page source --csv > gpurun_out/fwd_r1h_source.csv 2>/dev/null
tail -3 gpurun_out/tests_r1h.log
