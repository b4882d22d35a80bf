#!/bin/bash
O=gpurun_out; T=${1:-r2e}
timeout 1700 python -m pytest tests -m gpu -q > $O/tests_gpu_$T.log 2>&1; echo "gpu tests rc=$?"; tail -15 $O/tests_gpu_$T.log
