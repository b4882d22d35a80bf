#!/bin/bash
# ncu of the noise-layer kernels (Poisson re-sampling after its rewrite, Pearson draws) + launch list of one exposure
O=gpurun_out; T=${1:-r2w}
CMD="python bench.py --workload noiselayers --steps 1 --layers 2"
timeout 600 $CMD > $O/plain_layers_$T.log 2>&1 || { echo "plain failed"; tail -5 $O/plain_layers_$T.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_layers_$T.csv $CMD > $O/ncu_ll_$T.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:poisson_resample -c 1 -f -o $O/pois_$T $CMD > $O/ncu_pois_$T.log 2>&1; echo "pois rc=$?"
ncu -i $O/pois_$T.ncu-rep --page raw --csv > $O/pois_${T}_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pearson_noise -c 1 -f -o $O/pearson_$T $CMD > $O/ncu_pearson_$T.log 2>&1; echo "pearson rc=$?"
ncu -i $O/pearson_$T.ncu-rep --page raw --csv > $O/pearson_${T}_raw.csv 2>/dev/null
ls -la $O/*_$T*
