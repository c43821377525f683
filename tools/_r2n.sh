#!/bin/bash
set -x
C="python tools/conv_check.py --time --scan"
$C > gpurun_out/r2n_conv_check_scan.txt 2>&1; tail -4 gpurun_out/r2n_conv_check_scan.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:shpl_conv --csv --log-file gpurun_out/r2n_conv_launches.csv $C > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2n_conv_launches.csv > gpurun_out/r2n_conv_launches_summary.txt; cat gpurun_out/r2n_conv_launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:shpl_conv3x3_dense --launch-skip 100 -c 1 -o gpurun_out/r2n_conv_dense_scan -f $C > gpurun_out/ncu_r2n.log 2>&1
ncu -i gpurun_out/r2n_conv_dense_scan.ncu-rep --page source --csv > gpurun_out/r2n_conv_dense_scan_source.csv 2>/dev/null
python tools/ncu_metrics.py gpurun_out/r2n_conv_dense_scan.ncu-rep > gpurun_out/r2n_conv_dense_scan_ncu_metrics.csv
