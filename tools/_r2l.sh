#!/bin/bash
set -x
for v in A B C D; do
  SHPL_LIB=sparse_pooling_b200/libshpl_var$v.so python tools/sweep.py --full > gpurun_out/r2l_sweep_$v.json 2>> gpurun_out/r2l_sweep.err
done
C="python tools/conv_check.py --time"
$C > gpurun_out/r2l_conv_check.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:shpl_conv_z_tc --launch-skip 4 -c 1 -o gpurun_out/r2l_conv_z -f $C > gpurun_out/ncu_r2l.log 2>&1
python tools/ncu_metrics.py gpurun_out/r2l_conv_z.ncu-rep > gpurun_out/r2l_conv_z_ncu_metrics.csv
ncu -i gpurun_out/r2l_conv_z.ncu-rep --page source --csv > gpurun_out/r2l_conv_z_source.csv 2>/dev/null
ncu -i gpurun_out/r2l_conv_z.ncu-rep --page details > gpurun_out/r2l_conv_z_details.txt 2>/dev/null
tail -3 gpurun_out/r2l_sweep.err
