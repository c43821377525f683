#!/bin/bash
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_memsafety.py -x -q -m gpu > gpurun_out/r2j_gputests.log 2>&1; tail -3 gpurun_out/r2j_gputests.log
python tools/sweep.py --full > gpurun_out/r2j_sweep_product.json 2> gpurun_out/r2j_sweep.err
SHPL_LIB=sparse_pooling_b200/libshpl_exp.so SHPL_MAIN_KEEP=0 python tools/sweep.py --full --only zipf > gpurun_out/r2j_sweep_keep0.json 2>> gpurun_out/r2j_sweep.err
C="python tools/conv_check.py --time"
$C > gpurun_out/r2j_conv_check.txt 2>&1; tail -3 gpurun_out/r2j_conv_check.txt
ncu --set full --clock-control none --import-source on -k regex:shpl_conv3x3_dense --launch-skip 100 -c 1 -o gpurun_out/r2j_conv_dense_pooled -f $C > gpurun_out/ncu_r2j.log 2>&1
python tools/ncu_metrics.py gpurun_out/r2j_conv_dense_pooled.ncu-rep > gpurun_out/r2j_conv_dense_pooled_ncu_metrics.csv
ncu -i gpurun_out/r2j_conv_dense_pooled.ncu-rep --page source --csv > gpurun_out/r2j_conv_dense_pooled_source.csv 2>/dev/null
tail -3 gpurun_out/r2j_sweep.err
