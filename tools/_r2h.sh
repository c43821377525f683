#!/bin/bash
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv.py tests/test_gpu_memsafety.py -x -q -m gpu > gpurun_out/r2h_gputests.log 2>&1; tail -3 gpurun_out/r2h_gputests.log
python tools/conv_check.py --time > gpurun_out/r2h_conv_check.txt 2>&1; tail -4 gpurun_out/r2h_conv_check.txt
python tools/sweep.py --full > gpurun_out/r2h_sweep_product.json 2> gpurun_out/r2h_sweep.err
export SHPL_LIB=sparse_pooling_b200/libshpl_exp.so
SHPL_MAIN_KEEP=0 python tools/sweep.py --full --only zipf > gpurun_out/r2h_sweep_keep0.json 2>> gpurun_out/r2h_sweep.err
SHPL_STAGED_BATCHES=8 python tools/sweep.py --full --only uniform > gpurun_out/r2h_sweep_b8.json 2>> gpurun_out/r2h_sweep.err
unset SHPL_LIB
C="python tools/one_case.py 1000000 16 uniform"
$C > gpurun_out/r2h_case.json 2>&1 && ncu --set full --clock-control none --import-source on -k regex:shpl_pool_sparse -c 2 -o gpurun_out/r2h_staged_c16 -f $C > gpurun_out/ncu_r2h.log 2>&1
python tools/ncu_metrics.py gpurun_out/r2h_staged_c16.ncu-rep > gpurun_out/r2h_staged_c16_ncu_metrics.csv
ncu -i gpurun_out/r2h_staged_c16.ncu-rep --page details --launch-count 1 > gpurun_out/r2h_staged_c16_ncu_details.txt 2>/dev/null
tail -3 gpurun_out/r2h_sweep.err
