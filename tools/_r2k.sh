#!/bin/bash
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv.py tests/test_gpu_memsafety.py -x -q -m gpu > gpurun_out/r2k_gputests.log 2>&1; tail -3 gpurun_out/r2k_gputests.log
python tools/conv_check.py --time > gpurun_out/r2k_conv_check.txt 2>&1; tail -3 gpurun_out/r2k_conv_check.txt
python tools/sweep.py --full > gpurun_out/r2k_sweep_product.json 2> gpurun_out/r2k_sweep.err
SHPL_LIB=sparse_pooling_b200/libshpl_exp.so SHPL_MAIN_KEEP=0 python tools/sweep.py --full --only zipf > gpurun_out/r2k_sweep_keep0.json 2>> gpurun_out/r2k_sweep.err
tail -3 gpurun_out/r2k_sweep.err
