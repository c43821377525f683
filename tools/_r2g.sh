#!/bin/bash
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2g_gputests.log 2>&1; tail -3 gpurun_out/r2g_gputests.log
python tools/conv_check.py --time > gpurun_out/r2g_conv_check.txt 2>&1; tail -4 gpurun_out/r2g_conv_check.txt
python tools/sweep.py --full > gpurun_out/r2g_sweep_product.json 2> gpurun_out/r2g_sweep.err
export SHPL_LIB=sparse_pooling_b200/libshpl_exp.so
SHPL_MAIN_KEEP=0 python tools/sweep.py --full --only zipf > gpurun_out/r2g_sweep_keep0.json 2>> gpurun_out/r2g_sweep.err
SHPL_Q_SLICES=0 python tools/sweep.py --quick > gpurun_out/r2g_sweep_q0.json 2>> gpurun_out/r2g_sweep.err
SHPL_ENTRY_CHUNK=8 python tools/sweep.py --quick > gpurun_out/r2g_sweep_chunk8.json 2>> gpurun_out/r2g_sweep.err
SHPL_ENTRY_CHUNK=4 python tools/sweep.py --quick > gpurun_out/r2g_sweep_chunk4.json 2>> gpurun_out/r2g_sweep.err
tail -3 gpurun_out/r2g_sweep.err
