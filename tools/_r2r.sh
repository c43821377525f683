#!/bin/bash
set -x
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2r_tests.log 2>&1; tail -2 gpurun_out/r2r_tests.log
python tools/sweep.py --full --only uniform > gpurun_out/r2r_sweep_uniform.json 2> gpurun_out/r2r_sweep.err
python tools/sweep.py --full --only ground > gpurun_out/r2r_sweep_ground.json 2>> gpurun_out/r2r_sweep.err
SHPL_LIB=sparse_pooling_b200/libshpl_exp.so SHPL_ENTRY_CHUNK=32 python tools/sweep.py --quick > gpurun_out/r2r_sweep_chunk32.json 2>> gpurun_out/r2r_sweep.err
SHPL_LIB=sparse_pooling_b200/libshpl_exp.so SHPL_ENTRY_CHUNK=24 python tools/sweep.py --quick > gpurun_out/r2r_sweep_chunk24.json 2>> gpurun_out/r2r_sweep.err
