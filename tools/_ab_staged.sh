#!/bin/bash
# A/B of the staged entry walk's knobs through the experiment build (tools/sweep.py --full each)
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_memsafety.py -x -q -m gpu > gpurun_out/r2e_parity.log 2>&1; tail -3 gpurun_out/r2e_parity.log
python tools/sweep.py --full > gpurun_out/r2e_sweep_product.json 2> gpurun_out/r2e_sweep.err
export SHPL_LIB=sparse_pooling_b200/libshpl_exp.so
SHPL_STAGED=0 python tools/sweep.py --full > gpurun_out/r2e_sweep_staged0.json 2>> gpurun_out/r2e_sweep.err
SHPL_STAGED_BATCHES=1 python tools/sweep.py --full > gpurun_out/r2e_sweep_b1.json 2>> gpurun_out/r2e_sweep.err
SHPL_STAGED_BATCHES=4 python tools/sweep.py --full > gpurun_out/r2e_sweep_b4.json 2>> gpurun_out/r2e_sweep.err
SHPL_STAGED_DENSITY=64 python tools/sweep.py --full > gpurun_out/r2e_sweep_d64.json 2>> gpurun_out/r2e_sweep.err
SHPL_STAGED_ALL_VECS=16 python tools/sweep.py --full > gpurun_out/r2e_sweep_v16.json 2>> gpurun_out/r2e_sweep.err
tail -3 gpurun_out/r2e_sweep.err
