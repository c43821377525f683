#!/bin/bash
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/r2q_tests.log 2>&1; tail -2 gpurun_out/r2q_tests.log
python tools/sweep.py --full --only zipf > gpurun_out/r2q_sweep_zipf.json 2> gpurun_out/r2q_sweep.err
python tools/conv_bwd_bench.py > gpurun_out/r2q_conv_bwd.txt 2>&1; tail -1 gpurun_out/r2q_conv_bwd.txt
