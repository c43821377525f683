#!/bin/bash
set -x
python -m pytest tests/test_gpu_conv.py -x -q -m gpu > gpurun_out/r2p_tests.log 2>&1; tail -2 gpurun_out/r2p_tests.log
C="python tools/conv_bwd_bench.py"
$C > gpurun_out/r2p_conv_bwd.txt 2>&1; tail -2 gpurun_out/r2p_conv_bwd.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:shpl_ --csv --log-file gpurun_out/r2p_conv_bwd_launches.csv $C > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2p_conv_bwd_launches.csv > gpurun_out/r2p_conv_bwd_launches_summary.txt; cut -c1-60,76- gpurun_out/r2p_conv_bwd_launches_summary.txt
