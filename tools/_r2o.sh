#!/bin/bash
set -x
python -m pytest tests/test_gpu_conv.py tests/test_gpu_memsafety.py -x -q -m gpu > gpurun_out/r2o_tests.log 2>&1; tail -3 gpurun_out/r2o_tests.log
python tools/conv_check.py --time > gpurun_out/r2o_conv_check.txt 2>&1; tail -3 gpurun_out/r2o_conv_check.txt
python tools/conv_check.py --time --scan > gpurun_out/r2o_conv_check_scan.txt 2>&1; tail -3 gpurun_out/r2o_conv_check_scan.txt
