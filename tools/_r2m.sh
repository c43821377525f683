#!/bin/bash
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2m_gputests.log 2>&1; tail -3 gpurun_out/r2m_gputests.log
python tools/conv_check.py --time > gpurun_out/r2m_conv_check.txt 2>&1; tail -3 gpurun_out/r2m_conv_check.txt
python tools/sweep.py --full > gpurun_out/r2m_sweep_product.json 2> gpurun_out/r2m_sweep.err
python bench.py --steps 50 --warmup 5 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -c 300 gpurun_out/r2m_bench.json
tail -3 gpurun_out/r2m_sweep.err
