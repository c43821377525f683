#!/bin/bash
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2s_gputests.log 2>&1; tail -2 gpurun_out/r2s_gputests.log
python tools/sweep.py --full > gpurun_out/r2s_stress_sweep_full.json 2> gpurun_out/r2s_sweep.err
python bench.py --config 3 --steps 100 --warmup 5 > gpurun_out/r2s_bench_cfg3.json 2> gpurun_out/r2s_bench_cfg3.err
python bench.py --steps 100 --warmup 5 > gpurun_out/r2s_bench_cfg2.json 2> gpurun_out/r2s_bench_cfg2.err
python tools/mv3d_bench.py > gpurun_out/r2s_mv3d_bench.json 2>/dev/null
