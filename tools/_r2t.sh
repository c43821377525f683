#!/bin/bash
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2t_gputests.log 2>&1; tail -2 gpurun_out/r2t_gputests.log
python tools/sweep.py --quick > gpurun_out/r2t_sweep_quick.json 2> gpurun_out/r2t_sweep.err
SHPL_LIB=sparse_pooling_b200/libshpl_exp.so SHPL_ONE_WAVE=0 python tools/sweep.py --quick > gpurun_out/r2t_sweep_quick_nowave.json 2>> gpurun_out/r2t_sweep.err
python bench.py --steps 100 --warmup 5 > gpurun_out/r2t_bench_cfg2.json 2> gpurun_out/r2t_bench_cfg2.err
python bench.py --config 2p --steps 100 --warmup 5 > gpurun_out/r2t_bench_cfg2p.json 2> gpurun_out/r2t_bench_cfg2p.err
