#!/bin/bash
# ncu evidence for profiles/ (run under gpurun on ONE B200; every ncu command follows a plain run of the same
# command line that exited 0).  Usage: bash tools/profile_round.sh <tag>
tag=${1:-r1b}
set -x
B="python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline"
$B > gpurun_out/plain_${tag}_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches_bench.csv $B > gpurun_out/ncu_${tag}_bench.log 2>&1

M="python tools/microbench.py --config bench --iters 3"
$M > gpurun_out/plain_${tag}_micro.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:shpl_pool_sparse -c 12 -o gpurun_out/${tag}_pool_kernels -f $M > gpurun_out/ncu_${tag}_micro.log 2>&1

N="python tools/microbench.py --config b --iters 3"
$N > gpurun_out/plain_${tag}_micro_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:shpl_pool_sparse -c 6 -o gpurun_out/${tag}_narrow_kernels -f $N > gpurun_out/ncu_${tag}_micro_b.log 2>&1

F="python tools/feeder_bench.py"
$F > gpurun_out/plain_${tag}_feeder.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:shpl_bev -c 4 -o gpurun_out/${tag}_feeder_kernels -f $F > gpurun_out/ncu_${tag}_feeder.log 2>&1

V="python tools/mv3d_bench.py"
$V > gpurun_out/plain_${tag}_mv3d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"shpl_mv3d|shpl_radix" -c 7 -o gpurun_out/${tag}_mv3d_kernels -f $V > gpurun_out/ncu_${tag}_mv3d.log 2>&1
ls -la gpurun_out/${tag}_*
