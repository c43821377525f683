#!/bin/bash
# ncu evidence for profiles/ (run under gpurun on ONE B200; every ncu command follows a plain run of the same
# command line that exited 0).  Usage: bash tools/profile_round.sh <tag> [sections]   sections: any of b m n f v h (default all)
# Each --set full report is reduced on the box to a metrics CSV (tools/ncu_metrics.py) and a details text, then
# deleted: gpurun copies back at most 64 MiB and the reports of one round add up to more.
tag=${1:-r1b}
want=${2:-bmnfvh}
set -x
reduce() {      # reduce <name>: .ncu-rep -> _ncu_metrics.csv + _ncu_details.txt (first launch), drop the report
    python tools/ncu_metrics.py gpurun_out/$1.ncu-rep > gpurun_out/$1_ncu_metrics.csv
    ncu -i gpurun_out/$1.ncu-rep --page details --launch-count 1 > gpurun_out/$1_ncu_details.txt 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
if [[ $want == *b* ]]; then
B="python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline"
$B > gpurun_out/plain_${tag}_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches_bench.csv $B > gpurun_out/ncu_${tag}_bench.log 2>&1
fi
if [[ $want == *m* ]]; then
M="python tools/microbench.py --config bench --iters 3"
$M > gpurun_out/plain_${tag}_micro.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:shpl_pool_sparse -c 12 -o gpurun_out/${tag}_pool_kernels -f $M > gpurun_out/ncu_${tag}_micro.log 2>&1
reduce ${tag}_pool_kernels
fi
if [[ $want == *n* ]]; then
N="python tools/microbench.py --config b --iters 3"
$N > gpurun_out/plain_${tag}_micro_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:shpl_pool_sparse -c 6 -o gpurun_out/${tag}_narrow_kernels -f $N > gpurun_out/ncu_${tag}_micro_b.log 2>&1
reduce ${tag}_narrow_kernels
fi
if [[ $want == *f* ]]; then
F="python tools/feeder_bench.py"
$F > gpurun_out/plain_${tag}_feeder.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:shpl_bev -c 4 -o gpurun_out/${tag}_feeder_kernels -f $F > gpurun_out/ncu_${tag}_feeder.log 2>&1
reduce ${tag}_feeder_kernels
fi
if [[ $want == *v* ]]; then
V="python tools/mv3d_bench.py"
$V > gpurun_out/plain_${tag}_mv3d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"shpl_mv3d|shpl_radix" -c 7 -o gpurun_out/${tag}_mv3d_kernels -f $V > gpurun_out/ncu_${tag}_mv3d.log 2>&1
reduce ${tag}_mv3d_kernels
fi
if [[ $want == *h* ]]; then
H="python tools/heavy_bench.py"
$H > gpurun_out/plain_${tag}_heavy.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:shpl_pool_heavy -c 8 -o gpurun_out/${tag}_heavy_kernels -f $H > gpurun_out/ncu_${tag}_heavy.log 2>&1
reduce ${tag}_heavy_kernels
fi
ls -la gpurun_out/${tag}_*
du -sh gpurun_out
