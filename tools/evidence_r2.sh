#!/bin/bash
# Round-2 evidence pass (one B200 under gpurun): memory-safety run, conv launch list + ncu captures, the full config-5 sweep,
# the host-time profile of the reference-signature API.  Each ncu command follows a plain run of the same command.
tag=${1:-r2}
set -x
SHPL_LIB=sparse_pooling_b200/libshpl_debug.so python tools/memsafety_run.py > gpurun_out/${tag}_memsafety.txt 2>&1
C="python tools/conv_check.py --time"
$C > gpurun_out/${tag}_conv_check.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_conv_launches.csv $C > gpurun_out/ncu_${tag}_conv_l.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"shpl_conv3x3_dense|shpl_conv_z_tc" --launch-skip 12 -c 4 -o gpurun_out/${tag}_conv_kernels -f $C > gpurun_out/ncu_${tag}_conv.log 2>&1
python tools/ncu_metrics.py gpurun_out/${tag}_conv_kernels.ncu-rep > gpurun_out/${tag}_conv_kernels_ncu_metrics.csv
ncu -i gpurun_out/${tag}_conv_kernels.ncu-rep --page details > gpurun_out/${tag}_conv_kernels_ncu_details.txt 2>/dev/null
rm -f gpurun_out/${tag}_conv_kernels.ncu-rep
python tools/profile_api.py > gpurun_out/${tag}_profile_api.txt 2>&1
python tools/sweep.py --full > gpurun_out/${tag}_stress_sweep_full.json 2> gpurun_out/${tag}_sweep.err
ls -la gpurun_out/${tag}_*
