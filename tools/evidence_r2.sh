#!/bin/bash
# Round-2 evidence pass (one B200 under gpurun).  Every ncu command follows a plain run of the same command; numbers
# quoted as bench values come from the plain runs.  Usage: bash tools/evidence_r2.sh <tag> [sections]  (sections: t b s c m a d, default all but d)
tag=${1:-r2}
want=${2:-tbscma}
set -x
reduce() {      # reduce <name> [launches]: .ncu-rep -> _ncu_metrics.csv + _ncu_details.txt, drop the report
    python tools/ncu_metrics.py gpurun_out/$1.ncu-rep > gpurun_out/$1_ncu_metrics.csv
    ncu -i gpurun_out/$1.ncu-rep --page details --launch-count ${2:-1} > gpurun_out/$1_ncu_details.txt 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
}
if [[ $want == *t* ]]; then       # the GPU test suite + the memory-safety run's own log
python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_gputests.log 2>&1
SHPL_LIB=sparse_pooling_b200/libshpl_debug.so python tools/memsafety_run.py > gpurun_out/${tag}_memsafety.txt 2>&1
fi
if [[ $want == *b* ]]; then       # bench lines of every configuration, then the launch list + kernel captures of the default one
for c in 2 1 2p 3 4; do
python bench.py --config $c --steps 200 --warmup 10 > gpurun_out/${tag}_bench_cfg$c.json 2> gpurun_out/${tag}_bench_cfg$c.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err
B="python bench.py --steps 5 --warmup 3 --no-graph --no-cpu-baseline --no-feeder --no-conv --no-no-concat"
$B > gpurun_out/plain_${tag}_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches_bench.csv $B > gpurun_out/ncu_${tag}_bench.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_launches_bench.csv > gpurun_out/${tag}_launches_bench_summary.txt
M="python tools/microbench.py --config bench --iters 3"
$M > gpurun_out/plain_${tag}_micro.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:shpl_pool_sparse -c 12 -o gpurun_out/${tag}_pool_kernels -f $M > gpurun_out/ncu_${tag}_micro.log 2>&1
reduce ${tag}_pool_kernels
fi
if [[ $want == *s* ]]; then       # the config-5 sweep and captures of the staged instantiations
python tools/sweep.py --full > gpurun_out/${tag}_stress_sweep_full.json 2> gpurun_out/${tag}_sweep.err
for c in "1000000 16 uniform" "100000 64 zipf" "1000000 16 zipf"; do
t=$(echo $c | tr ' ' '_')
S="python tools/one_case.py $c"
$S > gpurun_out/${tag}_case_$t.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:shpl_pool --csv --log-file gpurun_out/${tag}_launches_$t.csv $S > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_launches_$t.csv > gpurun_out/${tag}_launches_${t}_summary.txt
ncu --set full --clock-control none --import-source on -k regex:shpl_pool_sparse -c 1 -o gpurun_out/${tag}_staged_$t -f $S > gpurun_out/ncu_${tag}_staged_$t.log 2>&1
reduce ${tag}_staged_$t
done
fi
if [[ $want == *c* ]]; then       # the fused conv: uniform and scan-pattern timings, launch list, captures of its two tcgen05 kernels
C="python tools/conv_check.py --time --scan"
python tools/conv_check.py --time > gpurun_out/${tag}_conv_check.txt 2>&1
$C > gpurun_out/${tag}_conv_check_scan.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:shpl_conv --csv --log-file gpurun_out/${tag}_conv_launches.csv $C > gpurun_out/ncu_${tag}_conv_l.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_conv_launches.csv > gpurun_out/${tag}_conv_launches_summary.txt
ncu --set full --clock-control none --import-source on -k regex:shpl_conv3x3_dense --launch-skip 100 -c 1 -o gpurun_out/${tag}_conv_dense_kernel -f $C > gpurun_out/ncu_${tag}_conv_d.log 2>&1
reduce ${tag}_conv_dense_kernel
ncu --set full --clock-control none --import-source on -k regex:shpl_conv_z_tc --launch-skip 4 -c 1 -o gpurun_out/${tag}_conv_z_kernel -f $C > gpurun_out/ncu_${tag}_conv_z.log 2>&1
reduce ${tag}_conv_z_kernel
W="python tools/conv_bwd_bench.py"
$W > gpurun_out/${tag}_conv_bwd.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:shpl_conv --csv --log-file gpurun_out/${tag}_conv_bwd_launches.csv $W > gpurun_out/ncu_${tag}_conv_b.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_conv_bwd_launches.csv > gpurun_out/${tag}_conv_bwd_launches_summary.txt
fi
if [[ $want == *d* ]]; then       # the bench line the driver's command prints (default flags) and configs[1] at 200 steps
python bench.py > gpurun_out/${tag}_bench_default_flags.json 2> gpurun_out/${tag}_bench_default_flags.err
python bench.py --config 2 --steps 200 --warmup 10 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err
fi
if [[ $want == *m* ]]; then       # feeders
python tools/feeder_bench.py > gpurun_out/${tag}_feeder_bench.json 2> gpurun_out/${tag}_feeder.err
python tools/mv3d_bench.py > gpurun_out/${tag}_mv3d_bench.json 2> gpurun_out/${tag}_mv3d.err
fi
if [[ $want == *a* ]]; then       # host time of the reference-signature API
python tools/profile_api.py > gpurun_out/${tag}_profile_api.txt 2>&1
fi
ls -la gpurun_out/${tag}_*
du -sh gpurun_out
