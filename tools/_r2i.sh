#!/bin/bash
set -x
python tools/conv_check.py --time > gpurun_out/r2i_conv_check.txt 2>&1; tail -4 gpurun_out/r2i_conv_check.txt
for lib in product keep0; do
  if [ $lib = keep0 ]; then export SHPL_LIB=sparse_pooling_b200/libshpl_exp.so SHPL_MAIN_KEEP=0; fi
  for c in "20000 16 zipf" "100000 64 zipf" "1000000 16 zipf"; do
    t=$(echo $c | tr ' ' '_')_$lib
    python tools/one_case.py $c > gpurun_out/r2i_case_$t.json 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:shpl_pool --csv --log-file gpurun_out/r2i_launches_$t.csv python tools/one_case.py $c > /dev/null 2>&1
    python tools/ncu_summary.py gpurun_out/r2i_launches_$t.csv > gpurun_out/r2i_launches_${t}_summary.txt
  done
done
