#!/bin/bash
set -x
python -m pytest tests/test_gpu_conv.py tests/test_gpu_memsafety.py -x -q -m gpu > gpurun_out/r2f_conv_tests.log 2>&1; tail -3 gpurun_out/r2f_conv_tests.log
python tools/conv_check.py --time > gpurun_out/r2f_conv_check.txt 2>&1; tail -4 gpurun_out/r2f_conv_check.txt
python bench.py --steps 50 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 600 gpurun_out/r2f_bench.json
for c in "1000000 64 zipf" "1000000 16 zipf" "1000000 16 uniform" "100000 64 zipf"; do
  t=$(echo $c | tr ' ' '_')
  python tools/one_case.py $c > gpurun_out/r2f_case_$t.json 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:shpl_pool --csv --log-file gpurun_out/r2f_launches_$t.csv python tools/one_case.py $c > /dev/null 2>&1
  python tools/ncu_summary.py gpurun_out/r2f_launches_$t.csv > gpurun_out/r2f_launches_${t}_summary.txt
done
