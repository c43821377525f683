for v in new base; do
  if [ $v = base ]; then export SHPL_LIB=build/base/libshpl_base.so; else unset SHPL_LIB; fi
  python tools/microbench.py --config b --iters 40 > gpurun_out/micro_ab_$v.json 2>> gpurun_out/micro_err.log
  python - <<PY
import json
d=json.load(open("gpurun_out/micro_ab_$v.json"))["pieces"]
print("$v", " ".join("%s=%.1f"%(k.split(".",1)[1], d[k]["us_median"]) for k in d if k.startswith("B_s1_c32.") and "nnz" not in k))
PY
done
unset SHPL_LIB
python tools/sweep.py > gpurun_out/sweep_r1d.json 2> gpurun_out/sweep_err.log; python - <<PY
import json
for r in json.load(open("gpurun_out/sweep_r1d.json"))["rows"]:
    print("%-30s nnz=%8d maxrow=%7d fwd %8.1fus %5d GB/s %.2f | bwd %8.1fus %5d GB/s %.2f" % (r["case"], r["nnz"], r["max_row"], r["fwd_us"], r["fwd_GBs"], r["fwd_frac"], r["bwd_us"], r["bwd_GBs"], r["bwd_frac"]))
PY
