run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>gpurun_out/ab_err.log | tail -1 > gpurun_out/ab_$name.json
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$name.json"))
r=d["roofline"]
print("$name", "value=%.0f"%d["value"], "ms=%.4f"%d["ms_per_step"], "fwdB=%.1fus"%r["us_per_launch"], "bwdB=%.1fus"%r["backward_kernel"]["us_per_launch"], "frac=%.3f"%r["frac"], "e2e=%.0f"%d["e2e"]["value"])
PY
}
run base SHPL_LIB=build/base/libshpl_base.so
run new X=1
run new12 SHPL_NARROW_CTAS_PER_SM=12
run new16 SHPL_NARROW_CTAS_PER_SM=16
run base2 SHPL_LIB=build/base/libshpl_base.so
