"""C = 16 Zipf forwards with the narrow-keep knob of the experiment build (SHPL_NARROW_KEEP)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import sweep  # noqa: E402
K = ((700, 800), (360, 1200))
for nnz in (5000, 20000, 100000, 1000000):
    for C in (8, 16):
        r = sweep.case("nnz%d C%d zipf" % (nnz, C), *K, C, nnz, "zipf", weights=True, seed=0)
        print(os.environ.get("SHPL_NARROW_KEEP", "default"), nnz, C, r["max_row"], r["fwd_us"], r["fwd_frac"], flush=True)
