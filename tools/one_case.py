#!/usr/bin/env python
"""One case of the stress sweep (for ncu launch lists / captures): python tools/one_case.py NNZ C SKEW [seed]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import sweep  # noqa: E402

if sys.argv[1] == "mv3d":      # BASELINE config 3's layer: MV3D middle-stage fusion, image stride 8 / BEV stride 2, C = 768
    print(json.dumps(sweep.case("cfg3 mv3d ped C768 (8,2)", (100, 120), (48, 160), 768, 20000, "ground", stride=(8, 2), weights=True)))
    sys.exit(0)
nnz, C, skew = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 0
print(json.dumps(sweep.case("cfg5 nnz%d C%d %s" % (nnz, C, skew), (700, 800), (360, 1200), C, nnz, skew, weights=True, seed=seed)))
