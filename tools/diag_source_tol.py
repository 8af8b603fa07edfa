#!/usr/bin/env python
"""Where SOURCE mode's correlation departs from the oracle's, stage by stage (GPU box)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import tdoa_b200 as T
from oracle import oracle
from helpers import load_golden, fm_capture

for case in ("fm_strong", "weak_noise", "fm_delays"):
    raws, _ = load_golden(case)
    with T.Engine(T.MODE_SOURCE) as e:
        for k, r in enumerate(raws):
            e.load_u8(k, r)
        for kind, name in ((T.KIND_REF, "REF"), (T.KIND_TGT, "TGT")):
            pk = e.xcorr(kind)[0]
            want = oracle.process_capture_source(raws)[0 if kind == T.KIND_REF else 1]
            for p, w in zip(pk, want):
                print(case, name, "lag", int(p["lag"]), w[0], "corr", float(p["corr"]), w[1], "rel", abs(float(p["corr"]) - w[1]) / max(1.0, abs(w[1])))
        # preprocessed samples of station 0's target signal
        n = raws[0].size // 2 // 3
        out, power, branch = e.preprocess(0, T.KIND_TGT, 0, n)
        d = oracle.unpack_u8(raws[0])
        want, br = oracle.preprocess_source(oracle.extract_target(d))
        diff = np.abs(out - want)
        ulps = np.abs(out.view(np.float32).view(np.int32).astype(np.int64) - want.view(np.float32).view(np.int32).astype(np.int64))
        print(case, "preprocess branch", branch, br, "max abs", float(diff.max()), "max |x|", float(np.abs(want).max()),
              "samples differing", int(np.count_nonzero(out != want)), "of", out.size, "max ulps", int(ulps.max()))
with T.Engine(T.MODE_BINARY) as e:
    bad = e.selftest(1)
    print("selftest(1) differing quads:", bad)
    print("selftest(0) div mismatches:", e.selftest(0))
