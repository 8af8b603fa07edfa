#!/usr/bin/env python
"""Discriminator kernels side by side (GPU box): the exhaustive 2^32-quad self-tests and the time of
one full-length launch per variant (fast_demod = 0 production, 2 round-1 kernel, 1 f32)."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
import tdoa_b200 as T
import bench

quick = "--quick" in sys.argv
import os
print("TDOA_DEMOD_VARIANT =", os.environ.get("TDOA_DEMOD_VARIANT", "0"))
with T.Engine(T.MODE_BINARY) as e:
    for which in (() if quick else (1, 2, 3)):
        t0 = time.time()
        v = e.selftest(which)
        print(f"selftest({which}) = {v}   [{e.last_error()}]  {time.time() - t0:.1f} s", flush=True)

dev = torch.device("cuda", 0)
block = 66_666_666
caps, delays = bench.synth_captures_gpu(torch, dev, block, 0)
for fd in ((0,) if quick else (0, 2, 1)):
    with T.Engine(T.MODE_BINARY, chunk_samples=0, use_fft=1, serial_kinds=1, fast_demod=fd) as e:
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        for k in range(3):
            e.load_u8_device(k, caps[k].data_ptr(), caps[k].numel(), keep=caps[k])
        acc = {"ms_demod": 0.0, "demod_launches": 0, "demod_samples": 0}
        for it in range(8):
            r = e.process(bench.STATION_LLH)
            if it >= 3:
                st = e.stats()
                for k in acc:
                    acc[k] += st[k]
        ms = acc["ms_demod"] / max(1, acc["demod_launches"])
        gbs = 6.0 * acc["demod_samples"] / max(1, acc["demod_launches"]) / (ms * 1e-3) / 1e9 if ms > 0 else 0
        print(f"fast_demod={fd}: {ms:.3f} ms per launch, {gbs:.0f} GB/s algorithmic = {gbs / 6557.4:.3f} of peak; "
              f"ref lags {[int(x) for x in r['ref']['lag']]} tgt corr {[float(x) for x in r['tgt']['corr']]}", flush=True)
