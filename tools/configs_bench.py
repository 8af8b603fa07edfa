#!/usr/bin/env python
"""Times the BASELINE.json configs other than the headline one (bench.py measures
configs[1]) and checks size-independent properties on them: injected integer delays are
recovered, every window agrees, the grid arg-min lands on the transmitter cell.

    python tools/configs_bench.py [--quick]      -> one JSON line per config
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import bench  # noqa: E402
import tdoa_b200 as T  # noqa: E402

FS, C = 2e6, 299792458.0


def ring_stations(n, seed=4242):
    """16-station layout of config 4: the 3 real collectors + synthetic ones on a ~25 km ring."""
    rng = np.random.default_rng(seed)
    st = [list(s) for s in bench.STATION_LLH]
    for k in range(n - 3):
        ang = 2 * np.pi * (k + rng.uniform(-0.2, 0.2)) / (n - 3)
        r_km = 25.0 * rng.uniform(0.8, 1.2)
        st.append([41.26 + r_km / 111.0 * np.cos(ang), -96.02 + r_km / (111.0 * np.cos(np.radians(41.26))) * np.sin(ang),
                   rng.uniform(300, 400)])
    return np.array(st[:n])


def delays_for(stations, tx=bench.TX_LLH):
    txe = bench.llh_to_ecef(*tx)
    d = np.array([np.linalg.norm(txe - bench.llh_to_ecef(*s)) for s in stations])
    k = np.rint(d / C * FS).astype(int)
    return k - k.min(), d


def synth(dev, n_st, block, delays, seed=0):
    """Mode-B captures for n_st stations (same generator as bench.py, any station count)."""
    g = torch.Generator(device=dev)
    pad = int(max(delays)) + 64

    def fm(n, sd, dv):
        g.manual_seed(sd)
        a = torch.randn(n + 64, device=dev, generator=g)
        cs = torch.cumsum(a, 0)
        a = (cs[64:] - cs[:-64]) / 64.0
        a = a / a.abs().max()
        ph = torch.remainder(torch.cumsum(a.double(), 0) * (2 * np.pi * dv / FS), 2 * np.pi).float()
        return 0.5 * torch.cos(ph), 0.5 * torch.sin(ph)

    ri, rq = fm(block + pad, 10 + seed, 75e3)
    ti, tq = fm(block + pad, 11 + seed, 60e3)
    caps = []
    for k in range(n_st):
        g.manual_seed(200 + k + 7 * seed)
        d = int(delays[k])
        raw = torch.empty(6 * block, dtype=torch.uint8, device=dev)
        for b, (si, sq) in enumerate(((ri, rq), (ti, tq), (ri, rq))):
            for comp, src in ((0, si), (1, sq)):
                x = src[pad - d:pad - d + block] + 0.02 * torch.randn(block, device=dev, generator=g)
                raw[2 * b * block + comp:2 * (b + 1) * block:2] = torch.clamp(x * 127.5 + 127.5, 0, 255).to(torch.uint8)
        caps.append(raw)
    return caps


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def cfg1(args, dev, res):
    # ---- config 1: 3-station 10 s capture, the reference's own chunked run (binary + source arithmetic)
    block = 2_000_000 if args.quick else 20_000_000
    delays, _ = delays_for(bench.STATION_LLH)
    caps = synth(dev, 3, block, delays)
    want = [int(delays[j] - delays[i]) for i in range(3) for j in range(i + 1, 3)]
    for mode, name in ((T.MODE_BINARY, "binary"), (T.MODE_SOURCE, "source")):
        with T.Engine(mode) as e:
            for k in range(3):
                e.load_u8_device(k, caps[k].data_ptr(), caps[k].numel(), keep=caps[k])
            dt, (r, t) = timed(lambda: (e.xcorr(T.KIND_REF)[0], e.xcorr(T.KIND_TGT)[0]))
            chunk = e.cfg.chunk_samples
        ok = [int(x) for x in t["lag"]] == (want if mode == T.MODE_BINARY else [0, 0, 0])
        res.append({"config": f"1 ({name} arithmetic, chunk {chunk})", "ms": dt * 1e3,
                    "pair_msamples_per_s": 6 * chunk / dt / 1e6, "lags": [int(x) for x in t["lag"]], "ok": bool(ok)})


def cfg2x(args, dev, res):
    # ---- config 2, extended variant: full-length target signals, +-50k lags, one fix
    delays, _ = delays_for(bench.STATION_LLH)
    want = [int(delays[j] - delays[i]) for i in range(3) for j in range(i + 1, 3)]
    block = 8_000_000 if args.quick else 66_666_666
    caps = synth(dev, 3, block, delays)
    L = 50_000
    with T.Engine(T.MODE_EXTENDED, max_lag=L, chunk_samples=0) as e:
        for k in range(3):
            e.load_u8_device(k, caps[k].data_ptr(), caps[k].numel(), keep=caps[k])
        dt, r = timed(lambda: e.process(bench.STATION_LLH), reps=args.reps)
    ok = [int(x) for x in r["tgt"]["lag"]] == want and [int(x) for x in r["ref"]["lag"]] == want
    res.append({"config": f"2x (EXTENDED, full length {2 * block}/{block} samples, +-{L} lags, f64 discriminator, 3 REF + 3 TGT pairs + fix)",
                "ms": dt * 1e3, "pair_msamples_per_s": 9 * block / dt / 1e6, "ok": bool(ok),
                "max_abs_frac": float(max(np.abs(r["tgt"]["frac"]).max(), np.abs(r["ref"]["frac"]).max()))})


def cfg3(args, dev, res):
    # ---- config 3: 100 s capture, sliding 1 s windows, +-50k lags, one fix per window
    delays, _ = delays_for(bench.STATION_LLH)
    want = [int(delays[j] - delays[i]) for i in range(3) for j in range(i + 1, 3)]
    block = 8_000_000 if args.quick else 66_666_666
    caps = synth(dev, 3, block, delays)
    W, L = 2_000_000, 50_000
    nw = block // W
    with T.Engine(T.MODE_EXTENDED, max_lag=L, fast_demod=1) as e:
        for k in range(3):
            e.load_u8_device(k, caps[k].data_ptr(), caps[k].numel(), keep=caps[k])

        def run3():
            pk = e.xcorr(T.KIND_TGT, 0, W, nw, W)
            rd = pk["lag"].astype(np.float64) / FS * C
            pos, status, _ = e.solve(bench.STATION_LLH, rd)
            return pk, pos

        dt, (pk, pos) = timed(run3, reps=args.reps)
    ok = all([int(x) for x in pk[w]["lag"]] == want for w in range(nw))
    res.append({"config": f"3 (EXTENDED, {nw} windows x 3 pairs, W=2e6, +-{L} lags)", "ms": dt * 1e3,
                "pair_msamples_per_s": nw * 3 * W / dt / 1e6, "fixes_per_s": nw / dt, "ok": bool(ok),
                "max_abs_frac": float(np.abs(pk["frac"]).max())})


def cfg4(args, dev, res):
    # ---- config 4: 16 stations (120 pairs), windowed
    W = 2_000_000
    st16 = ring_stations(16)
    d16, _ = delays_for(st16)
    block = 4_000_000 if args.quick else 66_666_666
    caps = synth(dev, 16, block, d16)
    nw = block // W
    want16 = [int(d16[j] - d16[i]) for i in range(16) for j in range(i + 1, 16)]
    with T.Engine(T.MODE_EXTENDED, n_stations=16, max_lag=2000, fast_demod=1) as e:
        for k in range(16):
            e.load_u8_device(k, caps[k].data_ptr(), caps[k].numel(), keep=caps[k])
        dt, pk = timed(lambda: e.xcorr(T.KIND_TGT, 0, W, nw, W), reps=args.reps)
    ok = all([int(x) for x in pk[w]["lag"]] == want16 for w in range(nw))
    res.append({"config": f"4 (EXTENDED, 16 stations, {nw} windows x 120 pairs, +-2000 lags)", "ms": dt * 1e3,
                "pair_msamples_per_s": nw * 120 * W / dt / 1e6, "fixes_per_s": nw / dt, "ok": bool(ok)})


def cfg5(args, dev, res):
    # ---- config 5: 1000 x 1000 grid, 16 stations, 64 windows
    st16 = ring_stations(16)
    _, dist = delays_for(st16)
    rd = np.array([dist[j] - dist[i] for i in range(16) for j in range(i + 1, 16)])
    rng = np.random.default_rng(5)
    sets = 8 if args.quick else 64
    rds = rd[None, :] + rng.normal(0, 50e-9 * C, (sets, rd.size))
    desc = [41.26 - 0.25, -96.02 - 0.25, 0.0005, 0.0005, 1000, 1000, 400.0]
    with T.Engine(T.MODE_BINARY) as e:
        dt, (out, cost, idx) = timed(lambda: e.grid(st16, desc, rds), reps=args.reps)
        e.solve_ls(st16, rds, init_llh=out, dims=2)   # first call loads the kernel
        t0 = time.perf_counter()
        fine, rms, status, iters = e.solve_ls(st16, rds, init_llh=out, dims=2)
        dt_ls = time.perf_counter() - t0
    err_m = [float(np.linalg.norm(bench.llh_to_ecef(*o) - bench.llh_to_ecef(*bench.TX_LLH))) for o in out]
    err_ls = [float(np.linalg.norm(bench.llh_to_ecef(o[0], o[1], bench.TX_LLH[2]) - bench.llh_to_ecef(*bench.TX_LLH)))
              for o in fine]
    res.append({"config": f"5 (grid 1000x1000, 16 stations, {sets} sets)", "ms": dt * 1e3,
                "cell_sets_per_s": 1e6 * sets / dt, "fixes_per_s": sets / dt, "median_err_m": float(np.median(err_m)),
                "ls_refine_ms": dt_ls * 1e3, "ls_median_err_m": float(np.median(err_ls)), "ls_median_rms_m": float(np.median(rms)),
                "ok": bool(np.median(err_m) < 60.0 and int(status.max()) == 0)})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", type=int, default=0, help="run one config only (1, 2, 3, 4 or 5)")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    res = []
    for n, fn in ((1, cfg1), (2, cfg2x), (3, cfg3), (4, cfg4), (5, cfg5)):
        if args.only in (0, n):
            fn(args, dev, res)
    for r in res:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
