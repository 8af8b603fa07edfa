#!/usr/bin/env python
"""bench.py -- station-pair cross-correlation throughput of the B200 TDOA engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): 3 stations x 100 s x 2 Msps dual-frequency uint8 IQ
captures (200 000 000 samples = 400 MB per station, blocks ref/target/ref), simulator.go
layout with FM content and integer sample delays (SURVEY.md 8d "Mode B"); every step runs
the full-length path: unpack -> FM discriminator -> DC -> box-car -> normalise per
station signal, all 3 station pairs of the reference signal (133 333 332 samples) and of
the target signal (66 666 666 samples) over lags 0..1999 with the reference binary's
block arithmetic, peak records, then the 2-D fix.

One step = one pass over one such capture set.  Metric: pair-Msamples/s =
sum over pairs and kinds of the signal length / step time.
`value` is timed with the captures already resident in HBM; `e2e` is the same step
through the C ABI with HOST (pinned) buffers: host->device copy of the three captures
and device->host read of the peak records and the fix inside the timed region.

N > 1 (torchrun): every rank processes its own capture set (weak scaling, no data-path
collective); the peak records are gathered with one NCCL all_gather.

Beside the headline the line carries
  parity_check.oracle_at_size  the CPU oracle (reference arithmetic) on the benchmark's own full-length captures:
                               every one of the six pairs at the reported peak lag +- 5 -- the lag must be
                               the oracle's arg-max and the correlation equal to 1e-6; a mismatch fails the run
  sharded                      BASELINE configs[3] (weak_signal_simulator content, 16 stations = 120 pairs, 66 REF +
                               33 TGT windows of 1 s) as ONE job dealt over the N GPUs inside the library
                               (tdoa_comm_init, in-place ncclAllGather of the peak records): fixed work, strong scaling
  cpu_restatement              BASELINE.md section 3, items 2-3: the C restatement of the binary's path on one host
                               thread and on all host cores ("restatement, not the reference build")
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

FS = 2e6
C_LIGHT = 299792458.0
STATION_LLH = np.array([
    [41.18660274289527, -95.96064116595667, 355.69],   # kx0u    (lat-lon-table.csv)
    [41.24669616513154, -96.08366304481238, 329.0],    # n3pay
    [41.32916620016985, -96.03513381562004, 373.18],   # kf0mtl
])
TX_LLH = np.array([41.20, -96.00, 400.0])              # simulator.go:229 usage example
STATION_NAMES = ["kx0u", "n3pay", "kf0mtl"]


def llh_to_ecef(lat, lon, h):
    a, f = 6378137.0, 1 / 298.257223563
    e2 = 2 * f - f * f
    la, lo = np.radians(lat), np.radians(lon)
    n = a / np.sqrt(1 - e2 * np.sin(la) ** 2)
    return np.array([(n + h) * np.cos(la) * np.cos(lo), (n + h) * np.cos(la) * np.sin(lo),
                     (n * (1 - e2) + h) * np.sin(la)])


def true_delays():
    tx = llh_to_ecef(*TX_LLH)
    d = np.array([np.linalg.norm(tx - llh_to_ecef(*s)) for s in STATION_LLH])
    k = np.rint(d / C_LIGHT * FS).astype(int)
    return k - k.min()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": sorted(reasons),
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


# ----------------------------------------------------------------------------- synthetic captures
def synth_captures_gpu(torch, device, block: int, seed: int):
    """Mode-B captures generated on the GPU (setup only, never timed): FM of box-car
    filtered Gaussian audio, amplitude 0.5, integer delays from the station geometry,
    AWGN sigma 0.02, quantised as simulator.go:150-160.  Returns uint8 tensors [3][6*block]."""
    g = torch.Generator(device=device)
    delays = true_delays()
    pad = int(delays.max()) + 64

    def fm(n, sd, dev):
        g.manual_seed(sd)
        a = torch.randn(n + 64, device=device, generator=g)
        cs = torch.cumsum(a, 0)
        a = (cs[64:] - cs[:-64]) / 64.0          # 64-tap box-car audio
        a = a / a.abs().max()
        ph = torch.cumsum(a.double(), 0) * (2 * np.pi * dev / FS)
        ph = torch.remainder(ph, 2 * np.pi).float()
        return 0.5 * torch.cos(ph), 0.5 * torch.sin(ph)

    ref_i, ref_q = fm(block + pad, 10 + seed, 75e3)
    tgt_i, tgt_q = fm(block + pad, 11 + seed, 60e3)
    caps = []
    for k in range(3):
        g.manual_seed(200 + k + 7 * seed)
        d = int(delays[k])
        raw = torch.empty(6 * block, dtype=torch.uint8, device=device)
        for b, (si, sq) in enumerate(((ref_i, ref_q), (tgt_i, tgt_q), (ref_i, ref_q))):
            for comp, src in ((0, si), (1, sq)):
                x = src[pad - d:pad - d + block] + 0.02 * torch.randn(block, device=device, generator=g)
                q = torch.clamp(x * 127.5 + 127.5, 0, 255).to(torch.uint8)  # truncating cast
                raw[2 * b * block + comp:2 * (b + 1) * block:2] = q
        caps.append(raw)
    return caps, delays


def synth_captures_cpu(block: int, seed: int = 0):
    from helpers import fm_capture
    delays = true_delays()
    return fm_capture(block, tuple(delays), tuple(delays), seed=seed), delays


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def run_reference_once(paths, csv, n_proc):
    """n_proc concurrent runs of the reference's shipped ELF (it is single-threaded);
    returns wall seconds of the slowest."""
    from oracle import oracle
    t0 = time.perf_counter()
    procs = [subprocess.Popen([str(oracle.REF_BINARY), "162400000", "92300000", str(csv), *map(str, paths)],
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for _ in range(n_proc)]
    outs = [p.communicate()[0] for p in procs]
    wall = time.perf_counter() - t0
    return wall, outs[0]


def reference_sample(block: int):
    """Bounded sample of the workload for the CPU arms: the same Mode-B content with
    `block`-sample blocks, written as .dat files the reference binary reads."""
    raws, delays = synth_captures_cpu(block)
    td = tempfile.mkdtemp(prefix="tdoa_ref_")
    paths = []
    for name, raw in zip(STATION_NAMES, raws):
        p = Path(td) / f"sim-{name}-1.dat"
        raw.tofile(p)
        paths.append(p)
    return paths, raws, delays


def cpu_baseline(block: int = 250_000, n_proc: int = 1):
    """Times the reference on this host: oracle/_ref/processor (the reference's own
    binary, kind "reference") if staged, else the C restatement (kind "port")."""
    from oracle import oracle
    paths, raws, _ = reference_sample(block)
    pair_samples = 3 * (2 * block + block)
    csv = ROOT / "tests" / "golden" / "stations.csv"
    if oracle.REF_BINARY.exists():
        wall, out = run_reference_once(paths, csv, n_proc)
        kind = "reference"
        ok = out.count("correlation=") >= 6
    else:
        t0 = time.perf_counter()
        oracle.process_capture_binary(raws)
        wall = time.perf_counter() - t0
        kind, n_proc, ok = "port", 1, True
    return {"value": n_proc * pair_samples / wall / 1e6, "unit": "pair-Msamples/s", "cores": n_proc, "kind": kind,
            "sample": f"3 stations x {3 * block} samples (blocks of {block}), 3 REF + 3 TGT pair correlations x 2000 lags, "
                      f"{n_proc} concurrent single-threaded run(s) of the reference's processor binary; wall {wall:.2f} s",
            "ok": ok}



# ----------------------------------------------------------------------------- oracle at size (checker, never timed)
def _oracle_signals(raws, block, kinds=(0, 1), workers=6):
    """The reference's preprocessing (oracle/tdoa_oracle.c: unpack, power, discriminator, DC, 11-tap
    box-car, normalise) of every station's REF (kind 0) / TGT (kind 1) signal, whole length, on host threads."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle

    def prep(job):
        k, kind = job
        raw = raws[k]
        if kind == 0:   # blocks 1 and 3 (processor.go:224-236)
            s = np.concatenate([oracle.unpack_u8(raw[:2 * block]), oracle.unpack_u8(raw[4 * block:6 * block])])
        else:           # block 2 (:258-263)
            s = oracle.unpack_u8(raw[2 * block:4 * block])
        y, branch = oracle.preprocess_binary(s)
        return job, y, branch

    jobs = [(k, kind) for kind in kinds for k in range(len(raws))]
    with ThreadPoolExecutor(max_workers=min(workers, len(jobs), os.cpu_count() or 1)) as pool:
        return {job: (y, br) for job, y, br in pool.map(prep, jobs)}


def oracle_check_at_size(raws, block, ref_pk, tgt_pk, max_lag=2000, span=5, tol=1e-6):
    """Config 2 at its full length against the oracle: for each of the 3 REF + 3 TGT pairs the
    reference's block-averaged correlation (binary_lag, ELF 0x49e027) at the engine's peak lag +- span.
    Asserts nothing itself; returns what bench.py puts into parity_check (and fails the run on)."""
    from oracle import oracle
    t0 = time.perf_counter()
    oracle.set_seq_dc_limit(4194304)   # beyond the reference's reach the engine's DC sum is exactly rounded
    try:
        sig = _oracle_signals(raws, block)
    finally:
        oracle.set_seq_dc_limit(-1)
    pairs = [(0, 1), (0, 2), (1, 2)]
    out, ok, worst = [], True, 0.0
    for kind, pk in ((0, ref_pk), (1, tgt_pk)):
        for p, (i, j) in enumerate(pairs):
            yi, yj = sig[(i, kind)][0], sig[(j, kind)][0]
            lag = int(pk[p]["lag"])
            d0, d1 = max(0, lag - span), min(max_lag - 1, lag + span)
            vals = oracle.tdcorr_binary_lags_mt(yi, yj[d0:], yi.size - max_lag, d1 - d0 + 1)   # equal lengths: template shortened by max_lag
            best = int(np.argmax(np.abs(vals)))   # first maximum, as the reference's strict `>` scan
            diff = abs(float(vals[lag - d0]) - float(pk[p]["corr"]))
            worst = max(worst, diff)
            good = best + d0 == lag and diff <= tol and sig[(i, kind)][1] == 0 and sig[(j, kind)][1] == 0
            ok &= good
            out.append({"kind": "REF" if kind == 0 else "TGT", "pair": [i, j], "engine_lag": lag, "oracle_argmax": best + d0,
                        "oracle_corr": float(vals[lag - d0]), "engine_corr": float(pk[p]["corr"]), "ok": bool(good)})
    return {"ok": bool(ok), "pairs": out, "lags_per_pair": 2 * span + 1, "max_abs_corr_diff": worst, "tolerance": tol,
            "samples_per_pair": [2 * block, block], "seconds": time.perf_counter() - t0,
            "what": "oracle/tdoa_oracle.c (the reference's arithmetic) on the benchmark's own full-length captures, "
                    "peak lag +- %d of all six pairs" % span}


def cpu_restatement(budget_block_st: int = 400_000, budget_block_mt: int = 6_000_000):
    """the C restatement of the shipped binary's path (oracle/tdoa_oracle.c) on
    whole signals -- no 1 000 000-sample truncation -- one host thread, then all host cores (OpenMP over lags,
    threads over signals).  Restatement, not the reference build; bounded samples of the workload."""
    from oracle import oracle
    res = {}
    for name, block, mt in (("single_thread", budget_block_st, False), ("all_cores", budget_block_mt, True)):
        raws, _ = synth_captures_cpu(block)
        t0 = time.perf_counter()
        sig = _oracle_signals(raws, block, workers=6 if mt else 1)
        n_lags, lag_found = 2000, []
        for kind in (0, 1):
            for i, j in ((0, 1), (0, 2), (1, 2)):
                yi, yj = sig[(i, kind)][0], sig[(j, kind)][0]
                if mt:
                    vals = oracle.tdcorr_binary_lags_mt(yi, yj, yi.size - n_lags, n_lags)
                    lag_found.append(int(np.argmax(np.abs(vals))))
                else:
                    lag_found.append(oracle.tdcorr_binary(yi, yj)[0])
        wall = time.perf_counter() - t0
        res[name] = {"value": 9 * block / wall / 1e6, "unit": "pair-Msamples/s", "cores": (os.cpu_count() or 1) if mt else 1,
                     "kind": "port", "wall_s": wall, "lags": lag_found,
                     "sample": f"3 stations x {3 * block} samples, whole-signal correlation of 3 REF ({2 * block}) + 3 TGT ({block}) "
                               f"pairs x {n_lags} lags, C restatement of the binary's path (restatement, not the reference build)"}
    return res


# ----------------------------------------------------------------------------- BASELINE configs[3], dealt over the GPUs
SHARD_W = 2_000_000


def sharded_block(args, torch, dist, T, device, local, rank, world):
    """weak_signal_simulator.go content (tools/simulators.py), 16 stations = 120 pairs, 1 s windows over the
    whole 100 s capture: 66 REF + 33 TGT windows x 120 pairs per step, EXTENDED mode (+-2000 lags, sub-sample
    vertex).  ONE job: the windows are dealt over the ranks inside tdoa_xcorr, all pairs of a window stay on one
    GPU, the records meet in one in-place ncclAllGather per call.  Strong scaling: the work does not grow with N."""
    sys.path.insert(0, str(ROOT / "tools"))
    import simulators as S
    block = args.shard_block
    stations = S.ring_stations(16)
    caps, _ = S.simulate_weak(stations, tuple(TX_LLH), 92300000.0, 10.0, 1000.0, block, seed=4242, device=device)
    eng = T.Engine(T.MODE_EXTENDED, n_stations=16, max_lag=2000, device=local)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    for k in range(16):
        eng.load_u8_device(k, caps[k].data_ptr(), caps[k].numel(), keep=caps[k])
    if world > 1:
        T.sharding.init_engine_comm(eng)
    nw_t, nw_r = block // SHARD_W, 2 * block // SHARD_W

    def step():
        # both pair loops in one call: the windows of both kinds are dealt over the ranks as one list (tdoa_xcorr_windows)
        return eng.xcorr_windows(0, SHARD_W, nw_r, nw_t, SHARD_W)

    steps, warm = max(1, min(args.steps, 5)), 3
    for _ in range(warm):
        r, t = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.stats()["launches_total"]
    e0.record()
    for _ in range(steps):
        r, t = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=device, dtype=torch.float64)
    launches = eng.stats()["launches_total"] - l0
    # every rank must hold the same table
    digest = torch.tensor([float(np.frombuffer(r.tobytes() + t.tobytes(), np.uint8).astype(np.int64).sum())], device=device,
                          dtype=torch.float64)
    same = True
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        lo, hi = digest.clone(), digest.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(lo.item() == hi.item())
    out = None
    if rank == 0:
        sec = float(ms.item()) * 1e-3
        units = (nw_r + nw_t) * 120
        # the oracle's statement of EXTENDED mode on one sampled TGT window, two pairs (CPU: seconds)
        verdict = sharded_oracle_sample(caps, block, t, nw_t) if world == 1 or args.shard_check else None
        out = {"workload": "configs[3]: weak_signal_simulator.go content (ref_power 10, tgt_power 1000; tools/simulators.py), 16 stations "
                           f"(120 pairs), {nw_r} REF + {nw_t} TGT windows of {SHARD_W} samples per step, EXTENDED mode, +-2000 lags",
               "scaling": "strong", "n_gpus": world, "ms_per_step": sec * 1e3, "value": units * SHARD_W / sec / 1e6,
               "unit": "pair-Msamples/s", "fixes_per_s": nw_t / sec, "window_pairs_per_step": units, "steps": steps, "warmup": warm,
               "gpu_launches_rank0": int(launches), "all_ranks_hold_the_same_table": same,
               "collective": "two in-place ncclAllGather of 32-byte peak records (REF table, TGT table) at the end of the one "
                             "tdoa_xcorr_windows call (inside the library); the ranks meet once per step"
                             if world > 1 else "none (one rank)",
               "windows_on_busiest_rank": -(-(nw_r + nw_t) // world),
               "oracle_sample": verdict}
    eng.close()
    del caps
    torch.cuda.empty_cache()
    return out


def sharded_oracle_sample(caps, block, tgt_table, nw_t, window=7, pairs=((0, 1), (3, 9))):
    from oracle import oracle
    t0 = time.perf_counter()
    w = min(window, nw_t - 1)
    oracle.set_seq_dc_limit(0)
    oracle.set_wide_boxcar_f64(33)
    try:
        ys = {}
        for k in sorted({s for pr in pairs for s in pr}):
            raw = caps[k][2 * block + 2 * w * SHARD_W:2 * block + 2 * (w + 1) * SHARD_W].cpu().numpy()
            ys[k] = oracle.preprocess_binary(oracle.unpack_u8(raw))
        res, ok = [], True
        all_pairs = [(i, j) for i in range(16) for j in range(i + 1, 16)]
        for i, j in pairs:
            c = oracle.xcorr_two_sided(ys[i][0], ys[j][0], 2000)
            idx, frac, val = oracle.peak_parabolic(c)
            got = tgt_table[w][all_pairs.index((i, j))]
            good = int(got["lag"]) == idx - 2000 and abs(float(got["frac"]) - frac) <= 1e-3 and abs(float(got["corr"]) - val) <= 1e-6
            ok &= good
            res.append({"pair": [i, j], "engine": [int(got["lag"]), float(got["frac"]), float(got["corr"])],
                        "oracle": [idx - 2000, float(frac), float(val)], "branches": [ys[i][1], ys[j][1]], "ok": bool(good)})
    finally:
        oracle.set_seq_dc_limit(-1)
        oracle.set_wide_boxcar_f64(0)
    return {"ok": bool(ok), "window": w, "pairs": res, "seconds": time.perf_counter() - t0,
            "what": "orc_preprocess_binary (EXTENDED arithmetic) + orc_xcorr_two_sided + orc_peak_parabolic on one TGT window"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_proc = max(1, min(os.cpu_count() or 1, 32))
    block = 250_000
    walls = []
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(block, n_proc)
        if not cb["ok"]:
            raise SystemExit("bench.py --impl reference: the reference binary did not print its six pair records")
        if i >= args.warmup:
            walls.append(3 * 3 * block * n_proc / (cb["value"] * 1e6))
    t = float(np.mean(walls))
    value = n_proc * 9 * block / t / 1e6
    line = {
        "impl": "reference", "metric": "station-pair xcorr throughput", "value": value, "unit": "pair-Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the workload is configs[1]; what this arm can run of it is a bounded sample (the binary truncates every
        # signal to 1 000 000 samples, processor.go:772-780 with the shipped chunk): the rate per pair-sample is
        # comparable (2000 lags on either side), the sizes are not -- both are stated
        "config": dict(WORKLOAD_CONFIG, samples_per_station=3 * block,
                       sample=f"each step: {n_proc} concurrent runs of the reference binary on 3 x {3 * block}-sample captures "
                              f"(blocks of {block}); the GPU arm runs 3 x 200 000 000-sample captures per step"),
        "cpu_baseline": {"value": value, "unit": "pair-Msamples/s", "cores": n_proc, "kind": cb["kind"],
                         "sample": cb["sample"]},
        "e2e": {"value": value, "unit": "pair-Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fixes_per_s": 0.0, "gpu_launches": 0,
    }
    OUT.emit(json.dumps(line))
    return 0


WORKLOAD_CONFIG = {
    "workload": "configs[1]: 3-station 100 s 2 Msps dual-frequency capture (3 x 200e6 samples, 400 MB each), "
                "full-length FM-demod correlation of 3 REF + 3 TGT pairs over lags 0..1999, single fix",
    "mode": "BINARY (reference binary arithmetic), chunk = whole signal",
    "stations": 3, "pairs": 3, "samples_per_station": 200_000_000, "max_lag": 2000, "block_size": 10000,
    "l2": "inputs (1.2 GB of captures per step) exceed the 126 MB L2; no explicit flush",
}


def bind_to_gpu_numa_node(index: int):
    """Run this rank on the CPUs next to its GPU (NVML's ideal affinity) before any pinned buffer is
    allocated, so the capture staging memory is first-touched on the GPU's own NUMA node and the
    host->device copies do not cross the socket interconnect.  Best effort; returns what was done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"cpus": len(allowed), "first": min(allowed)}
    except Exception as exc:  # no NVML, restricted cpuset, ...
        return {"error": str(exc)[:80]}
    return None


# ----------------------------------------------------------------------------- our arm
def main_ours(args):
    import torch
    import torch.distributed as dist
    import tdoa_b200 as T

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the engine has no CPU path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    if args.only_sharded:   # profiling aid: the sharded configs[3] block alone
        shard = sharded_block(args, torch, dist, T, device, local, rank, world)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        if rank == 0:
            OUT.emit(json.dumps({"sharded": shard}))
        return 0

    block = args.block
    n_samp = 3 * block
    caps, delays = synth_captures_gpu(torch, device, block, seed=rank)
    nbytes = caps[0].numel()
    # host (pinned) copies for the end-to-end leg
    pinned = [T.host_alloc(nbytes) for _ in range(3)]
    for k in range(3):
        pinned[k].array[:] = caps[k].cpu().numpy()

    eng = T.Engine(T.MODE_BINARY, chunk_samples=0, device=local, use_fft=1)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    # same engine settings, but the REF and TGT pair loops on ONE stream: in the product path their
    # kernels overlap (two streams), which is faster but makes a kernel's own duration unobservable;
    # the per-kernel roofline is measured on this serial twin, over the same number of steps
    eng_serial = T.Engine(T.MODE_BINARY, chunk_samples=0, device=local, use_fft=1, serial_kinds=1)
    eng_serial.set_stream(torch.cuda.current_stream().cuda_stream)
    pair_samples = 3 * (2 * block + block)
    pairs = [(0, 1), (0, 2), (1, 2)]
    gather_out = [torch.empty(6 * 32, dtype=torch.uint8, device=device) for _ in range(world)]

    stage = {k: 0.0 for k in ("ms_preprocess", "ms_fft", "ms_exact", "ms_fft_seg", "fft_launches", "fft_pair_samples",
                              "ms_demod", "ms_boxcar", "ms_cand", "demod_launches", "demod_samples",
                              "boxcar_launches", "boxcar_samples", "cand_launches", "cand_pair_samples")}
    collecting = [False]

    def collect():
        if collecting[0]:
            st = eng_serial.stats()
            for k in stage:
                stage[k] += st[k]

    def step_serial():
        eng_serial.process(STATION_LLH)
        collect()

    def step_resident():
        # tdoa_process = ProcessTDOA from the pair loops to the fix (processor.go:816-929): REF pair
        # loop, TGT pair loop, time / range differences, solveTDOA, queued on the device without
        # an intermediate host synchronisation; records and fix come back to the host
        r = eng.process(STATION_LLH)
        ref, tgt, pos, status = r["ref"], r["tgt"], r["position"], r["status"]
        if world > 1:
            # the path's single collective: every rank's peak records
            rec = torch.from_numpy(np.concatenate([ref, tgt]).view(np.uint8).copy()).to(device, non_blocking=True)
            dist.all_gather(gather_out, rec)
        return ref, tgt, pos, status

    def step_e2e():
        # host buffers in: the three captures sit in pinned host memory (tdoa_host_alloc);
        # tdoa_load_u8_pinned queues their copies (REF blocks of all stations first) and the
        # discriminator follows them chunk by chunk; records and fix come back to the host
        for k in range(3):
            eng.load_u8_pinned(k, pinned[k])
        return step_resident()

    for k in range(3):
        eng.load_u8_device(k, caps[k].data_ptr(), nbytes, keep=caps[k])
        eng_serial.load_u8_device(k, caps[k].data_ptr(), nbytes, keep=caps[k])

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness of the benchmarked path on this very input (not timed)
    ref, tgt, pos, status = step_resident()
    want = [int(delays[j] - delays[i]) for i, j in pairs]
    got_r, got_t = [int(x) for x in ref["lag"]], [int(x) for x in tgt["lag"]]
    clamp = [max(w, 0) for w in want]  # the reference searches non-negative lags only
    lags_ok = got_r == clamp and got_t == clamp
    # the reference's own arithmetic (CPU oracle) on these very captures, at their full length
    oracle_verdict = None
    if rank == 0 and world == 1 and not args.no_oracle_check:
        oracle_verdict = oracle_check_at_size([pb.array for pb in pinned], block, ref, tgt)

    # the same peaks through the engine's least-squares fix (all three range differences,
    # elevation held): the reference's own solver stops after ten half steps wherever it is
    # (processor.go:932-1020), which for these delays is nowhere near the transmitter
    ls_pos, ls_rms, ls_status, _ = eng.solve_ls(STATION_LLH, tgt["lag"].astype(np.float64) / FS * C_LIGHT,
                                                init_llh=[float(STATION_LLH[:, 0].mean()), float(STATION_LLH[:, 1].mean()), TX_LLH[2]], dims=2)
    ls_err = float(np.linalg.norm(llh_to_ecef(*ls_pos) - llh_to_ecef(*TX_LLH)))

    def timed(fn, steps, warmup, collect_stats=False):
        for _ in range(warmup):
            fn()
        sync_all()
        l0 = eng.stats()["launches_total"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        collecting[0] = collect_stats
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        collecting[0] = False
        sync_all()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), eng.stats()["launches_total"] - l0

    # clocks are sampled over both timed regions (nvidia-smi answers in ~100 ms, the
    # resident region alone is shorter than that at the default step count)
    with ClockSampler(local) as clocks:
        ms_res, launches = timed(step_resident, args.steps, args.warmup)
        ms_e2e, _ = timed(step_e2e, max(1, args.steps), max(3, args.warmup) if args.warmup else 0)
        ms_serial, _ = timed(step_serial, args.steps, args.warmup, collect_stats=True)

    t_step = ms_res / args.steps / 1e3
    t_e2e = ms_e2e / max(1, args.steps) / 1e3
    value = world * pair_samples / t_step / 1e6
    e2e_value = world * pair_samples / t_e2e / 1e6

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = json.loads(peaks_path.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    line = None
    if rank == 0:
        st = eng.stats()
        # per-kernel roofline: ALGORITHMIC bytes per launch = DISTINCT input + output bytes, each counted once
        # (SURVEY.md 8d's rule) / device time of the launch (CUDA events recorded by the
        # engine around the launch, on its own stream).  The correlation kernels read station planes that
        # several pairs share: 3 stations serve 3 pairs, so a pair-sample stands for 4 distinct bytes (one f32
        # of one plane), not the 8 a pair alone would read -- round 1 counted 8 and reported 0.42 / 1.23.
        traffic_db = {}
        tpath = ROOT / "profiles" / "traffic.json"
        if tpath.exists():
            traffic_db = json.loads(tpath.read_text())

        def kernel_line(name, bytes_per_unit, units_key, ms_key, n_key):
            n = stage[n_key]
            if n <= 0 or stage[ms_key] <= 0:
                return None
            alg = bytes_per_unit * stage[units_key] / n
            ms = stage[ms_key] / n
            ach = alg / (ms * 1e-3) / 1e9
            return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic_db.get(name), "traffic_source": traffic_db.get("_source") if traffic_db.get(name) else None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                    "kernel_ms_per_launch": ms, "share_of_step": stage[ms_key] / (ms_serial if ms_serial > 0 else 1.0),
                    "measured": "CUDA events around each launch, %d serial steps (serial_kinds=1) of the same workload" % args.steps}

        kernels = [k for k in (
            kernel_line("k_demod_df", 6.0, "demod_samples", "ms_demod", "demod_launches"),
            kernel_line("k_boxcar_small", 8.0, "boxcar_samples", "ms_boxcar", "boxcar_launches"),
            kernel_line("k_fft_tiles", 4.0, "fft_pair_samples", "ms_fft_seg", "fft_launches"),
            kernel_line("k_corr_candidates", 4.0, "cand_pair_samples", "ms_cand", "cand_launches"),
        ) if k]
        for k in kernels:
            if k["kernel"] == "k_fft_tiles":
                # the tile FFT is bound by FP32 issue, not by HBM (FFT is not a dense contraction: CUDA cores only).
                # Work per 6144-sample segment of a 3-station tile: two 8192-point complex transforms (5 N log2 N flop
                # each) + three cross-spectrum accumulations (4097 bins x 8 multiply-adds + the split of the packed spectra)
                seg = k["algorithmic_bytes_per_launch"] / 4.0 / 3.0 / 6144.0
                flop = seg * (2 * 5.0 * 8192 * 13 + 3 * 4097 * 2 * 8.0 + 4097 * 16.0)
                peak_tf = 148 * 128 * 2 * 1.965e9 / 1e12
                k["compute"] = {"bound": "fp32 (CUDA cores)", "flop_per_launch": flop, "achieved_tflops": flop / (k["kernel_ms_per_launch"] * 1e-3) / 1e12,
                                "peak_tflops": peak_tf, "frac": flop / (k["kernel_ms_per_launch"] * 1e-3) / 1e12 / peak_tf,
                                "note": "peak = 148 SMs x 128 FP32 lanes x 2 x 1.965 GHz; DESIGN.md section 4 has the instruction budget"}
        if not kernels:  # FFT path disabled: the exact every-lag correlator carries the step
            k_ms = stage["ms_exact"] / max(1, args.steps)
            ach = 8.0 * pair_samples / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
            kernels = [{"bound": "hbm", "kernel": "k_corr_brute", "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": 8.0 * pair_samples, "kernel_ms_per_launch": k_ms,
                        "share_of_step": 1.0}]
        dominant = max(kernels, key=lambda k: k["share_of_step"])
        cb = cpu_baseline() if world == 1 else None
        if cb is not None and not cb["ok"]:
            raise SystemExit("bench.py: the reference binary did not print its six pair records (cpu_baseline)")
        line = {
            "metric": "station-pair xcorr throughput", "value": value, "unit": "pair-Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(WORKLOAD_CONFIG, samples_per_station=n_samp,
                           parallelism=f"{world} rank(s), one capture set per rank, no data-path collective; "
                                       "one all_gather of 192-byte peak records per step" if world > 1 else "1 GPU"),
            "fixes_per_s": world / t_step,
            "e2e": {"value": e2e_value, "unit": "pair-Msamples/s", "h2d_bytes_per_step": 3 * nbytes,
                    # tdoa_process reads back: REF + TGT records, time and range differences, the fix
                    # (llh, status, iterations), the per-signal statistics and first-pass correlations
                    "d2h_bytes_per_step": 2 * 3 * 32 + 2 * 3 * 8 + 3 * 8 + 4 + 4 + 2 * 3 * 8 * 8 + 2 * 3 * 8,
                    "ms_per_step": t_e2e * 1e3,
                    "fixes_per_s": world / t_e2e,
                    # what the host path gives each GPU when all N pull at once (the copies are the step's floor)
                    "h2d_gbps_per_gpu": 3 * nbytes / t_e2e / 1e9},
            "gpu_launches": launches,
            "host_affinity": numa,
            "clocks": clocks.summary(),
            "roofline": dominant,
            "roofline_kernels": kernels,
            "stage_ms_per_step": {k: stage[k] / args.steps for k in ("ms_preprocess", "ms_fft", "ms_exact")},
            "serial_ms_per_step": ms_serial / args.steps,
            "overlap": "tdoa_process queues the TGT pair loop on a second stream beside the REF pair loop; value and e2e are "
                       "timed that way; roofline, roofline_kernels and stage_ms_per_step come from serial_ms_per_step's run",
            "parity_check": {"lags_match_injected_delays": bool(lags_ok), "ref_lags": got_r, "tgt_lags": got_t,
                             "injected": want, "n_candidates": [int(x) >> 16 & 255 for x in list(ref["flags"]) + list(tgt["flags"])], "fix_llh": [float(x) for x in pos], "fix_status": int(status),
                             "fix_note": "fix_llh is solveTDOA as the reference states it (its 10th half step, Z frozen); "
                                         "ls_fix_llh is tdoa_solve_ls on the same lags",
                             "ls_fix_llh": [float(x) for x in ls_pos], "ls_fix_error_m": ls_err, "ls_fix_rms_m": float(ls_rms),
                             "ls_fix_status": int(ls_status), "oracle_at_size": oracle_verdict},
            "cpu_baseline": {k: v for k, v in cb.items() if k != "ok"} if cb else None,
        }
    eng.close()
    eng_serial.close()
    del caps, pinned
    torch.cuda.empty_cache()
    # BASELINE configs[3] as one job over the N GPUs (every rank takes part)
    shard = None
    if not args.no_sharded:
        shard = sharded_block(args, torch, dist, T, device, local, rank, world)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    rc = 0
    if line is not None:
        line["sharded"] = shard
        if world == 1 and not args.no_cpu_restatement:
            line["cpu_restatement"] = cpu_restatement()
        bad = []
        if not lags_ok:
            bad.append("the benchmarked path did not recover the injected delays")
        if oracle_verdict is not None and not oracle_verdict["ok"]:
            bad.append("the benchmarked path disagrees with the CPU oracle at full size")
        if shard and (not shard["all_ranks_hold_the_same_table"] or (shard["oracle_sample"] and not shard["oracle_sample"]["ok"])):
            bad.append("the sharded configs[3] run failed its checks")
        if bad:   # a throughput of wrong results is not a result
            line["error"] = "; ".join(bad)
            line["value"] = None
            rc = 1
        OUT.emit(json.dumps(line))
    return rc


class OnlyJsonOnStdout:
    """Everything the libraries print while the run is in progress (e.g. NCCL's version
    banner) goes to stderr; stdout carries the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line: str):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)


OUT = None


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--block", type=int, default=66_666_666, help="samples per block (default: 100 s capture)")
    ap.add_argument("--shard-block", type=int, default=66_666_666, help="samples per block of the sharded configs[3] job")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded configs[3] block")
    ap.add_argument("--only-sharded", action="store_true", help="run the sharded configs[3] block alone (profiling aid)")
    ap.add_argument("--shard-check", action="store_true", help="run the oracle sample of the sharded block at N > 1 too")
    ap.add_argument("--no-oracle-check", action="store_true", help="skip the CPU oracle at full size (parity_check.oracle_at_size)")
    ap.add_argument("--no-cpu-restatement", action="store_true", help="skip the C restatement baselines (cpu_restatement)")
    args = ap.parse_args()
    with OnlyJsonOnStdout() as OUT:
        if args.impl == "reference":
            return main_reference(args)
        return main_ours(args)


if __name__ == "__main__":
    raise SystemExit(main())
