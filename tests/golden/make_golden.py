#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the reference's own
shipped binary (/root/reference/processor, staged as oracle/_ref/processor) on small
synthetic captures.  Run HERE (the build container); the GPU box only reads the
committed .npz/.json files.

    python tests/golden/make_golden.py

Each case stores the three station captures (uint8 IQ, the collector's .dat layout:
block1=ref, block2=tgt, block3=ref) and what the reference printed for them:
the six `REF|TGT a - b: delay=.. correlation=..` records (processor.go:826-828,
846-848) and the per-signal diagnostics (initial power, DC bias, pre-normalise power).
"""
from __future__ import annotations

import json
import re
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

FS = 2e6
STATIONS = ["kx0u", "n3pay", "kf0mtl"]

# lat-lon-table.csv of the reference (station coordinates are inputs, not code).
STATION_CSV = """Name,Latitude,Longitude,Elevation
KEVO,41.30888549464701,-96.02619229605524,356.0
162400000,41.25703803095629,-95.95512763589404,349.07
kx0u,41.18660274289527,-95.96064116595667,355.69
n3pay,41.24669616513154,-96.08366304481238,329.0
kf0mtl,41.32916620016985,-96.03513381562004,373.18
"""


def quantise(x: np.ndarray) -> np.ndarray:
    """simulator.go:150-160: byte(clamp(v*127.5+127.5, 0, 255)), truncating cast."""
    raw = np.empty(2 * len(x), np.uint8)
    raw[0::2] = np.clip(x.real * 127.5 + 127.5, 0, 255).astype(np.uint8)
    raw[1::2] = np.clip(x.imag * 127.5 + 127.5, 0, 255).astype(np.uint8)
    return raw


def audio(n, seed, taps=50):
    a = np.random.default_rng(seed).standard_normal(n + 4 * taps)
    a = np.convolve(a, np.ones(taps) / taps, "same")
    return a / np.abs(a).max()


def fm(n, seed, dev, amp=0.5):
    return amp * np.exp(1j * 2 * np.pi * np.cumsum(audio(n, seed)) * dev / FS)


def case_fm_strong(B=40000):
    """strong FM, integer delays 0/7/3 (ref and tgt alike)."""
    ref, tgt = fm(B, 10, 75e3), fm(B, 11, 75e3)
    out = {}
    for name, dl in zip(STATIONS, (0, 7, 3)):
        def blk(sig, seed):
            g = np.random.default_rng(seed)
            return sig[100 - dl:100 - dl + B] + 0.01 * (g.standard_normal(B) + 1j * g.standard_normal(B))
        out[name] = quantise(np.concatenate([blk(ref, 1), blk(tgt, 2), blk(ref, 3)]))
    return out


def case_fm_delays(B=50000):
    """Strong FM, distinct ref/tgt delays incl. one beyond the 120-sample sanity window
    (first-pass peak kept: nothing above half its height inside [0,120))."""
    ref, tgt = fm(B + 600, 20, 75e3), fm(B + 600, 21, 60e3)
    out = {}
    for k, (name, dr, dt) in enumerate(zip(STATIONS, (0, 5, 11), (0, 33, 301))):
        def blk(sig, d, seed):
            g = np.random.default_rng(seed)
            return sig[400 - d:400 - d + B] + 0.02 * (g.standard_normal(B) + 1j * g.standard_normal(B))
        out[name] = quantise(np.concatenate([blk(ref, dr, 30 + k), blk(tgt, dt, 40 + k), blk(ref, dr, 50 + k)]))
    return out


def case_moderate(B=40000):
    """power in (0.001, 0.01] -> envelope branch: AM carrier, amplitude ~0.07."""
    out = {}
    for k, (name, dl) in enumerate(zip(STATIONS, (0, 9, 4))):
        def blk(seed_a, seed_n):
            a = audio(B + 200, seed_a, taps=20)[:B + 200]
            env = 0.07 * (1.0 + 0.8 * a)
            ph = 2 * np.pi * 0.013 * np.arange(B + 200)
            s = (env * np.exp(1j * ph))[100 - dl:100 - dl + B]
            g = np.random.default_rng(seed_n)
            return s + 0.004 * (g.standard_normal(B) + 1j * g.standard_normal(B))
        out[name] = quantise(np.concatenate([blk(60, 70 + k), blk(61, 80 + k), blk(60, 90 + k)]))
    return out


def case_weak_tones(B=40000):
    """simulator.go:67-97,119-138 style tones (amp 0.01 ref / ~0.02 tgt, uniform noise
    +-0.01): power < 0.001 -> weak branch; periodic peaks exercise the sanity re-search."""
    out = {}
    i = np.arange(B)
    for k, (name, dist) in enumerate(zip(STATIONS, (9000.0, 14000.0, 21000.0))):
        g = np.random.default_rng(100 + k)
        def noise():
            return (g.random(B) - 0.5) * 2 * 0.01 + 1j * (g.random(B) - 0.5) * 2 * 0.01
        ref = 0.01 * np.exp(1j * (2 * np.pi * (162.4e6 % FS) * i / FS))
        amp_t = 0.1 * 1000.0 / dist * 2.0
        ph = 2 * np.pi * 92.3e6 * dist / 299792458.0
        tgt = amp_t * np.exp(1j * (2 * np.pi * (92.3e6 % FS) * i / FS + ph))
        out[name] = quantise(np.concatenate([ref + noise(), tgt + noise(), ref + noise()]))
    return out


def case_weak_noise(B=40000):
    """Weak broadband signal (band-limited noise, amp ~0.02, delays 0/6/2): weak branch
    with a genuine correlation peak."""
    out = {}
    def src(seed):
        g = np.random.default_rng(seed)
        s = g.standard_normal(B + 300) + 1j * g.standard_normal(B + 300)
        s = np.convolve(s, np.ones(6) / 6, "same")
        return 0.02 * s / np.sqrt(np.mean(np.abs(s) ** 2))
    r, t = src(300), src(301)
    for k, (name, dl) in enumerate(zip(STATIONS, (0, 6, 2))):
        def blk(sig, seed):
            g = np.random.default_rng(seed)
            return sig[100 - dl:100 - dl + B] + 0.003 * (g.standard_normal(B) + 1j * g.standard_normal(B))
        out[name] = quantise(np.concatenate([blk(r, 310 + k), blk(t, 320 + k), blk(r, 330 + k)]))
    return out


def case_fm_ragged(B=30001):
    """Strong FM in a capture whose length is not a multiple of three (N = 3 B + 2): the block
    is N / 3 in integer division and the two trailing samples belong to no block
    (processor.go:211-214, :244-246)."""
    ref, tgt = fm(B + 600, 22, 75e3), fm(B + 600, 23, 60e3)
    out = {}
    for k, (name, dr, dt) in enumerate(zip(STATIONS, (0, 4, 9), (0, 17, 2))):
        def blk(sig, d, seed, n=B):
            g = np.random.default_rng(seed)
            return sig[400 - d:400 - d + n] + 0.02 * (g.standard_normal(n) + 1j * g.standard_normal(n))
        tail = 0.3 * np.exp(1j * np.arange(2))   # two samples past the third block
        out[name] = quantise(np.concatenate([blk(ref, dr, 130 + k), blk(tgt, dt, 140 + k), blk(ref, dr, 150 + k), tail]))
    return out


def case_fm_truncated(B=1_050_000):
    """Blocks longer than the binary's 1 000 000-sample test chunk: REF (2 B samples) and TGT
    (B samples) are both cut to their first 1 000 000 samples before the pair loops
    (processor.go:772-780 with the shipped binary's chunk).  18 MB of captures: not stored --
    the tests regenerate them from the seeds and check the SHA-256 kept in the .json."""
    ref, tgt = fm(B + 600, 24, 75e3), fm(B + 600, 25, 60e3)
    out = {}
    for k, (name, dr, dt) in enumerate(zip(STATIONS, (0, 6, 13), (0, 41, 8))):
        def blk(sig, d, seed):
            g = np.random.default_rng(seed)
            return sig[400 - d:400 - d + B] + 0.02 * (g.standard_normal(B) + 1j * g.standard_normal(B))
        out[name] = quantise(np.concatenate([blk(ref, dr, 230 + k), blk(tgt, dt, 240 + k), blk(ref, dr, 250 + k)]))
    return out


def _simulators():
    sys.path.insert(0, str(ROOT / "tools"))
    import simulators
    return simulators


def case_sim_perfect(B=200_000):
    """simulator.go's own content (tools/simulators.py restates simulateStation, :100-177), with its
    usage example's arguments (:229: 41.20 -96.00 400 1000) and 200 000-sample blocks instead of
    20 000 000: literal tones + uniform noise, power ~2e-4 -> the weak branch on every signal, periodic
    peaks.  Regenerated from the seed (torch CPU generator), SHA-256 checked."""
    S = _simulators()
    caps, _ = S.simulate_perfect(list(S.STATIONS.values()), (41.20, -96.00, 400.0), 92300000.0, 1000.0, B, seed=100)
    return {name: c.numpy() for name, c in zip(STATIONS, caps)}


def case_sim_weak(B=200_000):
    """weak_signal_simulator.go's own content (tools/simulators.py restates simulateWeakSignalStation,
    :141-247) with its usage example's arguments (:298: ref_power 10, tgt_power 1000): the reference
    blocks carry 7e-5 ... 3e-4 of amplitude, under half a quantisation step, so every byte is 127, the
    signal after removeDCBias has (nearly) no power and the REF records are (0, 0.000000); the target
    blocks are a 1-3 count tone.  Regenerated from the seed, SHA-256 checked."""
    S = _simulators()
    caps, _ = S.simulate_weak(list(S.STATIONS.values()), (41.20, -96.00, 400.0), 92300000.0, 10.0, 1000.0, B, seed=300)
    return {name: c.numpy() for name, c in zip(STATIONS, caps)}


def _degenerate_third(third: np.ndarray):
    caps = case_fm_strong(B=20000)
    caps[STATIONS[2]] = third
    return caps


def case_tiny_third():
    """Third capture of two samples: fewer than three samples are "too small for dual-frequency
    extraction" and returned unchanged as REF and as TGT (processor.go:211-214, :244-246)."""
    return _degenerate_third(np.array([200, 90, 30, 160], np.uint8))


def case_three_sample_third():
    """Third capture of exactly three samples: one-sample blocks (REF two samples, TGT one)."""
    return _degenerate_third(np.array([200, 90, 30, 160, 140, 100], np.uint8))


def case_empty_third():
    """Third capture empty: crossCorrelate warns and returns (0, 0.0) (processor.go:622-625)."""
    return _degenerate_third(np.zeros(0, np.uint8))


def case_fm_reordered():
    """The fm_delays captures given to the command in another order (n3pay, kf0mtl, kx0u): pairs
    follow the order of the arguments, not of the station table (processor.go:816-817)."""
    caps = case_fm_delays()
    return {k: caps[k] for k in ("n3pay", "kf0mtl", "kx0u")}


def case_fm_two_valid(B=30000):
    """Exactly two valid range differences, close to the geometry of the start point: the one
    input class on which the shipped binary's solver (ELF 0x4a0360) runs to convergence
    (0.7-damped Newton steps, "Converged after 4 iterations") and prints a location.  Delays
    REF (0, 930, 6) / TGT (0, 900, 0): the third pair's true delay is negative on both signals,
    its peak is a spurious one far outside the 20.4 km limit."""
    ref, tgt = fm(B + 2900, 60, 75e3), fm(B + 2900, 61, 60e3)
    out = {}
    for k, (name, dr, dt) in enumerate(zip(STATIONS, (0, 930, 6), (0, 900, 0))):
        def blk(sig, d, seed):
            g = np.random.default_rng(seed)
            return sig[2600 - d:2600 - d + B] + 0.02 * (g.standard_normal(B) + 1j * g.standard_normal(B))
        out[name] = quantise(np.concatenate([blk(ref, dr, 700 + k), blk(tgt, dt, 710 + k), blk(ref, dr, 720 + k)]))
    return out


def case_four_stations(B=30000):
    """Four collectors (the table's KEVO as the fourth): six pairs per signal in i < j order; the
    binary's validation and solver stages see six range differences."""
    ref, tgt = fm(B + 600, 32, 75e3), fm(B + 600, 33, 60e3)
    out = {}
    for k, (name, dr, dt) in enumerate(zip(STATIONS + ["KEVO"], (0, 5, 11, 2), (0, 33, 14, 25))):
        def blk(sig, d, seed):
            g = np.random.default_rng(seed)
            return sig[400 - d:400 - d + B] + 0.02 * (g.standard_normal(B) + 1j * g.standard_normal(B))
        out[name] = quantise(np.concatenate([blk(ref, dr, 530 + k), blk(tgt, dt, 540 + k), blk(ref, dr, 550 + k)]))
    return out


def case_fm_close_lengths():
    """Lengths that differ by less than the 2000-lag search (blocks of 40000 / 40300 / 39800): the
    lag range shrinks to the length difference (processor.go:664-672) and true delays beyond it
    are out of reach."""
    return _uneven((40000, 40300, 39800), 28, (0, 8, 3), (0, 21, 5), 430)


def case_fm_uneven():
    """Captures of different lengths (blocks of 36000 / 41000 / 38500 samples, the second file one
    byte longer than a whole sample): the shorter signal of a pair is the template
    (processor.go:652-662), the lag range follows the length difference, and the odd trailing
    byte is dropped by the integer division in loadIQData (:188)."""
    return _uneven((36000, 41000, 38500), 26, (0, 8, 3), (0, 21, 5), 330, odd_byte=True)


def _uneven(Bs, seed, d_ref, d_tgt, noise_seed, odd_byte=False):
    ref, tgt = fm(max(Bs) + 600, seed, 75e3), fm(max(Bs) + 600, seed + 1, 60e3)
    out = {}
    for k, (name, B, dr, dt) in enumerate(zip(STATIONS, Bs, d_ref, d_tgt)):
        def blk(sig, d, s):
            g = np.random.default_rng(s)
            return sig[400 - d:400 - d + B] + 0.02 * (g.standard_normal(B) + 1j * g.standard_normal(B))
        raw = quantise(np.concatenate([blk(ref, dr, noise_seed + k), blk(tgt, dt, noise_seed + 10 + k), blk(ref, dr, noise_seed + 20 + k)]))
        if odd_byte and k == 1:
            raw = np.concatenate([raw, np.array([200], np.uint8)])
        out[name] = raw
    return out


# ---- station tables that drive the binary's solver into its rarely taken branches (captures:
# fm_two_valid, the one input class on which that solver runs)
def case_twin_stations():
    """n3pay and kf0mtl at the same coordinates: rows 1 and 2 of the Jacobian coincide, det = 0,
    "Singular matrix detected" and the single-equation step of equation 2, ten times."""
    return case_fm_two_valid()


def case_back_stations():
    """kf0mtl at kx0u's coordinates: the second row vanishes, det = 0, single-equation step of equation 1."""
    return case_fm_two_valid()


def case_close_stations():
    """Three collectors 2-3 km apart: "WARNING: Poor station geometry (small triangle area)"."""
    return case_fm_two_valid()


_ROWS = {"KEVO": "41.30888549464701,-96.02619229605524,356.0", "162400000": "41.25703803095629,-95.95512763589404,349.07",
         "kx0u": "41.18660274289527,-95.96064116595667,355.69", "n3pay": "41.24669616513154,-96.08366304481238,329.0",
         "kf0mtl": "41.32916620016985,-96.03513381562004,373.18"}


def _csv(**override):
    rows = dict(_ROWS, **override)
    return "Name,Latitude,Longitude,Elevation\n" + "".join(f"{k},{v}\n" for k, v in rows.items())


CASE_CSV = {   # case -> (file name under tests/golden/, contents)
    "twin_stations": ("stations_twin.csv", _csv(kf0mtl=_ROWS["n3pay"])),
    "back_stations": ("stations_back.csv", _csv(kf0mtl=_ROWS["kx0u"])),
    "close_stations": ("stations_close.csv", _csv(n3pay="41.20660274289527,-95.98064116595667,329.0",
                                                  kf0mtl="41.21660274289527,-95.95064116595667,373.18")),
}

# cases whose captures are regenerated from their seeds instead of being stored
REGENERATED = {"fm_truncated", "sim_perfect", "sim_weak"}

CASES = {
    "fm_strong": case_fm_strong,
    "fm_delays": case_fm_delays,
    "moderate": case_moderate,
    "weak_tones": case_weak_tones,
    "weak_noise": case_weak_noise,
    "fm_ragged": case_fm_ragged,
    "fm_truncated": case_fm_truncated,
    "fm_uneven": case_fm_uneven,
    "fm_close_lengths": case_fm_close_lengths,
    "fm_reordered": case_fm_reordered,
    "four_stations": case_four_stations,
    "fm_two_valid": case_fm_two_valid,
    "twin_stations": case_twin_stations,
    "back_stations": case_back_stations,
    "close_stations": case_close_stations,
    "tiny_third": case_tiny_third,
    "three_sample_third": case_three_sample_third,
    "empty_third": case_empty_third,
    "sim_perfect": case_sim_perfect,
    "sim_weak": case_sim_weak,
}

PAIR_RE = re.compile(r"^(REF|TGT) (\S+) - (\S+): delay=(-?\d+) samples \((-?[\d.]+) μs\), correlation=(-?[\d.]+)")


def parse_stdout(text: str):
    pairs, powers, dcs, norms, branches, reasonable = [], [], [], [], [], []
    for line in text.splitlines():
        line = line.split("\r")[-1]
        m = PAIR_RE.match(line)
        if m:
            pairs.append({"kind": m.group(1), "a": m.group(2), "b": m.group(3),
                          "delay": int(m.group(4)), "us": float(m.group(5)), "corr": float(m.group(6))})
        elif line.startswith("Initial signal power:"):
            powers.append(float(line.split(":")[1]))
        elif line.startswith("Removed DC bias:"):
            dcs.append(line.split(":", 1)[1].strip())
        elif line.startswith("Normalized signal power:"):
            norms.append(float(line.split(":")[1].split("→")[0]))
        elif line.startswith("Strong FM signal"):
            branches.append(0)
        elif line.startswith("Moderate signal"):
            branches.append(1)
        elif line.startswith("Weak signal"):
            branches.append(2)
        elif "reasonable" in line.lower() and "found" in line.lower():
            reasonable.append(line.strip())
    return {"pairs": pairs, "initial_power": powers, "dc_bias": dcs, "prenorm_power": norms,
            "branch": branches, "reasonable_lines": reasonable}


def main(only=None):
    import hashlib
    from oracle import oracle
    oracle.build()
    (HERE / "stations.csv").write_text(STATION_CSV)
    summary = {}
    for name, fn in CASES.items():
        if only and name not in only:
            continue
        caps = fn()
        with tempfile.TemporaryDirectory() as td:
            paths = []
            for st in caps:   # the order of the dictionary is the order of the arguments
                p = Path(td) / f"sim-{st}-1.dat"
                caps[st].tofile(p)
                paths.append(p)
            csv_name = "stations.csv"
            if name in CASE_CSV:
                csv_name, csv_text = CASE_CSV[name]
                (HERE / csv_name).write_text(csv_text)
            out, err, rc = oracle.run_reference_binary(paths, HERE / csv_name)
        parsed = parse_stdout(out)
        parsed["returncode"] = rc
        parsed["stderr_tail"] = err.strip().splitlines()[-1:] if err.strip() else []
        parsed["order"] = list(caps)
        parsed["csv"] = csv_name
        n_st = len(caps)
        assert len(parsed["pairs"]) == n_st * (n_st - 1), (name, out[-2000:])
        if name in CASE_CSV:
            parsed["captures"] = "fm_two_valid"   # the same captures: not stored twice
        elif name in REGENERATED:
            parsed["sha256"] = {st: hashlib.sha256(caps[st].tobytes()).hexdigest() for st in caps}
        else:
            np.savez_compressed(HERE / f"{name}.npz", **caps)
        (HERE / f"{name}.json").write_text(json.dumps(parsed, indent=1, ensure_ascii=False) + "\n")
        (HERE / f"{name}.stdout.txt").write_text(out)
        summary[name] = [(p["kind"], p["delay"], p["corr"]) for p in parsed["pairs"]]
        print(name, "branches", sorted(set(parsed["branch"])), summary[name])
    return summary


# ---- the command line: what the reference binary prints and returns before it touches a sample
CLI_FILES = {
    "bad_fields.csv": "Name,Latitude,Longitude,Elevation\n162400000,41.2,-95.9,349\nkx0u,41.1,-95.9\n",
    "bad_lat.csv": "Name,Latitude,Longitude,Elevation\n162400000,41.2,-95.9,349\nkx0u,abc,-95.9,300\n",
    "bad_lon.csv": "Name,Latitude,Longitude,Elevation\n162400000,41.2,-95.9,349\nkx0u,41.1,xyz,300\n",
    "bad_elev.csv": "Name,Latitude,Longitude,Elevation\n162400000,41.2,-95.9,349\nkx0u,41.1,-95.2,3o0\n",
    "header_only.csv": "Name,Latitude,Longitude,Elevation\n",
    "three_columns.csv": "Name,Latitude,Longitude\n162400000,41.2,-95.9\n",
    "bare_quote.csv": 'Name,Latitude,Longitude,Elevation\n\n162400000,41.2,-95.9,349\n"kx,0u",41.1,-95.9,3\nab"c,1,2,3\n',
    "quoted_ok.csv": 'Name,Latitude,Longitude,Elevation\r\n\r\n"162400000",41.2,-95.9,349\r\n"kx""0u",41.1,-95.9,3\r\n',
}
CLI_CASES = {   # {csv} = tests/golden/stations.csv, {dir} = a directory holding CLI_FILES
    "no_arguments": [],
    "three_arguments": ["162400000", "92300000", "{csv}"],
    "two_collectors": ["162400000", "92300000", "{csv}", "a-kx0u.dat", "b-n3pay.dat"],
    "one_collector": ["162400000", "92300000", "{csv}", "a-kx0u.dat"],
    "missing_csv": ["162400000", "92300000", "{dir}/nonexistent.csv", "a", "b", "c"],
    "reference_not_in_table": ["999", "92300000", "{csv}", "a", "b", "c"],
    "bad_reference_frequency": ["abc", "92300000", "{csv}", "a", "b", "c"],
    "bad_target_frequency": ["162400000", " 5", "{csv}", "a", "b", "c"],
    "frequency_out_of_range": ["1e400", "92300000", "{csv}", "a", "b", "c"],
    "csv_wrong_field_count": ["162400000", "92300000", "{dir}/bad_fields.csv", "a", "b", "c"],
    "csv_bad_latitude": ["162400000", "92300000", "{dir}/bad_lat.csv", "a", "b", "c"],
    "csv_bad_longitude": ["162400000", "92300000", "{dir}/bad_lon.csv", "a", "b", "c"],
    "csv_bad_elevation": ["162400000", "92300000", "{dir}/bad_elev.csv", "a", "b", "c"],
    "csv_header_only": ["162400000", "92300000", "{dir}/header_only.csv", "a", "b", "c"],
    "csv_three_columns": ["162400000", "92300000", "{dir}/three_columns.csv", "a", "b", "c"],
    "csv_bare_quote": ["162400000", "92300000", "{dir}/bare_quote.csv", "a", "b", "c"],
    "csv_quoted_fields_two_collectors": ["162400000", "92300000", "{dir}/quoted_ok.csv", "a", "b"],
}


# command lines that get as far as the captures (the mirrors need the engine for these: GPU tests).
# {dir} holds sim-kx0u-1.dat and sim-n3pay-1.dat (the fm_strong golden captures) and the directory
# sim-kf0mtl-dir.dat
CLI_ENGINE_CASES = {
    "unknown_station": ["162400000", "92300000", "{csv}", "{dir}/sim-kx0u-1.dat", "{dir}/sim-n3pay-1.dat", "{dir}/sim-nobody-1.dat"],
    "missing_third_file": ["162400000", "92300000", "{csv}", "{dir}/sim-kx0u-1.dat", "{dir}/sim-n3pay-1.dat", "{dir}/sim-kf0mtl-missing.dat"],
    "missing_first_file": ["162400000", "92300000", "{csv}", "{dir}/sim-kx0u-missing.dat", "{dir}/sim-n3pay-1.dat", "{dir}/sim-kf0mtl-1.dat"],
    "directory_as_capture": ["162400000", "92300000", "{csv}", "{dir}/sim-kx0u-1.dat", "{dir}/sim-n3pay-1.dat", "{dir}/sim-kf0mtl-dir.dat"],
}


def cli_engine_setup(td):
    """Files the CLI_ENGINE_CASES refer to (also called by the GPU test)."""
    caps = np.load(HERE / "fm_strong.npz")
    for st in STATIONS[:2]:
        caps[st].tofile(Path(td) / f"sim-{st}-1.dat")
    (Path(td) / "sim-kf0mtl-dir.dat").mkdir(exist_ok=True)


def cli_golden():
    """tests/golden/cli_errors.json: stdout, stderr (time stamp removed) and exit status of the
    reference binary for CLI_CASES; paths are kept as the {csv} / {dir} / {prog} place-holders."""
    from oracle import oracle
    oracle.build()
    exe = str(ROOT / "oracle" / "_ref" / "processor")
    out = {"files": CLI_FILES, "cases": {}}
    with tempfile.TemporaryDirectory() as td:
        for name, text in CLI_FILES.items():
            (Path(td) / name).write_bytes(text.encode())
        csv = str(HERE / "stations.csv")
        cli_engine_setup(td)
        out["engine_cases"] = sorted(CLI_ENGINE_CASES)
        for case, args in {**CLI_CASES, **CLI_ENGINE_CASES}.items():
            argv = [a.replace("{csv}", csv).replace("{dir}", td) for a in args]
            r = subprocess.run([exe, *argv], capture_output=True, text=True, cwd=td)
            def back(t):
                return t.replace(csv, "{csv}").replace(td, "{dir}").replace(exe, "{prog}")
            err = [back(l[20:]) for l in r.stderr.splitlines()]   # "2006/01/02 15:04:05 "
            assert all(re.match(r"\d{4}/\d\d/\d\d \d\d:\d\d:\d\d ", l) for l in r.stderr.splitlines()), r.stderr
            out["cases"][case] = {"args": args, "returncode": r.returncode, "stdout": [back(l) for l in r.stdout.splitlines()],
                                  "stderr": err}
            print(case, r.returncode, err[-1:] or out["cases"][case]["stdout"][-1:])
    (HERE / "cli_errors.json").write_text(json.dumps(out, indent=1, ensure_ascii=False) + "\n")


if __name__ == "__main__":
    if sys.argv[1:] == ["cli"]:
        cli_golden()
        sys.exit(0)
    main(only=set(sys.argv[1:]) or None)   # no arguments: every case
