"""Random capture sets for checking the CPU oracle against the reference's shipped binary
(oracle/_ref/processor) -- used by tests/test_oracle_fuzz.py on the CPU only.  A set mixes what the
golden captures hold one at a time: strong / moderate / weak amplitudes (all three preprocessing
branches, also within one run), equal or uneven block lengths, odd trailing bytes, random delays."""
from __future__ import annotations

import importlib.util
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_spec = importlib.util.spec_from_file_location("make_golden", HERE / "make_golden.py")
mg = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(mg)


def make(seed: int) -> dict:
    rng = np.random.default_rng(seed)
    kind = rng.integers(4)
    Bs = [int(rng.integers(12000, 30000)) for _ in range(3)] if rng.random() < 0.5 else [int(rng.integers(12000, 30000))] * 3
    n = max(Bs) + 700
    ref, tgt = mg.fm(n, int(rng.integers(1e6)), 75e3), mg.fm(n, int(rng.integers(1e6)), 60e3)
    amps = [0.5, 0.5, 0.5]
    if kind == 1:
        amps = [float(rng.choice([0.5, 0.07, 0.02])) for _ in range(3)]   # branches mixed within one run
    elif kind == 2:
        amps = [0.07] * 3
    elif kind == 3:
        amps = [0.02] * 3
    out = {}
    for k, name in enumerate(mg.STATIONS):
        B, dr, dt = Bs[k], int(rng.integers(0, 40)), int(rng.integers(0, 300))
        g = np.random.default_rng(int(rng.integers(1e6)))

        def blk(sig, d, a):
            s = sig[500 - d:500 - d + B] * (a / 0.5)
            if a < 0.1:
                s = s * (1 + 0.5 * np.sin(np.arange(B) * 0.01))
            return s + 0.1 * a * (g.standard_normal(B) + 1j * g.standard_normal(B))

        raw = mg.quantise(np.concatenate([blk(ref, dr, amps[k]), blk(tgt, dt, amps[k]), blk(ref, dr, amps[k])]))
        if rng.random() < 0.3:
            raw = np.concatenate([raw, rng.integers(0, 255, int(rng.integers(1, 6)), dtype=np.uint8)])
        out[name] = raw
    return out


def run_binary(caps: dict, exe: Path, csv: Path):
    """(parsed stdout, stdout text, return code, stderr line after the time stamp)"""
    with tempfile.TemporaryDirectory() as td:
        paths = []
        for st in caps:
            p = Path(td) / f"sim-{st}-1.dat"
            caps[st].tofile(p)
            paths.append(str(p))
        r = subprocess.run([str(exe), "162400000", "92300000", str(csv), *paths], capture_output=True, text=True)
    err = r.stderr.strip().splitlines()
    return mg.parse_stdout(r.stdout), r.stdout, r.returncode, (err[-1][20:] if err else "")
