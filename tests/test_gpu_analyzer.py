"""The per-capture analyzers (fast_analyzer.go, analyzer.go) on the GPU against their CPU
restatement (oracle/tdoa_oracle.c orc_analyze_samples, the reference's own O(M^2) DFT).
Integer statistics and flags exact; f64 derived values to 1e-9; the SNR (whose spectrum
is an FFT here and a direct DFT sum in the reference) to 1e-6 dB -- the reference prints
one decimal."""
import io

import numpy as np
import pytest

import tdoa_b200 as T
from oracle import oracle
from helpers import fm_capture

pytestmark = pytest.mark.gpu

EXACT = ("total_samples", "i_min", "i_max", "q_min", "q_max", "has_clipping", "has_overload", "has_dead_zones", "has_noise")
CLOSE = ("i_avg", "q_avg", "i_std", "q_std", "power_db", "dc_offset", "iq_imbalance")


def captures():
    rng = np.random.default_rng(7)
    fm = fm_capture(40000, (0, 7, 3), (0, 7, 3), seed=3)[0]                       # clean FM, 120 000 samples
    clipped = fm.copy(); clipped[1000:1100] = 255; clipped[5000] = 0              # clipping
    dead = rng.integers(100, 156, 2 * 3 * 30000, dtype=np.uint8)
    dead[70000:71500] = 0                                                         # 1500 zero bytes inside block 2
    dead[2 * 30000 - 600:2 * 30000] = 0; dead[4 * 30000:4 * 30000 + 600] = 0      # a run across the block-1/3 joint of REF
    quiet = np.full(2 * 3 * 20000, 127, np.uint8); quiet[::7] = 128                 # overload (tiny variation)
    noisy = rng.integers(0, 256, 2 * 3 * 25000, dtype=np.uint8)                   # excessive noise + clipping
    small = rng.integers(90, 170, 2 * 3 * 1000, dtype=np.uint8)                   # spectrum size 1000 / 2000: direct DFT
    tone = np.empty(2 * 3 * 50000, np.uint8)
    ph = 2 * np.pi * 0.05 * np.arange(3 * 50000)
    tone[0::2] = np.clip(60 * np.cos(ph) + 127.5 + rng.normal(0, 3, ph.size), 0, 255).astype(np.uint8)
    tone[1::2] = np.clip(60 * np.sin(ph) + 127.5 + rng.normal(0, 3, ph.size), 0, 255).astype(np.uint8)
    return {"fm": fm, "clipped": clipped, "dead": dead, "quiet": quiet, "noisy": noisy, "small": small, "tone": tone}


@pytest.fixture(scope="module")
def eng():
    with T.Engine(T.MODE_BINARY) as e:
        yield e


@pytest.mark.parametrize("name", list(captures()))
@pytest.mark.parametrize("fast", [True, False])
def test_quality_matches_the_restatement(eng, name, fast):
    raw = captures()[name]
    eng.load_u8(0, raw)
    got = eng.analyze(0, fast=fast)
    want = oracle.analyze_capture(raw, fast=fast)
    for g, w in zip(got, want):
        for k in EXACT:
            assert g[k] == w[k], (name, fast, k, g[k], w[k])
        for k in CLOSE:
            if np.isfinite(w[k]) or np.isfinite(g[k]):
                assert g[k] == pytest.approx(w[k], rel=1e-9, abs=1e-9), (name, fast, k)
        assert g["snr_db"] == pytest.approx(w["snr_db"], abs=1e-6), (name, fast, g["snr_db"], w["snr_db"])


def test_dead_zone_across_the_reference_joint(eng):
    """analyzer.go concatenates blocks 1 and 3 before scanning, so 600 + 600 zero bytes either
    side of the joint form one 1200-byte dead zone in REF; block 2 holds its own 1500."""
    raw = captures()["dead"]
    eng.load_u8(0, raw)
    ref, tgt = eng.analyze(0, fast=False)
    assert ref["has_dead_zones"] == 1 and tgt["has_dead_zones"] == 1
    ok = captures()["fm"]
    eng.load_u8(0, ok)
    ref, tgt = eng.analyze(0, fast=False)
    assert ref["has_dead_zones"] == 0 and tgt["has_dead_zones"] == 0


def test_fast_analyzer_cli_lines(eng, tmp_path):
    raw = captures()["tone"]
    f = tmp_path / "kx0u-1.dat"
    raw.tofile(f)
    buf = io.StringIO()
    assert T.analyzer.fast_main([str(f)], out=buf) == 0
    want = oracle.analyze_capture(raw, fast=True)
    lines = buf.getvalue().splitlines()
    for line, label, w in zip(lines, ("REF", "TGT"), want):
        assert line == "%s,%.1f,%.1f,%s,%s" % (label, w["snr_db"], w["power_db"], str(bool(w["has_clipping"])).lower(),
                                               str(bool(w["has_overload"])).lower())
    buf = io.StringIO()
    assert T.analyzer.main([str(f), "5"], out=buf) == 0
    text = buf.getvalue()
    assert "=== REFERENCE SIGNAL ANALYSIS ===" in text and "=== TDOA SUITABILITY ASSESSMENT ===" in text
    tiny = tmp_path / "tiny.dat"
    np.zeros(4, np.uint8).tofile(tiny)
    buf = io.StringIO()
    assert T.analyzer.fast_main([str(tiny)], out=buf) == 1 and "file too small" in buf.getvalue()


def test_cpp_fast_analyzer_command(tmp_path):
    """fast_analyzer_b200 (host/fast_analyzer_b200.cpp): the two CSV lines gain_calibrator.go:266-297
    parses, equal to the restatement's; the reference's error line and exit status for a tiny file."""
    import subprocess
    from pathlib import Path
    exe = Path(__file__).resolve().parent.parent / "tdoa-geolocation_b200" / "fast_analyzer_b200"
    for name in ("tone", "clipped"):
        raw = captures()[name]
        f = tmp_path / f"kx0u-{name}.dat"
        raw.tofile(f)
        r = subprocess.run([str(exe), str(f)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        want = oracle.analyze_capture(raw, fast=True)
        for line, label, w in zip(r.stdout.splitlines(), ("REF", "TGT"), want):
            assert line == "%s,%.1f,%.1f,%s,%s" % (label, w["snr_db"], w["power_db"], str(bool(w["has_clipping"])).lower(),
                                                   str(bool(w["has_overload"])).lower())
    tiny = tmp_path / "tiny.dat"
    np.zeros(4, np.uint8).tofile(tiny)
    r = subprocess.run([str(exe), str(tiny)], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("Error: ") and "file too small" in r.stdout
    r = subprocess.run([str(exe), str(tmp_path / "absent.dat")], capture_output=True, text=True)
    assert r.returncode == 1 and "failed to open file" in r.stdout


def test_cpp_analyzer_command_prints_what_the_mirror_prints(tmp_path):
    """analyzer_b200 (host/analyzer_b200.cpp, analyzer.go's report over the C ABI) against the Python
    mirror of the same report on the same file: byte for byte."""
    import subprocess
    from pathlib import Path
    exe = Path(__file__).resolve().parent.parent / "tdoa-geolocation_b200" / "analyzer_b200"
    for name in ("fm", "clipped", "dead", "quiet", "noisy"):
        f = tmp_path / f"kx0u-{name}.dat"
        captures()[name].tofile(f)
        r = subprocess.run([str(exe), str(f), "5"], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        buf = io.StringIO()
        assert T.analyzer.main([str(f), "5"], out=buf) == 0
        assert r.stdout == buf.getvalue(), name
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("Usage: analyzer <data_file.dat>")
