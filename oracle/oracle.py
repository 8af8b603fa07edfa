"""ctypes binding of the CPU oracle (oracle/tdoa_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(tdoa-geolocation_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libtdoa_oracle.so"
REF_BINARY = _HERE / "_ref" / "processor"


def build(force: bool = False) -> Path:
    """Compile the oracle (and stage oracle/_ref when /root/reference exists)."""
    src = _HERE / "tdoa_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "libtdoa_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    if Path("/root/reference/processor").exists():
        subprocess.run(["make", "-C", str(_HERE), "ref"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_i64 = C.c_int64
_f64 = C.c_double
_vp = C.c_void_p


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        L = _lib
        L.orc_unpack_u8.argtypes = [_vp, _i64, _vp]
        L.orc_extract_reference.argtypes = [_vp, _i64, _vp]
        L.orc_extract_reference.restype = _i64
        L.orc_extract_target.argtypes = [_vp, _i64, _vp]
        L.orc_extract_target.restype = _i64
        L.orc_signal_power.argtypes = [_vp, _i64]
        L.orc_signal_power.restype = _f64
        L.orc_remove_dc.argtypes = [_vp, _i64, _vp, _vp]
        L.orc_set_seq_dc_limit.argtypes = [_i64]
        L.orc_lowpass.argtypes = [_vp, _i64, C.c_int, _vp]
        L.orc_set_wide_boxcar_f64.argtypes = [C.c_int]
        L.orc_cutoff_window.argtypes = [_f64, _f64]
        L.orc_cutoff_window.restype = C.c_int
        L.orc_bandpass.argtypes = [_vp, _i64, _f64, _f64, _f64, _vp]
        L.orc_notch.argtypes = [_vp, _i64, _f64, _f64, _f64, _vp]
        L.orc_normalize.argtypes = [_vp, _i64, _vp]
        L.orc_normalize.restype = _f64
        L.orc_enhance_weak.argtypes = [_vp, _i64, _vp]
        L.orc_preprocess_source.argtypes = [_vp, _i64, _vp]
        L.orc_preprocess_source.restype = C.c_int
        L.orc_discriminator.argtypes = [_vp, _i64, _vp]
        L.orc_envelope.argtypes = [_vp, _i64, _vp]
        L.orc_preprocess_binary.argtypes = [_vp, _i64, _vp]
        L.orc_preprocess_binary.restype = C.c_int
        L.orc_tdcorr_source.argtypes = [_vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp, _vp]
        L.orc_tdcorr_source.restype = _i64
        L.orc_tdcorr_binary.argtypes = [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp]
        L.orc_tdcorr_binary.restype = _i64
        L.orc_cross_correlate_source.argtypes = [_vp, _i64, _vp, _i64, _vp, _vp]
        L.orc_cross_correlate_binary.argtypes = [_vp, _i64, _vp, _i64, _vp, _vp, _vp]
        L.orc_tdcorr_binary_lags_mt.argtypes = [_vp, _i64, _vp, _i64, _i64, _i64, _vp]
        L.orc_xcorr_two_sided.argtypes = [_vp, _vp, _i64, _i64, _vp]
        L.orc_peak_parabolic.argtypes = [_vp, _i64, _vp, _vp, _vp]
        L.orc_llh_to_ecef.argtypes = [_f64, _f64, _f64, _vp]
        L.orc_ecef_to_llh.argtypes = [_f64, _f64, _f64, _vp]
        L.orc_baseline.argtypes = [_vp, _vp]
        L.orc_baseline.restype = _f64
        L.orc_solve_tdoa.argtypes = [_vp, _vp, _vp, _vp]
        L.orc_solve_tdoa.restype = C.c_int
        L.orc_solve_binary.argtypes = [_vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp]
        L.orc_solve_binary.restype = C.c_int
        L.orc_grid_solve.argtypes = [_vp, C.c_int, _vp, _f64, _f64, _f64, _f64, C.c_int, C.c_int,
                                     _f64, _vp, _vp, _vp]
        L.orc_solve_ls.argtypes = [_vp, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp]
        L.orc_solve_ls.restype = C.c_int
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(_vp)


def _c64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.complex64)


# ------------------------------------------------------------------ load
def unpack_u8(raw: np.ndarray) -> np.ndarray:
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    n = raw.size // 2
    out = np.empty(n, np.complex64)
    lib().orc_unpack_u8(_p(raw), n, _p(out))
    return out


def extract_reference(data: np.ndarray) -> np.ndarray:
    data = _c64(data)
    out = np.empty(max(data.size, 1), np.complex64)
    n = lib().orc_extract_reference(_p(data), data.size, _p(out))
    return out[:n].copy()


def extract_target(data: np.ndarray) -> np.ndarray:
    data = _c64(data)
    out = np.empty(max(data.size, 1), np.complex64)
    n = lib().orc_extract_target(_p(data), data.size, _p(out))
    return out[:n].copy()


# ------------------------------------------------------------ primitives
def signal_power(s) -> float:
    s = _c64(s)
    return lib().orc_signal_power(_p(s), s.size)


def remove_dc(s):
    s = _c64(s)
    out = np.empty_like(s)
    dc = np.zeros(1, np.complex64)
    lib().orc_remove_dc(_p(s), s.size, _p(out), _p(dc))
    return out, dc[0]


def set_seq_dc_limit(n: int) -> None:
    """Signals longer than n samples get an exactly rounded DC sum (engine-defined
    extension, see tdoa_oracle.c); n < 0 restores the reference's arithmetic."""
    lib().orc_set_seq_dc_limit(int(n))


def set_wide_boxcar_f64(min_window: int) -> None:
    """EXTENDED mode's arithmetic (engine-defined): box-cars of >= min_window taps accumulate in
    f64 and round once; 0 restores the reference's f32 tap walk."""
    lib().orc_set_wide_boxcar_f64(int(min_window))


def lowpass(s, window: int):
    s = _c64(s)
    out = np.empty_like(s)
    lib().orc_lowpass(_p(s), s.size, int(window), _p(out))
    return out


def cutoff_window(cutoff: float, fs: float) -> int:
    return lib().orc_cutoff_window(cutoff, fs)


def bandpass(s, lo, hi, fs):
    s = _c64(s)
    out = np.empty_like(s)
    lib().orc_bandpass(_p(s), s.size, lo, hi, fs, _p(out))
    return out


def notch(s, f0, bw, fs):
    s = _c64(s)
    out = np.empty_like(s)
    lib().orc_notch(_p(s), s.size, f0, bw, fs, _p(out))
    return out


def normalize(s):
    s = _c64(s)
    out = np.empty_like(s)
    p = lib().orc_normalize(_p(s), s.size, _p(out))
    return out, p


def preprocess_source(s):
    s = _c64(s)
    out = np.empty_like(s)
    br = lib().orc_preprocess_source(_p(s), s.size, _p(out))
    return out, br


def discriminator(s):
    s = _c64(s)
    out = np.empty_like(s)
    lib().orc_discriminator(_p(s), s.size, _p(out))
    return out


def envelope(s):
    s = _c64(s)
    out = np.empty_like(s)
    lib().orc_envelope(_p(s), s.size, _p(out))
    return out


def preprocess_binary(s):
    s = _c64(s)
    out = np.empty_like(s)
    br = lib().orc_preprocess_binary(_p(s), s.size, _p(out))
    return out, br


# ----------------------------------------------------------- correlators
def preprocess_binary_dec(s, D: int):
    """Engine-defined: the binary's chain, decimating box-car (mean of D), normalise."""
    s = _c64(s)
    out = np.empty(max(s.size // max(D, 1), 1) if D > 1 else s.size, np.complex64)
    L = lib()
    L.orc_preprocess_binary_dec.argtypes = [_vp, _i64, C.c_int, _vp]
    L.orc_preprocess_binary_dec.restype = C.c_int
    br = L.orc_preprocess_binary_dec(_p(s), s.size, D, _p(out))
    return out[:s.size // D] if D > 1 else out, br


def tdcorr_source(s1, s2, max_lag=20000, block=1000, want_all=False):
    s1, s2 = _c64(s1), _c64(s2)
    d = np.zeros(1, np.int64)
    c = np.zeros(1, np.float64)
    allv = np.zeros(max(int(max_lag), 1), np.float64) if want_all else None
    n = lib().orc_tdcorr_source(_p(s1), s1.size, _p(s2), s2.size, max_lag, block, _p(d), _p(c),
                                _p(allv) if want_all else None)
    if want_all:
        return int(d[0]), float(c[0]), allv[:n]
    return int(d[0]), float(c[0])


def tdcorr_binary(s1, s2, max_lag=2000, block=10000, sanity=120, want_all=False):
    s1, s2 = _c64(s1), _c64(s2)
    d = np.zeros(1, np.int64)
    c = np.zeros(1, np.float64)
    r = np.zeros(1, np.int32)
    allv = np.zeros(max(int(max_lag), 1), np.float64) if want_all else None
    n = lib().orc_tdcorr_binary(_p(s1), s1.size, _p(s2), s2.size, max_lag, block, sanity,
                                _p(d), _p(c), _p(r), _p(allv) if want_all else None)
    if want_all:
        return int(d[0]), float(c[0]), bool(r[0]), allv[:n]
    return int(d[0]), float(c[0]), bool(r[0])


def cross_correlate_source(s1, s2):
    s1, s2 = _c64(s1), _c64(s2)
    d = np.zeros(1, np.int64)
    c = np.zeros(1, np.float64)
    lib().orc_cross_correlate_source(_p(s1), s1.size, _p(s2), s2.size, _p(d), _p(c))
    return int(d[0]), float(c[0])


def cross_correlate_binary(s1, s2):
    s1, s2 = _c64(s1), _c64(s2)
    d = np.zeros(1, np.int64)
    c = np.zeros(1, np.float64)
    r = np.zeros(1, np.int32)
    lib().orc_cross_correlate_binary(_p(s1), s1.size, _p(s2), s2.size, _p(d), _p(c), _p(r))
    return int(d[0]), float(c[0]), bool(r[0])


def tdcorr_binary_lags_mt(tpl, sig, tl_eff, n_lags, block=10000):
    tpl, sig = _c64(tpl), _c64(sig)
    out = np.zeros(n_lags, np.float64)
    lib().orc_tdcorr_binary_lags_mt(_p(tpl), tl_eff, _p(sig), sig.size, n_lags, block, _p(out))
    return out


def xcorr_two_sided(y1, y2, max_lag: int):
    y1, y2 = _c64(y1), _c64(y2)
    assert y1.size == y2.size
    out = np.zeros(2 * max_lag + 1, np.float64)
    lib().orc_xcorr_two_sided(_p(y1), _p(y2), y1.size, max_lag, _p(out))
    return out


def peak_parabolic(c):
    c = np.ascontiguousarray(c, np.float64)
    i = np.zeros(1, np.int64)
    f = np.zeros(1, np.float64)
    v = np.zeros(1, np.float64)
    lib().orc_peak_parabolic(_p(c), c.size, _p(i), _p(f), _p(v))
    return int(i[0]), float(f[0]), float(v[0])


# --------------------------------------------------------------- geodesy
def llh_to_ecef(lat, lon, elev):
    out = np.zeros(3, np.float64)
    lib().orc_llh_to_ecef(lat, lon, elev, _p(out))
    return out


def ecef_to_llh(x, y, z):
    out = np.zeros(3, np.float64)
    lib().orc_ecef_to_llh(x, y, z, _p(out))
    return out


def baseline(llh1, llh2) -> float:
    a = np.ascontiguousarray(llh1, np.float64)
    b = np.ascontiguousarray(llh2, np.float64)
    return lib().orc_baseline(_p(a), _p(b))


def solve_tdoa(stations_llh, range_diffs):
    st = np.ascontiguousarray(stations_llh, np.float64)
    rd = np.ascontiguousarray(range_diffs, np.float64)
    out = np.zeros(3, np.float64)
    it = np.zeros(1, np.int32)
    status = lib().orc_solve_tdoa(_p(st), _p(rd), _p(out), _p(it))
    return out, status, int(it[0])


def solve_binary(stations_llh, range_diffs):
    """solveTDOA of the shipped binary (orc_solve_binary): (llh, status, n_valid, n_iter, converged, trace)."""
    st = np.ascontiguousarray(stations_llh, np.float64)
    rd = np.ascontiguousarray(range_diffs, np.float64)
    out = np.zeros(3, np.float64)
    nv, ni, cv = (np.zeros(1, np.int32) for _ in range(3))
    trace = np.zeros((10, 5), np.float64)
    status = lib().orc_solve_binary(_p(st), len(st), _p(rd), len(rd), _p(out), _p(nv), _p(ni), _p(cv), _p(trace))
    return out, status, int(nv[0]), int(ni[0]), bool(cv[0]), trace[:int(ni[0])]


def solve_ls(stations_llh, range_diffs, init_llh=None, dims=2):
    """Engine-defined least-squares fix (orc_solve_ls; no reference equivalent)."""
    st = np.ascontiguousarray(stations_llh, np.float64)
    rd = np.ascontiguousarray(range_diffs, np.float64)
    init = None if init_llh is None else np.ascontiguousarray(init_llh, np.float64)
    out = np.zeros(3, np.float64)
    rms = np.zeros(1, np.float64)
    it = np.zeros(1, np.int32)
    status = lib().orc_solve_ls(_p(st), st.shape[0], _p(rd), _p(init) if init is not None else None, dims,
                                _p(out), _p(rms), _p(it))
    return out, float(rms[0]), status, int(it[0])


def grid_solve(stations_llh, range_diffs, lat0, lon0, dlat, dlon, nlat, nlon, elev):
    st = np.ascontiguousarray(stations_llh, np.float64)
    rd = np.ascontiguousarray(range_diffs, np.float64)
    idx = np.zeros(1, np.int64)
    cost = np.zeros(1, np.float64)
    llh = np.zeros(3, np.float64)
    lib().orc_grid_solve(_p(st), st.shape[0], _p(rd), lat0, lon0, dlat, dlon, nlat, nlon, elev,
                         _p(idx), _p(cost), _p(llh))
    return int(idx[0]), float(cost[0]), llh


# ------------------------------------------------ whole-capture drivers
def process_capture_binary(raws, chunk=1_000_000):
    """BINARY ProcessTDOA pair loops (ELF ProcessTDOA; pair order processor.go:816-817).

    raws: list of uint8 arrays (one .dat per station).  Returns (ref, tgt) lists of
    (delay, corr, researched) in i<j lexicographic order.
    """
    refs, tgts = [], []
    for raw in raws:
        data = unpack_u8(raw)
        refs.append(extract_reference(data)[:chunk])
        tgts.append(extract_target(data)[:chunk])
    out_ref, out_tgt = [], []
    n = len(raws)
    for i in range(n):
        for j in range(i + 1, n):
            out_ref.append(cross_correlate_binary(refs[i], refs[j]))
    for i in range(n):
        for j in range(i + 1, n):
            out_tgt.append(cross_correlate_binary(tgts[i], tgts[j]))
    return out_ref, out_tgt


def process_capture_source(raws, chunk=2_000_000):
    """SOURCE ProcessTDOA pair loops (processor.go:756-850)."""
    refs, tgts = [], []
    for raw in raws:
        data = unpack_u8(raw)
        refs.append(extract_reference(data)[:chunk])
        tgts.append(extract_target(data)[:chunk])
    out_ref, out_tgt = [], []
    n = len(raws)
    for i in range(n):
        for j in range(i + 1, n):
            out_ref.append(cross_correlate_source(refs[i], refs[j]))
    for i in range(n):
        for j in range(i + 1, n):
            out_tgt.append(cross_correlate_source(tgts[i], tgts[j]))
    return out_ref, out_tgt


def source_stdout(dat_files, names, station_rows, raws, ref_station_row, target_freq, n_table_rows, chunk=2_000_000):
    """What processor.go AS COMMITTED prints for a run (ProcessTDOA :739-929 and everything it
    calls), restated print statement by print statement with the numbers of this oracle.  The
    source cannot be run here (no Go toolchain): "parity unpinned" -- but every line it shares with
    the shipped binary (most of them) has the format the binary's golden stdout pins.
    dat_files: the command's file arguments; names / station_rows: station name and (lat, lon, elev)
    of each, in argument order; returns (text, error message or None)."""
    o = []
    P = o.append
    P("Loaded %d stations including reference %.0f MHz" % (n_table_rows, float(ref_station_row[0]) / 1e6))   # :105
    P("Processing TDOA for target frequency %.3f MHz" % (target_freq / 1e6))                                # :744
    P("Reference: %s at %.6f°, %.6f°, %.1fm" % tuple(ref_station_row))                                      # :745
    refs, tgts = [], []
    for f, name, row, raw in zip(dat_files, names, station_rows, raws):
        n = raw.size // 2
        P("Loading I/Q data from: %s" % f)                                   # :167
        P("File size: %d bytes, samples: %d" % (raw.size, n))                # :184
        P("Successfully loaded %d complex samples" % n)                      # :203
        data = unpack_u8(raw)
        b = n // 3
        for what, kind in (("reference", "reference samples from blocks 1 and 3"), ("target", "target samples from block 2")):
            P("Extracting %s signal from dual-frequency data" % what)       # :209, :242
            if n < 3:
                P("Warning: Data too small for dual-frequency extraction")  # :217, :251
            else:
                P("Total samples: %d, block size: %d" % (n, b))             # :221, :255
                P("Extracted %d %s" % (2 * b if what == "reference" else b, kind))   # :236, :265
        r, t = extract_reference(data), extract_target(data)
        if r.size > chunk:
            r = r[:chunk]
            P("Using test chunk: %d samples (%.1f ms)" % (chunk, chunk / 2e6 * 1000))          # :775
        if t.size > chunk:
            t = t[:chunk]
            P("Using target test chunk: %d samples (%.1f ms)" % (chunk, chunk / 2e6 * 1000))   # :779
        P("Coherent integration time: %.0f ms (expecting ~%.1f dB processing gain)"
          % (chunk / 2e6 * 1000, 10 * np.log10(chunk / 100000)))                              # :782
        refs.append(r)
        tgts.append(t)
        P("Loaded collector: %s at %.6f°, %.6f°, %.1fm" % (name, row[0], row[1], row[2]))      # :797
    S = len(raws)
    pairs = [(i, j) for i in range(S) for j in range(i + 1, S)]
    P("\nBaseline distances (3D):")                                          # :802
    for i, j in pairs:
        P("%s - %s: %.2f km" % (names[i], names[j], baseline(station_rows[i], station_rows[j]) / 1000))   # :806

    bp = "Bandpass filter: %.1f - %.1f Hz (at %.0f Hz sample rate)"          # :359

    def preprocess(sig, label):                                              # :469-499
        P("Preprocessing %s signal (%d samples)" % (label, sig.size))
        p0 = signal_power(sig)
        P("Initial signal power: %.9f" % p0)
        fs = 2000000.0
        if p0 < 0.001:
            P("Detected very weak signal - applying aggressive filtering")
            P("Enhancing weak signal: %s" % label)                          # :438
            x, dc = remove_dc(sig)
            P("Removed DC bias: %.6f + %.6fi" % (dc.real, dc.imag))         # :317
            for f0, bw in ((60, 5), (120, 5), (1000000, 50000)):            # :446-448
                P(bp % (max(f0 - bw / 2, 0), min(f0 + bw / 2, fs / 2), fs))
                x = notch(x, f0, bw, fs)
            P(bp % (100.0, 40000.0, fs))
            x = bandpass(x, 100.0, 40000.0, fs)
            x = lowpass(x, 50)
        else:
            P("Standard signal processing")
            x, dc = remove_dc(sig)
            P("Removed DC bias: %.6f + %.6fi" % (dc.real, dc.imag))
            P(bp % (500.0, 50000.0, fs))
            x = bandpass(x, 500.0, 50000.0, fs)
            x = lowpass(x, 100)
        y, p1 = normalize(x)
        if p1 > 0:                                                           # :338-340
            P("Normalized signal power: %.6f → 1.000000" % p1)
        return y

    def cross(s1, s2):                                                       # :619-643
        P("=== Cross-Correlation Analysis ===")
        if s1.size == 0 or s2.size == 0:
            P("Warning: Empty signals for correlation")
            return 0, 0.0
        P("\n--- Signal Preprocessing ---")
        y1, y2 = preprocess(s1, "Signal 1"), preprocess(s2, "Signal 2")
        P("\n--- Time Domain Correlation ---")
        P("Performing time domain correlation")                             # :647
        tl, sl = min(y1.size, y2.size), max(y1.size, y2.size)
        P("Template: %d samples, Signal: %d samples" % (tl, sl))            # :660
        ml = max(1, min(20000, sl - tl))                                     # :668-675
        P("Using coherent integration with %d-sample blocks" % 1000)        # :684
        nb = 0 if tl <= 1000 else (tl - 1000 + 999) // 1000                  # blocks of `for bs = 0; bs < tl - 1000; bs += 1000`
        prog = "".join("Time domain progress: %d/%d (coherent blocks: %d)\r" % (d, ml, nb) for d in range(0, ml, 2000))   # :729-731
        d, c = tdcorr_source(y1, y2)
        P(prog + "\nTime domain correlation: %.6f at delay %d samples" % (c, d))   # :734
        P("\n--- Result: Time Domain with Preprocessing ---")               # :639
        P("Correlation: %.6f at delay %d samples" % (c, d))                 # :640
        return d, c

    tds = {}
    for label, sigs, head, sub in (("REF", refs, "REFERENCE", "Testing weak 162.4 MHz NOAA weather signal:"),
                                   ("TGT", tgts, "TARGET", "Testing strong 92.3 MHz FM broadcast signal:")):
        P("\n=== %s SIGNAL CORRELATION TEST ===" % head)                     # :812, :832
        P(sub)                                                               # :813, :833
        tds[label] = []
        for i, j in pairs:
            d, c = cross(sigs[i], sigs[j])
            td = float(d) / 2e6                                              # :821
            tds[label].append(td)
            P("%s %s - %s: delay=%d samples (%.3f μs), correlation=%.6f" % (label, names[i], names[j], d, td * 1e6, c))
    td = tds["TGT"]                                                          # :853
    P("\n=== CORRELATION COMPARISON ===")
    P("Reference signal (162.4 MHz): Generally weaker correlation")
    P("Target signal (92.3 MHz): Should show stronger correlation")
    P("Using target signal for TDOA calculation")
    P("\nTDOA triangulation:")
    P("Time differences: %.3f μs, %.3f μs, %.3f μs" % (td[0] * 1e6, td[1] * 1e6, td[2] * 1e6))   # :869
    c0 = 299792458.0
    P("Distance differences: %.1f m, %.1f m, %.1f m" % (td[0] * c0, td[1] * c0, td[2] * c0))     # :878
    P("\nDiagnostic test with example delays:")
    P("Simulating 10 μs, 5 μs, -3 μs delays...")
    for k, dl in enumerate((10e-6, 5e-6, -3e-6)):
        P("Test delay %d: %.1f μs → %.1f m" % (k + 1, dl * 1e6, dl * c0))   # :888
    P("\n=== TDOA GEOLOCATION ===")
    rd = [t * c0 for t in td]
    P("Time differences (μs): " + "".join("%.3f " % (t * 1e6) for t in td))
    P("Range differences (m): " + "".join("%.1f " % r for r in rd))
    m = np.asarray(station_rows[:3], np.float64)
    P("Initial guess: %.6f°, %.6f°, %.1fm" % ((m[0, 0] + m[1, 0] + m[2, 0]) / 3.0, (m[0, 1] + m[1, 1] + m[2, 1]) / 3.0,
                                            (m[0, 2] + m[1, 2] + m[2, 2]) / 3.0))     # :957
    pos, status, iters = solve_tdoa(np.asarray(station_rows, np.float64), rd)
    if status != 0:
        return "\n".join(o) + "\n", "TDOA processing failed: TDOA solution failed: singular Jacobian matrix at iteration %d" % iters
    P("Converged after %d iterations" % iters if iters < 10 else "Maximum iterations reached")   # :971, :1013
    P("\n*** CALCULATED TRANSMITTER LOCATION ***")
    P("Latitude:  %.6f°" % pos[0])
    P("Longitude: %.6f°" % pos[1])
    P("Elevation: %.1f m" % pos[2])
    return "\n".join(o) + "\n", None


def run_reference_binary(dat_paths, csv_path, ref_hz="162400000", tgt_hz="92300000", timeout=600):
    """Run the staged reference ELF (oracle/_ref/processor); returns its stdout."""
    if not REF_BINARY.exists():
        raise FileNotFoundError(str(REF_BINARY))
    res = subprocess.run([str(REF_BINARY), ref_hz, tgt_hz, str(csv_path), *map(str, dat_paths)],
                         capture_output=True, text=True, timeout=timeout)
    return res.stdout, res.stderr, res.returncode


# ---- fast_analyzer.go / analyzer.go
class Quality(C.Structure):
    _fields_ = [("total_samples", C.c_int64), ("i_avg", C.c_double), ("q_avg", C.c_double), ("i_std", C.c_double),
                ("q_std", C.c_double), ("i_min", C.c_int32), ("i_max", C.c_int32), ("q_min", C.c_int32),
                ("q_max", C.c_int32), ("snr_db", C.c_double), ("power_db", C.c_double), ("dc_offset", C.c_double),
                ("iq_imbalance", C.c_double), ("has_clipping", C.c_int32), ("has_overload", C.c_int32),
                ("has_dead_zones", C.c_int32), ("has_noise", C.c_int32)]


def analyze_samples(sig_bytes: np.ndarray, fast: bool) -> dict:
    """fastAnalyzeSamples (fast_analyzer.go:113) / analyzeSamples (analyzer.go:130) of one
    signal's interleaved uint8 bytes."""
    b = np.ascontiguousarray(sig_bytes, np.uint8)
    q = Quality()
    L = lib()
    L.orc_analyze_samples.argtypes = [_vp, _i64, C.c_int, _vp]
    L.orc_analyze_samples(b.ctypes.data, b.size // 2, 1 if fast else 0, C.cast(C.byref(q), _vp))
    return {k: getattr(q, k) for k, _ in Quality._fields_}


def analyze_capture(raw: np.ndarray, fast: bool):
    """(ref, tgt) as fastAnalyzeDualFrequencyFile (fast_analyzer.go:54) / analyzeDualFrequencyFile (analyzer.go:85)."""
    raw = np.ascontiguousarray(raw, np.uint8)
    total = raw.size // 2
    block = total // 3
    if fast:
        a = min(32768, block)
        ref = np.concatenate([raw[0:2 * a], raw[4 * block:4 * block + 2 * a]])
        tgt = raw[2 * block:2 * block + 2 * a]
    else:
        ref = np.concatenate([raw[0:2 * block], raw[4 * block:2 * total]])[:4 * block]
        tgt = raw[2 * block:4 * block]
    return analyze_samples(ref, fast), analyze_samples(tgt, fast)
