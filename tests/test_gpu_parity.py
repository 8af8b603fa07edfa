"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden
vectors of the reference's shipped binary.  Tolerances (BASELINE.json north_star):
integer lag of every correlation peak bit-exact; sub-sample TDOA within 1e-3 samples;
position within 1 m.  Float intermediates: the f32 DC bias is the reference's sequential
f32 accumulator bit for bit (seqsum.cu) up to seq_dc_limit samples -- everything the
reference can reach -- and an exactly rounded sum beyond; the f64 power sums use a fixed
tree instead of the reference's running sum (<= 1e-15 relative), so a preprocessed sample
may differ in its last bit with probability ~1e-8 and a printed correlation by <= 1e-6."""
import io
import sys
import re

import numpy as np
import pytest

import tdoa_b200 as T
from oracle import oracle
from helpers import GOLDEN, GOLDEN_CASES, GOLDEN_DEGENERATE_CASES, GOLDEN_LONG_CASES, GOLDEN_SIM_CASES, GOLDEN_ORDER_CASES, GOLDEN_TABLE_CASES, STATION_LLH, fm_capture, load_golden, quantise

pytestmark = pytest.mark.gpu

CORR_TOL = 1e-6
SOURCE_TOL = 1e-12  # SOURCE mode, relative: the engine walks the taps, the blocks and the lags in the reference's order, so
                    # its numbers are the oracle's to the last bit (measured: 0 differing samples, identical correlations)
SAMPLE_TOL = 2e-6


@pytest.fixture(scope="module")
def eng_binary():
    with T.Engine(T.MODE_BINARY) as e:
        yield e


@pytest.fixture(scope="module")
def eng_source():
    with T.Engine(T.MODE_SOURCE) as e:
        yield e


def split(raw):
    d = oracle.unpack_u8(raw)
    return oracle.extract_reference(d), oracle.extract_target(d)


def load_all(e, raws):
    for k, r in enumerate(raws):
        e.load_u8(k, r)


# ------------------------------------------------------------------ K1 unpack
def test_unpack_bit_exact(eng_binary):
    raw = np.arange(512, dtype=np.uint8)  # every code in I and Q
    raw = np.concatenate([raw, np.random.default_rng(0).integers(0, 256, 100000, dtype=np.uint8)])
    eng_binary.load_u8(0, raw)
    got = eng_binary.unpack(0, 0, raw.size // 2)
    want = oracle.unpack_u8(raw)
    assert got.view(np.uint32).tolist() == want.view(np.uint32).tolist()
    part = eng_binary.unpack(0, 17, 1000)
    assert np.array_equal(part.view(np.uint32), want[17:1017].view(np.uint32))


# ------------------------------------------------------------------ preprocessing
@pytest.mark.parametrize("case,branch", [("fm_strong", 0), ("fm_delays", 0), ("moderate", 1),
                                         ("weak_noise", 2), ("weak_tones", 2)])
def test_preprocess_binary_branches(eng_binary, case, branch):
    raws, meta = load_golden(case)
    load_all(eng_binary, raws)
    for st in range(3):
        ref, tgt = split(raws[st])
        for kind, sig in ((T.KIND_REF, ref), (T.KIND_TGT, tgt)):
            got, power, br = eng_binary.preprocess(st, kind, 0, sig.size)
            want, wbr = oracle.preprocess_binary(sig)
            assert br == wbr == branch
            assert power == pytest.approx(oracle.signal_power(sig), rel=1e-12)
            assert np.max(np.abs(got - want)) <= SAMPLE_TOL


def test_preprocess_window_and_joint(eng_binary):
    """A window that straddles the block-1/block-3 joint of the reference signal."""
    raws, _ = load_golden("fm_strong")
    eng_binary.load_u8(0, raws[0])
    ref, _ = split(raws[0])
    b = ref.size // 2
    start, ln = b - 3000, 9000
    got, _, br = eng_binary.preprocess(0, T.KIND_REF, start, ln)
    want, _ = oracle.preprocess_binary(ref[start:start + ln])
    assert br == 0 and np.max(np.abs(got - want)) <= SAMPLE_TOL


@pytest.mark.parametrize("amp", [0.5, 0.01])
def test_preprocess_source_branches(eng_source, amp):
    rng = np.random.default_rng(5)
    n = 60000
    x = amp * np.exp(1j * np.cumsum(rng.standard_normal(n) * 0.1)) + 0.2 * amp * (
        rng.standard_normal(n) + 1j * rng.standard_normal(n))
    raw = quantise(np.concatenate([x, x[::-1], x]))
    eng_source.load_u8(0, raw)
    _, tgt = split(raw)
    got, power, br = eng_source.preprocess(0, T.KIND_TGT, 0, tgt.size)
    want, wbr = oracle.preprocess_source(tgt)
    assert (br != 0) == bool(wbr) == (amp < 0.03)
    assert np.max(np.abs(got - want)) <= 5e-6


# ------------------------------------------------------------------ golden pairs (shipped binary)
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_golden_pairs_binary_mode(eng_binary, case):
    raws, meta = load_golden(case)
    load_all(eng_binary, raws)
    want_ref, want_tgt = oracle.process_capture_binary(raws)
    got = list(eng_binary.xcorr(T.KIND_REF)[0]) + list(eng_binary.xcorr(T.KIND_TGT)[0])
    for pk, gold, orc in zip(got, meta["pairs"], want_ref + want_tgt):
        assert int(pk["lag"]) == gold["delay"] == orc[0], (case, gold)       # bit-exact lag
        assert abs(float(pk["corr"]) - gold["corr"]) <= 0.5e-6 + CORR_TOL   # reference prints %.6f
        assert abs(float(pk["corr"]) - orc[1]) <= CORR_TOL
        assert bool(pk["flags"] & 1) == orc[2]                               # sanity re-search taken


def test_cross_correlate_seam_binary(eng_binary):
    raws, _ = load_golden("fm_delays")
    t0, t1 = split(raws[0])[1], split(raws[2])[1]
    pk = eng_binary.cross_correlate(t0, t1)
    d, c, r = oracle.cross_correlate_binary(t0, t1)
    assert (pk.lag, pk.researched) == (d, r) and abs(pk.corr - c) <= CORR_TOL
    # unequal lengths: the shorter input is the template (processor.go:653-661)
    pk = eng_binary.cross_correlate(t0[:30000], t1)
    d, c, r = oracle.cross_correlate_binary(t0[:30000], t1)
    assert (pk.lag, pk.researched) == (d, r) and abs(pk.corr - c) <= CORR_TOL
    pk = eng_binary.cross_correlate(t1, t0[:30000])
    d, c, r = oracle.cross_correlate_binary(t1, t0[:30000])
    assert (pk.lag, pk.researched) == (d, r) and abs(pk.corr - c) <= CORR_TOL


def test_empty_and_tiny_inputs(eng_binary):
    z = np.zeros(0, np.complex64)
    x = (np.random.default_rng(1).standard_normal(5000) * 0.3).astype(np.complex64)
    pk = eng_binary.cross_correlate(z, x)
    assert (pk.lag, pk.corr) == (0, 0.0) and pk.flags & 0x8  # processor.go:622-625
    # shorter than one block: no whole block -> (0, 0.0) like the reference loop
    pk = eng_binary.cross_correlate(x, x)
    d, c, _ = oracle.cross_correlate_binary(x, x)
    assert (pk.lag, pk.corr) == (d, c) == (0, 0.0)


def test_ragged_station_lengths(eng_binary):
    raws = fm_capture(30000, (0, 4, 9), (0, 12, 30), seed=3)
    raws[1] = raws[1][: 2 * 3 * 27000]  # a shorter capture: block 27000
    load_all(eng_binary, raws)
    want_ref, want_tgt = oracle.process_capture_binary(raws)
    got = list(eng_binary.xcorr(T.KIND_REF)[0]) + list(eng_binary.xcorr(T.KIND_TGT)[0])
    for pk, orc in zip(got, want_ref + want_tgt):
        assert int(pk["lag"]) == orc[0] and abs(float(pk["corr"]) - orc[1]) <= CORR_TOL


def test_process_on_ragged_and_tiny_captures(eng_binary):
    """The one-call path on inputs the reference's loops treat specially: stations of different
    lengths (the shorter input is the template, processor.go:653-661), captures shorter than one
    block (no whole block -> (0, 0.0)), a 2-sample capture (N / 3 == 0: the data is returned
    unchanged, :211-214)."""
    raws = fm_capture(30000, (0, 4, 9), (0, 12, 30), seed=3)
    raws[1] = raws[1][: 2 * 3 * 27000]
    for variant in (raws, [r[: 2 * 3 * 2000] for r in raws], [raws[0], raws[1], raws[2][:4]]):
        load_all(eng_binary, variant)
        ref, tgt = eng_binary.xcorr(T.KIND_REF)[0], eng_binary.xcorr(T.KIND_TGT)[0]
        r = eng_binary.process(STATION_LLH)
        for name in ("lag", "corr", "flags", "n_blocks"):
            assert np.array_equal(r["ref"][name], ref[name]) and np.array_equal(r["tgt"][name], tgt[name]), name
        want_ref, want_tgt = oracle.process_capture_binary(variant)
        for pk, orc in zip(list(r["ref"]) + list(r["tgt"]), want_ref + want_tgt):
            assert int(pk["lag"]) == orc[0] and abs(float(pk["corr"]) - orc[1]) <= CORR_TOL


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_fft_path_equals_every_lag_path(case):
    """use_fft=1 (candidate search + exact re-evaluation) must pick the same peak as the
    exhaustive time-domain evaluation of every lag."""
    raws, _ = load_golden(case)
    res = []
    for use_fft in (0, 1):
        with T.Engine(T.MODE_BINARY, use_fft=use_fft) as e:
            load_all(e, raws)
            res.append(np.concatenate([e.xcorr(T.KIND_REF)[0], e.xcorr(T.KIND_TGT)[0]]))
    brute, fft = res
    assert np.all(brute["flags"] & 4) and not np.any(fft["flags"] & 4), "paths not exercised"
    assert np.array_equal(brute["lag"], fft["lag"])
    assert np.array_equal(brute["first_lag"], fft["first_lag"])
    assert np.array_equal(brute["flags"] & 1, fft["flags"] & 1)
    assert np.max(np.abs(brute["corr"] - fft["corr"])) <= 1e-12


@pytest.mark.parametrize("seed,noise,mode", [(31, 0.35, "BINARY"), (32, 0.6, "BINARY"), (33, 1.0, "BINARY"), (34, 1.5, "BINARY"),
                                             (35, 0.6, "EXTENDED"), (36, 1.2, "EXTENDED")])
def test_fft_ranking_survives_low_snr(seed, noise, mode):
    """Low-SNR stress of the FFT ranking (every path: BINARY one-sided lags, EXTENDED two-sided lags with the
    parabolic vertex, full windows and ragged ones): with noise up to 3 x the signal amplitude the correlation
    peak stands only a few per cent above its neighbours, so an f32 transform error beyond the ranking
    tolerance would drop the true maximum from the candidates.  The exhaustive every-lag path (use_fft = 0) is
    the reference's own loop; both must return the same records."""
    block = 90000 + 137 * seed
    raws = fm_capture(block, (0, 7, 19), (0, 23, 41), seed=seed, noise=noise)
    kw = dict(max_lag=500) if mode == "EXTENDED" else {}
    res = []
    for use_fft in (0, 1):
        with T.Engine(getattr(T, "MODE_" + mode), use_fft=use_fft, **kw) as e:
            load_all(e, raws)
            out = [e.xcorr(T.KIND_REF)[0], e.xcorr(T.KIND_TGT)[0]]
            if mode == "EXTENDED":
                out.append(e.xcorr(T.KIND_TGT, 1234, 20000, 3, 17001).reshape(-1))   # windows that straddle nothing round
            res.append(np.concatenate(out))
    brute, fft = res
    assert np.array_equal(brute["lag"], fft["lag"])
    assert np.array_equal(brute["first_lag"], fft["first_lag"])
    assert np.max(np.abs(brute["corr"] - fft["corr"])) <= 1e-12
    assert np.max(np.abs(brute["frac"] - fft["frac"])) <= 1e-9
    # the stress is real: the peak is small (noise dominates) in the noisiest cases
    if noise >= 1.0:
        assert np.max(np.abs(brute["corr"])) < 0.5


# ------------------------------------------------------------------ source mode (processor.go as committed)
def test_source_mode_pairs(eng_source):
    raws, _ = load_golden("fm_strong")
    load_all(eng_source, raws)
    want_ref, want_tgt = oracle.process_capture_source(raws)
    got = list(eng_source.xcorr(T.KIND_REF)[0]) + list(eng_source.xcorr(T.KIND_TGT)[0])
    for pk, orc in zip(got, want_ref + want_tgt):
        # equal lengths: the source evaluates lag 0 only (SURVEY.md finding 3)
        assert int(pk["lag"]) == orc[0] == 0
        assert abs(float(pk["corr"]) - orc[1]) <= SOURCE_TOL * max(1.0, abs(orc[1]))


def test_source_mode_unequal_lengths(eng_source):
    raws, _ = load_golden("fm_delays")
    t0, t1 = split(raws[0])[1], split(raws[1])[1]
    a, b = t0[:20000], t1[:26000]
    pk = eng_source.cross_correlate(a, b)
    d, c = oracle.cross_correlate_source(a, b)
    assert pk.lag == d
    assert abs(pk.corr - c) <= SOURCE_TOL * max(1.0, abs(c))


@pytest.mark.parametrize("case", ["fm_strong", "weak_noise", "fm_uneven"])
def test_source_command_prints_what_processor_go_prints(tmp_path, case):
    """`processor_b200 --source` and the Python mirror in MODE_SOURCE against processor.go's print
    statements restated in the oracle (oracle.source_stdout: ProcessTDOA :739-929, preprocessSignal
    :469-499, enhanceWeakSignal :437-466, timeDomainCorrelation :646-736, solveTDOA :957-1013), line
    for line -- the standard and the weak chain, equal and unequal lengths (progress lines, lags
    beyond 0).  The source cannot run here: the expected text is the oracle's, "parity unpinned"."""
    import subprocess
    raws, _ = load_golden(case)
    names = ["kx0u", "n3pay", "kf0mtl"]
    files = []
    for name, raw in zip(names, raws):
        f = tmp_path / f"sim-{name}-1.dat"
        raw.tofile(f)
        files.append(str(f))
    csv_file = str(GOLDEN / "stations.csv")
    want, err = oracle.source_stdout(files, names, [tuple(r) for r in STATION_LLH], raws,
                                     ("162400000", 41.25703803095629, -95.95512763589404, 349.07), 92300000.0, 5)
    assert err is None
    exe = GOLDEN.parent.parent / "tdoa-geolocation_b200" / "processor_b200"
    out = subprocess.run([str(exe), "--source", "162400000", "92300000", csv_file, *files], capture_output=True)   # bytes: the progress lines end in \r
    assert out.returncode == 0, out.stderr
    buf = io.StringIO()
    p = T.TDOAProcessor(162400000.0, 92300000.0, csv_file, mode=T.MODE_SOURCE, out=buf)
    p.process_tdoa(files)
    p.close()
    for who, text in (("processor_b200 --source", out.stdout.decode("utf-8")), ("python mirror", buf.getvalue())):
        ours, gold = text.split("\n"), want.split("\n")
        assert len(ours) == len(gold), (who, text[-1500:])
        for k, (a, b) in enumerate(zip(ours, gold)):
            assert _same_line(a, b, tol=SOURCE_TOL), f"{who}, line {k}: ours {a!r} != oracle {b!r}"


# ------------------------------------------------------------------ windows
def test_windows_equal_single_calls(eng_binary):
    raws = fm_capture(120000, (0, 3, 8), (0, 20, 41), seed=7)
    load_all(eng_binary, raws)
    W, hop, nw = 30000, 25000, 4
    multi = eng_binary.xcorr(T.KIND_TGT, 1000, W, nw, hop)
    for w in range(nw):
        one = eng_binary.xcorr(T.KIND_TGT, 1000 + w * hop, W, 1, 0)[0]
        assert np.array_equal(multi[w]["lag"], one["lag"])
        assert np.array_equal(multi[w]["corr"], one["corr"])
        for p, (i, j) in enumerate([(0, 1), (0, 2), (1, 2)]):
            ti = split(raws[i])[1][1000 + w * hop:1000 + w * hop + W]
            tj = split(raws[j])[1][1000 + w * hop:1000 + w * hop + W]
            d, c, _ = oracle.cross_correlate_binary(ti, tj)
            assert int(multi[w][p]["lag"]) == d and abs(float(multi[w][p]["corr"]) - c) <= CORR_TOL
    with pytest.raises(T.TdoaError):
        eng_binary.xcorr(T.KIND_TGT, 0, 100000, 3, 50000)  # runs past the block


# ------------------------------------------------------------------ extended mode (engine-defined)
def test_extended_two_sided_subsample():
    L, W = 300, 40000
    raws = fm_capture(60000, (40, 0, 17), (25, 60, 0), seed=11)  # negative true lags appear
    oracle.set_seq_dc_limit(0)  # EXTENDED mode: exactly rounded DC sum (engine-defined)
    with T.Engine(T.MODE_EXTENDED, max_lag=L) as e:
        load_all(e, raws)
        got = e.xcorr(T.KIND_TGT, 500, W, 1, 0)[0]
        for p, (i, j) in enumerate([(0, 1), (0, 2), (1, 2)]):
            yi, _ = oracle.preprocess_binary(split(raws[i])[1][500:500 + W])
            yj, _ = oracle.preprocess_binary(split(raws[j])[1][500:500 + W])
            c = oracle.xcorr_two_sided(yi, yj, L)
            idx, frac, val = oracle.peak_parabolic(c)
            assert int(got[p]["lag"]) == idx - L
            assert abs(float(got[p]["frac"]) - frac) <= 1e-3          # north_star: 1e-3 samples
            assert abs(float(got[p]["corr"]) - val) <= CORR_TOL
    oracle.set_seq_dc_limit(-1)
    assert [int(g["lag"]) for g in got] == [35, -25, -60]


def test_extended_wide_lags_use_the_big_transform():
    """>= 8192 lags: the ranking runs through the 2^21-point four-step transform
    (xcorr_big.cu) instead of 2048-lag chunks; candidates are still evaluated exactly, so
    the records must be the oracle's -- and identical to the chunked path's (use_fft=3)."""
    L, W = 5000, 60000
    raws = fm_capture(70000, (400, 0, 1700), (2500, 3100, 0), seed=12)
    oracle.set_seq_dc_limit(0)
    with T.Engine(T.MODE_EXTENDED, max_lag=L) as e, T.Engine(T.MODE_EXTENDED, max_lag=L, use_fft=3) as e3:
        load_all(e, raws)
        load_all(e3, raws)
        got = e.xcorr(T.KIND_TGT, 500, W, 1, 0)[0]
        old = e3.xcorr(T.KIND_TGT, 500, W, 1, 0)[0]
        ref = e.xcorr(T.KIND_REF, 0, 2 * W, 1, 0)[0]   # spans the block-1 / block-3 joint
        ref3 = e3.xcorr(T.KIND_REF, 0, 2 * W, 1, 0)[0]
        for name in ("lag", "corr", "frac", "flags"):
            assert np.array_equal(got[name], old[name]), name
            assert np.array_equal(ref[name], ref3[name]), name
        for p, (i, j) in enumerate([(0, 1), (0, 2), (1, 2)]):
            yi, _ = oracle.preprocess_binary(split(raws[i])[1][500:500 + W])
            yj, _ = oracle.preprocess_binary(split(raws[j])[1][500:500 + W])
            c = oracle.xcorr_two_sided(yi, yj, L)
            idx, frac, val = oracle.peak_parabolic(c)
            assert int(got[p]["lag"]) == idx - L
            assert abs(float(got[p]["frac"]) - frac) <= 1e-3
            assert abs(float(got[p]["corr"]) - val) <= CORR_TOL
    oracle.set_seq_dc_limit(-1)
    assert [int(g["lag"]) for g in got] == [-2500 + 3100, -2500, -3100]
    assert [int(g["lag"]) for g in ref] == [-400, 1300, 1700]


def test_extended_weak_branch_wide_boxcar():
    """EXTENDED mode on the reference simulators' content (every signal on the weak branch): the
    1001-tap high-pass takes its window sums from an f64 prefix sum (k_boxcar_slide) instead of
    1001 f32 additions per sample.  Against the oracle's statement of that arithmetic
    (orc_set_wide_boxcar_f64): samples equal to the f32 bit on all but a handful, lags identical,
    correlation <= 1e-6; against the reference's own f32 tap walk (what BINARY mode reproduces): the
    difference is that chain's rounding, < 1e-4 of the signal's RMS."""
    sys.path.insert(0, str(GOLDEN.parent.parent / "tools"))
    import simulators as S
    L, W = 300, 60000
    st = list(S.STATIONS.values())
    caps, _ = S.simulate_perfect(st, (41.20, -96.00, 400.0), 92300000.0, 1000.0, 70000, seed=500)
    raws = [c.numpy() for c in caps]
    oracle.set_seq_dc_limit(0)
    try:
        with T.Engine(T.MODE_EXTENDED, max_lag=L) as e:
            load_all(e, raws)
            got = e.xcorr(T.KIND_TGT, 500, W, 1, 0)[0]
            ys_ref, ys = [], []
            for k in range(3):
                x = split(raws[k])[1][500:500 + W]
                out, p0, br = e.preprocess(k, T.KIND_TGT, 500, W)
                assert br == 2
                oracle.set_wide_boxcar_f64(0)
                y_ref, wbr = oracle.preprocess_binary(x)          # the reference's f32 tap walk
                oracle.set_wide_boxcar_f64(33)
                y, _ = oracle.preprocess_binary(x)                # EXTENDED's statement
                assert wbr == 2
                assert np.count_nonzero(out != y) <= W // 1000, np.count_nonzero(out != y)
                assert np.max(np.abs(out - y)) <= 2e-6
                rms = float(np.sqrt(np.mean(np.abs(y_ref) ** 2)))
                assert np.max(np.abs(out - y_ref)) <= 1e-4 * rms
                ys.append(y)
            for p, (i, j) in enumerate([(0, 1), (0, 2), (1, 2)]):
                c = oracle.xcorr_two_sided(ys[i], ys[j], L)
                idx, frac, val = oracle.peak_parabolic(c)
                assert int(got[p]["lag"]) == idx - L
                assert abs(float(got[p]["frac"]) - frac) <= 1e-3
                assert abs(float(got[p]["corr"]) - val) <= CORR_TOL
    finally:
        oracle.set_seq_dc_limit(-1)
        oracle.set_wide_boxcar_f64(0)


def _weak_capture(block, seed):
    """A capture whose every block is on the weak branch (power <= 0.001): small tones + noise around a DC offset."""
    g = np.random.default_rng(seed)
    n = 3 * block
    t = np.arange(n)
    x = 0.012 * np.exp(2j * np.pi * 0.013 * t) + 0.006 * (g.standard_normal(n) + 1j * g.standard_normal(n)) + (0.004 - 0.003j)
    return quantise(x)


@pytest.mark.parametrize("kind,start,W", [(0, 0, 70001), (0, 46000, 9000), (1, 0, 50000), (1, 137, 6144), (1, 49000, 1000),
                                           (0, 49990, 20), (1, 5, 3), (1, 0, 12289)])
def test_extended_weak_fused_pass_equals_the_four_kernel_chain(kind, start, W):
    """EXTENDED mode, weak branch: k_raw_stats + k_weak_fused (two passes over the capture bytes, preprocess_weak.cu)
    against the four kernels over complex planes they replace (use_fft = 6: k_power, k_unpack, k_boxcar_slide,
    k_boxcar) and against the oracle's statement of the chain.  Views of both kinds: across the joint of the REF
    signal's two blocks (50 000), ragged lengths, windows shorter than the filter, a tile boundary (6144)."""
    block = 50000
    raws = [_weak_capture(block, 900 + k) for k in range(3)]
    oracle.set_seq_dc_limit(0)
    oracle.set_wide_boxcar_f64(33)
    try:
        with T.Engine(T.MODE_EXTENDED, max_lag=8) as e, T.Engine(T.MODE_EXTENDED, max_lag=8, use_fft=6) as e4:
            load_all(e, raws)
            load_all(e4, raws)
            for k in range(3):
                e.preprocess(k, kind, start, W)    # the first call of a (station, kind) guesses "strong FM"; the
                e4.preprocess(k, kind, start, W)   # branch is remembered from then on
                l0, l4 = e.stats()["launches_total"], e4.stats()["launches_total"]
                a, pa, ba = e.preprocess(k, kind, start, W)
                b, pb, bb = e4.preprocess(k, kind, start, W)
                assert ba == 2 and bb == 2
                assert e.stats()["launches_total"] - l0 == e4.stats()["launches_total"] - l4 - 2   # two kernels instead of four
                assert abs(pa - pb) <= 1e-15 * abs(pb)
                # the window sums come from prefix sums of different tiles: a sample may differ in its last f32 bit
                assert np.count_nonzero(a != b) <= max(1, W // 1000), np.count_nonzero(a != b)
                assert np.max(np.abs(a - b)) <= 2e-6
                sig = split(raws[k])[0 if kind == 0 else 1][start:start + W]
                y, br = oracle.preprocess_binary(sig)
                assert br == 2
                assert np.count_nonzero(a != y) <= max(1, W // 1000), np.count_nonzero(a != y)
                assert np.max(np.abs(a - y)) <= 2e-6
            if W > 64:
                ra = e.xcorr(kind, start, W, 1, 0)[0]
                rb = e4.xcorr(kind, start, W, 1, 0)[0]
                assert [int(x) for x in ra["lag"]] == [int(x) for x in rb["lag"]]
                assert np.max(np.abs(ra["corr"] - rb["corr"])) <= CORR_TOL
                assert sorted({int(x["branch"]) for x in e.xcorr_info(kind)[0]}) == [2]
    finally:
        oracle.set_seq_dc_limit(-1)
        oracle.set_wide_boxcar_f64(0)


@pytest.mark.parametrize("D", [4, 8])
def test_extended_decimating_boxcar(D):
    """decimate = D (EXTENDED only, engine-defined): the binary's chain, then the mean of D
    consecutive samples, then normalise; correlation at fs / D over max_lag / D lags; records in
    samples of the capture.  Oracle: orc_preprocess_binary_dec + orc_xcorr_two_sided."""
    L, W = 320, 48000
    true_t = (25, 60, 0)
    raws = fm_capture(60000, (40, 0, 17), true_t, seed=13, dev_tgt=20e3)
    oracle.set_seq_dc_limit(0)
    with T.Engine(T.MODE_EXTENDED, max_lag=L, decimate=D) as e:
        load_all(e, raws)
        for st in range(3):
            got, p0, br = e.preprocess(st, T.KIND_TGT, 500, W)
            want, wbr = oracle.preprocess_binary_dec(split(raws[st])[1][500:500 + W], D)
            assert br == wbr == 0 and got.size == W // D
            assert np.max(np.abs(got - want)) <= 4e-6
        pk = e.xcorr(T.KIND_TGT, 500, W, 1, 0)[0]
        for p, (i, j) in enumerate([(0, 1), (0, 2), (1, 2)]):
            yi, _ = oracle.preprocess_binary_dec(split(raws[i])[1][500:500 + W], D)
            yj, _ = oracle.preprocess_binary_dec(split(raws[j])[1][500:500 + W], D)
            c = oracle.xcorr_two_sided(yi, yj, L // D)
            idx, frac, val = oracle.peak_parabolic(c)
            want_total = D * (idx - L // D + frac)
            got_total = int(pk[p]["lag"]) + float(pk[p]["frac"])
            assert abs(got_total - want_total) <= 1e-3 * D       # 1e-3 samples at the decimated rate
            assert abs(float(pk[p]["corr"]) - val) <= CORR_TOL
            assert abs(float(pk[p]["frac"])) <= 0.5 + 1e-6
            assert abs(got_total - (true_t[j] - true_t[i])) <= 0.75   # the injected delay, to sub-sample accuracy
    oracle.set_seq_dc_limit(-1)
    with pytest.raises(T.TdoaError):
        T.Engine(T.MODE_BINARY, decimate=4)


@pytest.mark.parametrize("L", [300, 1500, 3000])
def test_many_stations_spectra_path(L):
    """>= 10 pairs of one window: every station-segment is transformed once and the pairs are
    formed from the parked spectra (xcorr_spec.cu); 2049..4096 lags run as one chunk of
    4096-sample segments.  Same records as the 2 x 2 tile path (use_fft=4) and as the oracle."""
    W = 40000 + 2 * L
    d = (0, 35, 11, 140, 77, 260)
    raws = fm_capture(W + 3000, d, d, seed=31)
    oracle.set_seq_dc_limit(0)
    with T.Engine(T.MODE_EXTENDED, n_stations=6, max_lag=L) as e, \
            T.Engine(T.MODE_EXTENDED, n_stations=6, max_lag=L, use_fft=4) as e4:
        load_all(e, raws)
        load_all(e4, raws)
        got = e.xcorr(T.KIND_TGT, 700, W, 2, 1500)
        old = e4.xcorr(T.KIND_TGT, 700, W, 2, 1500)
        for name in ("lag", "corr", "frac", "flags"):
            assert np.array_equal(got[name], old[name]), name
        pairs = [(i, j) for i in range(6) for j in range(i + 1, 6)]
        assert [int(x) for x in got[0]["lag"]] == [d[j] - d[i] if abs(d[j] - d[i]) <= L else int(x)
                                                   for (i, j), x in zip(pairs, got[0]["lag"])]
        for p in (0, 7, 14):
            i, j = pairs[p]
            yi, _ = oracle.preprocess_binary(split(raws[i])[1][700:700 + W])
            yj, _ = oracle.preprocess_binary(split(raws[j])[1][700:700 + W])
            idx, frac, val = oracle.peak_parabolic(oracle.xcorr_two_sided(yi, yj, L))
            assert int(got[0][p]["lag"]) == idx - L
            assert abs(float(got[0][p]["frac"]) - frac) <= 1e-3 and abs(float(got[0][p]["corr"]) - val) <= CORR_TOL
    oracle.set_seq_dc_limit(-1)


# ------------------------------------------------------------------ geodesy + solvers
def test_baselines_and_solver(eng_binary):
    base = eng_binary.baselines(STATION_LLH)
    want = [oracle.baseline(STATION_LLH[i], STATION_LLH[j]) for i in range(3) for j in range(i + 1, 3)]
    assert np.max(np.abs(base - want)) < 1e-6
    rng = np.random.default_rng(2)
    rds = np.concatenate([np.array([[0, 0, 0], [3.5, 6.0, 2.5], [10, 5, -3], [-3.5, 2, 0]]) * 1e-6 * 299792458.0,
                          rng.uniform(-8000, 8000, (60, 3))])
    out, status, iters = eng_binary.solve(STATION_LLH, rds)
    for k in range(len(rds)):
        w, ws, wi = oracle.solve_tdoa(STATION_LLH, rds[k])
        assert (status[k] != 0) == (ws != 0) and iters[k] == wi
        if ws == 0:
            ga, wa = oracle.llh_to_ecef(*out[k]), oracle.llh_to_ecef(*w)
            assert np.linalg.norm(ga - wa) < 1e-3  # north_star: position within 1 m


def test_grid_multilateration(eng_binary):
    rng = np.random.default_rng(4)
    st = np.vstack([STATION_LLH, STATION_LLH[:2] + rng.uniform(-0.1, 0.1, (2, 3)) * [1, 1, 100]])
    tx = np.array([41.25, -96.0, 350.0])
    r = [np.linalg.norm(oracle.llh_to_ecef(*tx) - oracle.llh_to_ecef(*s)) for s in st]
    rd = np.array([r[j] - r[i] for i in range(5) for j in range(i + 1, 5)])
    rds = np.stack([rd, rd + rng.normal(0, 15, rd.size), rd + rng.normal(0, 40, rd.size)])
    desc = [41.20, -96.06, 0.002, 0.002, 60, 70, 350.0]
    out, cost, idx = eng_binary.grid(st, desc, rds)
    for k in range(3):
        wi, wc, wl = oracle.grid_solve(st, rds[k], *desc[:4], int(desc[4]), int(desc[5]), desc[6])
        assert idx[k] == wi
        assert cost[k] == pytest.approx(wc, rel=1e-9)
        assert np.allclose(out[k], wl, atol=1e-12)


@pytest.mark.parametrize("n_st,n_sets,nlat,nlon,dlat", [(3, 4, 90, 110, 0.002), (5, 7, 257, 129, 0.001), (16, 9, 300, 300, 0.0015),
                                                        (16, 3, 64, 500, 0.0), (9, 700, 40, 37, 0.003), (2, 2, 50, 50, 0.002)])
def test_grid_ranked_equals_exhaustive(n_st, n_sets, nlat, nlon, dlat):
    """tdoa_grid ranks the cells by the expanded cost (16 FMA per cell and set, solve.cu) and lets the statement
    settle the survivors; use_fft = 0 evaluates every cell by the statement.  Same index and the same cost, bit
    for bit: noisy sets, noiseless sets (cost ~ 0 at the minimum), a grid of identical rows (dlat = 0: every
    cell ties with nlat - 1 others and the lowest index has to win), more sets than one launch takes (512)."""
    rng = np.random.default_rng(100 * n_st + n_sets)
    st = np.array([[41.26 + rng.uniform(-0.2, 0.2), -96.02 + rng.uniform(-0.25, 0.25), rng.uniform(300, 400)] for _ in range(n_st)])
    desc = [41.26 - 0.5 * nlat * dlat, -96.02 - 0.5 * nlon * 0.002, dlat, 0.002, nlat, nlon, 350.0]
    rds = []
    for k in range(n_sets):
        tx = np.array([41.26 + rng.uniform(-0.05, 0.05), -96.02 + rng.uniform(-0.08, 0.08), 350.0])
        r = [np.linalg.norm(oracle.llh_to_ecef(*tx) - oracle.llh_to_ecef(*s)) for s in st]
        rd = np.array([r[j] - r[i] for i in range(n_st) for j in range(i + 1, n_st)])
        rds.append(rd + rng.normal(0, [0.0, 15.0, 150.0][k % 3], rd.size))
    rds = np.stack(rds)
    with T.Engine(T.MODE_BINARY) as e, T.Engine(T.MODE_BINARY, use_fft=0) as ex:
        l0, l1 = e.stats()["launches_total"], ex.stats()["launches_total"]
        out, cost, idx = e.grid(st, desc, rds)
        out0, cost0, idx0 = ex.grid(st, desc, rds)
        launches = e.stats()["launches_total"] - l0, ex.stats()["launches_total"] - l1
    assert launches == (5 * ((n_sets + 511) // 512), 2)   # ranked (five kernels per 512 sets), exhaustive (two)
    assert idx.tolist() == idx0.tolist()
    assert cost.view(np.uint64).tolist() == cost0.view(np.uint64).tolist()
    assert np.array_equal(out, out0)
    if dlat == 0.0:
        assert all(int(i) < nlon for i in idx)   # first row wins the ties
    k = 1
    wi, wc, wl = oracle.grid_solve(st, rds[k], *desc[:4], int(desc[4]), int(desc[5]), desc[6])
    assert idx[k] == wi and cost[k] == pytest.approx(wc, rel=1e-9)


def test_grid_ranked_falls_back_on_sets_that_are_not_numbers():
    """A set of range differences holding a NaN leaves the ranked path without a survivor (every comparison with
    its bound fails); tdoa_grid then lets the exhaustive kernel decide, so both engines answer alike."""
    st = np.vstack([STATION_LLH, STATION_LLH[:1] + [0.05, 0.07, 10.0]])
    tx = np.array([41.25, -96.0, 350.0])
    r = [np.linalg.norm(oracle.llh_to_ecef(*tx) - oracle.llh_to_ecef(*s)) for s in st]
    rd = np.array([r[j] - r[i] for i in range(4) for j in range(i + 1, 4)])
    rds = np.stack([rd, rd.copy(), rd + 5.0])
    rds[1, 2] = np.nan
    desc = [41.20, -96.06, 0.002, 0.002, 60, 70, 350.0]
    with T.Engine(T.MODE_BINARY) as e, T.Engine(T.MODE_BINARY, use_fft=0) as ex:
        out, cost, idx = e.grid(st, desc, rds)
        out0, cost0, idx0 = ex.grid(st, desc, rds)
    assert idx.tolist() == idx0.tolist()
    assert cost.view(np.uint64)[[0, 2]].tolist() == cost0.view(np.uint64)[[0, 2]].tolist()
    assert np.isnan(cost[1]) and np.isnan(cost0[1])


# ------------------------------------------------------------------ the reference-interface mirror
def test_processor_mirror_stdout(tmp_path):
    raws, meta = load_golden("fm_strong")
    files = []
    for name, raw in zip(["kx0u", "n3pay", "kf0mtl"], raws):
        f = tmp_path / f"sim-{name}-1.dat"
        raw.tofile(f)
        files.append(str(f))
    # the binary's revision: its own solver refuses three valid measurements (ELF 0x4a0360)
    buf = io.StringIO()
    p = T.TDOAProcessor(162400000.0, 92300000.0, str(GOLDEN / "stations.csv"), out=buf)
    with pytest.raises(RuntimeError, match="TDOA solution failed: no valid range difference measurements remain"):
        p.process_tdoa(files)
    p.close()
    text = buf.getvalue()
    gold = (GOLDEN / "fm_strong.stdout.txt").read_text()
    for line in gold.splitlines():
        if line.startswith(("REF ", "TGT ")) or line.endswith(" km"):
            assert line in text, line
    assert "*** CALCULATED TRANSMITTER LOCATION ***" not in text
    # processor.go as committed: solveTDOA on the target differences (processor.go:853, :932-1020)
    buf = io.StringIO()
    p = T.TDOAProcessor(162400000.0, 92300000.0, str(GOLDEN / "stations.csv"), mode=T.MODE_SOURCE, out=buf)
    res = p.process_tdoa(files)
    p.close()
    assert "*** CALCULATED TRANSMITTER LOCATION ***" in buf.getvalue()
    rd = np.array(res["range_differences"])
    want, status, _ = oracle.solve_tdoa(STATION_LLH, rd)
    assert status == 0 and np.allclose(res["position"], want, atol=1e-9)


def _same_line(a: str, b: str, tol: float = 1.5e-6) -> bool:
    """Text equality, or equality of the words with numbers compared to `tol` (a printed
    %.6f may flip its last digit on a 1e-16 difference in the value)."""
    import re
    if a == b:
        return True
    num = re.compile(r"-?\d+\.\d+|-?\d+")
    if num.sub("#", a) != num.sub("#", b):
        return False
    return all(abs(float(x) - float(y)) <= tol * max(1.0, abs(float(y))) for x, y in zip(num.findall(a), num.findall(b)))


def _same_outcome(failure, meta):
    """The binary's own solver (ELF 0x4a0360) gives a fix only with exactly two valid range
    differences; otherwise the command ends with its error text and status 1."""
    want = [l[20:] for l in meta["stderr_tail"]]   # after the log time stamp
    if meta["returncode"] == 0:
        assert failure is None
    else:
        assert failure is not None and want == ["TDOA processing failed: " + failure]


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_processor_stdout_is_the_shipped_binarys(tmp_path, case):
    """the host mirror prints, line for line, what the reference's
    shipped `processor` printed for the same captures (tests/golden/<case>.stdout.txt),
    up to the point where that build's solver aborts; only the temp-file paths differ."""
    raws, meta = load_golden(case)
    files = []
    for name, raw in zip(["kx0u", "n3pay", "kf0mtl"], raws):
        f = tmp_path / f"sim-{name}-1.dat"
        raw.tofile(f)
        files.append(str(f))
    buf = io.StringIO()
    p = T.TDOAProcessor(162400000.0, 92300000.0, str(GOLDEN / "stations.csv"), out=buf)
    failure = None
    try:
        p.process_tdoa(files)
    except RuntimeError as exc:
        failure = str(exc)
    p.close()
    _same_outcome(failure, meta)
    skip = "Loading I/Q data from:"
    ours = [l for l in buf.getvalue().splitlines() if not l.startswith(skip)]
    gold = [l for l in (GOLDEN / f"{case}.stdout.txt").read_text().splitlines() if not l.startswith(skip)]
    assert len(ours) == len(gold)
    for k, (a, b) in enumerate(zip(ours, gold)):
        assert _same_line(a, b), f"line {k}: ours {a!r} != reference {b!r}"


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_cpp_host_mirror_stdout_is_the_shipped_binarys(tmp_path, case):
    """processor_b200 (host/processor_b200.cpp: the reference's command over the C ABI) prints, line
    for line, what the reference's shipped `processor` printed for the same captures."""
    import subprocess
    raws, meta = load_golden(case)
    files = []
    for name, raw in zip(["kx0u", "n3pay", "kf0mtl"], raws):
        f = tmp_path / f"sim-{name}-1.dat"
        raw.tofile(f)
        files.append(str(f))
    exe = GOLDEN.parent.parent / "tdoa-geolocation_b200" / "processor_b200"
    r = subprocess.run([str(exe), "162400000", "92300000", str(GOLDEN / "stations.csv"), *files], capture_output=True, text=True)
    assert r.returncode == meta["returncode"], r.stderr
    assert [l[20:] for l in r.stderr.splitlines()] == [l[20:] for l in meta["stderr_tail"]]
    skip = "Loading I/Q data from:"
    ours = [l for l in r.stdout.splitlines() if not l.startswith(skip)]
    gold = [l for l in (GOLDEN / f"{case}.stdout.txt").read_text().splitlines() if not l.startswith(skip)]
    assert len(ours) == len(gold)
    for k, (a, b) in enumerate(zip(ours, gold)):
        assert _same_line(a, b), f"line {k}: ours {a!r} != reference {b!r}"


@pytest.mark.parametrize("case", GOLDEN_LONG_CASES + GOLDEN_DEGENERATE_CASES + GOLDEN_ORDER_CASES + GOLDEN_TABLE_CASES + GOLDEN_SIM_CASES)
def test_blocks_longer_than_the_test_chunk(tmp_path, case):
    """Blocks of 1 050 000 samples: the shipped binary cuts REF and TGT to their first 1 000 000
    samples before the pair loops (processor.go:772-780).  Degenerate cases: a third capture of 2
    samples (returned unchanged as REF and TGT), of 3 samples (one-sample blocks), empty
    (crossCorrelate warns and returns (0, 0.0)).  Captures given in another order, and four
    collectors: pairs i < j in the order of the arguments.  The reference's own simulators' content
    (sim_perfect, sim_weak: weak branch on every signal).  Station tables with coincident or
    close stations: the solver's single-equation fall-back and poor-geometry warning.  Records against what the binary
    printed (bit-exact lags, correlation to the printed 6 decimals), and the C++ command's
    stdout against the binary's, line for line."""
    import subprocess
    raws, meta = load_golden(case)
    names = meta.get("order", ["kx0u", "n3pay", "kf0mtl"])
    csv_file = str(GOLDEN / meta.get("csv", "stations.csv"))
    table = {row.split(",")[0]: [float(v) for v in row.split(",")[1:]]
             for row in open(csv_file).read().splitlines()[1:]}
    llh = np.array([table[n] for n in names])
    with T.Engine(T.MODE_BINARY, n_stations=len(raws)) as e:
        load_all(e, raws)
        r = e.process(llh)
    assert len(r["ref"]) + len(r["tgt"]) == len(meta["pairs"])
    for pk, gold in zip(list(r["ref"]) + list(r["tgt"]), meta["pairs"]):
        assert int(pk["lag"]) == gold["delay"], gold
        assert abs(float(pk["corr"]) - gold["corr"]) <= 0.5e-6 + CORR_TOL
    files = []
    for name, raw in zip(names, raws):
        f = tmp_path / f"sim-{name}-1.dat"
        raw.tofile(f)
        files.append(str(f))
    exe = GOLDEN.parent.parent / "tdoa-geolocation_b200" / "processor_b200"
    out = subprocess.run([str(exe), "162400000", "92300000", csv_file, *files], capture_output=True, text=True)
    assert out.returncode == meta["returncode"], out.stderr
    assert [l[20:] for l in out.stderr.splitlines()] == [l[20:] for l in meta["stderr_tail"]]
    skip = "Loading I/Q data from:"
    ours = [l for l in out.stdout.splitlines() if not l.startswith(skip)]
    gold = [l for l in (GOLDEN / f"{case}.stdout.txt").read_text().splitlines() if not l.startswith(skip)]
    assert len(ours) == len(gold), out.stdout[-2000:]
    for k, (a, b) in enumerate(zip(ours, gold)):
        assert _same_line(a, b), f"line {k}: ours {a!r} != reference {b!r}"
    buf = io.StringIO()
    p = T.TDOAProcessor(162400000.0, 92300000.0, csv_file, out=buf)
    failure = None
    try:
        p.process_tdoa(files)
    except RuntimeError as exc:
        failure = str(exc)
    p.close()
    _same_outcome(failure, meta)
    ours = [l for l in buf.getvalue().splitlines() if not l.startswith(skip)]
    assert len(ours) == len(gold)
    for k, (a, b) in enumerate(zip(ours, gold)):
        assert _same_line(a, b), f"python mirror, line {k}: ours {a!r} != reference {b!r}"


def _check_binary_solver(eng, llh, rd):
    got = eng.solve_binary(llh, rd)
    want = oracle.solve_binary(llh, rd)
    assert got[1:5] == want[1:5], (rd, got[1:5], want[1:5])           # status, n_valid, n_iter, converged
    assert got[5].shape == want[5].shape
    if len(want[5]):
        assert np.array_equal(got[5][:, 4], want[5][:, 4])             # which branch every iteration took
        assert np.allclose(got[5][:, :4], want[5][:, :4], rtol=1e-7, atol=1e-6)   # metres: ECEF coordinates of 6.4e6 m carry 1e-9 m of f64 rounding
    assert np.allclose(got[0][:2], want[0][:2], rtol=0, atol=1e-10) and abs(got[0][2] - want[0][2]) <= 1e-5
    return want


def test_binary_solver_equals_oracle(eng_binary):
    """tdoa_solve_binary (solveTDOA of the shipped binary, ELF 0x4a0360) against orc_solve_binary --
    itself pinned by the binary's printed traces (tests/test_oracle_golden.py) -- over all its
    branches: fewer / more than two valid measurements, limited and damped steps, convergence,
    ten iterations and both single-equation fall-backs."""
    rng = np.random.default_rng(12)
    seen = set()
    for _ in range(60):
        rd = rng.uniform(-30000.0, 30000.0, 3)
        if rng.random() < 0.5:
            rd[rng.integers(3)] = 25000.0 + rng.uniform(0, 1e4)        # one filtered: the solver runs
        if rng.random() < 0.3:
            rd[:2] = np.array([-4450.9, -952.5]) + rng.normal(0, 200.0, 2)   # near the start point: converges
            rd[2] = 4e4
        w = _check_binary_solver(eng_binary, STATION_LLH, rd)
        seen.add((w[1], w[4], bool(len(w[5]) and (w[5][:, 4] == 1).any()), w[3] == 10))
    assert {s[0] for s in seen} == {0, 1, 2} and any(s[1] for s in seen) and any(s[2] for s in seen) and any(s[3] for s in seen)
    back = np.array([STATION_LLH[0], STATION_LLH[1], STATION_LLH[0]])
    w = _check_binary_solver(eng_binary, back, [100.0, 300.0, 5e4])                        # det = 0: equation 1 alone
    assert w[1] == 0 and (w[5][:, 4] == 2).all()
    twin = np.array([STATION_LLH[0], STATION_LLH[1], STATION_LLH[1]])
    w = _check_binary_solver(eng_binary, twin, [100.0, 300.0, 5e4])                        # det = 0: equation 2 alone
    assert w[1] == 0 and (w[5][:, 4] == 3).all()
    four = np.vstack([STATION_LLH, [[41.30888549464701, -96.02619229605524, 356.0]]])
    assert _check_binary_solver(eng_binary, four, [100.0, 5e4, 5e4, 5e4, -200.0, 5e4])[2] == 2   # six measurements, two valid


def _cli_engine_cases():
    import json
    return json.loads((GOLDEN / "cli_errors.json").read_text())


@pytest.mark.parametrize("case", _cli_engine_cases()["engine_cases"])
def test_capture_file_errors_are_the_reference_binarys(tmp_path, capsys, case):
    """A capture whose name matches no station, a missing capture, a directory in place of one:
    stdout, stderr (after the log time stamp) and exit status of both host mirrors against the
    reference binary's (tests/golden/cli_errors.json; processor.go:110-122, :166-191, :757-765)."""
    import importlib.util
    import subprocess
    spec = importlib.util.spec_from_file_location("make_golden", GOLDEN / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    mg.cli_engine_setup(tmp_path)
    want = _cli_engine_cases()["cases"][case]
    csv_path = str(GOLDEN / "stations.csv")
    argv = [a.replace("{csv}", csv_path).replace("{dir}", str(tmp_path)) for a in want["args"]]
    size_line = re.compile(r"^File size: \d+ bytes, samples: \d+$")   # a directory's size is the file system's business

    def same(got_lines, want_lines):
        assert len(got_lines) == len(want_lines), (got_lines[-3:], want_lines[-3:])
        for a, b in zip(got_lines, want_lines):
            if "sim-kf0mtl-dir.dat" in " ".join(want["args"]) and size_line.match(a) and size_line.match(b):
                continue
            assert a == b

    exe = str(GOLDEN.parent.parent / "tdoa-geolocation_b200" / "processor_b200")
    r = subprocess.run([exe, *argv], capture_output=True, text=True, cwd=tmp_path)
    back = lambda t: t.replace(csv_path, "{csv}").replace(str(tmp_path), "{dir}").replace(exe, "{prog}")
    assert r.returncode == want["returncode"]
    same([back(l) for l in r.stdout.splitlines()], want["stdout"])
    same([back(l[20:]) for l in r.stderr.splitlines()], want["stderr"])
    from importlib import import_module
    proc = import_module("tdoa-geolocation_b200.processor")
    capsys.readouterr()
    rc = proc.main(argv, prog="{prog}")
    got = capsys.readouterr()
    assert rc == want["returncode"]
    same([back(l) for l in got.out.splitlines()], want["stdout"])
    same([back(l[20:]) for l in got.err.splitlines()], want["stderr"])


# ------------------------------------------------------------------ discriminator bit parity
def test_discriminator_bits_match_oracle(eng_binary):
    """Strong-FM branch, sample by sample: the custom f64 arctangent rounds to the same
    f32 as the reference's math.Atan2 (up to the ~1e-8/sample chance that two correct
    f64 implementations straddle an f32 rounding boundary)."""
    raws, _ = load_golden("fm_delays")
    eng_binary.load_u8(0, raws[0])
    ref, tgt = split(raws[0])
    for kind, sig in ((T.KIND_REF, ref), (T.KIND_TGT, tgt)):
        got, _, br = eng_binary.preprocess(0, kind, 0, sig.size)
        want, _ = oracle.preprocess_binary(sig)
        assert br == 0
        mism = int(np.count_nonzero(got.view(np.uint32) != want.view(np.uint32)))
        assert mism <= 2, f"{mism} of {sig.size} samples differ"


def test_fast_demod_keeps_the_lags():
    raws, meta = load_golden("fm_delays")
    with T.Engine(T.MODE_BINARY, fast_demod=1) as e:
        load_all(e, raws)
        got = list(e.xcorr(T.KIND_REF)[0]) + list(e.xcorr(T.KIND_TGT)[0])
    for pk, gold in zip(got, meta["pairs"]):
        assert int(pk["lag"]) == gold["delay"]
        assert abs(float(pk["corr"]) - gold["corr"]) <= 5e-6


def test_constant_divisor_division_is_exact(eng_binary):
    """The small box-car divides by its tap count with a reciprocal + FMA correction; the
    device checks it against a correctly rounded divide for every float bit pattern."""
    assert eng_binary.selftest(0) == 0


def test_lean_discriminator_all_byte_quads(eng_binary):
    """The production discriminator (k_demod_df: f64 products rounded to f32, f32 double-float
    arctangent two samples per packed instruction, full-accuracy f64 fall-back wherever the fast
    value sits within 2^-40 of an f32 rounding boundary) against the reference statement (F2F
    conversions, gates, the older f64 arctangent) over every (previous, current) byte quad -- the
    discriminator is a function of four bytes, so this is exhaustive.  Measured: 0 differing quads
    of 2^32, 132 432 fall-backs of which 112 changed the value (profiles/r2_demod.md), and 0 is what
    is asserted; selftest(3) is the same check of the round-1 kernel (fast_demod = 2).  What it proves:
    the production kernel equals the engine's plain statement of the reference's discriminator
    (demod_one<false>: f64 products, F2F roundings, the gates, a <= 1 ulp f64 arctangent); the link
    from that statement to Go's math.Atan2 is test_discriminator_bits_match_oracle (libm's atan2
    on the CPU: <= 2 differing samples of 150 000, each where two correct f64 arctangents
    straddle an f32 rounding boundary)."""
    bad = eng_binary.selftest(1)
    print("lean discriminator: differing byte quads =", bad, eng_binary.last_error() if bad else "")
    assert bad == 0
    assert eng_binary.selftest(3) == 0
    falls = eng_binary.selftest(2)
    assert 0 < falls < 1 << 20       # the exact fall-back exists and is rare (3e-5 of the quads)


def test_production_discriminator_on_random_windows():
    """k_demod_df stages 16-byte chunks from the 16-byte-aligned address below a tile's previous sample and
    takes the fast path only for tiles that lie, with 8 samples of margin on both sides, inside one run of the
    capture; everything else goes sample by sample.  Random (start, length) views -- every 16-byte
    misalignment class, lengths below one tile, REF views across the block-1 / block-3 joint, views ending at
    the capture's last sample -- must preprocess to the bits of the round-1 kernel (fast_demod = 2: the f64
    arctangent on every sample, staged word by word), power and DC included."""
    rng = np.random.default_rng(77)
    block = 61003                      # odd block length: the joint and the tail sit at odd byte offsets
    raws = fm_capture(block, (0, 5, 11), (0, 17, 30), seed=23)
    cases = [(T.KIND_REF, 0, 2 * block), (T.KIND_TGT, 0, block), (T.KIND_REF, block - 4100, 8300), (T.KIND_TGT, block - 5000, 5000)]
    for _ in range(36):
        kind = int(rng.integers(0, 2))
        n = (2 if kind == T.KIND_REF else 1) * block
        length = int(rng.choice([3, 17, 4095, 4096, 4097, 9000, 20000, 40000]))
        start = int(rng.integers(0, n - length + 1))
        cases.append((kind, start, length))
    with T.Engine(T.MODE_BINARY, chunk_samples=0) as new, T.Engine(T.MODE_BINARY, chunk_samples=0, fast_demod=2) as old:
        load_all(new, raws)
        load_all(old, raws)
        for kind, start, length in cases:
            a, pa, ba = new.preprocess(1, kind, start, length)
            b, pb, bb = old.preprocess(1, kind, start, length)
            assert ba == bb == 0, (kind, start, length)
            assert abs(pa - pb) <= 4e-16 * abs(pb), (kind, start, length, pa, pb)   # the same terms, another (fixed) order
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (kind, start, length)


# ------------------------------------------------------------------ lazy pinned load (copy-following discriminator)
@pytest.mark.parametrize("copy_chunk", [0, 65536, 8192])
def test_pinned_lazy_load_equals_synchronous_load(copy_chunk):
    """tdoa_load_u8_pinned queues the host->device copies in chunks (REF blocks of every
    station first when REF is asked for first) and the discriminator follows them; the
    records must be those of the synchronous tdoa_load_u8 path.  The power / DC sums are
    added chunk by chunk, i.e. in another (fixed) order: lags identical, correlation to
    1e-12; and against the oracle as everywhere else."""
    block = 150000
    raws = fm_capture(block, (0, 5, 11), (0, 17, 30), seed=21)
    with T.Engine(T.MODE_BINARY, chunk_samples=0) as a:
        load_all(a, raws)
        want = [a.xcorr(T.KIND_REF)[0], a.xcorr(T.KIND_TGT)[0]]
    bufs = []
    for r in raws:
        b = T.host_alloc(r.size)
        b.array[:] = r
        bufs.append(b)
    for order in ((T.KIND_REF, T.KIND_TGT), (T.KIND_TGT, T.KIND_REF)):
        with T.Engine(T.MODE_BINARY, chunk_samples=0, copy_chunk=copy_chunk) as e:
            for rep in range(2):  # second round: reload while the chunk events of the first still exist
                for k, b in enumerate(bufs):
                    e.load_u8_pinned(k, b)
                got = {kind: e.xcorr(kind)[0] for kind in order}
                for kind in (T.KIND_REF, T.KIND_TGT):
                    assert np.array_equal(got[kind]["lag"], want[kind]["lag"])
                    assert np.allclose(got[kind]["corr"], want[kind]["corr"], rtol=0, atol=1e-12)
                    assert np.array_equal(got[kind]["n_blocks"], want[kind]["n_blocks"])
            # the one-call path on a capture that is still arriving (what bench.py's e2e leg does)
            for k, b in enumerate(bufs):
                e.load_u8_pinned(k, b)
            r = e.process(STATION_LLH)
            for kind, key in ((T.KIND_REF, "ref"), (T.KIND_TGT, "tgt")):
                assert np.array_equal(r[key]["lag"], want[kind]["lag"])
                assert np.allclose(r[key]["corr"], want[kind]["corr"], rtol=0, atol=1e-12)
            # other readers of the capture wait for the copies too
            for k, b in enumerate(bufs):
                e.load_u8_pinned(k, b)
            assert np.array_equal(e.unpack(2, 3 * block - 64, 64), oracle.unpack_u8(raws[2])[3 * block - 64:])
            sig, p0, br = e.preprocess(1, T.KIND_TGT, 4096, 50000)
            with T.Engine(T.MODE_BINARY) as s:
                load_all(s, raws)
                sig2, p02, br2 = s.preprocess(1, T.KIND_TGT, 4096, 50000)
            assert br == br2 and abs(p0 - p02) <= 1e-12 and np.allclose(sig, sig2, rtol=0, atol=2e-6)
    ref, tgt = oracle.process_capture_binary(raws)  # 300 000 / 150 000 samples: under the 1 M chunk
    for got, w in zip(list(want[0]) + list(want[1]), ref + tgt):
        assert int(got["lag"]) == w[0] and abs(float(got["corr"]) - w[1]) <= CORR_TOL
    for b in bufs:
        b.free()


# ------------------------------------------------------------------ streaming file loader, guard samples
def test_load_file_streams_the_capture(tmp_path, eng_binary):
    """tdoa_load_file (loadIQData, processor.go:166-205): same device bytes as tdoa_load_u8,
    the reference's sample count (:183) and its error text for a missing file (:170-172)."""
    raws = fm_capture(60001, (0, 4, 9), (0, 11, 25), seed=33)   # odd sizes: N % 3 != 0 and a ragged last piece
    for k, r in enumerate(raws):
        p = tmp_path / f"sim-st{k}-1.dat"
        r.tofile(p)
        assert eng_binary.load_file(k, p) == r.size // 2
        n = r.size // 2
        assert np.array_equal(eng_binary.unpack(k, n - 1000, 1000), oracle.unpack_u8(r)[n - 1000:])
        assert np.array_equal(eng_binary.unpack(k, 0, 1000), oracle.unpack_u8(r)[:1000])
    got = eng_binary.xcorr(T.KIND_TGT)[0]
    load_all(eng_binary, raws)
    want = eng_binary.xcorr(T.KIND_TGT)[0]
    assert np.array_equal(got["lag"], want["lag"]) and np.array_equal(got["corr"], want["corr"])
    with pytest.raises(T.TdoaError) as ei:
        eng_binary.load_file(0, tmp_path / "absent.dat")
    assert ei.value.code == -7 and "failed to open file" in str(ei.value)
    big = np.random.default_rng(5).integers(0, 256, 70_000_000, dtype=np.uint8)   # three 32 MB pieces
    big.tofile(tmp_path / "big.dat")
    assert eng_binary.load_file(1, tmp_path / "big.dat") == big.size // 2
    for first in (0, (32 << 19) - 8, (32 << 20) - 8, big.size // 2 - 16):
        assert np.array_equal(eng_binary.unpack(1, first, 16), oracle.unpack_u8(big[2 * first:2 * first + 32]))


def test_guard_samples_trim_the_retuned_blocks():
    """guard_samples = G (engine-defined; 0 is the reference's split): REF = block 1 ++ block 3[G:],
    TGT = block 2[G:]; everything downstream is the reference's arithmetic on those signals."""
    G = 5000
    raws = fm_capture(90000, (0, 6, 13), (0, 15, 33), seed=44)
    with T.Engine(T.MODE_BINARY, guard_samples=G) as e:
        load_all(e, raws)
        ref = e.xcorr(T.KIND_REF)[0]
        tgt = e.xcorr(T.KIND_TGT)[0]
    sigs = []
    for r in raws:
        d = oracle.unpack_u8(r)
        b = len(d) // 3
        sigs.append((np.concatenate([d[:b], d[2 * b + G:3 * b]]), d[b + G:2 * b]))
    for p, (i, j) in enumerate([(0, 1), (0, 2), (1, 2)]):
        for got, k in ((ref[p], 0), (tgt[p], 1)):
            d, c, _ = oracle.cross_correlate_binary(sigs[i][k], sigs[j][k])
            assert int(got["lag"]) == d and abs(float(got["corr"]) - c) <= CORR_TOL


# ------------------------------------------------------------------ ProcessTDOA as one call
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_process_equals_the_separate_calls(case):
    """tdoa_process queues both pair loops and the fix without synchronising in between and
    checks its branch guesses afterwards; the weak / moderate captures take its fall-back.
    Records, diagnostics, differences and fix must be those of tdoa_xcorr x 2 + tdoa_solve."""
    raws, meta = load_golden(case)
    for mode in (T.MODE_BINARY, T.MODE_SOURCE):
        with T.Engine(mode) as a, T.Engine(mode) as b:
            load_all(a, raws)
            load_all(b, raws)
            ref, tgt = a.xcorr(T.KIND_REF)[0], a.xcorr(T.KIND_TGT)[0]
            info = [a.xcorr_info(k) for k in (T.KIND_REF, T.KIND_TGT)]
            for rep in range(2):   # second call: the branch memo is warm
                r = b.process(STATION_LLH)
                for name in ("lag", "corr", "flags", "first_lag", "n_blocks"):
                    assert np.array_equal(r["ref"][name], ref[name]), (name, rep)
                    assert np.array_equal(r["tgt"][name], tgt[name]), (name, rep)
                fs = a.cfg.sample_rate
                tt, tr = tgt["lag"].astype(np.float64) / fs, ref["lag"].astype(np.float64) / fs
                td = tt if mode == T.MODE_SOURCE else tt - tr       # processor.go:853 / the binary's correction
                assert np.array_equal(r["time_differences"], td)
                assert np.array_equal(r["range_differences"], td * 299792458.0)
                pos, status, iters = a.solve(STATION_LLH, td * 299792458.0)
                assert r["status"] == status and r["iters"] == iters   # (fm_delays: the reference's solver goes singular)
                assert np.array_equal(r["position"], pos)
                for k in (T.KIND_REF, T.KIND_TGT):
                    got_sig, got_first = b.xcorr_info(k)
                    assert got_sig == info[k][0] and np.array_equal(got_first, info[k][1])
    if case == "fm_strong":
        want_ref, want_tgt = oracle.process_capture_binary(raws)
        with T.Engine(T.MODE_BINARY) as e:
            load_all(e, raws)
            r = e.process(STATION_LLH)
        for got, want in zip(list(r["ref"]) + list(r["tgt"]), want_ref + want_tgt):
            assert int(got["lag"]) == want[0] and abs(float(got["corr"]) - want[1]) <= CORR_TOL


# ------------------------------------------------------------------ every search path against the every-lag path
@pytest.mark.parametrize("S,L,W,hop,nw", [
    (5, 700, 30000, 9000, 2),      # 10 pairs: parked spectra, 2048-lag chunks, odd station count (a transform with one plane)
    (17, 1100, 26000, 0, 1),       # 136 pairs: two accumulation jobs share one set of spectra; 4096-lag chunk
    (3, 4500, 50000, 20000, 2),    # 9001 lags: the 2^21-point transform
    (4, 2500, 40000, 0, 1),        # 6 pairs, 5001 lags: 2 x 2 tiles, three 2048-lag chunks
    (20, 300, 12000, 0, 1),        # 190 pairs on 20 stations: more packed transforms than one spectra job holds -> tiles
])
def test_search_paths_agree_with_every_lag_evaluation(S, L, W, hop, nw):
    """Whatever ranks the lags (tiles, parked spectra, the big transform), the record must be the one the
    exhaustive time-domain evaluation of every lag (use_fft=0, the reference's order) produces."""
    rng = np.random.default_rng(S * 1000 + L)
    d = tuple(int(x) for x in rng.integers(0, min(L, 400), S))
    raws = fm_capture(W + hop * (nw - 1) + 2000, d, d, seed=S)
    with T.Engine(T.MODE_EXTENDED, n_stations=S, max_lag=L) as e, T.Engine(T.MODE_EXTENDED, n_stations=S, max_lag=L, use_fft=0) as b:
        load_all(e, raws)
        load_all(b, raws)
        got = e.xcorr(T.KIND_TGT, 300, W, nw, hop)
        want = b.xcorr(T.KIND_TGT, 300, W, nw, hop)
    assert np.array_equal(got["lag"], want["lag"])
    assert np.allclose(got["corr"], want["corr"], rtol=0, atol=1e-12)
    assert np.allclose(got["frac"], want["frac"], rtol=0, atol=1e-6)
    pairs = [(i, j) for i in range(S) for j in range(i + 1, S)]
    for w in range(nw):
        assert [int(x) for x in got[w]["lag"]] == [d[j] - d[i] for i, j in pairs]


# ------------------------------------------------------------------ the sequential DC chain, chunk-parallel
@pytest.mark.parametrize("case", ["fm_strong", "moderate", "weak_noise", "big_fm", "big_weak"])
def test_chunked_sequential_dc_is_the_plain_walk(case):
    """removeDCBias's sequential f32 accumulator (processor.go:304-309): the chunk-parallel
    evaluation (seqsum.cu) must give the bits of the plain one-add-at-a-time walk (use_fft=5
    keeps the old kernel) -- DC of every signal, hence every record."""
    if case == "big_fm":
        raws = fm_capture(700_000, (0, 9, 31), (0, 9, 31), seed=5)                      # 1.4 M / 0.7 M samples per signal
    elif case == "big_weak":
        rng = np.random.default_rng(9)
        raws = [rng.integers(118, 138, 2 * 3 * 500_000, dtype=np.uint8) for _ in range(3)]   # weak branch: complex, zero-mean
    else:
        raws, _ = load_golden(case)
    for mode in (T.MODE_BINARY, T.MODE_SOURCE):
        # (SOURCE arithmetic never uses the FFT search, so the switch only selects the DC kernel there)
        with T.Engine(mode, chunk_samples=0) as a, T.Engine(mode, chunk_samples=0, use_fft=5) as b:
            load_all(a, raws)
            load_all(b, raws)
            for kind in (T.KIND_REF, T.KIND_TGT):
                ga, gb = a.xcorr(kind)[0], b.xcorr(kind)[0]
                ia, ib = a.xcorr_info(kind)[0], b.xcorr_info(kind)[0]
                for sa, sb in zip(ia, ib):
                    assert sa["dc_re"] == sb["dc_re"] and sa["dc_im"] == sb["dc_im"], (case, mode, kind)
                    assert sa["power1"] == sb["power1"]
                assert np.array_equal(ga["lag"], gb["lag"]) and np.array_equal(ga["corr"], gb["corr"])
