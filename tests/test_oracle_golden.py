"""Pins the CPU oracle (oracle/tdoa_oracle.c) against what the reference's own shipped
binary printed for the golden captures (tests/golden/*.json, produced by
tests/golden/make_golden.py), and against the known answers in the reference's docs."""
import numpy as np
import pytest

from oracle import oracle
from helpers import GOLDEN_CASES, GOLDEN_DEGENERATE_CASES, GOLDEN_LONG_CASES, GOLDEN_SIM_CASES, GOLDEN_ORDER_CASES, GOLDEN_TABLE_CASES, STATION_LLH, load_golden


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_binary_pairs_match_reference_stdout(case):
    raws, meta = load_golden(case)
    ref, tgt = oracle.process_capture_binary(raws)
    got = [("REF",) + r for r in ref] + [("TGT",) + r for r in tgt]
    assert len(got) == len(meta["pairs"]) == 6
    for (kind, delay, corr, _), want in zip(got, meta["pairs"]):
        assert kind == want["kind"]
        assert delay == want["delay"], (case, want)
        # the reference prints %.6f
        assert abs(corr - want["corr"]) <= 0.51e-6, (case, corr, want)


@pytest.mark.parametrize("case", GOLDEN_SIM_CASES)
def test_binary_on_the_reference_simulators_content(case):
    """simulator.go / weak_signal_simulator.go content (BASELINE configs[0] and configs[3] name them;
    restated in tools/simulators.py): every signal takes the weak branch (1001-tap high-pass, 5-tap
    low-pass), the tones give periodic peaks and sanity re-searches, the weak simulator's reference
    blocks quantise to a constant.  Records as the binary printed them."""
    raws, meta = load_golden(case)
    assert set(meta["branch"]) == {2}
    ref, tgt = oracle.process_capture_binary(raws)
    got = [("REF",) + r for r in ref] + [("TGT",) + r for r in tgt]
    assert len(got) == len(meta["pairs"]) == 6
    for (kind, delay, corr, _), want in zip(got, meta["pairs"]):
        assert (kind, delay) == (want["kind"], want["delay"]), (case, want)
        assert abs(corr - want["corr"]) <= 0.51e-6, (case, corr, want)


@pytest.mark.parametrize("case", GOLDEN_DEGENERATE_CASES + GOLDEN_ORDER_CASES)
def test_binary_degenerate_captures(case):
    """A capture of 2 samples (returned unchanged as REF and TGT), of 3 samples (one-sample
    blocks) and an empty one; captures given in another order; four collectors (pairs i < j in
    the order of the arguments, processor.go:816-817): the records the binary printed."""
    raws, meta = load_golden(case)
    ref, tgt = oracle.process_capture_binary(raws)
    got = [("REF",) + r for r in ref] + [("TGT",) + r for r in tgt]
    assert len(got) == len(meta["pairs"]) == len(raws) * (len(raws) - 1)
    for (kind, delay, corr, _), want in zip(got, meta["pairs"]):
        assert (kind, delay) == (want["kind"], want["delay"]), (case, want)
        assert abs(corr - want["corr"]) <= 0.51e-6, (case, corr, want)


@pytest.mark.parametrize("case", GOLDEN_LONG_CASES)
def test_binary_truncates_to_its_test_chunk(case):
    """Blocks longer than 1 000 000 samples: the binary cuts REF and TGT to their first
    1 000 000 samples (processor.go:772-780 with the shipped chunk); records as printed."""
    raws, meta = load_golden(case)
    assert len(raws[0]) // 2 // 3 > 1_000_000
    ref, tgt = oracle.process_capture_binary(raws)
    got = [("REF",) + r for r in ref] + [("TGT",) + r for r in tgt]
    for (kind, delay, corr, _), want in zip(got, meta["pairs"]):
        assert (kind, delay) == (want["kind"], want["delay"]), (case, want)
        assert abs(corr - want["corr"]) <= 0.51e-6, (case, corr, want)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_binary_preprocess_diagnostics(case):
    """Initial power / branch of every preprocessSignal call the binary printed."""
    raws, meta = load_golden(case)
    sigs = []
    for kind in ("ref", "tgt"):
        per = []
        for raw in raws:
            d = oracle.unpack_u8(raw)
            s = oracle.extract_reference(d) if kind == "ref" else oracle.extract_target(d)
            per.append(s[:1_000_000])
        for i in range(3):
            for j in range(i + 1, 3):
                sigs += [per[i], per[j]]
    assert len(sigs) == len(meta["initial_power"]) == 12
    for s, p_want, br_want in zip(sigs, meta["initial_power"], meta["branch"]):
        p = oracle.signal_power(s)
        assert abs(p - p_want) <= 0.51e-9 * max(1.0, abs(p_want) / 1e-9 * 1e-9) + 1e-9
        _, br = oracle.preprocess_binary(s)
        assert br == br_want


def _golden_solver_inputs(case):
    import json
    from helpers import GOLDEN, STATIONS
    meta = json.loads((GOLDEN / f"{case}.json").read_text())
    names = meta.get("order", STATIONS)
    table = {row.split(",")[0]: [float(v) for v in row.split(",")[1:]]
             for row in (GOLDEN / meta.get("csv", "stations.csv")).read_text().splitlines()[1:]}
    llh = np.array([table[n] for n in names])
    P = len(names) * (len(names) - 1) // 2
    ref = [q["delay"] for q in meta["pairs"][:P]]
    tgt = [q["delay"] for q in meta["pairs"][P:]]
    rd = [(t / 2e6 - r / 2e6) * 299792458.0 for t, r in zip(tgt, ref)]   # the binary's correction, then * c
    return meta, llh, rd, (GOLDEN / f"{case}.stdout.txt").read_text()


ALL_GOLDEN = GOLDEN_CASES + GOLDEN_DEGENERATE_CASES + GOLDEN_LONG_CASES + GOLDEN_ORDER_CASES + GOLDEN_TABLE_CASES + GOLDEN_SIM_CASES


def check_binary_solver_against_stdout(llh, rd, text, returncode, err):
    """orc_solve_binary against what the binary printed (text) and how it ended (returncode, err =
    its stderr line after the time stamp): outcome, iteration trace, every branch line, location."""
    import re
    out, status, n_valid, n_iter, conv, trace = oracle.solve_binary(llh, rd)
    assert n_valid == sum(1 for l in text.splitlines() if l.startswith("VALID: Range difference"))
    if status == 1:
        assert err == ("TDOA processing failed: TDOA solution failed: insufficient valid measurements: only %d of %d "
                       "range differences are reliable" % (n_valid, len(rd)))
        return status
    if status == 2:
        assert err == "TDOA processing failed: TDOA solution failed: no valid range difference measurements remain"
        assert "Iteration 0" not in text
        return status
    assert status == 0 and returncode == 0
    lines = ["Iteration %d: det=%.2e, residuals=[%.1f, %.1f]" % (k, t[0], t[1], t[2]) for k, t in enumerate(trace)]
    assert lines == [l for l in text.splitlines() if l.startswith("Iteration ")]
    big = ["Large step detected (%.1fm) - limiting to %.1fm" % (t[3], 1000.0 * (1000.0 / t[3] * 0.7)) for t in trace if t[4] == 1]
    assert big == [l for l in text.splitlines() if l.startswith("Large step detected")]
    single = ["Using single equation approach (equation %d)" % (t[4] - 1) for t in trace if t[4] >= 2]
    assert single == [l for l in text.splitlines() if l.startswith("Using single equation approach")]
    singular = ["Singular matrix detected (det=%.2e) - trying alternative approach" % t[0] for t in trace if t[4] >= 2]
    assert singular == [l for l in text.splitlines() if l.startswith("Singular matrix detected")]
    assert ("Converged after %d iterations" % n_iter in text) == conv
    assert ("Maximum iterations reached" in text) == (n_iter == 10)
    lat, lon, elev = (float(re.search(p, text).group(1)) for p in
                      (r"Latitude:\s+(-?[\d.]+)°", r"Longitude:\s+(-?[\d.]+)°", r"Elevation:\s+(-?[\d.]+) m"))
    assert "%.6f" % out[0] == "%.6f" % lat and "%.6f" % out[1] == "%.6f" % lon and "%.1f" % out[2] == "%.1f" % elev
    return status


@pytest.mark.parametrize("case", ALL_GOLDEN)
def test_binary_solver_outcome_and_trace(case):
    """solveTDOA of the shipped binary (ELF 0x4a0360, orc_solve_binary) against what the binary
    printed for every golden capture: which of its three outcomes (fewer than two valid range
    differences / more than two / a fix), every "Iteration k: det=..., residuals=[...]" line, every
    "Large step detected" / single-equation line, the convergence line and the location to the
    printed digits."""
    meta, llh, rd, text = _golden_solver_inputs(case)
    err = meta["stderr_tail"][0][20:] if meta["stderr_tail"] else ""
    check_binary_solver_against_stdout(llh, rd, text, meta["returncode"], err)


def test_binary_solver_gives_a_fix_on_five_goldens():
    fixes = [c for c in ALL_GOLDEN if _golden_solver_inputs(c)[0]["returncode"] == 0]
    # ten limited steps / four damped steps to convergence / single equation 1 and 2 / poor geometry
    assert sorted(fixes) == ["back_stations", "close_stations", "fm_reordered", "fm_two_valid", "twin_stations"]


def test_baselines_known_answers():
    # PROJECT_NOTES.md:25-27: 12.29 / 17.02 / 10.02 km
    want = [12.29, 17.02, 10.02]
    got = [oracle.baseline(STATION_LLH[i], STATION_LLH[j]) / 1000.0
           for i in range(3) for j in range(i + 1, 3)]
    for g, w in zip(got, want):
        assert abs(g - w) < 0.005


def test_microsecond_to_metre_diagnostic():
    # processor.go:885-889: 10 / 5 / -3 us -> 2997.9 / 1499.0 / -899.4 m
    for us, m in ((10, 2997.9), (5, 1499.0), (-3, -899.4)):
        assert abs(us * 1e-6 * 299792458.0 - m) < 0.05


def test_source_solver_survey_values():
    # SURVEY.md 8c cross-check values of the f64 restatement of processor.go:932-1045
    cases = {
        (0.0, 0.0, 0.0): (41.262323848, -95.982993948, -681.535),
        (3.5, 6.0, 2.5): (41.254478374, -95.979774992, 311.680),
        (10.0, 5.0, -3.0): (41.265694717, -95.943691803, -1108.147),
        (-3.5, 2.0, 0.0): (41.254686678, -95.998887381, 285.304),
    }
    for us, want in cases.items():
        rd = np.array(us) * 1e-6 * 299792458.0
        llh, status, _ = oracle.solve_tdoa(STATION_LLH, rd)
        assert status == 0
        assert abs(llh[0] - want[0]) < 5e-9 and abs(llh[1] - want[1]) < 5e-9
        assert abs(llh[2] - want[2]) < 5e-3


def test_ecef_round_trip():
    for llh in STATION_LLH:
        xyz = oracle.llh_to_ecef(*llh)
        back = oracle.ecef_to_llh(*xyz)
        assert np.allclose(back[:2], llh[:2], atol=1e-9) and abs(back[2] - llh[2]) < 1e-3


def test_self_correlation_sanity():
    # correlation_sanity.go:48-63 / simple_corr.go:31-44: crossCorrelate(x, x) > 0.5 at delay 0
    raws, _ = load_golden("fm_strong")
    ref = oracle.extract_reference(oracle.unpack_u8(raws[0]))
    d, c, _ = oracle.cross_correlate_binary(ref, ref)
    assert d == 0 and c > 0.8
    d, c = oracle.cross_correlate_source(ref, ref)
    assert d == 0 and c > 0.5
