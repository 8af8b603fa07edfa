"""Host-side mirror of the reference's per-capture analyzers, numbers from the CUDA engine:

    fast_analyzer <file.dat>            -> "REF,snr,power,clip,overload" / "TGT,..."   (fast_analyzer.go:27-52)
    analyzer <file.dat> [duration_s]    -> statistics, quality flags, gain verdict, REF/TGT comparison
                                           (analyzer.go:44-83, :350-487)

The verdict lines that depend on the measurements are reproduced; the reference's static
advice paragraphs (lists of future collector options, analyzer.go:489-628) are not.
"""
from __future__ import annotations

import sys
from typing import Optional, Tuple

import numpy as np

from . import _native as N


def _gobool(x) -> str:
    return "true" if x else "false"


def _load(filename: str, engine: Optional[N.Engine]) -> Tuple[N.Engine, bool, int]:
    try:
        raw = np.fromfile(filename, dtype=np.uint8)
    except OSError as exc:
        raise RuntimeError(f"failed to open file: {exc}") from exc
    own = engine is None
    if own:
        engine = N.Engine(N.MODE_BINARY, n_stations=2)
    engine.load_u8(0, raw)
    return engine, own, raw.size // 2


def fast_analyze_dual_frequency_file(filename: str, engine: Optional[N.Engine] = None):
    """fast_analyzer.go:54-111: (ref, tgt) quality of the first 32768 samples of each block."""
    eng, own, _ = _load(filename, engine)
    try:
        return eng.analyze(0, fast=True)
    except N.TdoaError as exc:
        raise RuntimeError(str(exc)) from exc
    finally:
        if own:
            eng.close()


def fast_main(argv=None, out=sys.stdout) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 1:
        print("Usage: fast_analyzer <data_file.dat>\nFast signal quality analyzer for gain sweeps", file=out)
        return 1
    try:
        ref, tgt = fast_analyze_dual_frequency_file(argv[0])
    except RuntimeError as exc:
        print(f"Error: {exc}", file=out)
        return 1
    for label, a in (("REF", ref), ("TGT", tgt)):
        print("%s,%.1f,%.1f,%s,%s" % (label, a["snr_db"], a["power_db"], _gobool(a["has_clipping"]),
                                      _gobool(a["has_overload"])), file=out)
    return 0


# ---- analyzer.go
def count_quality_issues(a) -> int:  # analyzer.go:448-457
    return (int(bool(a["has_clipping"])) + int(bool(a["has_overload"])) + int(bool(a["has_dead_zones"])) +
            int(bool(a["has_noise"])) + int(a["dc_offset"] > 10) + int(a["iq_imbalance"] > 0.1))


def assess_tdoa_suitability(a) -> bool:  # analyzer.go:459-470
    if a["has_clipping"] or a["has_overload"] or a["has_dead_zones"]:
        return False
    if a["snr_db"] < 15:
        return False
    return not (a["dc_offset"] > 15 or a["iq_imbalance"] > 0.15)


def gain_verdict(a) -> str:  # analyzer.go:472-499, first line of each case
    if a["has_clipping"]:
        return "🔻 REDUCE GAIN: Signal clipping detected"
    if a["has_overload"]:
        return "🔻 REDUCE GAIN: Signal appears overloaded"
    if a["power_db"] < -60:
        return "🔺 INCREASE GAIN: Signal level very low"
    if a["power_db"] < -40:
        return "🔺 INCREASE GAIN: Signal level low"
    if a["i_std"] > 50 and a["q_std"] > 50:
        return "✅ GAIN OK: Good signal level, no clipping"
    return "🔧 FINE-TUNE GAIN: Signal usable but could be optimized"


def summary_verdict(a) -> str:  # analyzer.go:618-627
    if a["has_clipping"] or a["has_overload"]:
        return "❌ CRITICAL: Adjust gain immediately - signal distortion present"
    if a["power_db"] < -50:
        return "⚠️  WARNING: Signal very weak - increase gain or check antenna"
    if a["dc_offset"] > 10 or a["iq_imbalance"] > 0.1:
        return "🔧 HARDWARE: RTL-SDR calibration issues detected"
    return "✅ ACCEPTABLE: Signal quality adequate for TDOA processing"


def print_analysis_results(a, out=sys.stdout):  # analyzer.go:350-371
    P = lambda *x: print(*x, file=out)
    P("=== Signal Statistics ===")
    P("Total Samples: %d" % a["total_samples"])
    P("I Channel: min=%d, max=%d, avg=%.1f, σ=%.1f" % (a["i_min"], a["i_max"], a["i_avg"], a["i_std"]))
    P("Q Channel: min=%d, max=%d, avg=%.1f, σ=%.1f" % (a["q_min"], a["q_max"], a["q_avg"], a["q_std"]))
    P("\n=== Signal Quality Metrics ===")
    P("DC Offset: %.1f (should be ~0)" % a["dc_offset"])
    P("IQ Imbalance: %.3f (should be <0.1)" % a["iq_imbalance"])
    P("Estimated SNR: %.1f dB" % a["snr_db"])
    P("Power Level: %.1f dB" % a["power_db"])
    P("\n=== Quality Flags ===")
    for name, key in (("Clipping/Saturation", "has_clipping"), ("Overload (too low variation)", "has_overload"),
                      ("Dead zones detected", "has_dead_zones"), ("Excessive noise", "has_noise")):
        P("%s %s: %s" % (("⚠️", name, "DETECTED") if a[key] else ("✅", name, "OK")))


def compare_signals(ref, tgt, out=sys.stdout):  # analyzer.go:398-446
    P = lambda *x: print(*x, file=out)
    P("\n=== SIGNAL COMPARISON ===")
    P("SNR Comparison:")
    P("  Reference: %.1f dB" % ref["snr_db"])
    P("  Target:    %.1f dB" % tgt["snr_db"])
    if ref["snr_db"] > tgt["snr_db"] + 10:
        P("  ⚠️  Reference significantly stronger - consider reducing reference gain")
    elif tgt["snr_db"] > ref["snr_db"] + 10:
        P("  ⚠️  Target significantly stronger - consider reducing target gain")
    else:
        P("  ✅ Signal levels reasonably balanced")
    P("\nPower Level Comparison:")
    P("  Reference: %.1f dB" % ref["power_db"])
    P("  Target:    %.1f dB" % tgt["power_db"])
    P("\nQuality Issues:")
    ri, ti = count_quality_issues(ref), count_quality_issues(tgt)
    P("  Reference: %d issues detected" % ri)
    P("  Target:    %d issues detected" % ti)
    if ri == 0 and ti == 0:
        P("  ✅ Both signals appear suitable for TDOA processing")
    elif ri > ti:
        P("  ⚠️  Reference signal needs more attention")
    elif ti > ri:
        P("  ⚠️  Target signal needs more attention")
    P("\n=== TDOA SUITABILITY ASSESSMENT ===")
    rs, ts = assess_tdoa_suitability(ref), assess_tdoa_suitability(tgt)
    if rs and ts:
        P("✅ EXCELLENT: Both signals suitable for TDOA correlation")
    elif not rs and not ts:
        P("❌ POOR: Both signals need improvement before TDOA processing")
    elif not rs:
        P("⚠️  MARGINAL: Reference signal needs improvement")
    else:
        P("⚠️  MARGINAL: Target signal needs improvement")


def analyze_dual_frequency_file(filename: str, engine: Optional[N.Engine] = None, out=sys.stdout):
    """analyzer.go:85-128: whole-block analysis; prints the file-structure header."""
    eng, own, total = _load(filename, engine)
    try:
        ref, tgt = eng.analyze(0, fast=False)
    except N.TdoaError as exc:
        raise RuntimeError(str(exc)) from exc
    finally:
        if own:
            eng.close()
    P = lambda *x: print(*x, file=out)
    P("=== File Structure Analysis ===")
    P("Total samples: %d" % total)
    P("Samples per frequency block: %d" % (total // 3))
    P("Reference samples: %d (blocks 1+3)" % ref["total_samples"])
    P("Target samples: %d (block 2)\n" % tgt["total_samples"])
    return ref, tgt


def main(argv=None, out=sys.stdout) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 1:
        print("Usage: analyzer <data_file.dat> [expected_duration_seconds]", file=out)
        return 1
    duration = 30
    if len(argv) > 1:
        try:
            duration = int(argv[1])
        except ValueError:
            pass
    P = lambda *x: print(*x, file=out)
    P("=== Advanced Signal Quality Analysis ===")
    P("File: %s" % argv[0])
    P("Expected Duration: %d seconds\n" % duration)
    try:
        ref, tgt = analyze_dual_frequency_file(argv[0], out=out)
    except RuntimeError as exc:
        P("Error analyzing file: %s" % exc)
        return 1
    for title, label, a in (("REFERENCE", "Reference", ref), ("TARGET", "Target", tgt)):
        P(("" if title == "REFERENCE" else "\n") + "=== %s SIGNAL ANALYSIS ===" % title)
        print_analysis_results(a, out)
        P("\n=== %s SIGNAL RECOMMENDATIONS ===" % label)
        P("\n--- Gain Recommendations ---")
        P(gain_verdict(a))
        if a["snr_db"] < 10:
            P("📡 SNR TOO LOW (%.1f dB): Increase gain or improve antenna" % a["snr_db"])
        elif a["snr_db"] > 40:
            P("📡 SNR HIGH (%.1f dB): Consider reducing gain to prevent overload" % a["snr_db"])
        P("\n=== SUMMARY ===")
        P(summary_verdict(a))
    compare_signals(ref, tgt, out)
    return 0


if __name__ == "__main__":
    raise SystemExit(fast_main() if "--fast" in sys.argv else main([a for a in sys.argv[1:] if a != "--fast"]))
