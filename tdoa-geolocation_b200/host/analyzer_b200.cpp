// analyzer_b200 -- the reference's `analyzer` command (analyzer.go) on the B200 engine:
//
//   analyzer_b200 <data_file.dat> [expected_duration_seconds]
//
// File structure, per-signal statistics / quality metrics / flags (analyzer.go:350-371), the
// first line of each gain / SNR / summary verdict (:472-499, :618-627), the signal comparison
// and the TDOA suitability assessment (:398-470).  Every number is computed on the GPU
// (tdoa_load_file + tdoa_analyze, fast = 0: whole blocks, dead-zone scan, 16384-point
// DC-corrected Blackman-Harris spectrum); no CPU path.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "tdoa_b200.h"

namespace {

using Q = tdoa_signal_quality;

int countQualityIssues(const Q &a)  // analyzer.go:448-457
{
    return (a.has_clipping != 0) + (a.has_overload != 0) + (a.has_dead_zones != 0) + (a.has_noise != 0) + (a.dc_offset > 10) +
           (a.iq_imbalance > 0.1);
}

bool assessTDOASuitability(const Q &a)  // analyzer.go:459-470
{
    if (a.has_clipping || a.has_overload || a.has_dead_zones) return false;
    if (a.snr_db < 15) return false;
    return !(a.dc_offset > 15 || a.iq_imbalance > 0.15);
}

const char *gainVerdict(const Q &a)  // analyzer.go:472-499, first line of each case
{
    if (a.has_clipping) return "🔻 REDUCE GAIN: Signal clipping detected";
    if (a.has_overload) return "🔻 REDUCE GAIN: Signal appears overloaded";
    if (a.power_db < -60) return "🔺 INCREASE GAIN: Signal level very low";
    if (a.power_db < -40) return "🔺 INCREASE GAIN: Signal level low";
    if (a.i_std > 50 && a.q_std > 50) return "✅ GAIN OK: Good signal level, no clipping";
    return "🔧 FINE-TUNE GAIN: Signal usable but could be optimized";
}

const char *summaryVerdict(const Q &a)  // analyzer.go:618-627
{
    if (a.has_clipping || a.has_overload) return "❌ CRITICAL: Adjust gain immediately - signal distortion present";
    if (a.power_db < -50) return "⚠️  WARNING: Signal very weak - increase gain or check antenna";
    if (a.dc_offset > 10 || a.iq_imbalance > 0.1) return "🔧 HARDWARE: RTL-SDR calibration issues detected";
    return "✅ ACCEPTABLE: Signal quality adequate for TDOA processing";
}

void flag(const char *name, int v)
{
    if (v) printf("⚠️ %s: DETECTED\n", name);
    else printf("✅ %s: OK\n", name);
}

void printAnalysisResults(const Q &a)  // analyzer.go:350-371
{
    printf("=== Signal Statistics ===\n");
    printf("Total Samples: %lld\n", (long long)a.total_samples);
    printf("I Channel: min=%d, max=%d, avg=%.1f, σ=%.1f\n", a.i_min, a.i_max, a.i_avg, a.i_std);
    printf("Q Channel: min=%d, max=%d, avg=%.1f, σ=%.1f\n", a.q_min, a.q_max, a.q_avg, a.q_std);
    printf("\n=== Signal Quality Metrics ===\n");
    printf("DC Offset: %.1f (should be ~0)\n", a.dc_offset);
    printf("IQ Imbalance: %.3f (should be <0.1)\n", a.iq_imbalance);
    printf("Estimated SNR: %.1f dB\n", a.snr_db);
    printf("Power Level: %.1f dB\n", a.power_db);
    printf("\n=== Quality Flags ===\n");
    flag("Clipping/Saturation", a.has_clipping);
    flag("Overload (too low variation)", a.has_overload);
    flag("Dead zones detected", a.has_dead_zones);
    flag("Excessive noise", a.has_noise);
}

void compareSignals(const Q &ref, const Q &tgt)  // analyzer.go:398-446
{
    printf("\n=== SIGNAL COMPARISON ===\n");
    printf("SNR Comparison:\n");
    printf("  Reference: %.1f dB\n", ref.snr_db);
    printf("  Target:    %.1f dB\n", tgt.snr_db);
    if (ref.snr_db > tgt.snr_db + 10) printf("  ⚠️  Reference significantly stronger - consider reducing reference gain\n");
    else if (tgt.snr_db > ref.snr_db + 10) printf("  ⚠️  Target significantly stronger - consider reducing target gain\n");
    else printf("  ✅ Signal levels reasonably balanced\n");
    printf("\nPower Level Comparison:\n");
    printf("  Reference: %.1f dB\n", ref.power_db);
    printf("  Target:    %.1f dB\n", tgt.power_db);
    printf("\nQuality Issues:\n");
    const int ri = countQualityIssues(ref), ti = countQualityIssues(tgt);
    printf("  Reference: %d issues detected\n", ri);
    printf("  Target:    %d issues detected\n", ti);
    if (ri == 0 && ti == 0) printf("  ✅ Both signals appear suitable for TDOA processing\n");
    else if (ri > ti) printf("  ⚠️  Reference signal needs more attention\n");
    else if (ti > ri) printf("  ⚠️  Target signal needs more attention\n");
    printf("\n=== TDOA SUITABILITY ASSESSMENT ===\n");
    const bool rs = assessTDOASuitability(ref), ts = assessTDOASuitability(tgt);
    if (rs && ts) printf("✅ EXCELLENT: Both signals suitable for TDOA correlation\n");
    else if (!rs && !ts) printf("❌ POOR: Both signals need improvement before TDOA processing\n");
    else if (!rs) printf("⚠️  MARGINAL: Reference signal needs improvement\n");
    else printf("⚠️  MARGINAL: Target signal needs improvement\n");
}

}  // namespace

int main(int argc, char **argv)
{
    if (argc < 2) {
        printf("Usage: analyzer <data_file.dat> [expected_duration_seconds]\n");
        return 1;
    }
    int duration = 30;
    if (argc > 2) {
        char *end = nullptr;
        const long v = strtol(argv[2], &end, 10);
        if (end && *end == '\0' && end != argv[2]) duration = (int)v;
    }
    printf("=== Advanced Signal Quality Analysis ===\n");
    printf("File: %s\n", argv[1]);
    printf("Expected Duration: %d seconds\n\n", duration);
    tdoa_config cfg;
    tdoa_default_config(TDOA_MODE_BINARY, &cfg);
    cfg.n_stations = 2;
    tdoa_engine *e = nullptr;
    if (tdoa_create(&e, &cfg) != TDOA_OK) {
        printf("Error analyzing file: %s\n", tdoa_last_error(nullptr));
        return 1;
    }
    Q ref, tgt;
    int64_t total = 0;
    int rc = tdoa_load_file(e, 0, argv[1], &total);
    if (rc == TDOA_OK) rc = tdoa_analyze(e, 0, 0, &ref, &tgt);
    if (rc != TDOA_OK) {
        printf("Error analyzing file: %s\n", tdoa_last_error(e));
        tdoa_destroy(e);
        return 1;
    }
    tdoa_destroy(e);
    printf("=== File Structure Analysis ===\n");
    printf("Total samples: %lld\n", (long long)total);
    printf("Samples per frequency block: %lld\n", (long long)(total / 3));
    printf("Reference samples: %lld (blocks 1+3)\n", (long long)ref.total_samples);
    printf("Target samples: %lld (block 2)\n\n", (long long)tgt.total_samples);
    const struct { const char *title, *label; const Q *a; } parts[2] = {{"REFERENCE", "Reference", &ref}, {"TARGET", "Target", &tgt}};
    for (int k = 0; k < 2; k++) {
        const Q &a = *parts[k].a;
        printf("%s=== %s SIGNAL ANALYSIS ===\n", k == 0 ? "" : "\n", parts[k].title);
        printAnalysisResults(a);
        printf("\n=== %s SIGNAL RECOMMENDATIONS ===\n", parts[k].label);
        printf("\n--- Gain Recommendations ---\n");
        printf("%s\n", gainVerdict(a));
        if (a.snr_db < 10) printf("📡 SNR TOO LOW (%.1f dB): Increase gain or improve antenna\n", a.snr_db);
        else if (a.snr_db > 40) printf("📡 SNR HIGH (%.1f dB): Consider reducing gain to prevent overload\n", a.snr_db);
        printf("\n=== SUMMARY ===\n");
        printf("%s\n", summaryVerdict(a));
    }
    compareSignals(ref, tgt);
    return 0;
}
