// fast_analyzer_b200 -- the reference's `fast_analyzer` command (fast_analyzer.go:26-52) on the
// B200 engine: prints the two CSV lines gain_calibrator.go:266-297 parses,
//     REF,<snr dB>,<power dB>,<clipping>,<overload>
//     TGT,...
// The capture is streamed to the GPU (tdoa_load_file) and analysed there (tdoa_analyze,
// fast = 1: first 32768 samples of each block, 8192-point Hanning spectrum); no CPU path.
#include <cstdio>

#include "tdoa_b200.h"

static const char *gobool(int v) { return v ? "true" : "false"; }  // Go's %t

int main(int argc, char **argv)
{
    if (argc < 2) {
        printf("Usage: %s <data_file.dat>\n", argv[0]);
        printf("Fast signal quality analyzer for gain sweeps\n");
        return 1;
    }
    tdoa_config cfg;
    tdoa_default_config(TDOA_MODE_BINARY, &cfg);
    cfg.n_stations = 2;
    tdoa_engine *e = nullptr;
    if (tdoa_create(&e, &cfg) != TDOA_OK) {
        printf("Error: %s\n", tdoa_last_error(nullptr));
        return 1;
    }
    tdoa_signal_quality ref, tgt;
    int rc = tdoa_load_file(e, 0, argv[1], nullptr);
    if (rc == TDOA_OK) rc = tdoa_analyze(e, 0, 1, &ref, &tgt);
    if (rc != TDOA_OK) {
        printf("Error: %s\n", tdoa_last_error(e));  // fast_analyzer.go:38-41
        tdoa_destroy(e);
        return 1;
    }
    printf("REF,%.1f,%.1f,%s,%s\n", ref.snr_db, ref.power_db, gobool(ref.has_clipping), gobool(ref.has_overload));
    printf("TGT,%.1f,%.1f,%s,%s\n", tgt.snr_db, tgt.power_db, gobool(tgt.has_clipping), gobool(tgt.has_overload));
    tdoa_destroy(e);
    return 0;
}
