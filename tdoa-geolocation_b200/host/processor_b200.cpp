// processor_b200 -- the reference's `processor` command on the B200 engine.
//
//   processor_b200 [--source] <ref_freq_hz> <target_freq_hz> <csv_file> <dat_file1> [dat_file2] [dat_file3] ...
//
// Host-side mirror of processor.go in the reference's own shape (Go is not installed in this
// image, the reference is compiled code, so the mirror is C++): a TDOAProcessor with the
// reference's methods -- loadStations (:52), getStationFromFilename (:110), loadIQData
// (:166), ProcessTDOA (:739), solveTDOA (:932) -- same argument meaning, same error texts,
// same stdout, line for line (the default follows the shipped binary's revision, --source
// follows processor.go as committed).  Every numeric step is a call into the C ABI of
// libtdoa_b200.so (include/tdoa_b200.h); there is no arithmetic on samples here and no CPU
// fallback: without a B200 tdoa_create fails and so does this program.
#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <sys/stat.h>
#include <unistd.h>

#include "tdoa_b200.h"

namespace {

constexpr double kSpeedOfLight = 299792458.0;  // processor.go:899

std::string fmt(const char *f, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof(buf), f, ap);
    va_end(ap);
    return buf;
}

// Go's text for an errno ("no such file or directory"): the C library's, first letter lower case
std::string goStrerror(int e)
{
    std::string m = strerror(e);
    if (!m.empty()) m[0] = (char)tolower((unsigned char)m[0]);
    return m;
}

// strconv.ParseFloat(s, 64) as far as the command line needs it: the whole string must be a number
bool parseFloat(const std::string &s, double *out, std::string *err)
{
    errno = 0;
    char *end = nullptr;
    const double v = strtod(s.c_str(), &end);
    const bool all = !s.empty() && end == s.c_str() + s.size() && !isspace((unsigned char)s[0]);
    if (!all) { *err = "strconv.ParseFloat: parsing \"" + s + "\": invalid syntax"; return false; }
    if (errno == ERANGE && std::isinf(v)) { *err = "strconv.ParseFloat: parsing \"" + s + "\": value out of range"; return false; }
    *out = v;
    return true;
}

// encoding/csv's ReadAll as far as a station table needs it: records of comma-separated fields,
// "quoted" fields with "" for a quote, blank lines skipped, every record as long as the first
// (FieldsPerRecord = 0) -- with the package's error text otherwise
std::vector<std::vector<std::string>> readCsvAll(std::istream &f)
{
    std::vector<std::vector<std::string>> records;
    std::string line;
    int lineNo = 0;
    while (std::getline(f, line)) {
        lineNo++;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        std::vector<std::string> rec;
        std::string cell;
        size_t i = 0;
        while (true) {
            cell.clear();
            if (i < line.size() && line[i] == '"') {
                i++;
                while (true) {
                    if (i >= line.size()) throw std::runtime_error(fmt("failed to read CSV: record on line %d; parse error on line %d, column %zu: extraneous or missing \" in quoted-field", lineNo, lineNo, i + 1));
                    if (line[i] == '"') {
                        if (i + 1 < line.size() && line[i + 1] == '"') { cell += '"'; i += 2; continue; }
                        i++;
                        break;
                    }
                    cell += line[i++];
                }
                if (i < line.size() && line[i] != ',')
                    throw std::runtime_error(fmt("failed to read CSV: parse error on line %d, column %zu: extraneous or missing \" in quoted-field", lineNo, i + 1));
            } else {
                while (i < line.size() && line[i] != ',') {
                    if (line[i] == '"') throw std::runtime_error(fmt("failed to read CSV: parse error on line %d, column %zu: bare \" in non-quoted-field", lineNo, i + 1));
                    cell += line[i++];
                }
            }
            rec.push_back(cell);
            if (i >= line.size()) break;
            i++;   // the comma
            if (i == line.size()) { rec.push_back(""); break; }
        }
        if (!records.empty() && rec.size() != records[0].size())
            throw std::runtime_error(fmt("failed to read CSV: record on line %d: wrong number of fields", lineNo));
        records.push_back(rec);
    }
    return records;
}

struct Station {  // processor.go:15-20
    std::string name;
    double latitude = 0, longitude = 0, elevation = 0;
};

class TDOAProcessor {  // processor.go:29-33
public:
    TDOAProcessor(double refFreq, double targetFreq, const std::string &csv, int mode)
        : referenceFreq(refFreq), targetFreq(targetFreq), mode(mode)
    {
        loadStations(csv);
    }
    ~TDOAProcessor() { if (engine) tdoa_destroy(engine); }

    // processor.go:52-107
    void loadStations(const std::string &csv)
    {
        std::ifstream f(csv);
        if (!f) throw std::runtime_error("failed to load stations: failed to open CSV file: open " + csv + ": " + goStrerror(errno));
        std::vector<std::vector<std::string>> records;
        try { records = readCsvAll(f); }
        catch (const std::exception &e) { throw std::runtime_error(std::string("failed to load stations: ") + e.what()); }
        bool haveRef = false;
        const std::string refName = fmt("%.0f", referenceFreq);  // :96
        for (size_t i = 1; i < records.size(); i++) {            // header skipped (:66); message lines are i + 2 of records[1:]
            const std::vector<std::string> &rec = records[i];
            const int lineNo = (int)i + 1;
            if (rec.size() != 4) throw std::runtime_error(fmt("failed to load stations: invalid CSV format at line %d", lineNo));
            Station st;
            st.name = rec[0];
            std::string err;
            if (!parseFloat(rec[1], &st.latitude, &err)) throw std::runtime_error(fmt("failed to load stations: invalid latitude at line %d: %s", lineNo, err.c_str()));
            if (!parseFloat(rec[2], &st.longitude, &err)) throw std::runtime_error(fmt("failed to load stations: invalid longitude at line %d: %s", lineNo, err.c_str()));
            if (!parseFloat(rec[3], &st.elevation, &err)) throw std::runtime_error(fmt("failed to load stations: invalid elevation at line %d: %s", lineNo, err.c_str()));
            stations[st.name] = st;
            if (st.name == refName) { refStation = st; haveRef = true; }
        }
        if (!haveRef)
            throw std::runtime_error(fmt("failed to load stations: reference frequency %.0f not found in stations", referenceFreq));
        printf("Loaded %zu stations including reference %.0f MHz\n", stations.size(), referenceFreq / 1e6);
    }

    // processor.go:110-122.  The reference ranges over a Go map (random order); nested names
    // are resolved here by preferring the longest name.
    Station getStationFromFilename(const std::string &filename) const
    {
        const size_t slash = filename.find_last_of('/');
        const std::string base = slash == std::string::npos ? filename : filename.substr(slash + 1);
        std::vector<std::string> names;
        for (const auto &kv : stations) names.push_back(kv.first);
        std::sort(names.begin(), names.end(), [](const std::string &a, const std::string &b) {
            return a.size() != b.size() ? a.size() > b.size() : a < b;
        });
        for (const auto &n : names)
            if (base.find(n) != std::string::npos) return stations.at(n);
        throw std::runtime_error("could not identify station from filename: " + filename);
    }

    // processor.go:166-205: the bytes go to the GPU (tdoa_load_file); the complex64 samples stay there
    long long loadIQData(int slot, const std::string &filename)
    {
        printf("Loading I/Q data from: %s\n", filename.c_str());
        // the reference prints the size between opening and reading (:176-183); a file that cannot
        // be opened is reported by tdoa_load_file with the reference's text
        struct stat sb;
        const bool have = stat(filename.c_str(), &sb) == 0 && access(filename.c_str(), R_OK) == 0;
        if (have) printf("File size: %lld bytes, samples: %lld\n", (long long)sb.st_size, (long long)sb.st_size / 2);
        int64_t n = 0;
        if (tdoa_load_file(engine, slot, filename.c_str(), &n) != TDOA_OK) throw std::runtime_error(tdoa_last_error(engine));
        if (!have) printf("File size: %lld bytes, samples: %lld\n", (long long)n * 2, (long long)n);
        printf("Successfully loaded %lld complex samples\n", (long long)n);
        return n;
    }

    // processor.go:739-929; by default every stdout line of the shipped binary
    void ProcessTDOA(const std::vector<std::string> &datFiles)
    {
        if (datFiles.size() < 3)
            throw std::runtime_error(fmt("need at least 3 collector stations, got %zu", datFiles.size()));  // :740-742
        const bool binary = mode == TDOA_MODE_BINARY;
        printf("Processing TDOA for target frequency %.3f MHz\n", targetFreq / 1e6);
        printf("Reference: %s at %.6f°, %.6f°, %.1fm\n", refStation.name.c_str(), refStation.latitude, refStation.longitude,
               refStation.elevation);
        const int S = (int)datFiles.size(), P = S * (S - 1) / 2;
        tdoa_config cfg;
        tdoa_default_config(mode, &cfg);
        cfg.n_stations = S;
        if (tdoa_create(&engine, &cfg) != TDOA_OK) throw std::runtime_error(tdoa_last_error(nullptr));
        std::vector<Station> st;
        for (int slot = 0; slot < S; slot++) {
            Station s;
            try { s = getStationFromFilename(datFiles[slot]); }
            catch (const std::exception &e) { throw std::runtime_error("failed to identify station for " + datFiles[slot] + ": " + e.what()); }
            long long n = 0;
            try { n = loadIQData(slot, datFiles[slot]); }
            catch (const std::exception &e) { throw std::runtime_error("failed to load data from " + datFiles[slot] + ": " + e.what()); }
            {
                const long long b = n / 3;  // processor.go:211-236, :244-265
                const long long nRef = n < 3 ? n : 2 * b, nTgt = n < 3 ? n : b;   // fewer than 3 samples: returned unchanged
                printf("Extracting reference signal from dual-frequency data\n");
                if (n < 3) printf("Warning: Data too small for dual-frequency extraction\n");
                else {
                    printf("Total samples: %lld, block size: %lld\n", n, b);
                    printf("Extracted %lld reference samples from blocks 1 and 3\n", 2 * b);
                }
                printf("Extracting target signal from dual-frequency data\n");
                if (n < 3) printf("Warning: Data too small for dual-frequency extraction\n");
                else {
                    printf("Total samples: %lld, block size: %lld\n", n, b);
                    printf("Extracted %lld target samples from block 2\n", b);
                }
                // processor.go:772-783 (testChunkSize 2 000 000; the shipped binary: 1 000 000): the
                // engine cuts both signals the same way (chunk_samples of the mode)
                const long long chunk = cfg.chunk_samples;
                if (nRef > chunk) printf("Using test chunk: %lld samples (%.1f ms)\n", chunk, (double)chunk / 2e6 * 1000);
                if (nTgt > chunk) printf("Using target test chunk: %lld samples (%.1f ms)\n", chunk, (double)chunk / 2e6 * 1000);
                printf("Coherent integration time: %.0f ms (expecting ~%.1f dB processing gain)\n", (double)chunk / 2e6 * 1000,
                       10 * std::log10((double)chunk / 100000));
            }
            st.push_back(s);
            printf("Loaded collector: %s at %.6f°, %.6f°, %.1fm\n", s.name.c_str(), s.latitude, s.longitude, s.elevation);
        }
        std::vector<double> llh;
        for (const auto &s : st) { llh.push_back(s.latitude); llh.push_back(s.longitude); llh.push_back(s.elevation); }
        printf("\nBaseline distances (3D):\n");
        std::vector<double> base(P);
        check(tdoa_baselines(engine, llh.data(), S, base.data()));
        std::vector<std::pair<int, int>> pairs;
        for (int i = 0; i < S; i++)
            for (int j = i + 1; j < S; j++) pairs.push_back({i, j});
        for (int p = 0; p < P; p++)
            printf("%s - %s: %.2f km\n", st[pairs[p].first].name.c_str(), st[pairs[p].second].name.c_str(), base[p] / 1000);
        const double fs = cfg.sample_rate;
        // the whole numeric path in one engine call: both pair loops, time and range differences, solveTDOA
        std::vector<tdoa_peak> peaks[2] = {std::vector<tdoa_peak>(P), std::vector<tdoa_peak>(P)};
        std::vector<double> dev_td(P), dev_rd(P);
        double fix[3] = {0, 0, 0};
        int32_t fixStatus = 0, fixIters = 0;
        check(tdoa_process(engine, llh.data(), peaks[0].data(), peaks[1].data(), dev_td.data(), dev_rd.data(), fix, &fixStatus,
                           &fixIters));
        std::vector<double> tds[2];
        for (int kind = 0; kind < 2; kind++) {
            const char *label = kind == TDOA_KIND_REF ? "REF" : "TGT";
            if (kind == TDOA_KIND_REF) {
                printf("\n=== REFERENCE SIGNAL CORRELATION TEST ===\n");
                // processor.go:813 prints the frequency as a literal; the shipped binary formats it
                if (binary) printf("Testing weak %.1f MHz NOAA weather signal:\n", referenceFreq / 1e6);
                else printf("Testing weak 162.4 MHz NOAA weather signal:\n");
            } else {
                printf("\n=== TARGET SIGNAL CORRELATION TEST ===\n");
                if (binary) printf("Testing strong %.1f MHz FM broadcast signal:\n", targetFreq / 1e6);
                else printf("Testing strong 92.3 MHz FM broadcast signal:\n");   // processor.go:833
            }
            std::vector<tdoa_signal_info> info(S);
            std::vector<double> first(P);
            check(tdoa_xcorr_info(engine, kind, info.data(), first.data()));
            for (int p = 0; p < P; p++) {
                const tdoa_peak &pk = peaks[kind][p];
                const double td = (double)pk.lag / fs;  // processor.go:821
                tds[kind].push_back(td);
                if (binary) printPairBinary(info[pairs[p].first], info[pairs[p].second], pk, first[p], fs, cfg);
                else printPairSource(info[pairs[p].first], info[pairs[p].second], pk, cfg);
                printf("%s %s - %s: delay=%d samples (%.3f μs), correlation=%.6f\n", label, st[pairs[p].first].name.c_str(),
                       st[pairs[p].second].name.c_str(), pk.lag, td * 1e6, pk.corr);
            }
        }
        std::vector<double> td(P);
        if (!binary) {
            td = tds[1];  // processor.go:853: target differences only
            printf("\n=== CORRELATION COMPARISON ===\n");   // :855-858
            printf("Reference signal (162.4 MHz): Generally weaker correlation\n");
            printf("Target signal (92.3 MHz): Should show stronger correlation\n");
            printf("Using target signal for TDOA calculation\n");
        } else {
            printf("\n=== REFERENCE SIGNAL SYNCHRONIZATION ===\n");
            printf("Using reference signal to synchronize collector timing...\n");
            for (int k = 0; k < P; k++) printf("Reference timing offset %d: %.3f μs\n", k, tds[0][k] * 1e6);
            printf("\n=== APPLYING TIMING CORRECTIONS TO TARGET SIGNAL ===\n");
            for (int k = 0; k < P; k++) {
                td[k] = tds[1][k] - tds[0][k];
                printf("Target delay %d: %.3f μs (raw) - %.3f μs (ref offset) = %.3f μs (corrected)\n", k, tds[1][k] * 1e6,
                       tds[0][k] * 1e6, td[k] * 1e6);
            }
            printf("\n=== CORRELATION COMPARISON ===\n");
            printf("Reference signal (%.1f MHz): Used for timing synchronization\n", referenceFreq / 1e6);
            printf("Target signal (%.1f MHz): Corrected with reference timing offsets\n", targetFreq / 1e6);
            printf("Using corrected target signal for TDOA calculation\n");
        }
        std::vector<double> rd(P);
        for (int k = 0; k < P; k++) rd[k] = td[k] * kSpeedOfLight;  // :899-903
        {
            printf("\nTDOA triangulation:\n");
            // both revisions print the first three here, whatever the number of pairs (:869-879)
            const std::vector<double> td3(td.begin(), td.begin() + 3), rd3(rd.begin(), rd.begin() + 3);
            printf("%s: %s\n", binary ? "Corrected time differences" : "Time differences", join(td3, 1e6, "%.3f μs").c_str());
            printf("%s: %s\n", binary ? "Corrected distance differences" : "Distance differences", join(rd3, 1.0, "%.1f m").c_str());
            printf("\nDiagnostic test with example delays:\n");  // processor.go:882-889
            printf("Simulating 10 μs, 5 μs, -3 μs delays...\n");
            const double us[3] = {10.0, 5.0, -3.0};
            for (int k = 0; k < 3; k++) printf("Test delay %d: %.1f μs → %.1f m\n", k + 1, us[k], us[k] * 1e-6 * kSpeedOfLight);
        }
        printf("\n=== TDOA GEOLOCATION ===\n");
        printf("Time differences (μs): ");
        for (double t : td) printf("%.3f ", t * 1e6);
        printf("\nRange differences (m): ");
        for (double r : rd) printf("%.1f ", r);
        printf("\n");
        if (binary) {
            // shipped binary: measurements beyond 1.2 x 17 km are dropped before its solver, which
            // then works only with exactly two left (ELF 0x4a0360)
            printf("Validating range differences against baseline distances...\n");
            const double limit = 20400.0;
            int valid = 0;
            for (int k = 0; k < P; k++) {
                if (std::fabs(rd[k]) <= limit) {
                    printf("VALID: Range difference %d: %.1fm (within ±%.1fm limit)\n", k, rd[k], limit);
                    valid++;
                } else {
                    printf("FILTERING OUT: Range difference %d: %.1fm exceeds expected maximum %.1fm\n", k, rd[k], limit);
                    printf("This measurement is unreliable and will be excluded\n");
                }
            }
            // the binary's own solveTDOA (ELF 0x4a0360) from here on: tdoa_solve_binary
            if (valid < 2)
                throw std::runtime_error(fmt("TDOA solution failed: insufficient valid measurements: only %d of %d range differences are reliable", valid, P));
            printf("Using %d of %d range difference measurements\n", valid, P);
            double e[3][3];
            for (int k = 0; k < 3; k++) ecef(st[k], e[k]);
            const double area = 0.5 * std::fabs((e[1][0] - e[0][0]) * (e[2][1] - e[0][1]) - (e[2][0] - e[0][0]) * (e[1][1] - e[0][1]));
            printf("Station geometry triangle area: %.1f m²\n", area);
            if (area < 1e7) {
                printf("WARNING: Poor station geometry (small triangle area)\n");
                printf("This may cause TDOA solution instability\n");
            }
            printf("Initial guess: %.6f°, %.6f°, %.1fm\n", (st[0].latitude + st[1].latitude + st[2].latitude) / 3,
                   (st[0].longitude + st[1].longitude + st[2].longitude) / 3, (st[0].elevation + st[1].elevation + st[2].elevation) / 3);
        }
        fflush(stdout);
        // the device did the same arithmetic (delay / fs, target - reference, * c) before its solveTDOA
        for (int k = 0; k < P; k++)
            if (rd[k] != dev_rd[k]) throw std::runtime_error(fmt("range difference %d: host %.17g, device %.17g", k, rd[k], dev_rd[k]));
        if (binary) {
            int32_t bStatus = 0, nValid = 0, nIter = 0, conv = 0;
            double trace[10][5];
            check(tdoa_solve_binary(engine, llh.data(), S, rd.data(), P, fix, &bStatus, &nValid, &nIter, &conv, &trace[0][0]));
            if (bStatus == 2) throw std::runtime_error("TDOA solution failed: no valid range difference measurements remain");
            for (int k = 0; k < nIter; k++) {
                const double *t = trace[k];
                printf("Iteration %d: det=%.2e, residuals=[%.1f, %.1f]\n", k, t[0], t[1], t[2]);
                const int code = (int)t[4];
                if (code == 1) printf("Large step detected (%.1fm) - limiting to %.1fm\n", t[3], 1000.0 * (1000.0 / t[3] * 0.7));
                if (code >= 2 || (bStatus == 3 && k == nIter - 1))
                    printf("Singular matrix detected (det=%.2e) - trying alternative approach\n", t[0]);
                if (code >= 2) printf("Using single equation approach (equation %d)\n", code - 1);
                if (bStatus == 3 && k == nIter - 1)
                    throw std::runtime_error(fmt("TDOA solution failed: singular Jacobian matrix at iteration %d (det=%.2e)", k, t[0]));
                if (k == 9) printf("Maximum iterations reached\n");
            }
            if (conv) printf("Converged after %d iterations\n", nIter);
        } else {
            // processor.go:957, :971, :998, :1013 (solveTDOA ran inside tdoa_process; fixIters = the
            // iteration it stopped at: converged, singular, or 10 = all ten steps taken)
            printf("Initial guess: %.6f°, %.6f°, %.1fm\n", (st[0].latitude + st[1].latitude + st[2].latitude) / 3.0,
                   (st[0].longitude + st[1].longitude + st[2].longitude) / 3.0, (st[0].elevation + st[1].elevation + st[2].elevation) / 3.0);
            if (fixStatus != 0)
                throw std::runtime_error(fmt("TDOA solution failed: singular Jacobian matrix at iteration %d", fixIters));  // :997-999, :920
            if (fixIters < 10) printf("Converged after %d iterations\n", fixIters);
            else printf("Maximum iterations reached\n");
        }
        printf("\n*** CALCULATED TRANSMITTER LOCATION ***\n");
        printf("Latitude:  %.6f°\n", fix[0]);
        printf("Longitude: %.6f°\n", fix[1]);
        printf("Elevation: %.1f m\n", fix[2]);
    }

private:
    void check(int rc) const
    {
        if (rc != TDOA_OK) throw std::runtime_error(tdoa_last_error(engine));
    }

    static std::string join(const std::vector<double> &v, double scale, const char *f)
    {
        std::string s;
        for (size_t k = 0; k < v.size(); k++) s += (k ? ", " : "") + fmt(f, v[k] * scale);
        return s;
    }

    // processor.go:125-148 latLonToECEF (WGS-84): host copy for the printed triangle area only
    static void ecef(const Station &s, double *out)
    {
        const double a = 6378137.0, f = 1.0 / 298.257223563, e2 = 2 * f - f * f;
        const double la = s.latitude * M_PI / 180.0, lo = s.longitude * M_PI / 180.0;
        const double n = a / std::sqrt(1 - e2 * std::sin(la) * std::sin(la));
        out[0] = (n + s.elevation) * std::cos(la) * std::cos(lo);
        out[1] = (n + s.elevation) * std::cos(la) * std::sin(lo);
        out[2] = (n * (1 - e2) + s.elevation) * std::sin(la);
    }

    // stdout of the shipped binary while it works on one pair (ELF 0x49cd40, 0x49d6a0)
    void printPairBinary(const tdoa_signal_info &s1, const tdoa_signal_info &s2, const tdoa_peak &pk, double firstCorr, double fs,
                         const tdoa_config &cfg) const
    {
        static const char *branchText[3] = {"Strong FM signal - using instantaneous frequency correlation approach",
                                            "Moderate signal - envelope correlation approach",
                                            "Weak signal - standard processing with timing preservation"};
        printf("=== Cross-Correlation Analysis ===\n");
        if (s1.n == 0 || s2.n == 0) {   // processor.go:622-625
            printf("Warning: Empty signals for correlation\n");
            return;
        }
        printf("\n--- Signal Preprocessing ---\n");
        const tdoa_signal_info *sg[2] = {&s1, &s2};
        for (int k = 0; k < 2; k++) {
            printf("Preprocessing Signal %d signal (%lld samples)\n", k + 1, (long long)sg[k]->n);
            printf("Initial signal power: %.9f\n", sg[k]->power0);
            printf("%s\n", branchText[sg[k]->branch]);
            printf("Removed DC bias: %.6f + %.6fi\n", sg[k]->dc_re, sg[k]->dc_im);
            if (sg[k]->branch == 2) printf("Bandpass filter: %.1f - %.1f Hz (at %.0f Hz sample rate)\n", 100.0, 200000.0, 2000000.0);
            if (sg[k]->power1 > 0) printf("Normalized signal power: %.6f → 1.000000\n", sg[k]->power1);   // processor.go:338-340, :349
        }
        printf("\n--- Time Domain Correlation ---\n");
        printf("Performing time domain correlation\n");
        const long long tl = std::min(s1.n, s2.n), sl = std::max(s1.n, s2.n);
        printf("Template: %lld samples, Signal: %lld samples\n", tl, sl);
        if (tl == sl) printf("Reduced template to %lld samples to allow %d sample delay search\n", tl - cfg.max_lag, cfg.max_lag);
        printf("Using coherent integration with %d-sample blocks\n", cfg.block_size);
        const long long nLags = tl == sl ? cfg.max_lag : std::max<long long>(1, std::min<long long>(cfg.max_lag, sl - tl));
        printf("Time domain progress: 0/%lld (coherent blocks: %d)\n", nLags, pk.n_blocks);
        printf("Time domain correlation: %.6f at delay %d samples\n", firstCorr, pk.first_lag);
        if (cfg.sanity_lag > 0 && pk.first_lag > cfg.sanity_lag) {
            printf("WARNING: Delay %d samples (%.1f μs) exceeds reasonable range for baseline distances\n", pk.first_lag,
                   pk.first_lag / fs * 1e6);
            printf("Maximum expected delay: 56.7 μs for 17 km baseline\n");
            printf("This suggests correlation algorithm found wrong peak\n");
            if (pk.flags & TDOA_PEAK_RESEARCHED)
                printf("Found better peak within reasonable range: delay=%d samples (%.1f μs), correlation=%.6f\n", pk.lag,
                       pk.lag / fs * 1e6, pk.corr);
        }
        printf("\n--- Result: Time Domain with Preprocessing ---\n");
        printf("Correlation: %.6f at delay %d samples\n", pk.corr, pk.lag);
    }

    // stdout of processor.go while it works on one pair (crossCorrelate :619-643, preprocessSignal
    // :469-499, enhanceWeakSignal :437-466, timeDomainCorrelation :646-736)
    void printPairSource(const tdoa_signal_info &s1, const tdoa_signal_info &s2, const tdoa_peak &pk, const tdoa_config &cfg) const
    {
        printf("=== Cross-Correlation Analysis ===\n");
        if (s1.n == 0 || s2.n == 0) {   // :622-625
            printf("Warning: Empty signals for correlation\n");
            return;
        }
        printf("\n--- Signal Preprocessing ---\n");
        const tdoa_signal_info *sg[2] = {&s1, &s2};
        for (int k = 0; k < 2; k++) {
            printf("Preprocessing Signal %d signal (%lld samples)\n", k + 1, (long long)sg[k]->n);
            printf("Initial signal power: %.9f\n", sg[k]->power0);
            const char *bp = "Bandpass filter: %.1f - %.1f Hz (at %.0f Hz sample rate)\n";
            if (sg[k]->branch == 2) {   // power < 0.001 (:476)
                printf("Detected very weak signal - applying aggressive filtering\n");
                printf("Enhancing weak signal: Signal %d\n", k + 1);
                printf("Removed DC bias: %.6f + %.6fi\n", sg[k]->dc_re, sg[k]->dc_im);
                printf(bp, 57.5, 62.5, 2000000.0);          // notch 60 / 5 (:446)
                printf(bp, 117.5, 122.5, 2000000.0);        // notch 120 / 5 (:447)
                printf(bp, 975000.0, 1000000.0, 2000000.0); // notch 1 MHz / 50 kHz, upper edge clamped to fs / 2 (:448, :420-422)
                printf(bp, 100.0, 40000.0, 2000000.0);      // :457
            } else {
                printf("Standard signal processing\n");
                printf("Removed DC bias: %.6f + %.6fi\n", sg[k]->dc_re, sg[k]->dc_im);
                printf(bp, 500.0, 50000.0, 2000000.0);      // :489
            }
            if (sg[k]->power1 > 0) printf("Normalized signal power: %.6f → 1.000000\n", sg[k]->power1);   // :338-340, :349
        }
        printf("\n--- Time Domain Correlation ---\n");
        printf("Performing time domain correlation\n");
        const long long tl = std::min(s1.n, s2.n), sl = std::max(s1.n, s2.n);
        printf("Template: %lld samples, Signal: %lld samples\n", tl, sl);
        long long maxLag = std::min<long long>(cfg.max_lag, sl - tl);   // :668-675
        if (maxLag < 1) maxLag = 1;
        printf("Using coherent integration with %d-sample blocks\n", cfg.block_size);
        // :729-731: every 2000th delay, carriage return instead of a new line
        for (long long d = 0; d < maxLag; d += 2000) printf("Time domain progress: %lld/%lld (coherent blocks: %d)\r", d, maxLag, pk.n_blocks);
        printf("\nTime domain correlation: %.6f at delay %d samples\n", pk.corr, pk.lag);
        printf("\n--- Result: Time Domain with Preprocessing ---\n");
        printf("Correlation: %.6f at delay %d samples\n", pk.corr, pk.lag);
    }

    double referenceFreq, targetFreq;
    int mode;
    std::map<std::string, Station> stations;
    Station refStation;
    tdoa_engine *engine = nullptr;
};

}  // namespace

// log.Fatalf: "2006/01/02 15:04:05 " + message on stderr, exit status 1
static int fatalf(const std::string &msg)
{
    fflush(stdout);
    char stamp[32];
    const time_t now = time(nullptr);
    struct tm tmv;
    localtime_r(&now, &tmv);
    strftime(stamp, sizeof(stamp), "%Y/%m/%d %H:%M:%S", &tmv);
    fprintf(stderr, "%s %s\n", stamp, msg.c_str());
    return 1;
}

// processor.go:1047-1076
int main(int argc, char **argv)
{
    int mode = TDOA_MODE_BINARY;
    std::vector<std::string> args(argv + 1, argv + argc);
    if (!args.empty() && args[0] == "--source") { mode = TDOA_MODE_SOURCE; args.erase(args.begin()); }
    if (args.size() < 4) {   // :1048-1052
        printf("Usage: %s <ref_freq_hz> <target_freq_hz> <csv_file> <dat_file1> [dat_file2] [dat_file3] ...\n", argv[0]);
        printf("Example: ./processor 162400000 101700000 lat-lon-table.csv kx0u-data.dat n3pay-data.dat kf0mtl-data.dat\n");
        return 1;
    }
    double refFreq = 0, tgtFreq = 0;
    std::string err;
    if (!parseFloat(args[0], &refFreq, &err)) return fatalf("Invalid reference frequency: " + err);   // :1054-1057
    if (!parseFloat(args[1], &tgtFreq, &err)) return fatalf("Invalid target frequency: " + err);      // :1059-1062
    std::unique_ptr<TDOAProcessor> p;
    try {
        p.reset(new TDOAProcessor(refFreq, tgtFreq, args[2], mode));
    } catch (const std::exception &e) {
        return fatalf(std::string("Failed to create processor: ") + e.what());   // :1067-1070
    }
    try {
        p->ProcessTDOA(std::vector<std::string>(args.begin() + 3, args.end()));
    } catch (const std::exception &e) {
        return fatalf(std::string("TDOA processing failed: ") + e.what());       // :1072-1075
    }
    return 0;
}
