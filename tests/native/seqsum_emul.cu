// Host check of csrc/seqsum_core.cuh: the chunk-parallel evaluation of the reference's sequential
// f32 accumulator (processor.go:304-309) against the plain loop, bit for bit, on signals that
// stress it: discriminator-like noise, DC-heavy envelopes, values that tie at every step, sign
// changes, wide dynamic range, zeros.  Prints "mismatch <count> fast_fraction <min over cases>".
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <random>
#include <vector>

#include "seqsum_core.cuh"

using namespace tdoa::seqsum;

static long g_scans = 0, g_chunks = 0;

static float plain(const std::vector<float> &x)
{
    float s = 0.f;
    for (float v : x) { volatile float t = s + v; s = t; }
    return s;
}

// the device walk of k_seq_apply (seqsum.cu), lane by lane: batches of 16 chunks, an inclusive
// Hillis-Steele scan of the hops for the sum's binade and sign, the first chunk that is not
// admitted added sample by sample, the rest of the batch scanned again
static float walk_scan(const std::vector<ChunkRule> &rules, const std::vector<float> &x, long *scans)
{
    const int kBatch = 16, kLanes = 32;
    const int n = (int)x.size(), nc = (n + kChunk - 1) / kChunk;
    float s = 0.f;
    for (int c0 = 0; c0 < nc; c0 += kBatch) {
        const int chunks = std::min(kBatch, nc - c0);
        int pos = 0;
        while (pos < chunks) {
            int es;
            bool neg;
            const int m = mantissa_signed(s, &es, &neg);
            Hop f[kLanes];
            for (int lane = 0; lane < kLanes; lane++)
                f[lane] = lane < pos ? hop_identity()
                                     : (lane < chunks ? hop_from_rules(&rules[(size_t)(c0 + lane) * kGuesses], es, neg) : hop_empty());
            for (int off = 1; off < kBatch; off <<= 1) {
                Hop g[kLanes];
                for (int lane = 0; lane < kLanes; lane++) g[lane] = lane >= off ? hop_compose(f[lane - off], f[lane]) : f[lane];
                for (int lane = 0; lane < kLanes; lane++) f[lane] = g[lane];
            }
            (*scans)++;
            int upto = pos;
            while (upto < kLanes && (upto < pos || hop_admits(f[upto], m))) upto++;
            upto = std::min(upto, chunks);
            if (upto > pos) {
                s = hop_apply(f[upto - 1], s, m);
                pos = upto;
            }
            if (pos >= chunks) break;
            const int count = std::min(kChunk, n - (c0 + pos) * kChunk);
            for (int i = 0; i < count; i++) { volatile float t = s + x[(size_t)(c0 + pos) * kChunk + i]; s = t; }
            pos++;
        }
    }
    return s;
}

static float chunked(const std::vector<float> &x, double *fast_fraction)
{
    const int n = (int)x.size(), nc = (n + kChunk - 1) / kChunk;
    // guesses from exact f64 prefix sums (what the device scan produces)
    std::vector<double> pre(nc);
    double acc = 0.0;
    for (int c = 0; c < nc; c++) {
        pre[c] = acc;
        for (int i = c * kChunk; i < std::min(n, (c + 1) * kChunk); i++) acc += (double)x[i];
    }
    // kGuesses binades around the one the prefix sum suggests (the f32 chain and the exact sum may sit
    // on different sides of a power of two)
    std::vector<ChunkInfo> info((size_t)nc * kGuesses);
    for (int c = 0; c < nc; c++)
        for (int k = 0; k < kGuesses; k++)
            info[(size_t)c * kGuesses + k] = chunk_analyse(x.data() + (size_t)c * kChunk, std::min(kChunk, n - c * kChunk),
                                                           guess_binade(pre[c], k));
    float s = 0.f;
    int fast = 0;
    // the device walk uses the packed 32-bit rules: run those here, and the 64-bit summaries beside them
    std::vector<ChunkRule> rules((size_t)nc * kGuesses);
    for (size_t k = 0; k < rules.size(); k++) rules[k] = chunk_rule(info[k]);
    float s2 = 0.f;
    for (int c = 0; c < nc; c++) {
        const int count = std::min(kChunk, n - c * kChunk);
        s = chunk_apply(s, &info[(size_t)c * kGuesses], x.data() + (size_t)c * kChunk, count, &fast);
        int done = 0;
        s2 = rule_apply(s2, &rules[(size_t)c * kGuesses], &done);
        if (!done)
            for (int i = 0; i < count; i++) { volatile float t = s2 + x[(size_t)c * kChunk + i]; s2 = t; }
    }
    if (f2u(s) != f2u(s2)) return std::nanf("");   // the two forms of the walk must agree
    long scans = 0;
    const float s3 = walk_scan(rules, x, &scans);
    if (f2u(s) != f2u(s3)) return std::nanf("");   // and so must the scanned walk the device runs
    g_scans += scans; g_chunks += nc;
    *fast_fraction = nc ? (double)fast / nc : 1.0;
    return s;
}

int main()
{
    std::mt19937_64 g(7);
    std::normal_distribution<double> nrm(0.0, 1.0);
    std::uniform_real_distribution<double> uni(-1.0, 1.0);
    long mism = 0;
    double min_fast_typical = 1.0;
    int cases = 0;
    auto run = [&](const std::vector<float> &x, bool typical) {
        double ff;
        const float a = plain(x), b = chunked(x, &ff);
        if (f2u(a) != f2u(b)) { mism++; printf("case %d: plain %.9g chunked %.9g\n", cases, a, b); }
        if (typical && ff < min_fast_typical) min_fast_typical = ff;
        if (cases < 14) printf("case %d n %zu fast %.4f\n", cases, x.size(), ff);
        cases++;
    };
    for (int rep = 0; rep < 6; rep++) {
        const int n = rep < 4 ? 300000 + 777 * rep : 1000000 + rep;
        std::vector<float> x(n);
        // discriminator-like: zero-mean noise, sigma 0.15 (the running sum random-walks across binades)
        for (auto &v : x) v = (float)(0.15 * nrm(g));
        run(x, false);
        // envelope-like: DC 0.5 + ripple (the sum grows steadily: almost every chunk is fast)
        for (auto &v : x) v = (float)(0.5 + 0.05 * nrm(g));
        run(x, true);
        // negative DC
        for (auto &v : x) v = (float)(-0.3 + 0.1 * nrm(g));
        run(x, true);
        // coarse values: multiples of 2^-6 -- ties at every step once the sum is large
        for (auto &v : x) v = (float)(std::floor(uni(g) * 64.0 + 40.0) / 64.0);
        run(x, true);
        // alternating ties: +-(odd multiple of 2^-12) around a large DC
        for (int i = 0; i < n; i++) x[i] = (float)(3.0 + ((i & 1) ? 1 : -1) * (2 * (i % 5) + 1) * 0.000244140625);
        run(x, true);
        // wide dynamic range with sign changes and zeros
        for (int i = 0; i < n; i++) {
            const double m = std::pow(10.0, 4.0 * uni(g));
            x[i] = (i % 97 == 0) ? 0.f : (float)(uni(g) * m);
        }
        run(x, false);
        // slow sign change of the sum: drifts up, then down through zero
        for (int i = 0; i < n; i++) x[i] = (float)((i < n / 2 ? 0.01 : -0.0201) + 0.001 * nrm(g));
        run(x, false);
    }
    // tiny inputs
    for (int n : {0, 1, 2, 255, 256, 257, 513}) {
        std::vector<float> x(n);
        for (auto &v : x) v = (float)(0.4 + 0.2 * nrm(g));
        run(x, false);
    }
    printf("scans %ld chunks %ld\n", g_scans, g_chunks);
    printf("mismatch %ld fast_fraction %.4f cases %d\n", mism, min_fast_typical, cases);
    return mism ? 1 : 0;
}
