// analyze.cu -- per-capture signal quality analysis on the GPU: the numeric core of the
// reference's fast_analyzer.go (gain sweeps: "REF,snr,power,clip,overload") and
// analyzer.go (full report).  Two kernels per call, batched over the REF and TGT signal
// of a capture:
//   k_quality_stats    byte statistics exactly as the reference accumulates them (sums of
//                      small integers: exact in u64, hence equal to its f64 sums), min/max,
//                      and the longest closed run of zero bytes (checkForDeadZones) as an
//                      associative (prefix, suffix, longest) summary per thread -> CTA -> grid;
//   k_quality_spectrum the windowed spectrum of the middle M samples.  The reference does an
//                      O(M^2) DFT in complex128 (fast_analyzer.go:229-252, analyzer.go:315-330);
//                      here a radix-2 f64 FFT when M is a power of two (M = 8192 / 16384 for
//                      any real capture), the direct sum otherwise; |X|^2, a bitonic sort in
//                      shared memory for the percentile thresholds, and the reference's
//                      signal / noise means.
// Everything in f64; differences from the reference are the summation order of the DFT
// (~1e-13 relative on a bin) -- far below the one decimal it prints.
#include "kernels.h"

namespace tdoa {

namespace {

constexpr int kQThreads = 256;
constexpr int kQTile = 8192;           // samples staged per CTA step (16 KB of bytes)
constexpr int kQPerThread = kQTile / kQThreads;  // 32 samples = 64 bytes per thread

// zero-byte runs of a byte range, closed runs only (a run is closed by a non-zero byte)
struct ZeroRuns {
    long long pre;   // leading zeros (the run continues to the left)
    long long suf;   // trailing zeros (still open)
    long long mx;    // longest closed run strictly inside
    long long len;   // bytes covered
    int full;        // every byte zero
};

__device__ __forceinline__ ZeroRuns zr_empty() { return ZeroRuns{0, 0, 0, 0, 1}; }

__device__ __forceinline__ ZeroRuns zr_join(const ZeroRuns &L, const ZeroRuns &R)
{
    ZeroRuns o;
    o.len = L.len + R.len;
    if (L.full && R.full) { o.full = 1; o.pre = o.suf = 0; o.mx = 0; return o; }
    o.full = 0;
    if (L.full) { o.pre = L.len + R.pre; o.suf = R.suf; o.mx = R.mx; }
    else if (R.full) { o.pre = L.pre; o.suf = L.suf + R.len; o.mx = L.mx; }
    else {
        o.pre = L.pre; o.suf = R.suf;
        const long long mid = L.suf + R.pre;
        o.mx = L.mx > R.mx ? L.mx : R.mx;
        if (mid > o.mx) o.mx = mid;
    }
    return o;
}

struct QualPart {
    unsigned long long isum, qsum, isq, qsq;
    int imin, imax, qmin, qmax;
    ZeroRuns zr;
};

__global__ void __launch_bounds__(kQThreads) k_quality_stats(const QualJob *jobs)
{
    __shared__ unsigned char s_b[2 * kQTile];
    __shared__ QualPart s_part[kQThreads];
    const QualJob &J = jobs[blockIdx.y];
    const int tid = threadIdx.x;
    const i64 n = J.n;
    // contiguous span of tiles per CTA, so the zero-run summaries join in order
    const i64 tiles = (n + kQTile - 1) / kQTile;
    const i64 per = (tiles + gridDim.x - 1) / gridDim.x;
    const i64 t_begin = (i64)blockIdx.x * per, t_end = min(tiles, t_begin + per);
    QualPart acc;
    acc.isum = acc.qsum = acc.isq = acc.qsq = 0;
    acc.imin = acc.qmin = 255; acc.imax = acc.qmax = 0;
    acc.zr = zr_empty();
    const uchar2 *raw = reinterpret_cast<const uchar2 *>(J.src.raw);
    for (i64 tile = t_begin; tile < t_end; tile++) {
        const i64 i0 = tile * kQTile;
        __syncthreads();
        for (int u = 0; u < kQPerThread; u++) {
            const int m = tid + kQThreads * u;
            const i64 i = i0 + m;
            uchar2 v = make_uchar2(1, 1);
            if (i < n) v = raw[raw_index(J.src, i)];
            reinterpret_cast<uchar2 *>(s_b)[m] = v;
        }
        __syncthreads();
        // this thread's 32 consecutive samples
        QualPart p;
        p.isum = p.qsum = p.isq = p.qsq = 0;
        p.imin = p.qmin = 255; p.imax = p.qmax = 0;
        ZeroRuns z = zr_empty();
        long long run = 0;
        bool seen_nonzero = false;
        const i64 first = i0 + (i64)kQPerThread * tid;
        const int cnt = (int)max((i64)0, min((i64)kQPerThread, n - first));
        for (int k = 0; k < 2 * cnt; k++) {
            const unsigned b = s_b[2 * kQPerThread * tid + k];
            if (k & 1) { p.qsum += b; p.qsq += b * b; p.qmin = min(p.qmin, (int)b); p.qmax = max(p.qmax, (int)b); }
            else { p.isum += b; p.isq += b * b; p.imin = min(p.imin, (int)b); p.imax = max(p.imax, (int)b); }
            if (b == 0) run++;
            else {
                if (!seen_nonzero) { z.pre = run; seen_nonzero = true; }
                else if (run > z.mx) z.mx = run;
                run = 0;
            }
        }
        z.len = 2 * cnt;
        z.full = seen_nonzero ? 0 : 1;
        z.suf = seen_nonzero ? run : 0;
        p.zr = z;
        s_part[tid] = p;
        __syncthreads();
        if (tid == 0) {
            for (int t = 0; t < kQThreads; t++) {
                const QualPart &q = s_part[t];
                acc.isum += q.isum; acc.qsum += q.qsum; acc.isq += q.isq; acc.qsq += q.qsq;
                acc.imin = min(acc.imin, q.imin); acc.imax = max(acc.imax, q.imax);
                acc.qmin = min(acc.qmin, q.qmin); acc.qmax = max(acc.qmax, q.qmax);
                acc.zr = zr_join(acc.zr, q.zr);
            }
        }
    }
    if (tid == 0) reinterpret_cast<QualPart *>(J.parts)[blockIdx.x] = acc;
}

__device__ __forceinline__ double block_sum_1024(double v, double *scratch)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double r = 0.0;
    if (wid == 0) {
        r = lane < (int)(blockDim.x >> 5) ? scratch[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    if (threadIdx.x == 0) scratch[0] = r;
    __syncthreads();
    r = scratch[0];
    __syncthreads();
    return r;
}

// one CTA (1024 threads) per signal
__global__ void __launch_bounds__(1024) k_quality_spectrum(const QualJob *jobs, int n_stat_cta)
{
    extern __shared__ double s_sorted[];   // next power of two >= M doubles
    __shared__ double s_red[32];
    const QualJob &J = jobs[blockIdx.x];
    const int tid = threadIdx.x;
    const i64 n = J.n;
    const int M = J.m;
    tdoa_signal_quality *out = reinterpret_cast<tdoa_signal_quality *>(J.out);
    // ---- statistics (fast_analyzer.go:137-151, analyzer.go:158-186)
    if (tid == 0) {
        const QualPart *parts = reinterpret_cast<const QualPart *>(J.parts);
        QualPart a = parts[0];
        for (int c = 1; c < n_stat_cta; c++) {
            const QualPart &q = parts[c];
            a.isum += q.isum; a.qsum += q.qsum; a.isq += q.isq; a.qsq += q.qsq;
            a.imin = min(a.imin, q.imin); a.imax = max(a.imax, q.imax);
            a.qmin = min(a.qmin, q.qmin); a.qmax = max(a.qmax, q.qmax);
            a.zr = zr_join(a.zr, q.zr);
        }
        tdoa_signal_quality r;
        memset(&r, 0, sizeof(r));
        r.total_samples = n;
        const double dn = (double)n;
        r.i_avg = (double)a.isum / dn; r.q_avg = (double)a.qsum / dn;
        r.i_std = sqrt(__dsub_rn(__ddiv_rn((double)a.isq, dn), __dmul_rn(r.i_avg, r.i_avg)));
        r.q_std = sqrt(__dsub_rn(__ddiv_rn((double)a.qsq, dn), __dmul_rn(r.q_avg, r.q_avg)));
        r.i_min = a.imin; r.i_max = a.imax; r.q_min = a.qmin; r.q_max = a.qmax;
        const double mag = sqrt(__dadd_rn(__dmul_rn(r.i_std, r.i_std), __dmul_rn(r.q_std, r.q_std)));
        r.power_db = (J.fast && mag <= 1e-10) ? -100.0 : 20.0 * log10(mag);
        r.has_clipping = a.imin == 0 || a.imax == 255 || a.qmin == 0 || a.qmax == 255;
        r.has_overload = r.i_std < 2.0 || r.q_std < 2.0;
        if (!J.fast) {
            const double di = r.i_avg - 127.5, dq = r.q_avg - 127.5;
            r.dc_offset = sqrt(__dadd_rn(__dmul_rn(di, di), __dmul_rn(dq, dq)));
            r.iq_imbalance = fabs(r.i_std - r.q_std) / fmax(r.i_std, r.q_std);
            const long long longest = a.zr.full ? 0 : (a.zr.pre > a.zr.mx ? a.zr.pre : a.zr.mx);
            r.has_dead_zones = longest > 1000;
            r.has_noise = r.i_std > 60.0 || r.q_std > 60.0;
        }
        r.snr_db = -20.0;
        *out = r;
    }
    if (M <= 0) return;
    // ---- the middle M samples, windowed (fast: Hanning on (b - 127.5) / 127.5; full: DC-corrected, Blackman-Harris)
    const uchar2 *raw = reinterpret_cast<const uchar2 *>(J.src.raw);
    const i64 start = (n - M) / 2;
    double idc = 127.5, qdc = 127.5;
    if (!J.fast) {
        double si = 0.0, sq = 0.0;
        for (int i = tid; i < M; i += blockDim.x) {
            const uchar2 v = raw[raw_index(J.src, start + i)];
            si += (double)v.x; sq += (double)v.y;   // integers: exact in any order
        }
        si = block_sum_1024(si, s_red);
        sq = block_sum_1024(sq, s_red);
        idc = si / (double)M; qdc = sq / (double)M;
    }
    double2 *a = reinterpret_cast<double2 *>(J.fft_a), *b = reinterpret_cast<double2 *>(J.fft_b);
    for (int i = tid; i < M; i += blockDim.x) {
        const uchar2 v = raw[raw_index(J.src, start + i)];
        const double iv = ((double)v.x - idc) / 127.5, qv = ((double)v.y - qdc) / 127.5;
        const double x = (double)i / (double)(M - 1);
        double w;
        if (J.fast) w = 0.5 - 0.5 * cospi(2.0 * x);
        else w = 0.35875 - 0.48829 * cospi(2.0 * x) + 0.14128 * cospi(4.0 * x) - 0.01168 * cospi(6.0 * x);
        a[i] = make_double2(w * iv, w * qv);
    }
    __syncthreads();
    // ---- spectrum
    const bool pow2 = (M & (M - 1)) == 0;
    double2 *X = a;
    if (pow2 && M >= 2) {
        // radix-2 Stockham, ping-pong between the two global scratch arrays
        for (int ns = 1; ns < M; ns <<= 1) {
            for (int j = tid; j < M / 2; j += blockDim.x) {
                const int k = j & (ns - 1);
                double s, c;
                sincospi(-(double)k / (double)ns, &s, &c);
                const double2 u = a[j], v = a[j + M / 2];
                const double2 t = make_double2(v.x * c - v.y * s, v.x * s + v.y * c);
                const int o = ((j - k) << 1) + k;
                b[o] = make_double2(u.x + t.x, u.y + t.y);
                b[o + ns] = make_double2(u.x - t.x, u.y - t.y);
            }
            __syncthreads();
            double2 *sw = a; a = b; b = sw;
        }
        X = a;
    } else if (M >= 2) {
        for (int k = tid; k < M; k += blockDim.x) {
            double sr = 0.0, si = 0.0;
            for (int i = 0; i < M; i++) {
                const long long idx = ((long long)k * i) % M;
                double s, c;
                sincospi(-2.0 * (double)idx / (double)M, &s, &c);
                const double2 v = a[i];
                sr += v.x * c - v.y * s;
                si += v.x * s + v.y * c;
            }
            b[k] = make_double2(sr, si);
        }
        __syncthreads();
        X = b;
    }
    // ---- |X|^2, sorted copy (bitonic, padded with +inf)
    int P2 = 1;
    while (P2 < M) P2 <<= 1;
    double *psd = J.psd;
    for (int k = tid; k < P2; k += blockDim.x) {
        double p = __longlong_as_double(0x7ff0000000000000LL);
        if (k < M) {
            const double m = hypot(X[k].x, X[k].y);   // cmplx.Abs
            p = m * m;
            psd[k] = p;
        }
        s_sorted[k] = p;
    }
    __syncthreads();
    for (int size = 2; size <= P2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < P2 / 2; t += blockDim.x) {
                const int lo = ((t / stride) * stride * 2) + (t % stride), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const double x = s_sorted[lo], y = s_sorted[hi];
                if ((x > y) == up) { s_sorted[lo] = y; s_sorted[hi] = x; }
            }
            __syncthreads();
        }
    // ---- signal / noise means (fast_analyzer.go:194-226, analyzer.go:237-271)
    const double sig_thr = s_sorted[(int)(0.9 * (double)M)];
    double sp = 0.0, np_ = 0.0, sc = 0.0, nc = 0.0;
    if (J.fast) {
        const double noise_thr = s_sorted[(int)(0.4 * (double)M)];
        for (int k = tid; k < M; k += blockDim.x) {
            const double p = psd[k];
            if (p >= sig_thr) { sp += p; sc += 1.0; }
            else if (p <= noise_thr) { np_ += p; nc += 1.0; }
        }
    } else {
        const int ne = (int)(0.5 * (double)M);
        for (int k = tid; k < M; k += blockDim.x) {
            const double p = psd[k];
            if (p >= sig_thr) { sp += p; sc += 1.0; }
            if (k < ne) np_ += s_sorted[k];
        }
        nc = (double)ne;
    }
    sp = block_sum_1024(sp, s_red); sc = block_sum_1024(sc, s_red);
    np_ = block_sum_1024(np_, s_red);
    if (J.fast) nc = block_sum_1024(nc, s_red);
    if (tid == 0) {
        if (sc > 0.0) sp /= sc;
        if (nc > 0.0) np_ /= nc;
        out->snr_db = (np_ > 0.0 && sp > np_) ? 10.0 * log10(sp / np_) : -20.0;
    }
}

}  // namespace

size_t quality_part_bytes() { return sizeof(QualPart); }

int quality_setup()
{
    return cudaFuncSetAttribute(k_quality_spectrum, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * (int)sizeof(double)) ==
                   cudaSuccess
               ? 0
               : -1;
}

void launch_quality(const QualJob *d_jobs, int n_jobs, int n_stat_cta, int max_m, cudaStream_t st)
{
    if (n_jobs <= 0) return;
    k_quality_stats<<<dim3(n_stat_cta, n_jobs), kQThreads, 0, st>>>(d_jobs);
    int p2 = 1;
    while (p2 < max_m) p2 <<= 1;
    k_quality_spectrum<<<n_jobs, 1024, (size_t)p2 * sizeof(double), st>>>(d_jobs, n_stat_cta);
}

}  // namespace tdoa
