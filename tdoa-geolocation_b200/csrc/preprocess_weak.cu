// preprocess_weak.cu -- the weak-signal branch of the shipped binary's preprocessing in EXTENDED mode,
// as two passes over the capture bytes instead of four kernels over complex f32 planes.
//
// The branch (ELF 0x49cd40, "weak": initial power <= 0.001) is removeDCBias (processor.go:299-319) ->
// applyBandpassFilter(100 Hz, 200 kHz) = 1001-tap high-pass then 5-tap low-pass (:354-394, :270-296) ->
// normalizeSignal (:336-351).  The generic path runs k_power (2 B/sample), k_unpack (2 + 8),
// k_boxcar_slide (8 + 8) and k_boxcar (8 + 8): 44 B per sample at 0.5-1.8 TB/s, 47 ms of the 66 ms that
// BASELINE configs[3] takes on weak_signal_simulator.go content.  Here:
//
//   k_raw_stats   one read of the bytes: initial power (selects the branch) AND the two DC sums     2 B/sample
//   k_weak_fused  bytes -> (x - dc) -> high-pass -> low-pass -> real plane + power of both parts   2 + 4 B/sample
//
// The correlators of the binary's revision read real parts only (ELF 0x49d6a0; CORR_BINARY / CORR_EXTENDED),
// so the imaginary plane is written only when a caller asks for it (the tdoa_preprocess probe, the decimator);
// its power still enters the normalisation, as in the reference.
//
// Arithmetic: EXTENDED mode's own statement of the chain (oracle: orc_preprocess_binary with
// orc_set_wide_boxcar_f64) -- the 1001-tap window sum is the difference of two entries of an f64 prefix sum
// of the staged tile, rounded to f32 once after the divide, as k_boxcar_slide does it (full windows multiply by
// RN(1 / taps) instead: an ulp of f64 from the quotient, 2^-29 of the f32 step, far inside what the prefix sums'
// own rounding already moves); the 5-tap low-pass is the reference's sequential f32 sum and f32 divide, bit for
// bit; DC and power as in k_unpack / k_boxcar.  Against the oracle's statement: at most a handful of samples per
// 10^5 differ in their last f32 bit (tests/test_gpu_parity.py).
// BINARY mode keeps the tap-by-tap kernels: only that walk reproduces the reference's f32 rounding chain.
#include "kernels.h"

namespace tdoa {

namespace {

// ---------------------------------------------------------------- byte -> f32 without a divide
// processor.go:198-199 (f32(b) - 127.5) / 127.5 with a true f32 division.  q = a * RN(1/127.5), one FMA
// residual, one FMA correction (Markstein): checked against the divide for all 256 codes by the engine's
// unpack self-test at creation (weak_unpack_selftest).
__device__ __forceinline__ float unpack_fast(unsigned b)
{
    constexpr float d = 127.5f, r = 1.0f / 127.5f;
    const float a = __fsub_rn((float)b, d);
    const float q = __fmul_rn(a, r);
    const float e = __fmaf_rn(-q, d, a);
    return __fmaf_rn(e, r, q);
}

__global__ void k_weak_unpack_selftest(int *bad)
{
    const unsigned b = threadIdx.x;
    if (__float_as_uint(unpack_fast(b)) != __float_as_uint(unpack_byte(b))) atomicAdd(bad, 1);
}

// x / 5 as in preprocess_fast.cu's div_small (exhaustively proven equal to __fdiv_rn by tdoa_selftest(e, 0))
__device__ __forceinline__ float div5(float x)
{
    constexpr float r = 1.0f / 5.0f;
    const float q = __fmul_rn(x, r);
    const float e = __fmaf_rn(-q, 5.0f, x);
    return __fmaf_rn(e, r, q);
}

// ---------------------------------------------------------------- statistics of the raw signal
constexpr int kStatThreads = 256;

struct RawAcc { double pw, sr, si; };

__device__ __forceinline__ void raw_acc_sample(RawAcc &a, unsigned bi, unsigned bq)
{
    const float x = unpack_fast(bi), y = unpack_fast(bq);
    a.pw += (double)mag2_f32(x, y);   // processor.go:328
    a.sr += (double)x;
    a.si += (double)y;
}

__device__ __forceinline__ void raw_acc_word(RawAcc &a, unsigned w)
{
    raw_acc_sample(a, w & 0xffu, (w >> 8) & 0xffu);
    raw_acc_sample(a, (w >> 16) & 0xffu, w >> 24);
}

// One run of the view: `len` samples from raw sample `rs`.  16-byte loads over the aligned body (8 samples
// per load, grid-stride), the unaligned head and tail sample by sample in CTA 0.
__device__ __forceinline__ void raw_acc_run(RawAcc &a, const uint8_t *__restrict__ raw, i64 rs, i64 len)
{
    if (len <= 0) return;
    const uint8_t *p = raw + 2 * rs;
    i64 head = (i64)((16 - (reinterpret_cast<uintptr_t>(p) & 15)) & 15) / 2;
    if (head > len) head = len;
    const i64 units = (len - head) / 8;
    const i64 tail0 = head + 8 * units;
    const uint4 *__restrict__ body = reinterpret_cast<const uint4 *>(p + 2 * head);
    const i64 stride = (i64)gridDim.x * kStatThreads;
    i64 u = (i64)blockIdx.x * kStatThreads + threadIdx.x;
    for (; u + 3 * stride < units; u += 4 * stride) {   // four 16-byte loads in flight per thread
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = __ldg(body + u + k * stride);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            raw_acc_word(a, v[k].x);
            raw_acc_word(a, v[k].y);
            raw_acc_word(a, v[k].z);
            raw_acc_word(a, v[k].w);
        }
    }
    for (; u < units; u += stride) {
        const uint4 v = __ldg(body + u);
        raw_acc_word(a, v.x);
        raw_acc_word(a, v.y);
        raw_acc_word(a, v.z);
        raw_acc_word(a, v.w);
    }
    if (blockIdx.x == 0) {
        const uchar2 *__restrict__ s2 = reinterpret_cast<const uchar2 *>(p);
        if ((i64)threadIdx.x < head) {
            const uchar2 v = s2[threadIdx.x];
            raw_acc_sample(a, v.x, v.y);
        }
        const i64 t = tail0 + ((i64)threadIdx.x - 32);
        if (threadIdx.x >= 32 && t < len) {
            const uchar2 v = s2[t];
            raw_acc_sample(a, v.x, v.y);
        }
    }
}

__global__ void __launch_bounds__(kStatThreads) k_raw_stats(const SigJob *jobs)
{
    __shared__ double scratch[32];
    const SigJob &J = jobs[blockIdx.y];
    RawAcc a{0.0, 0.0, 0.0};
    if (J.src.raw) {
        const i64 n0 = min(J.n, J.src.run0_len);
        raw_acc_run(a, J.src.raw, J.src.run0_start, n0);
        raw_acc_run(a, J.src.raw, J.src.run1_start, J.n - n0);
    } else {
        const i64 stride = (i64)gridDim.x * kStatThreads;
        for (i64 i = (i64)blockIdx.x * kStatThreads + threadIdx.x; i < J.n; i += stride) {
            const float2 v = load_sample(J.src, i);
            a.pw += (double)mag2_f32(v.x, v.y);
            a.sr += (double)v.x;
            a.si += (double)v.y;
        }
    }
    double part[3], total[3];
    part[0] = block_sum(a.pw, scratch);
    part[1] = block_sum(a.sr, scratch);
    part[2] = block_sum(a.si, scratch);
    if (grid_sum_last<3>(part, J.partials, J.counter, gridDim.x, blockIdx.x, scratch, total)) {
        const i64 n = J.n;
        J.stats[ST_POWER0] = n > 0 ? total[0] / (double)n : 0.0;
        J.stats[ST_SUM_RE] = total[1];
        J.stats[ST_SUM_IM] = total[2];
        // processor.go:309: the f32 accumulator divided by f32(n); the accumulator is the exactly rounded sum
        J.stats[ST_DC_RE] = n > 0 ? (double)__fdiv_rn((float)total[1], (float)n) : 0.0;
        J.stats[ST_DC_IM] = n > 0 ? (double)__fdiv_rn((float)total[2], (float)n) : 0.0;
    }
}

// ---------------------------------------------------------------- the fused chain
constexpr int kWfTile = 6144;            // outputs per CTA
constexpr int kWfThreads = 512;
constexpr int kWfPer = 15;               // consecutive staged entries per thread in the scan (odd: conflict-free)
constexpr int kWfLen = kWfThreads * kWfPer;          // 7680 staged entries at most
constexpr int kWfOut = kWfTile / kWfThreads;         // 12 outputs per thread
constexpr int kWfSmem = (kWfLen + 2) * (int)sizeof(double) + kWfLen * (int)sizeof(float) + kWfLen;   // prefix sums, samples, Q bytes

}  // namespace

int weak_fused_max_half_wide() { return 500; }                                    // processor.go:404 clamps the window to 1000
int weak_fused_max_half_small() { return (kWfLen - kWfTile) / 2 - 500; }          // what the staged halo leaves: 268

namespace {

// s_p carries a leading zero: s_p[j + 1] = x[0] + ... + x[j], so a window sum is s_p[jb + 1] - s_p[ja] with no test.
// Tiles away from the ends of the signal (all but the first and the last of a signal) take loops without clamps.
__global__ void __launch_bounds__(kWfThreads, 2) k_weak_fused(const SigJob *jobs)
{
    extern __shared__ __align__(16) unsigned char wf_raw[];
    double *s_p = reinterpret_cast<double *>(wf_raw);            // exclusive-indexed prefix sums of s_x (kWfLen + 1)
    float *s_x = reinterpret_cast<float *>(s_p + kWfLen + 2);    // x - dc, then the high-pass output in place
    uint8_t *s_q = reinterpret_cast<uint8_t *>(s_x + kWfLen);
    __shared__ double s_w[kWfThreads / 32];
    __shared__ double scratch[32];
    const SigJob &J = jobs[blockIdx.y];
    const i64 n = J.n;
    const int h = J.window / 2;       // wide window (high-pass), >= 1
    const int h2 = J.window2 / 2;     // small window (low-pass), >= 1
    const i64 i0 = (i64)blockIdx.x * kWfTile;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double pacc = 0.0;
    if (i0 < n) {   // uniform per CTA
        const i64 iend = min(n, i0 + kWfTile);
        const i64 hp_lo = max((i64)0, i0 - h2), hp_hi = min(n, iend + h2);   // high-pass outputs the low-pass needs
        const i64 lo = max((i64)0, hp_lo - h), hi = min(n, hp_hi + h);       // samples those need
        // everything below in staged (tile-relative) 32-bit indices.  Clamping a window to [0, len) or to
        // [o_lo, o_hi) is clamping it to the signal: a staged range ends short of its halo only at an end of the signal.
        const int len = (int)(hi - lo);
        const int o_lo = (int)(hp_lo - lo), o_hi = (int)(hp_hi - lo);   // high-pass outputs
        const int t_lo = (int)(i0 - lo), t_n = (int)(iend - i0);        // the tile's outputs
        // no window of this tile is cut short, the tile is whole and the low-pass has the binary's five taps
        const bool inner = o_lo == h && len - o_hi == h && t_lo - o_lo == h2 && o_hi - (t_lo + t_n) == h2 &&
                           t_n == kWfTile && h2 == 2;
        const double rc_full = 1.0 / (double)(2 * h + 1);
        float *__restrict__ g_re = J.p_re + i0;
        float *__restrict__ g_im = J.p_im ? J.p_im + i0 : nullptr;
        if (tid == 0) s_p[0] = 0.0;
#pragma unroll 1
        for (int comp = 0; comp < 2; comp++) {
            const float dc = (float)J.stats[comp == 0 ? ST_DC_RE : ST_DC_IM];
            // ---- stage x - dc (processor.go:313-316)
            if (comp == 0) {
                const uchar2 *__restrict__ raw2 = reinterpret_cast<const uchar2 *>(J.src.raw);
                const i64 split64 = J.src.run0_len - lo;   // staged entries served by run 0
                const int split = (int)max((i64)0, min((i64)len, split64));
                const uchar2 *__restrict__ r0 = raw2 + (J.src.run0_start + lo);
                const uchar2 *__restrict__ r1 = raw2 + (J.src.run1_start - split64);
                // every load of the thread issued before the first use: one trip to memory per tile, not fifteen
                uchar2 v[kWfPer];
                if (split >= len || split <= 0) {   // one run serves the tile (all tiles but the one on the joint)
                    const uchar2 *__restrict__ rp = (split >= len ? r0 : r1) + tid;
#pragma unroll
                    for (int k = 0; k < kWfPer; k++) {
                        v[k] = make_uchar2(0, 0);
                        if (tid + k * kWfThreads < len) v[k] = rp[k * kWfThreads];
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < kWfPer; k++) {
                        const int j = tid + k * kWfThreads;
                        v[k] = make_uchar2(0, 0);
                        if (j < len) v[k] = j < split ? r0[j] : r1[j];
                    }
                }
#pragma unroll
                for (int k = 0; k < kWfPer; k++) {
                    const int j = tid + k * kWfThreads;
                    if (j < len) {
                        s_x[j] = __fsub_rn(unpack_fast(v[k].x), dc);
                        s_q[j] = v[k].y;
                    }
                }
            } else {
                __syncthreads();   // the low-pass of the real part has read s_x
#pragma unroll
                for (int k = 0; k < kWfPer; k++) {
                    const int j = tid + k * kWfThreads;
                    if (j < len) s_x[j] = __fsub_rn(unpack_fast(s_q[j]), dc);
                }
            }
            __syncthreads();
            // ---- block-wide scan in f64: a thread's 15 entries, the warp, the 16 warp totals
            {
                const int j0 = tid * kWfPer;
                const bool whole = j0 + kWfPer <= len;
                double run = 0.0;
                if (whole) {
#pragma unroll
                    for (int k = 0; k < kWfPer; k++) run += (double)s_x[j0 + k];
                } else {
#pragma unroll
                    for (int k = 0; k < kWfPer; k++) run += j0 + k < len ? (double)s_x[j0 + k] : 0.0;
                }
                double inc = run;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                if (lane == 31) s_w[wid] = inc;
                __syncthreads();
                if (wid == 0) {   // exclusive scan of the warp totals, fixed order
                    const double mine = lane < kWfThreads / 32 ? s_w[lane] : 0.0;
                    double w = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
                    if (lane < kWfThreads / 32) s_w[lane] = w - mine;
                }
                __syncthreads();
                // the thread's entries again, carried on from its exclusive prefix (no 15 doubles kept in registers)
                double acc = (inc - run) + s_w[wid];
                if (whole) {
#pragma unroll
                    for (int k = 0; k < kWfPer; k++) { acc += (double)s_x[j0 + k]; s_p[j0 + k + 1] = acc; }
                } else {
#pragma unroll
                    for (int k = 0; k < kWfPer; k++)
                        if (j0 + k < len) { acc += (double)s_x[j0 + k]; s_p[j0 + k + 1] = acc; }
                }
            }
            __syncthreads();
            // ---- high-pass in place: x - LP_wide(x) (processor.go:384-394), the window sum from the prefix sums
            if (inner) {
                const double *__restrict__ pa = s_p + (o_lo - h) + tid, *__restrict__ pb = s_p + (o_lo + h + 1) + tid;
                float *__restrict__ px = s_x + o_lo + tid;
#pragma unroll
                for (int k = 0; k < (kWfTile + 4 + kWfThreads - 1) / kWfThreads; k++) {
                    if (k * kWfThreads + kWfThreads <= kWfTile + 4 || tid + k * kWfThreads < kWfTile + 4) {
                        // the full window's sum times RN(1 / taps): within an ulp of the f64 quotient, 2^-29 of the f32 step
                        const double q = (pb[k * kWfThreads] - pa[k * kWfThreads]) * rc_full;
                        px[k * kWfThreads] = __fsub_rn(px[k * kWfThreads], (float)q);
                    }
                }
            } else {
                for (int j = o_lo + tid; j < o_hi; j += kWfThreads) {
                    const int ja = max(0, j - h), jb = min(len - 1, j + h), cnt = jb - ja + 1;
                    const double sum = s_p[jb + 1] - s_p[ja];
                    const double q = cnt == 2 * h + 1 ? sum * rc_full : sum / (double)cnt;
                    s_x[j] = __fsub_rn(s_x[j], (float)q);
                }
            }
            __syncthreads();
            // ---- low-pass (sequential f32 taps, ascending, f32 divide: processor.go:270-296) + outputs; the power
            // (processor.go:328) takes the real part's output back from the plane this thread wrote it to
            // (twelve registers fewer than carrying it across the imaginary part's pass)
            float *__restrict__ g = comp == 0 ? g_re : g_im;
            if (inner) {
                const float *__restrict__ px = s_x + (t_lo - 2) + tid;
                float out[kWfOut];
#pragma unroll
                for (int k = 0; k < kWfOut; k++) {
                    const float *q = px + k * kWfThreads;
                    float acc = q[0];
                    acc = __fadd_rn(acc, q[1]);
                    acc = __fadd_rn(acc, q[2]);
                    acc = __fadd_rn(acc, q[3]);
                    acc = __fadd_rn(acc, q[4]);
                    out[k] = div5(acc);
                }
                if (g) {
#pragma unroll
                    for (int k = 0; k < kWfOut; k++) g[tid + k * kWfThreads] = out[k];
                }
                if (comp == 1) {
                    float re[kWfOut];
#pragma unroll
                    for (int k = 0; k < kWfOut; k++) re[k] = g_re[tid + k * kWfThreads];
#pragma unroll
                    for (int k = 0; k < kWfOut; k++) pacc += (double)mag2_f32(re[k], out[k]);
                }
            } else {
                for (int t = tid; t < t_n; t += kWfThreads) {
                    const int j = t_lo + t;
                    const int ja = max(o_lo, j - h2), cnt = min(o_hi - 1, j + h2) - ja + 1;
                    float acc = 0.f;
                    for (int k = 0; k < cnt; k++) acc = __fadd_rn(acc, s_x[ja + k]);
                    const float out = cnt == 5 ? div5(acc) : __fdiv_rn(acc, (float)cnt);
                    if (g) g[t] = out;
                    if (comp == 1) pacc += (double)mag2_f32(g_re[t], out);
                }
            }
        }
    }
    double part[1] = {block_sum(pacc, scratch)}, total[1];
    if (grid_sum_last<1>(part, J.partials, J.counter, gridDim.x, blockIdx.x, scratch, total)) {
        const double pw = n > 0 ? total[0] / (double)n : 0.0;
        J.stats[ST_POWER1] = pw;
        // processor.go:343-345  scale = f32(1/sqrt(power)); power <= 0 leaves the signal alone
        J.stats[ST_SCALE] = pw > 0.0 ? (double)(float)(1.0 / sqrt(pw)) : 1.0;
    }
}

}  // namespace

int raw_stats_grid_x(i64 n)
{
    const i64 want = (n / 8 + kStatThreads * 4 - 1) / (kStatThreads * 4);
    const i64 cap = 148 * 4;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

void launch_raw_stats(const SigJob *d_jobs, int n_jobs, i64 max_n, cudaStream_t st)
{
    k_raw_stats<<<dim3(raw_stats_grid_x(max_n), n_jobs), kStatThreads, 0, st>>>(d_jobs);
}

void launch_weak_fused(const SigJob *d_jobs, int n_jobs, i64 max_n, cudaStream_t st)
{
    static bool opted_in[64] = {false};   // per device: more than 48 KB of dynamic shared memory
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !opted_in[dev]) {
        cudaFuncSetAttribute(k_weak_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, kWfSmem);
        opted_in[dev] = true;
    }
    const i64 g = (max_n + kWfTile - 1) / kWfTile;
    k_weak_fused<<<dim3((unsigned)(g < 1 ? 1 : g), n_jobs), kWfThreads, kWfSmem, st>>>(d_jobs);
}

int weak_unpack_selftest(cudaStream_t st)
{
    int *d_bad = nullptr, h_bad = -1;
    if (cudaMalloc(&d_bad, sizeof(int)) != cudaSuccess) return -1;
    cudaMemsetAsync(d_bad, 0, sizeof(int), st);
    k_weak_unpack_selftest<<<1, 256, 0, st>>>(d_bad);
    cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
    const cudaError_t err = cudaStreamSynchronize(st);
    cudaFree(d_bad);
    return err == cudaSuccess ? h_bad : -1;
}

}  // namespace tdoa
