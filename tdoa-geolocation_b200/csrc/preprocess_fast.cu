// preprocess_fast.cu -- the hot "strong FM" branch of preprocessSignal (shipped binary, ELF
// 0x49cd40) as two streaming kernels:
//
//   k_demod_fused : uint8 IQ -> [unpack (processor.go:198-199)] -> initial power
//                   (calculateSignalPower, :322-333) + FM discriminator (ELF 0x49d120) +
//                   DC sum, one pass: 2 B/sample in, 4 B/sample out.
//   k_boxcar_small: removeDCBias subtract (:313-316) + box-car low-pass (:270-296, window
//                   <= 17) + pre-normalise power (:336-351): 4 B in, 4 B out.
//
// Both keep the reference's arithmetic: f32 unpack via table, f64 products rounded once
// to f32, f64 arctangent rounded to f32, sequential f32 tap sums in ascending order.
#include "atan2_core.cuh"
#include "kernels.h"

namespace tdoa {

namespace {

constexpr int kThreads = 256;
constexpr int kTile = 2048;  // samples per CTA step (8 per thread)

__device__ __forceinline__ float dc_from_sum(double sum, i64 n) { return __fdiv_rn((float)sum, (float)n); }

// f32 arctangent for the fast_demod path: same octant/table reduction, 2-term polynomial
__device__ __forceinline__ float atan2_fast(float y, float x, const float *table)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float q = __fdividef(mn, mx);
    int k = (int)(q * 8.0f + 0.5f);
    k = min(max(k, 0), 8);
    const float c = (float)k * 0.125f;
    const float z = __fdividef(fmaf(-c, mx, mn), fmaf(c, mn, mx));
    const float w = z * z;
    const float p = fmaf(w, 0.2f, -0.33333333333f);
    float r = table[k] + fmaf(z * w, p, z);
    if (ay > ax) r = 1.57079632679489661923f - r;
    if (x < 0.f) r = 3.14159265358979323846f - r;
    return y < 0.f ? -r : r;
}

struct DemodLuts {
    double lut[256];   // unpacked sample value, widened (exact)
    float lutf[256];
    double atan_d[9];
    float atan_f[9];
};

// one discriminator output from the raw bytes of samples i-1 (prv) and i (cur)
template <bool FAST>
__device__ __forceinline__ float demod_one(const DemodLuts &L, uchar2 prv, uchar2 cur)
{
    if (FAST) {
        const float pr = L.lutf[prv.x], pi = L.lutf[prv.y], cr = L.lutf[cur.x], ci = L.lutf[cur.y];
        const float fre = fmaf(pr, cr, pi * ci), fim = fmaf(ci, pr, -(pi * cr));
        const float m = fre * fre + fim * fim;
        const float y = atan2_fast(fim, fre, L.atan_f);
        return m > 1e-10f ? y : 0.f;
    } else {
        const double pr = L.lut[prv.x], pi = L.lut[prv.y], cr = L.lut[cur.x], ci = L.lut[cur.y];
        // products of f32 values are exact in f64, so one fused rounding equals the
        // reference's  pr*cr - ci*(-pi)  and  (-pi)*cr + ci*pr
        const double re = fma(pr, cr, __dmul_rn(ci, pi));
        const double im = fma(ci, pr, -__dmul_rn(pi, cr));
        const float fre = (float)re, fim = (float)im;
        // gates of the reference (p == 0, |p|^2 <= 1e-10f) as a select: no divergence
        const float m = __fadd_rn(__fmul_rn(fre, fre), __fmul_rn(fim, fim));
        const float y = (float)atan2_octant((double)fim, (double)fre, fim, fre, L.atan_d);
        return m > 1e-10f ? y : 0.f;
    }
}

template <bool FAST>
__global__ void __launch_bounds__(kThreads) k_demod_fused(const SigJob *jobs)
{
    __shared__ DemodLuts L;
    __shared__ double scratch[32];
    const SigJob &J = jobs[blockIdx.y];
    const int tid = threadIdx.x;
    {
        const float v = unpack_byte((unsigned)tid);
        L.lutf[tid] = v;
        L.lut[tid] = (double)v;
        if (tid < 9) { L.atan_d[tid] = atan_k8(tid); L.atan_f[tid] = (float)atan_k8(tid); }
    }
    __syncthreads();
    const i64 n = J.n;
    const uchar2 *__restrict__ raw = reinterpret_cast<const uchar2 *>(J.src.raw);
    float *__restrict__ out = J.p_re;
    const i64 run0 = J.src.run0_len;
    double pw = 0.0, sr = 0.0;
    for (i64 i0 = (i64)blockIdx.x * kTile; i0 < n; i0 += (i64)gridDim.x * kTile) {
        // whole tile (and the sample before it) inside one run of the capture: plain
        // 32-bit indexing from one base pointer, no per-sample bounds or run tests
        const bool one_run = i0 > 0 && i0 + kTile <= n && (i0 - 1 >= run0 || i0 + kTile <= run0);
        if (one_run) {
            const uchar2 *__restrict__ base = raw + raw_index(J.src, i0) + tid;
            float *__restrict__ o = out + i0 + tid;
            uchar2 cur[8], prv[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                cur[u] = base[256 * u];
                prv[u] = base[256 * u - 1];
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                pw += (double)mag2_f32(L.lutf[cur[u].x], L.lutf[cur[u].y]);
                const float y = demod_one<FAST>(L, prv[u], cur[u]);
                o[256 * u] = y;
                sr += (double)y;
            }
        } else {
            for (int u = 0; u < 8; u++) {
                const i64 i = i0 + tid + 256 * u;
                if (i >= n) break;
                const i64 k = i == 0 ? 1 : i;  // out[0] = out[1]
                const uchar2 cur = raw[raw_index(J.src, k)], prv = raw[raw_index(J.src, k - 1)];
                const uchar2 me = i == 0 ? prv : cur;  // initial power is of sample i itself
                pw += (double)mag2_f32(L.lutf[me.x], L.lutf[me.y]);
                const float y = demod_one<FAST>(L, prv, cur);
                out[i] = y;
                sr += (double)y;
            }
        }
    }
    double part[2], total[2];
    part[0] = block_sum(pw, scratch);
    part[1] = block_sum(sr, scratch);
    if (grid_sum_last<2>(part, J.partials, J.counter, gridDim.x, blockIdx.x, scratch, total)) {
        J.stats[ST_POWER0] = n > 0 ? total[0] / (double)n : 0.0;
        J.stats[ST_SUM_RE] = total[1];
        J.stats[ST_SUM_IM] = 0.0;
        J.stats[ST_DC_RE] = n > 0 ? (double)dc_from_sum(total[1], n) : 0.0;
        J.stats[ST_DC_IM] = 0.0;
    }
}

// ---------------------------------------------------------------- small box-car, real signal
constexpr int kHalo = 16;   // staged halo each side (multiple of 4 for float4 staging)
constexpr int kHalfMax = 8; // window / 2 served by this kernel

// x / d for a small odd integer d, correctly rounded: q = x * RN(1/d), one FMA residual,
// one FMA correction (Markstein).  Exhaustively checked against __fdiv_rn over all 2^32
// float inputs for every d this file uses (tdoa_selftest / tests/test_gpu_parity.py).
template <int D>
__device__ __forceinline__ float div_small(float x)
{
    constexpr float r = 1.0f / (float)D;
    const float q = __fmul_rn(x, r);
    const float e = __fmaf_rn(-q, (float)D, x);
    return __fmaf_rn(e, r, q);
}

// Leading / trailing out-of-range taps are staged as 0.f: 0.f + x == x and acc + 0.f == acc
// exactly, so the tap sum equals the reference's sum over the in-range taps only.
// INTERIOR: every output has all 2H+1 taps in range -> constant divisor.
template <int H, bool INTERIOR>
__device__ __forceinline__ void box_outputs(const float (&w)[24], float (&out)[8], i64 ib, i64 n)
{
#pragma unroll
    for (int o = 0; o < 8; o++) {
        float acc = w[8 + o - H];
#pragma unroll
        for (int j = 1; j <= 2 * H; j++) acc = __fadd_rn(acc, w[8 + o - H + j]);
        if (INTERIOR) {
            out[o] = div_small<2 * H + 1>(acc);
        } else {
            const i64 i = ib + o;
            const i64 a = max((i64)0, i - H), b = min(n - 1, i + H);
            out[o] = __fdiv_rn(acc, (float)(int)(b - a + 1));  // processor.go:289 divides by the tap count
        }
    }
}

template <bool INTERIOR>
__device__ __forceinline__ void box_dispatch(int h, const float (&w)[24], float (&out)[8], i64 ib, i64 n)
{
    switch (h) {
        case 0:
#pragma unroll
            for (int o = 0; o < 8; o++) out[o] = w[8 + o];
            break;
        case 1: box_outputs<1, INTERIOR>(w, out, ib, n); break;
        case 2: box_outputs<2, INTERIOR>(w, out, ib, n); break;
        case 3: box_outputs<3, INTERIOR>(w, out, ib, n); break;
        case 4: box_outputs<4, INTERIOR>(w, out, ib, n); break;
        case 5: box_outputs<5, INTERIOR>(w, out, ib, n); break;
        case 6: box_outputs<6, INTERIOR>(w, out, ib, n); break;
        case 7: box_outputs<7, INTERIOR>(w, out, ib, n); break;
        default: box_outputs<8, INTERIOR>(w, out, ib, n); break;
    }
}

__global__ void __launch_bounds__(kThreads, 4) k_boxcar_small(const SigJob *jobs)
{
    __shared__ __align__(16) float s_x[kTile + 2 * kHalo];
    __shared__ double scratch[32];
    const SigJob &J = jobs[blockIdx.y];
    const i64 n = J.n;
    const int tid = threadIdx.x;
    const int h = J.window <= 1 ? 0 : J.window / 2;
    const float dc = J.sub_dc ? (float)J.stats[ST_DC_RE] : 0.f;
    const float *__restrict__ q = J.q_re;
    float *__restrict__ p = J.p_re;
    double pacc = 0.0;
    // software pipeline: the next tile's global loads are issued before this tile's taps
    // are summed, so their latency hides behind the arithmetic (3 float4 per thread)
    constexpr int kStage4 = (kTile + 2 * kHalo) / 4;          // 520 float4 per tile
    constexpr int kPer = (kStage4 + kThreads - 1) / kThreads;  // 3
    float4 pf[kPer];
    auto tile_interior = [&](i64 t0) { return t0 >= kHalo && t0 + kTile + kHalo <= n; };
    auto prefetch = [&](i64 t0) {
        if (t0 < n && tile_interior(t0)) {
            const float4 *__restrict__ src = reinterpret_cast<const float4 *>(q + (t0 - kHalo));
#pragma unroll
            for (int k = 0; k < kPer; k++) {
                const int j4 = tid + k * kThreads;
                if (j4 < kStage4) pf[k] = src[j4];
            }
        }
    };
    prefetch((i64)blockIdx.x * kTile);
    for (i64 i0 = (i64)blockIdx.x * kTile; i0 < n; i0 += (i64)gridDim.x * kTile) {
        const bool interior = tile_interior(i0);
        __syncthreads();
        // stage [i0 - kHalo, i0 + kTile + kHalo) minus dc
        if (interior) {
#pragma unroll
            for (int k = 0; k < kPer; k++) {
                const int j4 = tid + k * kThreads;
                if (j4 < kStage4) {
                    float4 v = pf[k];
                    v.x = __fsub_rn(v.x, dc); v.y = __fsub_rn(v.y, dc); v.z = __fsub_rn(v.z, dc); v.w = __fsub_rn(v.w, dc);
                    *reinterpret_cast<float4 *>(s_x + 4 * j4) = v;
                }
            }
        } else {
            for (int j = tid; j < kTile + 2 * kHalo; j += kThreads) {
                const i64 g = i0 - kHalo + j;
                s_x[j] = (g >= 0 && g < n) ? __fsub_rn(q[g], dc) : 0.f;
            }
        }
        prefetch(i0 + (i64)gridDim.x * kTile);
        __syncthreads();
        const i64 ib = i0 + 8 * tid;  // first output of this thread
        if (ib < n) {
            // outputs ib..ib+7 need s_x[kHalo + 8 tid - h .. kHalo + 8 tid + 7 + h], inside
            // the aligned window s_x[8 tid + 8 .. 8 tid + 32)
            float w[24];
#pragma unroll
            for (int v4 = 0; v4 < 6; v4++) {
                const float4 v = *reinterpret_cast<const float4 *>(s_x + 8 * tid + 8 + 4 * v4);
                w[4 * v4] = v.x; w[4 * v4 + 1] = v.y; w[4 * v4 + 2] = v.z; w[4 * v4 + 3] = v.w;
            }
            float out[8];
            if (interior) box_dispatch<true>(h, w, out, ib, n);
            else box_dispatch<false>(h, w, out, ib, n);
            if (ib + 7 < n) {
#pragma unroll
                for (int o = 0; o < 8; o++) pacc += (double)__fmul_rn(out[o], out[o]);
                *reinterpret_cast<float4 *>(p + ib) = make_float4(out[0], out[1], out[2], out[3]);
                *reinterpret_cast<float4 *>(p + ib + 4) = make_float4(out[4], out[5], out[6], out[7]);
            } else {
#pragma unroll
                for (int o = 0; o < 8; o++)
                    if (ib + o < n) {
                        pacc += (double)__fmul_rn(out[o], out[o]);
                        p[ib + o] = out[o];
                    }
            }
        }
    }
    if (J.want_power) {
        double part[1] = {block_sum(pacc, scratch)}, total[1];
        if (grid_sum_last<1>(part, J.partials, J.counter, gridDim.x, blockIdx.x, scratch, total)) {
            const double pw = n > 0 ? total[0] / (double)n : 0.0;
            J.stats[ST_POWER1] = pw;
            J.stats[ST_SCALE] = pw > 0.0 ? (double)(float)(1.0 / sqrt(pw)) : 1.0;
        }
    }
}

// exhaustive check of div_small<D> against __fdiv_rn over every float bit pattern
template <int D>
__device__ __forceinline__ unsigned div_mismatch(float x)
{
    const float a = div_small<D>(x), b = __fdiv_rn(x, (float)D);
    return (__float_as_uint(a) != __float_as_uint(b) && !(a != a && b != b)) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_div_selftest(unsigned long long *bad)
{
    unsigned cnt = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * 256;
    for (unsigned long long v = (unsigned long long)blockIdx.x * 256 + threadIdx.x; v < (1ull << 32); v += stride) {
        const float x = __uint_as_float((unsigned)v);
        if (fabsf(x) < 1e-35f || fabsf(x) > 1e35f) continue;  // box-car sums live far inside this range
        cnt += div_mismatch<3>(x) + div_mismatch<5>(x) + div_mismatch<7>(x) + div_mismatch<9>(x) + div_mismatch<11>(x) +
               div_mismatch<13>(x) + div_mismatch<15>(x) + div_mismatch<17>(x);
    }
    if (cnt) atomicAdd(bad, (unsigned long long)cnt);
}

}  // namespace

int fast_grid_x(i64 n)
{
    const i64 tiles = (n + kTile - 1) / kTile;
    const i64 cap = 148 * 8;
    return (int)(tiles < 1 ? 1 : (tiles > cap ? cap : tiles));
}

void launch_demod_fused(const SigJob *d_jobs, int n_jobs, i64 max_n, int fast, cudaStream_t st)
{
    const dim3 grid(fast_grid_x(max_n), n_jobs);
    if (fast) k_demod_fused<true><<<grid, kThreads, 0, st>>>(d_jobs);
    else k_demod_fused<false><<<grid, kThreads, 0, st>>>(d_jobs);
}

void launch_boxcar_small(const SigJob *d_jobs, int n_jobs, i64 max_n, cudaStream_t st)
{
    k_boxcar_small<<<dim3(fast_grid_x(max_n), n_jobs), kThreads, 0, st>>>(d_jobs);
}

int boxcar_small_max_half() { return kHalfMax; }

long long div_selftest(cudaStream_t st)
{
    unsigned long long *d_bad = nullptr, h_bad = ~0ull;
    if (cudaMalloc(&d_bad, sizeof(*d_bad)) != cudaSuccess) return -1;
    cudaMemsetAsync(d_bad, 0, sizeof(*d_bad), st);
    k_div_selftest<<<148 * 16, 256, 0, st>>>(d_bad);
    cudaMemcpyAsync(&h_bad, d_bad, sizeof(h_bad), cudaMemcpyDeviceToHost, st);
    const cudaError_t err = cudaStreamSynchronize(st);
    cudaFree(d_bad);
    return err == cudaSuccess ? (long long)h_bad : -1;
}

}  // namespace tdoa
