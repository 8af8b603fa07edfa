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

template <bool FAST>
__global__ void __launch_bounds__(kThreads) k_demod_fused(const SigJob *jobs)
{
    __shared__ double s_lut[256];   // unpacked sample value, widened (exact)
    __shared__ float s_lutf[256];
    __shared__ double s_atan[9];
    __shared__ float s_atanf[9];
    __shared__ double scratch[32];
    const SigJob &J = jobs[blockIdx.y];
    const int tid = threadIdx.x;
    {
        const float v = unpack_byte((unsigned)tid);
        s_lutf[tid] = v;
        s_lut[tid] = (double)v;
        if (tid < 9) { s_atan[tid] = atan_k8(tid); s_atanf[tid] = (float)atan_k8(tid); }
    }
    __syncthreads();
    const i64 n = J.n;
    const uchar2 *raw = reinterpret_cast<const uchar2 *>(J.src.raw);
    double pw = 0.0, sr = 0.0;
    for (i64 i0 = (i64)blockIdx.x * kTile; i0 < n; i0 += (i64)gridDim.x * kTile) {
        uchar2 cur[8], prv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const i64 i = i0 + tid + 256 * u;
            const i64 k = i == 0 ? 1 : i;  // out[0] = out[1]
            cur[u] = make_uchar2(128, 128);
            prv[u] = cur[u];
            if (i < n) {
                cur[u] = raw[raw_index(J.src, k)];
                prv[u] = raw[raw_index(J.src, k - 1)];
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const i64 i = i0 + tid + 256 * u;
            if (i >= n) continue;
            // initial power of sample i (f32 re*re + im*im, widened)
            const uchar2 me = i == 0 ? prv[u] : cur[u];
            pw += (double)mag2_f32(s_lutf[me.x], s_lutf[me.y]);
            float y;
            if (FAST) {
                const float pr = s_lutf[prv[u].x], pi = s_lutf[prv[u].y], cr = s_lutf[cur[u].x], ci = s_lutf[cur[u].y];
                const float fre = fmaf(pr, cr, pi * ci), fim = fmaf(ci, pr, -(pi * cr));
                const float m = fre * fre + fim * fim;
                y = m > 1e-10f ? atan2_fast(fim, fre, s_atanf) : 0.f;
            } else {
                const double pr = s_lut[prv[u].x], pi = s_lut[prv[u].y], cr = s_lut[cur[u].x], ci = s_lut[cur[u].y];
                // products of f32 values are exact in f64, so one fused rounding equals the
                // reference's  pr*cr - ci*(-pi)  and  (-pi)*cr + ci*pr
                const double re = fma(pr, cr, __dmul_rn(ci, pi));
                const double im = fma(ci, pr, -__dmul_rn(pi, cr));
                const float fre = (float)re, fim = (float)im;
                y = 0.f;
                if (!(fre == 0.f && fim == 0.f)) {
                    const float m = __fadd_rn(__fmul_rn(fre, fre), __fmul_rn(fim, fim));
                    if (m > 1e-10f) y = (float)atan2_octant((double)fim, (double)fre, s_atan);
                }
            }
            J.p_re[i] = y;
            sr += (double)y;
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

// Leading / trailing out-of-range taps are staged as 0.f: 0.f + x == x and acc + 0.f == acc
// exactly, so the tap sum equals the reference's sum over the in-range taps only.
template <int H>
__device__ __forceinline__ void box_outputs(const float (&w)[24], float (&out)[8])
{
#pragma unroll
    for (int o = 0; o < 8; o++) {
        float acc = w[8 + o - H];
#pragma unroll
        for (int j = 1; j <= 2 * H; j++) acc = __fadd_rn(acc, w[8 + o - H + j]);
        out[o] = acc;
    }
}

__global__ void __launch_bounds__(kThreads) k_boxcar_small(const SigJob *jobs)
{
    __shared__ __align__(16) float s_x[kTile + 2 * kHalo];
    __shared__ double scratch[32];
    const SigJob &J = jobs[blockIdx.y];
    const i64 n = J.n;
    const int tid = threadIdx.x;
    const int h = J.window <= 1 ? 0 : J.window / 2;
    const float dc = J.sub_dc ? (float)J.stats[ST_DC_RE] : 0.f;
    const float *__restrict__ q = J.q_re;
    double pacc = 0.0;
    for (i64 i0 = (i64)blockIdx.x * kTile; i0 < n; i0 += (i64)gridDim.x * kTile) {
        __syncthreads();
        // stage [i0 - kHalo, i0 + kTile + kHalo) minus dc; float4 where whole and aligned
        for (int j4 = tid; j4 < (kTile + 2 * kHalo) / 4; j4 += kThreads) {
            const i64 g = i0 - kHalo + 4 * (i64)j4;
            float4 v;
            if (g >= 0 && g + 3 < n) {
                v = *reinterpret_cast<const float4 *>(q + g);
                v.x = __fsub_rn(v.x, dc); v.y = __fsub_rn(v.y, dc); v.z = __fsub_rn(v.z, dc); v.w = __fsub_rn(v.w, dc);
            } else {
                v.x = (g >= 0 && g < n) ? __fsub_rn(q[g], dc) : 0.f;
                v.y = (g + 1 >= 0 && g + 1 < n) ? __fsub_rn(q[g + 1], dc) : 0.f;
                v.z = (g + 2 >= 0 && g + 2 < n) ? __fsub_rn(q[g + 2], dc) : 0.f;
                v.w = (g + 3 >= 0 && g + 3 < n) ? __fsub_rn(q[g + 3], dc) : 0.f;
            }
            *reinterpret_cast<float4 *>(s_x + 4 * j4) = v;
        }
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
            switch (h) {
                case 0:
#pragma unroll
                    for (int o = 0; o < 8; o++) out[o] = w[8 + o];
                    break;
                case 1: box_outputs<1>(w, out); break;
                case 2: box_outputs<2>(w, out); break;
                case 3: box_outputs<3>(w, out); break;
                case 4: box_outputs<4>(w, out); break;
                case 5: box_outputs<5>(w, out); break;
                case 6: box_outputs<6>(w, out); break;
                case 7: box_outputs<7>(w, out); break;
                default: box_outputs<8>(w, out); break;
            }
#pragma unroll
            for (int o = 0; o < 8; o++) {
                const i64 i = ib + o;
                if (h > 0) {
                    const i64 a = max((i64)0, i - h), b = min(n - 1, i + h);
                    out[o] = __fdiv_rn(out[o], (float)(int)(b - a + 1));
                }
                if (i < n) pacc += (double)__fmul_rn(out[o], out[o]);
            }
            if (ib + 7 < n) {
                *reinterpret_cast<float4 *>(J.p_re + ib) = make_float4(out[0], out[1], out[2], out[3]);
                *reinterpret_cast<float4 *>(J.p_re + ib + 4) = make_float4(out[4], out[5], out[6], out[7]);
            } else {
#pragma unroll
                for (int o = 0; o < 8; o++)
                    if (ib + o < n) J.p_re[ib + o] = out[o];
            }
        }
    }
    if (J.want_power) {
        double part[1] = {block_sum(pacc, scratch)}, total[1];
        if (grid_sum_last<1>(part, J.partials, J.counter, gridDim.x, blockIdx.x, scratch, total)) {
            const double p = n > 0 ? total[0] / (double)n : 0.0;
            J.stats[ST_POWER1] = p;
            J.stats[ST_SCALE] = p > 0.0 ? (double)(float)(1.0 / sqrt(p)) : 1.0;
        }
    }
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

}  // namespace tdoa
