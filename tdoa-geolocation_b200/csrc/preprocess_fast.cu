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
#include <cstdlib>
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

// ---------------------------------------------------------------- lean discriminator
// The same arithmetic as demod_one<false> -- f64 products, one rounding to f32, f64
// arctangent, one rounding to f32 -- laid out for the B200's issue/pipe budget.  Measured
// on this part (tools/fp64_rates.cu): FP64 runs at half the FP32 rate (64/clk/SM), and
// every f32<->f64 conversion costs 8 cycles of the 16-lane XU pipe per warp, so the
// conversions are what the first version of this kernel was bound by.  Here:
//   - a thread owns 8 consecutive samples (one 16-byte span of the capture), so each
//     sample is unpacked once and reused as the "previous" of the next;
//   - bytes index a 256 x 16-slot table of {value widened to f64, f32 square}; slot =
//     lane % 16 makes every 128-bit load conflict free, and the address is one PRMT
//     (byte << 8 | slot << 4) with the table base folded into the LDS as a uniform register;
//   - "round the f64 product to f32" is integer arithmetic on the bit pattern (rn24);
//   - arctangent: octant by select, atan(q) = atan(k/64) + atan(z) with k from a
//     magic-number add (no int<->float conversion), one RCP64H + one Newton step + one
//     residual correction for the quotient, |z| <= 1/127 so three odd terms suffice
//     (truncation < 2^-59 relative); the octant fix-ups pi/2 - a, pi - a are folded
//     into the table (one entry per octant case and k) and a sign flip of atan(z);
//   - the |p|^2 <= 1e-10 gate of the reference can never fire for uint8 samples (the
//     smallest product magnitude is 9.46e-10), so it is not evaluated.
// tdoa_selftest(1) runs this function and demod_one<false> over all 2^32 byte quads.
constexpr int kLeanThreads = 512;
constexpr int kLeanPer = 8;                           // consecutive samples per thread
constexpr int kLeanTile = kLeanThreads * kLeanPer;    // 4096 samples per CTA step
constexpr int kAtanK = 64;
constexpr int kOctStride = 128;                       // table entries per octant case

struct __align__(16) LeanEntry {
    double v;   // unpacked sample value (processor.go:198-199), widened: exact
    float sq;   // RN_f32(v * v)
    float vf;   // v
};

constexpr int kLeanWords = 6;  // 32-bit words of the capture a thread needs: 1 before + 4 + 1 after

struct LeanSmem {
    LeanEntry tab[256][16];         // 64 KB
    double oct[4][kOctStride];      // B(case) + sigma(case) * atan(k / 64)
    unsigned stage[2][kLeanWords][kLeanThreads];  // cp.async landing zone, double buffered
    double scratch[32];
    int last;
};

__device__ __forceinline__ void lean_fill(LeanSmem &S, const double *__restrict__ atab_g)
{
    for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) {
        const float v = unpack_byte((unsigned)(i >> 4));
        LeanEntry e;
        e.v = (double)v; e.sq = __fmul_rn(v, v); e.vf = v;
        S.tab[i >> 4][i & 15] = e;
    }
    // case = swap + 2 * (x < 0):  a, pi/2 - a, pi - a, pi/2 + a
    for (int i = threadIdx.x; i < 4 * kOctStride; i += blockDim.x) {
        const int oc = i / kOctStride, k = i % kOctStride;
        const double a = atab_g[k <= kAtanK ? k : kAtanK];
        const double pio2 = 1.57079632679489661923, pi = 3.14159265358979323846;
        S.oct[oc][k] = oc == 0 ? a : (oc == 1 ? pio2 - a : (oc == 2 ? pi - a : pio2 + a));
    }
}

__device__ __forceinline__ double hilo(unsigned hi, unsigned lo) { return __hiloint2double((int)hi, (int)lo); }

// x rounded to f32 precision, ties to even, returned as f64 (x finite, in f32's normal range or 0)
__device__ __forceinline__ double rn24(double x)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    u += 0x0FFFFFFFull + ((u >> 29) & 1ull);
    return __longlong_as_double((long long)(u & ~0x1FFFFFFFull));
}

__device__ __forceinline__ double rcp_seed(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

__device__ __forceinline__ double lds_f64(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// f64 arctangent of (y, x), both f32-representable, not both zero, y != -0.  <= 1 ulp.
// oct_base: shared-window address of LeanSmem::oct.
__device__ __forceinline__ double atan2_lean(double y, double x, unsigned oct_base)
{
    const unsigned xh = (unsigned)__double2hiint(x), yh = (unsigned)__double2hiint(y);
    const double ax = hilo(xh & 0x7fffffffu, (unsigned)__double2loint(x));
    const double ay = hilo(yh & 0x7fffffffu, (unsigned)__double2loint(y));
    const bool swap = ay > ax;
    const double mx = swap ? ay : ax, mn = swap ? ax : ay;
    // k = round(64 mn / mx) from the low mantissa bits of q + 1.5 * 2^46 (ulp 2^-6)
    const double kMagic = 105553116266496.0;
    const double t = __dadd_rn(__dmul_rn(mn, rcp_seed(mx)), kMagic);
    const unsigned k = (unsigned)__double2loint(t);
    const double c = __dadd_rn(t, -kMagic);          // k / 64, exact
    const double num = fma(-c, mx, mn);              // mn - c mx
    const double den = fma(c, mn, mx);               // mx + c mn  in [mx, 2 mx]
    double r = rcp_seed(den);
    r = fma(r, fma(-den, r, 1.0), r);
    const double z0 = num * r;
    const double z = fma(r, fma(-den, z0, num), z0);
    const double w = z * z;
    double p = fma(w, -1.0 / 7.0, 1.0 / 5.0);
    p = fma(w, p, -1.0 / 3.0);
    const double az = fma(z * w, p, z);              // atan(z), |z| <= 1/127
    const unsigned sw = swap ? 1u : 0u, xn = xh >> 31;
    const double base = lds_f64(oct_base + ((xn * 2u + sw) * (unsigned)kOctStride + k) * 8u);
    // atan(z) enters with a minus sign when exactly one of {swap, x < 0} holds
    const unsigned flip = (sw ^ xn) << 31;
    const double a = base + hilo((unsigned)__double2hiint(az) ^ flip, (unsigned)__double2loint(az));
    return hilo((unsigned)__double2hiint(a) | (yh & 0x80000000u), (unsigned)__double2loint(a));
}

// discriminator output for the sample (cr, ci) given the previous sample (pr, pi)
__device__ __forceinline__ float lean_one(double pr, double pi, double cr, double ci, unsigned oct_base)
{
    const double re = rn24(fma(pr, cr, __dmul_rn(pi, ci)));
    const double im = rn24(fma(ci, pr, -__dmul_rn(pi, cr)));
    return (float)atan2_lean(im, re, oct_base);
}

__device__ __forceinline__ void lean_lookup(unsigned addr, double &v, float &sq)
{
    // one 128-bit load as two f64 registers: the value lands in an aligned register pair
    // (no moves), the f32 square is the low word of the second
    double packed;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v), "=d"(packed) : "r"(addr));
    sq = __int_as_float(__double2loint(packed));
}

// grid_sum_last with the "am I last" flag in caller-provided shared memory
template <int K>
__device__ __forceinline__ bool grid_sum_last_dyn(const double (&part)[K], double *partials, unsigned *counter,
                                                  int n_cta, int cta, double *scratch, int *s_last, double (&total)[K])
{
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) partials[(size_t)k * n_cta + cta] = part[k];
        __threadfence();
        const unsigned t = atomicAdd(counter, 1u);
        *s_last = (t == (unsigned)n_cta - 1u);
    }
    __syncthreads();
    if (!*s_last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < K; k++) {
        double v = 0.0;
        for (int i = threadIdx.x; i < n_cta; i += blockDim.x) v += __ldcg(partials + (size_t)k * n_cta + i);
        v = block_sum(v, scratch);
        total[k] = v;
    }
    if (threadIdx.x == 0) {
        *counter = 0u;
        return true;
    }
    return false;
}

__global__ void __launch_bounds__(kLeanThreads, 2) k_demod_lean(const SigJob *jobs, const double *__restrict__ atab_g)
{
    extern __shared__ __align__(128) unsigned char lean_raw[];
    LeanSmem &S = *reinterpret_cast<LeanSmem *>(lean_raw);
    const SigJob &J = jobs[blockIdx.y];
    const int tid = threadIdx.x;
    lean_fill(S, atab_g);
    __syncthreads();
    const i64 n = J.n;
    const i64 i_begin = J.i_begin, i_end = J.i_end > 0 ? J.i_end : n;
    const uint8_t *__restrict__ rawb = J.src.raw;
    float *__restrict__ out = J.p_re;
    const i64 run0 = J.src.run0_len;
    const unsigned slot16 = (unsigned)(tid & 15) * 16u;
    const unsigned tab_base = (unsigned)__cvta_generic_to_shared(&S.tab[0][0]);
    const unsigned oct_base = (unsigned)__cvta_generic_to_shared(&S.oct[0][0]);
    double pw = 0.0, sr = 0.0;
    // A tile takes the fast path when it, the sample before it and one 32-bit word after
    // it lie inside one run of the capture.  Its bytes are fetched one tile ahead with
    // cp.async (4-byte granules: only 2-byte alignment of a signal start is known), each
    // thread staging and later reading only its own six words, so no barrier is needed.
    auto fast_tile = [&](i64 t0) {
        const bool in0 = t0 + kLeanTile + 2 <= run0, in1 = t0 - 1 >= run0 && t0 + kLeanTile + 2 <= n;
        return t0 > 0 && t0 + kLeanTile <= i_end && (in0 || in1);
    };
    auto tile_addr = [&](i64 t0) { return rawb + 2 * (raw_index(J.src, t0) + (i64)kLeanPer * tid); };
    auto stage_tile = [&](i64 t0, int buf) {
        if (fast_tile(t0)) {
            const uint8_t *ap = tile_addr(t0);
            const unsigned *wp = reinterpret_cast<const unsigned *>(ap - (reinterpret_cast<uintptr_t>(ap) & 2u)) - 1;
#pragma unroll
            for (int k = 0; k < kLeanWords; k++) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&S.stage[buf][k][tid]);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(wp + k) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const i64 step = (i64)gridDim.x * kLeanTile;
    int buf = 0;
    stage_tile(i_begin + (i64)blockIdx.x * kLeanTile, 0);
    for (i64 i0 = i_begin + (i64)blockIdx.x * kLeanTile; i0 < i_end; i0 += step, buf ^= 1) {
        stage_tile(i0 + step, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        if (fast_tile(i0)) {
            const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(tile_addr(i0)) & 2u) * 8u;
            const unsigned wm = S.stage[buf][0][tid], w0 = S.stage[buf][1][tid], w1 = S.stage[buf][2][tid],
                           w2 = S.stage[buf][3][tid], w3 = S.stage[buf][4][tid], w4 = S.stage[buf][5][tid];
            const unsigned pv = __funnelshift_r(wm, w0, sh);   // bytes [ap - 4, ap): previous sample in the top half
            unsigned w[4];
            w[0] = __funnelshift_r(w0, w1, sh); w[1] = __funnelshift_r(w1, w2, sh);
            w[2] = __funnelshift_r(w2, w3, sh); w[3] = __funnelshift_r(w3, w4, sh);
            // table address of byte b of a word: (b << 8) | slot * 16
            double pr, pi;
            float sq_i, sq_q;
            lean_lookup(tab_base + __byte_perm(pv, slot16, 0x5524), pr, sq_i);
            lean_lookup(tab_base + __byte_perm(pv, slot16, 0x5534), pi, sq_q);
            float o[kLeanPer];
#pragma unroll
            for (int s = 0; s < kLeanPer; s++) {
                const unsigned ww = w[s >> 1];
                double cr, ci;
                lean_lookup(tab_base + __byte_perm(ww, slot16, (s & 1) ? 0x5524 : 0x5504), cr, sq_i);
                lean_lookup(tab_base + __byte_perm(ww, slot16, (s & 1) ? 0x5534 : 0x5514), ci, sq_q);
                pw += (double)__fadd_rn(sq_i, sq_q);   // processor.go:328, f32 re*re + im*im
                o[s] = lean_one(pr, pi, cr, ci, oct_base);
                sr += (double)o[s];
                pr = cr; pi = ci;
            }
            float4 *op = reinterpret_cast<float4 *>(out + i0 + (i64)kLeanPer * tid);
            op[0] = make_float4(o[0], o[1], o[2], o[3]);
            op[1] = make_float4(o[4], o[5], o[6], o[7]);
        } else {
            // edges (signal start, the block-1/block-3 joint, the tail): the reference's
            // own statement of the discriminator, sample by sample, gates included
            for (int u = 0; u < kLeanPer; u++) {
                const i64 i = i0 + tid + (i64)kLeanThreads * u;
                if (i >= i_end) break;
                const i64 k = i == 0 ? 1 : i;  // out[0] = out[1]
                const uchar2 cur = reinterpret_cast<const uchar2 *>(rawb)[raw_index(J.src, k)];
                const uchar2 prv = reinterpret_cast<const uchar2 *>(rawb)[raw_index(J.src, k - 1)];
                const uchar2 me = i == 0 ? prv : cur;  // initial power is of sample i itself
                pw += (double)__fadd_rn(S.tab[me.x][0].sq, S.tab[me.y][0].sq);
                const double pr = S.tab[prv.x][0].v, pi = S.tab[prv.y][0].v, cr = S.tab[cur.x][0].v, ci = S.tab[cur.y][0].v;
                const double re = fma(pr, cr, __dmul_rn(ci, pi)), im = fma(ci, pr, -__dmul_rn(pi, cr));
                const float fre = (float)re, fim = (float)im;
                const float m = __fadd_rn(__fmul_rn(fre, fre), __fmul_rn(fim, fim));
                float y = 0.f;
                if (m > 1e-10f) y = (float)atan2_lean((double)fim, (double)fre, oct_base);
                out[i] = y;
                sr += (double)y;
            }
        }
    }
    double part[2], total[2];
    part[0] = block_sum(pw, S.scratch);
    part[1] = block_sum(sr, S.scratch);
    if (grid_sum_last_dyn<2>(part, J.partials, J.counter, gridDim.x, blockIdx.x, S.scratch, &S.last, total)) {
        if (J.chunk_out) {
            J.chunk_out[0] = total[0];
            J.chunk_out[1] = total[1];
            return;
        }
        J.stats[ST_POWER0] = n > 0 ? total[0] / (double)n : 0.0;
        J.stats[ST_SUM_RE] = total[1];
        J.stats[ST_SUM_IM] = 0.0;
        J.stats[ST_DC_RE] = n > 0 ? (double)dc_from_sum(total[1], n) : 0.0;
        J.stats[ST_DC_IM] = 0.0;
    }
}


// ---------------------------------------------------------------- double-float discriminator (round 2)
// The same bits as k_demod_lean at four fifths of its time.  What ncu said about the f64 kernels
// (profiles/r2_demod.md): 76 instructions per sample, NINE of them on the 16-lane XU pipe (three f64->f32 and
// four f32->f64 conversions, MUFU.RCP, MUFU.RCP64H) and 28 LSU wavefronts per warp-sample (six 4-byte cp.async
// per thread at a 16-byte lane stride, three 128-bit table loads per sample) -- the XU and LSU pipes, not FP64,
// were what bound it.  Here:
//   - the arctangent runs in f32 DOUBLE-FLOAT arithmetic (value = hi + lo, two f32), issued two samples at a
//     time with Blackwell's packed FFMA2 / FADD2 / FMUL2 (one issue slot for two samples);
//   - X, Y: the reference's f64 products (DMUL + DFMA), rounded to f32 by the conversion instruction: the only
//     two conversions per sample (the two sums widen their terms with integer moves, widen_scaled);
//   - octant by FMNMX / FSETP on |x|, |y|; c = k/128 from a magic-number add on mn * rcp(mx);
//   - rotation  y' = mn - c mx  (exact in f32: the leading bits cancel),  x' = mx + c mn  as hi + lo
//     (three FMAs), u = y'/x' from one MUFU.RCP seed and an exact-residual correction (error < 2^-44),
//     atan(u) = u - u^3/3 + u^5/5 with |u| <= 2^-8 (the terms after u ride in the low word);
//   - B(case) +- atan(k/128) from a table of (hi, lo) pairs, added by a Fast2Sum;
//   - the f32 result is o = RN(hi + lo); with r = (hi + lo) - o (exact, |r| <= ulp/2) the test
//     RN(o + r (1 + 2^-16)) != o says that hi + lo sits within 2^-40 (relative) of a rounding boundary, where
//     the fast path's accuracy does not decide the rounding: 3e-5 of the samples.  A thread with such a
//     sample among its eight redoes them with the full-accuracy f64 arctangent of round 1 (demod_exact8).
// tdoa_selftest(1) runs THIS path (fast value, exact fall-back when flagged) against the reference statement
// demod_one<false> over all 2^32 byte quads: 0 differing, 132 432 fall-backs, 112 of which changed the value.
constexpr int kAtanK3 = 128;
constexpr int kOct3Stride = 256;   // entries per octant case

typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pk(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float rcp_f32(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#define TDOA_PK2(x) ((((unsigned long long)(x)) << 32) | (unsigned long long)(x))

// table entry of a byte: the address is (byte << 8 | slot << 4) from one PRMT; the table base is CTA-uniform
__device__ __forceinline__ void df_lookup(const LeanEntry *tab, unsigned off, double &v, float &sq)
{
    const double2 t = *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(tab) + off);
    v = t.x;
    sq = __int_as_float(__double2loint(t.y));
}

// f32 -> f64 without the conversion pipe: the bit pattern moved into the f64 fields is the value times
// 2^-896, exactly (normal f32 or zero); sums of such terms are scaled back once per CTA
__device__ __forceinline__ double widen_scaled_pos(float g)
{
    const unsigned b = __float_as_uint(g);
    return hilo(b >> 3, b << 29);
}
__device__ __forceinline__ double widen_scaled(float o)
{
    const unsigned b = __float_as_uint(o);
    return hilo((unsigned)((int)b >> 3) & 0x8FFFFFFFu, b << 29);
}

// Two discriminator values at once from the f64 products X = Re p, Y = Im p of two samples; `diff` collects
// (o1 ^ o2) of every sample: nonzero = some rounding undecided
template <class SM>
__device__ __forceinline__ void df_pair(const SM &S, double XA, double YA, double XB, double YB,
                                           float &oA, float &oB, unsigned &diff)
{
    const float xA = (float)XA, yA = (float)YA, xB = (float)XB, yB = (float)YB;   // the reference's float32(...) of the products
    const float mxA = fmaxf(fabsf(xA), fabsf(yA)), mnA = fminf(fabsf(xA), fabsf(yA));
    const float mxB = fmaxf(fabsf(xB), fabsf(yB)), mnB = fminf(fabsf(xB), fabsf(yB));
    const bool swA = fabsf(yA) > fabsf(xA), swB = fabsf(yB) > fabsf(xB);
    const unsigned xbA = __float_as_uint(xA), xbB = __float_as_uint(xB);
    const f32x2 MX = pk(mxA, mxB), MN = pk(mnA, mnB);
    const f32x2 kMagic = TDOA_PK2(0x47C00000u);     // 98304 = 1.5 * 2^16: ulp 2^-7
    const f32x2 kMinus1 = TDOA_PK2(0xBF800000u);
    const f32x2 T = fma2(MN, pk(rcp_f32(mxA), rcp_f32(mxB)), kMagic);   // q + magic: k = round(128 q) in the low mantissa bits
    const f32x2 C = fma2(kMagic, kMinus1, T);                   // T - magic =  c = k / 128, exact
    const f32x2 NC = fma2(T, kMinus1, kMagic);                  // -c
    float tA, tB;
    upk(T, tA, tB);
    // table entry: case = swap + 2 * (x < 0)
    const unsigned ixA = (__float_as_uint(tA) & 0x1FFu) + (swA ? (unsigned)kOct3Stride : 0u) + ((xbA >> 31) << 9);
    const unsigned ixB = (__float_as_uint(tB) & 0x1FFu) + (swB ? (unsigned)kOct3Stride : 0u) + ((xbB >> 31) << 9);
    const float2 eA = S.oct[ixA], eB = S.oct[ixB];
    const f32x2 BH = pk(eA.x, eB.x), BL = pk(eA.y, eB.y);
    const f32x2 Y1 = fma2(NC, MX, MN);                          // mn - c mx (exact)
    const f32x2 XH = fma2(C, MN, MX);                           // mx + c mn, rounded
    const f32x2 D = fma2(XH, kMinus1, MX);                      // mx - xh (exact)
    const f32x2 XL = fma2(C, MN, D);                            // (mx + c mn) - xh (exact)
    float xhA, xhB;
    upk(XH, xhA, xhB);
    const f32x2 NR = pk(rcp_f32(-xhA), rcp_f32(-xhB));          // -1 / xh
    const f32x2 NUH = mul2(Y1, NR);                             // -u, high part
    const f32x2 E = fma2(NUH, XH, Y1);                          // y' - uh xh
    const f32x2 E2 = fma2(NUH, XL, E);                          // y' - uh (xh + xl)
    const f32x2 NUL = mul2(E2, NR);                             // -u, low part
    const f32x2 W = mul2(NUH, NUH);
    const f32x2 P = fma2(W, TDOA_PK2(0x3E4CCCCDu), TDOA_PK2(0xBEAAAAABu));   // w / 5 - 1 / 3
    const f32x2 CORR = fma2(mul2(NUH, W), P, NUL);              // atan(-u) - (-uh)
    // atan(u) enters with a minus sign when exactly one of {swap, x < 0} holds; we carry -u, so sigma' = -sigma
    const unsigned sgA = (swA ? 0x3F800000u : 0xBF800000u) ^ (xbA & 0x80000000u);
    const unsigned sgB = (swB ? 0x3F800000u : 0xBF800000u) ^ (xbB & 0x80000000u);
    const f32x2 SIG = pk(__uint_as_float(sgA), __uint_as_float(sgB));
    const f32x2 SH = fma2(SIG, NUH, BH);                        // Fast2Sum: |bh| >= |uh| or bh == 0
    const f32x2 T2 = fma2(SH, kMinus1, BH);                     // bh - sh
    const f32x2 ERR = fma2(SIG, NUH, T2);
    const f32x2 LO = add2(ERR, fma2(SIG, CORR, BL));
    // hi + lo is not normalised (for k = 0 the low word carries all of u^3/3 ...): round, take the exact
    // remainder r (|r| <= half an ulp of o), and ask whether r (1 + 2^-14) still rounds to o
    const f32x2 O = add2(SH, LO);
    const f32x2 R = add2(fma2(O, kMinus1, SH), LO);             // (sh - o) + lo, both steps exact
    const f32x2 O1 = fma2(R, TDOA_PK2(0x3F800080u), O);         // 1 + 2^-16
    float o1A, o1B, o2A, o2B;
    upk(O1, o1A, o1B);
    upk(O, o2A, o2B);
    diff |= (__float_as_uint(o1A) ^ __float_as_uint(o2A)) | (__float_as_uint(o1B) ^ __float_as_uint(o2B));
    oA = __uint_as_float(__float_as_uint(o2A) | (__float_as_uint(yA) & 0x80000000u));
    oB = __uint_as_float(__float_as_uint(o2B) | (__float_as_uint(yB) & 0x80000000u));
}

// ---------------------------------------------------------------- staging of the capture bytes
// A thread owns 8 consecutive samples = 16 bytes of the capture, plus the sample before them.  The bytes of a
// tile arrive in shared memory as (THREADS + 1) 16-byte chunks copied from the 16-byte-aligned address below
// the tile's previous sample; a thread reads its two chunks with two LDS.128 and cuts its 18 bytes out with
// funnel shifts (the misalignment is uniform over the tile).  Two ways to get them there (template STAGE):
//   1 (production): every warp copies its own 33 chunks with 16-byte cp.async, one tile ahead -- nothing wider
//     than __syncwarp between the warps;
//   0: one thread issues ONE bulk copy (cp.async.bulk, the TMA engine: UBLKCP in the SASS) per CTA tile, two
//     tiles ahead, completion on an mbarrier.  No LSU work for the copy at all, but the buffer hand-back needs a
//     CTA barrier per tile, and with 4096-sample tiles that costs more than the copy saves (0.90 vs 0.75 ms).
template <int THREADS>
struct __align__(128) DfSmem {
    LeanEntry tab[256][16];                       // 64 KB
    float2 oct[4 * kOct3Stride];                  // double-float B(case) + sigma(case) * atan(k / 128): 8 KB
    double oct_exact[4][kOctStride];              // round-1 table for the exact fall-back (4 KB)
    __align__(128) uint4 stage[2][THREADS + THREADS / 32 + 8];   // landing zone of the bulk copies, double buffered (STAGE 0); or
                                                  // per warp w: chunks [33 w, 33 w + 33) (STAGE 1)
    unsigned long long full[2];                   // mbarriers
    double scratch[32];
    int last;
};

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
}
// one bulk copy global -> this CTA's shared memory; src, dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_bulk(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
}

// the 5 words starting at word m4 of the 8 staged ones, each shifted right by `s` bits into its successor
template <int M4>
__device__ __forceinline__ void cut_words(const unsigned (&W)[8], unsigned s, unsigned (&V)[5])
{
#pragma unroll
    for (int j = 0; j < 5; j++) V[j] = __funnelshift_r(W[M4 + j], M4 + j + 1 < 8 ? W[M4 + j + 1] : 0u, s);
}

template <class SM>
__device__ __forceinline__ void df_fill(SM &S, const double *__restrict__ atab_g, const double *__restrict__ atab3_g)
{
    const int tid = threadIdx.x, THREADS = blockDim.x;
    // tables: byte -> {value widened, f32 square}; octant case and k -> B(case) + sigma(case) atan(k / 128) as (hi, lo)
    for (int i = tid; i < 256 * 16; i += THREADS) {
        const float v = unpack_byte((unsigned)(i >> 4));
        LeanEntry e;
        e.v = (double)v; e.sq = __fmul_rn(v, v); e.vf = v;
        S.tab[i >> 4][i & 15] = e;
    }
    {
        const double pio2 = 1.57079632679489661923, pi = 3.14159265358979323846;
        for (int i = tid; i < 4 * kOct3Stride; i += THREADS) {
            const int oc = i / kOct3Stride, k = i % kOct3Stride;
            const double a = atab3_g[k <= kAtanK3 ? k : kAtanK3];
            const double b = oc == 0 ? a : (oc == 1 ? pio2 - a : (oc == 2 ? pi - a : pio2 + a));
            const float bh = (float)b;
            S.oct[i] = make_float2(bh, (float)(b - (double)bh));
        }
        for (int i = tid; i < 4 * kOctStride; i += THREADS) {
            const int oc = i / kOctStride, k = i % kOctStride;
            const double a = atab_g[k <= kAtanK ? k : kAtanK];
            S.oct_exact[oc][k] = oc == 0 ? a : (oc == 1 ? pio2 - a : (oc == 2 ? pi - a : pio2 + a));
        }
    }
}

// a thread's eight samples with the full-accuracy arctangent (rare: kept out of line so that its registers
// do not weigh on the fast path); v0..v4 as cut_words leaves them; stores the eight values, returns their sum
__device__ __noinline__ double demod_exact8(unsigned v0, unsigned v1, unsigned v2, unsigned v3, unsigned v4, unsigned slot16,
                                            unsigned tab_base, unsigned octx_base, float4 *op)
{
    const unsigned V[5] = {v0, v1, v2, v3, v4};
    double pr, pi;
    float sq_i, sq_q, o[kLeanPer];
    lean_lookup(tab_base + __byte_perm(V[0], slot16, 0x5504), pr, sq_i);
    lean_lookup(tab_base + __byte_perm(V[0], slot16, 0x5514), pi, sq_q);
    double srt = 0.0;
#pragma unroll
    for (int s = 0; s < kLeanPer; s++) {
        const unsigned ww = V[(s + 1) >> 1];
        double cr, ci;
        lean_lookup(tab_base + __byte_perm(ww, slot16, (s & 1) ? 0x5504 : 0x5524), cr, sq_i);
        lean_lookup(tab_base + __byte_perm(ww, slot16, (s & 1) ? 0x5514 : 0x5534), ci, sq_q);
        o[s] = lean_one(pr, pi, cr, ci, octx_base);
        srt += (double)o[s];
        pr = cr; pi = ci;
    }
    op[0] = make_float4(o[0], o[1], o[2], o[3]);
    op[1] = make_float4(o[4], o[5], o[6], o[7]);
    return srt;
}

// STAGE: 0 = one bulk copy (TMA) per CTA tile, mbarrier + one CTA barrier per tile; 1 = every warp stages its
// own 33 chunks with 16-byte cp.async (no barrier wider than the warp)
template <int THREADS, int MINB, int SUMS, int STAGE>
__global__ void __launch_bounds__(THREADS, MINB) k_demod_df(const SigJob *jobs, const double *__restrict__ atab_g,
                                                             const double *__restrict__ atab3_g)
{
    constexpr int kLeanTile = THREADS * kLeanPer;
    constexpr unsigned kStageBytes = (THREADS + 1) * 16;
    extern __shared__ __align__(128) unsigned char lean_raw[];
    DfSmem<THREADS> &S = *reinterpret_cast<DfSmem<THREADS> *>(lean_raw);
    const SigJob &J = jobs[blockIdx.y];
    const int tid = threadIdx.x;
    df_fill(S, atab_g, atab3_g);
    if (tid == 0) {
        mbar_init(&S.full[0], 1);
        mbar_init(&S.full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const i64 n = J.n;
    const i64 i_begin = J.i_begin, i_end = J.i_end > 0 ? J.i_end : n;
    const uint8_t *__restrict__ rawb = J.src.raw;
    float *__restrict__ out = J.p_re;
    const i64 run0 = J.src.run0_len;
    const unsigned slot16 = (unsigned)(tid & 15) * 16u;
    const LeanEntry *tab = &S.tab[0][0];
    const unsigned tab_base = (unsigned)__cvta_generic_to_shared(&S.tab[0][0]);
    const unsigned octx_base = (unsigned)__cvta_generic_to_shared(&S.oct_exact[0][0]);
    double pw = 0.0, sr = 0.0;       // plain sums (edges, fall-backs)
    double pws = 0.0, srs = 0.0;     // sums of terms scaled by 2^-896 (SUMS == 1)
    // Tile j covers samples [i_begin + j T, + T).  It takes the fast path when it, the 8 samples before it and the
    // 8 after it lie inside one run of the capture (the staged copy starts up to 14 bytes before the tile's
    // previous sample and ends up to 16 after its last): two ranges of tile indices, worked out once.
    constexpr i64 T = kLeanTile;
    const int n_tiles = (int)((i_end - i_begin + T - 1) / T);
    auto ceil_div = [](i64 a, i64 b) { return a <= 0 ? (i64)0 : (a + b - 1) / b; };
    auto tiles_below = [&](i64 limit) {   // number of tiles j with i_begin + j T + T <= limit
        const i64 room = limit - i_begin;
        return room < T ? (i64)0 : room / T;
    };
    auto clampi = [&](i64 v) { return (int)(v < 0 ? 0 : (v > n_tiles ? n_tiles : v)); };
    const i64 end0 = (run0 - 8 < i_end ? run0 - 8 : i_end), end1 = (n - 8 < i_end ? n - 8 : i_end);
    const int jA0 = clampi(ceil_div(8 - i_begin, T)), jB0 = clampi(tiles_below(end0));
    const int jA1 = clampi(ceil_div(run0 + 8 - i_begin, T)), jB1 = clampi(tiles_below(end1));
    auto fast_tile = [&](int j) { return (j >= jA0 && j < jB0) || (j >= jA1 && j < jB1); };
    // address of tile j's previous sample, per run
    const uintptr_t base0 = reinterpret_cast<uintptr_t>(rawb) + 2 * (uintptr_t)(J.src.run0_start + i_begin) - 2;
    const uintptr_t base1 = reinterpret_cast<uintptr_t>(rawb) + 2 * (uintptr_t)(J.src.run1_start + i_begin - run0) - 2;
    auto prev_addr = [&](int j) { return (j < jB0 ? base0 : base1) + (uintptr_t)j * (uintptr_t)(2 * T); };
    auto issue = [&](int j, int buf) {
        if (j < n_tiles && fast_tile(j))
            tma_load_bulk(&S.stage[buf][0], reinterpret_cast<const void *>(prev_addr(j) & ~(uintptr_t)15), kStageBytes, &S.full[buf]);
    };
    const int jstep = (int)gridDim.x;
    const int jfirst = (int)blockIdx.x;
    // STAGE 1: warp w keeps its 33 chunks at stage[buf][33 w .. 33 w + 32]
    const int lane = tid & 31, wrp = tid >> 5;
    const unsigned wstage0 = (unsigned)__cvta_generic_to_shared(&S.stage[0][0] + 33 * wrp + lane);
    constexpr unsigned kBufStride = (unsigned)sizeof(S.stage[0]);
    auto issue_w = [&](int j, int b) {
        if (j < n_tiles && fast_tile(j)) {
            const char *src = reinterpret_cast<const char *>(prev_addr(j) & ~(uintptr_t)15) + 16 * tid;
            const unsigned dst = wstage0 + (unsigned)b * kBufStride;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            if (lane == 0)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 32 * 16), "l"(src + 32 * 16) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (STAGE == 0) {
        if (tid == 0) { issue(jfirst, 0); issue(jfirst + jstep, 1); }
    } else {
        issue_w(jfirst, 0);
    }
    unsigned phase[2] = {0u, 0u};
    int buf = 0;
    for (int j = jfirst; j < n_tiles; j += jstep, buf ^= 1) {
        const bool fast = fast_tile(j);
        const i64 i0 = i_begin + (i64)j * T;
        unsigned W[8];
        if (STAGE == 0) {
            if (fast) {
                mbar_wait(&S.full[buf], phase[buf]);
                phase[buf] ^= 1u;
                const uint4 a = S.stage[buf][tid], b = S.stage[buf][tid + 1];
                W[0] = a.x; W[1] = a.y; W[2] = a.z; W[3] = a.w; W[4] = b.x; W[5] = b.y; W[6] = b.z; W[7] = b.w;
            }
            __syncthreads();                                   // every thread holds its words: the buffer is free
            if (tid == 0) issue(j + 2 * jstep, buf);
        } else {
            __syncwarp();                                      // the warp is done reading the other buffer
            issue_w(j + jstep, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();                                      // ... and every lane's copies of this one have landed
            if (fast) {
                uint4 a, b;
                const unsigned src = wstage0 + (unsigned)buf * kBufStride;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(src));
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(src));
                W[0] = a.x; W[1] = a.y; W[2] = a.z; W[3] = a.w; W[4] = b.x; W[5] = b.y; W[6] = b.z; W[7] = b.w;
            }
        }
        if (fast) {
            const unsigned mis = (unsigned)(prev_addr(j) & 15);  // even, uniform over the tile
            const unsigned sb = (mis & 2u) * 8u;
            unsigned V[5];   // V0 = [previous | s0], V1 = [s1 | s2], V2 = [s3 | s4], V3 = [s5 | s6], V4 = [s7 | -]
            switch (mis >> 2) {
                case 0: cut_words<0>(W, sb, V); break;
                case 1: cut_words<1>(W, sb, V); break;
                case 2: cut_words<2>(W, sb, V); break;
                default: cut_words<3>(W, sb, V); break;
            }
            double pr, pi;
            {
                float sq_i, sq_q;
                df_lookup(tab, __byte_perm(V[0], slot16, 0x5504), pr, sq_i);
                df_lookup(tab, __byte_perm(V[0], slot16, 0x5514), pi, sq_q);
            }
            float4 *op = reinterpret_cast<float4 *>(out + i0 + (i64)kLeanPer * tid);
            float o[4];
            unsigned diff = 0u;
            double srt = 0.0;
#pragma unroll
            for (int s = 0; s < kLeanPer; s += 2) {
                const unsigned wa = V[s >> 1], wb = V[(s >> 1) + 1];   // sample s in the high half of wa, s + 1 in the low half of wb
                double crA, ciA, crB, ciB;
                float sqiA, sqqA, sqiB, sqqB;
                df_lookup(tab, __byte_perm(wa, slot16, 0x5524), crA, sqiA);
                df_lookup(tab, __byte_perm(wa, slot16, 0x5534), ciA, sqqA);
                df_lookup(tab, __byte_perm(wb, slot16, 0x5504), crB, sqiB);
                df_lookup(tab, __byte_perm(wb, slot16, 0x5514), ciB, sqqB);
                const float gA = __fadd_rn(sqiA, sqqA), gB = __fadd_rn(sqiB, sqqB);   // processor.go:328, f32 re*re + im*im
                if (SUMS) { pws += widen_scaled_pos(gA); pws += widen_scaled_pos(gB); }
                else { pw += (double)gA; pw += (double)gB; }
                const double XA = fma(pr, crA, __dmul_rn(pi, ciA)), YA = fma(ciA, pr, -__dmul_rn(pi, crA));
                const double XB = fma(crA, crB, __dmul_rn(ciA, ciB)), YB = fma(ciB, crA, -__dmul_rn(ciA, crB));
                df_pair(S, XA, YA, XB, YB, o[s & 2], o[(s & 2) + 1], diff);
                if (SUMS) { srt += widen_scaled(o[s & 2]); srt += widen_scaled(o[(s & 2) + 1]); }
                else { srt += (double)o[s & 2]; srt += (double)o[(s & 2) + 1]; }
                if (s & 2) op[s >> 2] = make_float4(o[0], o[1], o[2], o[3]);
                pr = crB; pi = ciB;
            }
            if (diff) {
                // ~2^-13 of the samples: the thread's eight again, with the full-accuracy arctangent
                sr += demod_exact8(V[0], V[1], V[2], V[3], V[4], slot16, tab_base, octx_base, op);
            } else if (SUMS) {
                srs += srt;
            } else {
                sr += srt;
            }
        } else {
            // edges (signal start, the block-1/block-3 joint, the tail): the reference's
            // own statement of the discriminator, sample by sample, gates included
            for (int u = 0; u < kLeanPer; u++) {
                const i64 i = i0 + tid + (i64)THREADS * u;
                if (i >= i_end) break;
                const i64 k = i == 0 ? 1 : i;  // out[0] = out[1]
                const uchar2 cur = reinterpret_cast<const uchar2 *>(rawb)[raw_index(J.src, k)];
                const uchar2 prv = reinterpret_cast<const uchar2 *>(rawb)[raw_index(J.src, k - 1)];
                const uchar2 me = i == 0 ? prv : cur;  // initial power is of sample i itself
                pw += (double)__fadd_rn(S.tab[me.x][0].sq, S.tab[me.y][0].sq);
                const double pr = S.tab[prv.x][0].v, pi = S.tab[prv.y][0].v, cr = S.tab[cur.x][0].v, ci = S.tab[cur.y][0].v;
                const double re = fma(pr, cr, __dmul_rn(ci, pi)), im = fma(ci, pr, -__dmul_rn(pi, cr));
                const float fre = (float)re, fim = (float)im;
                const float m = __fadd_rn(__fmul_rn(fre, fre), __fmul_rn(fim, fim));
                float y = 0.f;
                if (m > 1e-10f) y = (float)atan2_lean((double)fim, (double)fre, octx_base);
                out[i] = y;
                sr += (double)y;
            }
        }
    }
    if (SUMS) {
        const double k2p896 = hilo(0x77F00000u, 0u);
        pw += pws * k2p896;
        sr += srs * k2p896;
    }
    double part[2], total[2];
    part[0] = block_sum(pw, S.scratch);
    part[1] = block_sum(sr, S.scratch);
    if (grid_sum_last_dyn<2>(part, J.partials, J.counter, gridDim.x, blockIdx.x, S.scratch, &S.last, total)) {
        if (J.chunk_out) {
            J.chunk_out[0] = total[0];
            J.chunk_out[1] = total[1];
            return;
        }
        J.stats[ST_POWER0] = n > 0 ? total[0] / (double)n : 0.0;
        J.stats[ST_SUM_RE] = total[1];
        J.stats[ST_SUM_IM] = 0.0;
        J.stats[ST_DC_RE] = n > 0 ? (double)dc_from_sum(total[1], n) : 0.0;
        J.stats[ST_DC_IM] = 0.0;
    }
}

// every (previous, current) byte quad, two at a time: the third-generation path (fast value, exact fall-back
// when flagged) against demod_one<false>; also counts the fall-backs and the fast values that would have been
// wrong without their flag
__global__ void __launch_bounds__(256) k_demod_selftest_df(const double *__restrict__ atab_g, const double *__restrict__ atab3_g,
                                                         unsigned long long *counts, unsigned *first_bad)
{
    extern __shared__ __align__(128) unsigned char lean_raw[];
    DfSmem<kLeanThreads> &S = *reinterpret_cast<DfSmem<kLeanThreads> *>(lean_raw);
    __shared__ DemodLuts L;
    df_fill(S, atab_g, atab3_g);
    {
        const float v = unpack_byte((unsigned)threadIdx.x);
        L.lutf[threadIdx.x] = v;
        L.lut[threadIdx.x] = (double)v;
        if (threadIdx.x < 9) { L.atan_d[threadIdx.x] = atan_k8(threadIdx.x); L.atan_f[threadIdx.x] = (float)atan_k8(threadIdx.x); }
    }
    __syncthreads();
    const unsigned octx_base = (unsigned)__cvta_generic_to_shared(&S.oct_exact[0][0]);
    const int slot = threadIdx.x & 15;
    unsigned bad = 0, flagged = 0, saved = 0;
    const unsigned long long half = 1ull << 31;
    const unsigned long long stride = (unsigned long long)gridDim.x * 256;
    for (unsigned long long q0 = (unsigned long long)blockIdx.x * 256 + threadIdx.x; q0 < half; q0 += stride) {
        unsigned long long qq[2] = {q0, q0 + half};
        double pr[2], pi[2], cr[2], ci[2], X[2], Y[2];
        float want[2], got[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const unsigned long long q = qq[j];
            const uchar2 prv = make_uchar2((unsigned char)(q & 255), (unsigned char)(q >> 8 & 255));
            const uchar2 cur = make_uchar2((unsigned char)(q >> 16 & 255), (unsigned char)(q >> 24 & 255));
            want[j] = demod_one<false>(L, prv, cur);
            pr[j] = S.tab[prv.x][slot].v; pi[j] = S.tab[prv.y][slot].v; cr[j] = S.tab[cur.x][slot].v; ci[j] = S.tab[cur.y][slot].v;
            X[j] = fma(pr[j], cr[j], __dmul_rn(pi[j], ci[j]));
            Y[j] = fma(ci[j], pr[j], -__dmul_rn(pi[j], cr[j]));
        }
        unsigned diff = 0u;
        df_pair(S, X[0], Y[0], X[1], Y[1], got[0], got[1], diff);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            // the production kernel redoes all eight samples of a flagged thread: here the pair
            if (diff) {
                flagged++;
                const float exact = lean_one(pr[j], pi[j], cr[j], ci[j], octx_base);
                if (__float_as_uint(exact) != __float_as_uint(got[j])) saved++;
                got[j] = exact;
            }
            if (__float_as_uint(want[j]) != __float_as_uint(got[j])) {
                bad++;
                const unsigned at = atomicAdd(first_bad, 1u);
                if (at < 63) first_bad[1 + at] = (unsigned)qq[j];
            }
        }
    }
    if (bad) atomicAdd(counts, (unsigned long long)bad);
    if (flagged) atomicAdd(counts + 1, (unsigned long long)flagged);
    if (saved) atomicAdd(counts + 2, (unsigned long long)saved);
}

// statistics of a signal whose discriminator ran in chunks (fixed order: chunk 0, 1, ...)
__global__ void k_demod_finish(const double *__restrict__ chunk_sums, int n_chunks, i64 n, double *stats)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double pw = 0.0, sr = 0.0;
    for (int c = 0; c < n_chunks; c++) { pw += chunk_sums[2 * c]; sr += chunk_sums[2 * c + 1]; }
    stats[ST_POWER0] = n > 0 ? pw / (double)n : 0.0;
    stats[ST_SUM_RE] = sr;
    stats[ST_SUM_IM] = 0.0;
    stats[ST_DC_RE] = n > 0 ? (double)dc_from_sum(sr, n) : 0.0;
    stats[ST_DC_IM] = 0.0;
}

// every (previous, current) byte quad: lean_one against demod_one<false>
__global__ void __launch_bounds__(256) k_demod_selftest(const double *__restrict__ atab_g, unsigned long long *bad,
                                                        unsigned *first_bad)
{
    extern __shared__ __align__(128) unsigned char lean_raw[];
    LeanSmem &S = *reinterpret_cast<LeanSmem *>(lean_raw);
    __shared__ DemodLuts L;
    lean_fill(S, atab_g);
    {
        const float v = unpack_byte((unsigned)threadIdx.x);
        L.lutf[threadIdx.x] = v;
        L.lut[threadIdx.x] = (double)v;
        if (threadIdx.x < 9) { L.atan_d[threadIdx.x] = atan_k8(threadIdx.x); L.atan_f[threadIdx.x] = (float)atan_k8(threadIdx.x); }
    }
    __syncthreads();
    const unsigned oct_base = (unsigned)__cvta_generic_to_shared(&S.oct[0][0]);
    const int slot = threadIdx.x & 15;
    unsigned cnt = 0;
    const unsigned long long stride = (unsigned long long)gridDim.x * 256;
    for (unsigned long long q = (unsigned long long)blockIdx.x * 256 + threadIdx.x; q < (1ull << 32); q += stride) {
        const uchar2 prv = make_uchar2((unsigned char)(q & 255), (unsigned char)(q >> 8 & 255));
        const uchar2 cur = make_uchar2((unsigned char)(q >> 16 & 255), (unsigned char)(q >> 24 & 255));
        const float want = demod_one<false>(L, prv, cur);
        const float got = lean_one(S.tab[prv.x][slot].v, S.tab[prv.y][slot].v, S.tab[cur.x][slot].v, S.tab[cur.y][slot].v,
                                   oct_base);
        if (__float_as_uint(want) != __float_as_uint(got)) {
            cnt++;
            const unsigned at = atomicAdd(first_bad, 1u);
            if (at < 63) first_bad[1 + at] = (unsigned)q;
        }
    }
    if (cnt) atomicAdd(bad, (unsigned long long)cnt);
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

// atan(k / 64) for the lean discriminator, correctly rounded by the host's libm
// (one table per device: an engine with n_devices > 1 has peers on other GPUs of this process)
constexpr int kMaxDevices = 64;
static double *g_atab_dev[kMaxDevices] = {nullptr};
static double *g_atab3_dev[kMaxDevices] = {nullptr};   // atan(k / 128), third-generation discriminator

static double *atab_here()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < kMaxDevices ? g_atab_dev[dev] : nullptr;
}

static double *atab3_here()
{
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < kMaxDevices ? g_atab3_dev[dev] : nullptr;
}

int demod_setup(cudaStream_t st)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
    if (!g_atab_dev[dev]) {
        double h[kAtanK + 2];
        for (int k = 0; k <= kAtanK + 1; k++) h[k] = atan((double)k / (double)kAtanK);
        double *d = nullptr;
        if (cudaMalloc(&d, sizeof(h)) != cudaSuccess) return -1;
        if (cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
        if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
        g_atab_dev[dev] = d;
    }
    if (!g_atab3_dev[dev]) {
        double h[kAtanK3 + 2];
        for (int k = 0; k <= kAtanK3 + 1; k++) h[k] = atan((double)k / (double)kAtanK3);
        double *d = nullptr;
        if (cudaMalloc(&d, sizeof(h)) != cudaSuccess) return -1;
        if (cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
        if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
        g_atab3_dev[dev] = d;
    }
    if (cudaFuncSetAttribute(k_demod_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LeanSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(k_demod_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LeanSmem)) != cudaSuccess ||
        cudaFuncSetAttribute(k_demod_df<512, 2, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DfSmem<512>)) != cudaSuccess ||
        cudaFuncSetAttribute(k_demod_df<512, 2, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DfSmem<512>)) != cudaSuccess ||
        cudaFuncSetAttribute(k_demod_df<512, 2, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DfSmem<512>)) != cudaSuccess ||
        cudaFuncSetAttribute(k_demod_df<384, 2, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DfSmem<384>)) != cudaSuccess ||
        cudaFuncSetAttribute(k_demod_selftest_df, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DfSmem<kLeanThreads>)) != cudaSuccess)
        return -1;
    return 0;
}

int lean_grid_x(i64 n, int n_jobs)
{
    const i64 tiles = (n + kLeanTile - 1) / kLeanTile;
    i64 cap = (2 * 148) / (n_jobs > 0 ? n_jobs : 1);  // all jobs together: one resident wave, 2 CTAs per SM
    if (cap < 1) cap = 1;
    return (int)(tiles < 1 ? 1 : (tiles > cap ? cap : tiles));
}

// fast = 0: the production discriminator (k_demod_df: double-float arctangent, exact fall-back where its accuracy
// does not decide the rounding -- the reference's bits); 2: the round-1 kernel (every sample through the f64
// arctangent; same bits, test switch); 1: f32 arctangent (not the parity path).
// TDOA_DEMOD_VARIANT (experiment switch, fast = 0 only): 1 = tiles staged by one TMA bulk copy per CTA,
// 2 = 384-thread CTAs, 3 = the two sums' conversions on the conversion pipe.
void launch_demod_fused(const SigJob *d_jobs, int n_jobs, i64 max_n, int fast, cudaStream_t st)
{
    if (fast == 2) {
        k_demod_lean<<<dim3(lean_grid_x(max_n, n_jobs), n_jobs), kLeanThreads, sizeof(LeanSmem), st>>>(d_jobs, atab_here());
    } else if (fast) {
        k_demod_fused<true><<<dim3(fast_grid_x(max_n), n_jobs), kThreads, 0, st>>>(d_jobs);
    } else {
        static const int variant = getenv("TDOA_DEMOD_VARIANT") ? atoi(getenv("TDOA_DEMOD_VARIANT")) : 0;
        auto gx = [&](int threads, int per_sm) {
            const i64 tiles = (max_n + (i64)threads * kLeanPer - 1) / ((i64)threads * kLeanPer);
            i64 cap = (i64)(per_sm * 148) / (n_jobs > 0 ? n_jobs : 1);   // all jobs together: one resident wave
            if (cap < 1) cap = 1;
            return (int)(tiles < 1 ? 1 : (tiles > cap ? cap : tiles));
        };
        const double *a = atab_here(), *a3 = atab3_here();
        switch (variant) {
            case 1: k_demod_df<512, 2, 1, 0><<<dim3(gx(512, 2), n_jobs), 512, sizeof(DfSmem<512>), st>>>(d_jobs, a, a3); break;
            case 2: k_demod_df<384, 2, 1, 1><<<dim3(gx(384, 2), n_jobs), 384, sizeof(DfSmem<384>), st>>>(d_jobs, a, a3); break;
            case 3: k_demod_df<512, 2, 0, 1><<<dim3(gx(512, 2), n_jobs), 512, sizeof(DfSmem<512>), st>>>(d_jobs, a, a3); break;
            default: k_demod_df<512, 2, 1, 1><<<dim3(gx(512, 2), n_jobs), 512, sizeof(DfSmem<512>), st>>>(d_jobs, a, a3); break;
        }
    }
}

void launch_demod_finish(const double *chunk_sums, int n_chunks, i64 n, double *stats, cudaStream_t st)
{
    k_demod_finish<<<1, 32, 0, st>>>(chunk_sums, n_chunks, n, stats);
}

// which = 1: the production discriminator (fast value + exact fall-back); which = 3: the round-1 kernel's
// arithmetic (every sample through the full-accuracy arctangent).  extra[0] = fall-backs taken, extra[1] =
// fall-backs whose fast value would have been wrong (which = 1 only).
long long demod_selftest(cudaStream_t st, unsigned *first_bad_out, int which, long long *extra)
{
    unsigned long long *d_cnt = nullptr, h_cnt[3] = {~0ull, 0, 0};
    unsigned *d_first = nullptr;
    if (cudaMalloc(&d_cnt, sizeof(h_cnt)) != cudaSuccess) return -1;
    if (cudaMalloc(&d_first, 64 * sizeof(unsigned)) != cudaSuccess) { cudaFree(d_cnt); return -1; }
    cudaMemsetAsync(d_cnt, 0, sizeof(h_cnt), st);
    cudaMemsetAsync(d_first, 0, 64 * sizeof(unsigned), st);
    if (which == 3) k_demod_selftest<<<148 * 2, 256, sizeof(LeanSmem), st>>>(atab_here(), d_cnt, d_first);
    else k_demod_selftest_df<<<148 * 2, 256, sizeof(DfSmem<kLeanThreads>), st>>>(atab_here(), atab3_here(), d_cnt, d_first);
    cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st);
    if (first_bad_out) cudaMemcpyAsync(first_bad_out, d_first, 64 * sizeof(unsigned), cudaMemcpyDeviceToHost, st);
    const cudaError_t err = cudaStreamSynchronize(st);
    cudaFree(d_cnt);
    cudaFree(d_first);
    if (extra) { extra[0] = (long long)h_cnt[1]; extra[1] = (long long)h_cnt[2]; }
    return err == cudaSuccess ? (long long)h_cnt[0] : -1;
}

void launch_boxcar_small(const SigJob *d_jobs, int n_jobs, i64 max_n, cudaStream_t st)
{
    // all jobs together: six waves of the 4 x 148 resident CTAs, each CTA walking its tiles with the next one's loads
    // in flight.  (One CTA per tile for every job -- 62 528 single-tile CTAs for 64 windows of 2 000 000 samples --
    // left nothing for the software pipeline to hide: 2.35 TB/s where the three long signals of config 2 reach 5;
    // two waves of long CTAs cost those three signals 15 %: the tail.)
    const i64 tiles = (max_n + kTile - 1) / kTile;
    i64 per_job = (148 * 24 + n_jobs - 1) / (n_jobs > 0 ? n_jobs : 1);
    if (per_job < 1) per_job = 1;
    if (per_job > 148 * 8) per_job = 148 * 8;
    const i64 gx = tiles < 1 ? 1 : (tiles > per_job ? per_job : tiles);   // <= fast_grid_x(max_n): the partial-sum slots suffice
    k_boxcar_small<<<dim3((unsigned)gx, n_jobs), kThreads, 0, st>>>(d_jobs);
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
