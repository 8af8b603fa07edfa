// fft_core.cuh -- the 8192-point complex FFT that lives in one CTA's shared memory.
//
// Three Stockham passes (radix 32, 16, 16) over 256 threads, 32 points per thread:
// reads are always stride N/R across consecutive threads (bank-conflict free), writes
// go to the auto-sort position so the result is in natural order.  Butterflies are
// fully unrolled radix-2 DIF networks in registers with compile-time twiddles; the
// inter-pass twiddles W^(r*k) come from one table lookup (W^k) and a short product
// tree.  Everything is __host__ __device__ so tests/native/fft_emul.cu can run the
// exact per-thread phases on the CPU (one "thread" after another between the
// barriers) and compare them with a direct DFT.
#pragma once
#include <cuda_runtime.h>

namespace tdoa {
namespace fft {

#define TDOA_HD __host__ __device__ __forceinline__

constexpr int kN = 8192;        // FFT size (complex points)
constexpr int kThreads = 256;   // threads per CTA
constexpr int kPad = kN + kN / 32;  // padded shared-memory length in float2

TDOA_HD int pad(int i) { return i + (i >> 5); }

TDOA_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
TDOA_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
TDOA_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// W_32^k = exp(-2 pi i k / 32), k = 0..15
__host__ __device__ constexpr float w32_re(int k)
{
    constexpr float c[16] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                             0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f,
                             0.0f, -0.19509032201612825f, -0.38268343236508977f, -0.55557023301960218f,
                             -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
    return c[k];
}
__host__ __device__ constexpr float w32_im(int k)
{
    constexpr float s[16] = {0.0f, -0.19509032201612825f, -0.38268343236508977f, -0.55557023301960218f,
                             -0.70710678118654752f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f,
                             -1.0f, -0.98078528040323043f, -0.92387953251128674f, -0.83146961230254524f,
                             -0.70710678118654752f, -0.55557023301960218f, -0.38268343236508977f, -0.19509032201612825f};
    return s[k];
}

// v * W_32^K with the trivial cases folded at compile time
template <int K>
TDOA_HD float2 mul_w32(float2 v)
{
    if constexpr (K == 0) {
        return v;
    } else if constexpr (K == 8) {  // -i
        return make_float2(v.y, -v.x);
    } else if constexpr (K == 4) {  // (1 - i)/sqrt2
        constexpr float h = 0.70710678118654752f;
        return make_float2((v.x + v.y) * h, (v.y - v.x) * h);
    } else if constexpr (K == 12) {  // (-1 - i)/sqrt2
        constexpr float h = 0.70710678118654752f;
        return make_float2((v.y - v.x) * h, -(v.x + v.y) * h);
    } else {
        return cmul(v, make_float2(w32_re(K), w32_im(K)));
    }
}

// One DIF stage of span LEN over an R-point register array.
template <int R, int LEN, int B, int K>
struct DifStage {
    static TDOA_HD void run(float2 (&v)[R])
    {
        constexpr int i = B * LEN + K;
        const float2 a = v[i], b = v[i + LEN / 2];
        v[i] = cadd(a, b);
        v[i + LEN / 2] = mul_w32<K * (32 / LEN)>(csub(a, b));
        if constexpr (K + 1 < LEN / 2) DifStage<R, LEN, B, K + 1>::run(v);
        else if constexpr (B + 1 < R / LEN) DifStage<R, LEN, B + 1, 0>::run(v);
    }
};

template <int R, int LEN>
TDOA_HD void dif_all(float2 (&v)[R])
{
    DifStage<R, LEN, 0, 0>::run(v);
    if constexpr (LEN > 2) dif_all<R, LEN / 2>(v);
}

template <int R>
__host__ __device__ constexpr int bitrev(int i)
{
    int r = 0;
    for (int b = 1; b < R; b <<= 1) { r = (r << 1) | (i & 1); i >>= 1; }
    return r;
}

template <int R, int I>
TDOA_HD void unscramble(const float2 (&v)[R], float2 (&o)[R])
{
    o[bitrev<R>(I)] = v[I];
    if constexpr (I + 1 < R) unscramble<R, I + 1>(v, o);
}

// forward DFT of R points (R = 2..32), natural order in and out
template <int R>
TDOA_HD void dft(float2 (&v)[R])
{
    dif_all<R, R>(v);
    float2 o[R];
    unscramble<R, 0>(v, o);
#pragma unroll
    for (int i = 0; i < R; i++) v[i] = o[i];
}

// v[r] *= w^r for r = 1..15 given w (product tree of depth <= 4)
TDOA_HD void twiddle16(float2 (&v)[16], float2 w1)
{
    const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
    const float2 w5 = cmul(w4, w1), w6 = cmul(w3, w3), w7 = cmul(w4, w3), w8 = cmul(w4, w4);
    v[1] = cmul(v[1], w1);
    v[2] = cmul(v[2], w2);
    v[3] = cmul(v[3], w3);
    v[4] = cmul(v[4], w4);
    v[5] = cmul(v[5], w5);
    v[6] = cmul(v[6], w6);
    v[7] = cmul(v[7], w7);
    v[8] = cmul(v[8], w8);
    v[9] = cmul(v[9], cmul(w8, w1));
    v[10] = cmul(v[10], cmul(w8, w2));
    v[11] = cmul(v[11], cmul(w8, w3));
    v[12] = cmul(v[12], cmul(w8, w4));
    v[13] = cmul(v[13], cmul(w8, w5));
    v[14] = cmul(v[14], cmul(w8, w6));
    v[15] = cmul(v[15], cmul(w8, w7));
}

// ---- the per-thread phases; barriers go between them (see xcorr_fft.cu)
//
// pass 1: radix 32, Ns = 1.  in[tid + 256 r] -> out[32 tid + r]
TDOA_HD void pass1_store(float2 (&v)[32], int tid, float2 *sm)
{
    dft<32>(v);
#pragma unroll
    for (int r = 0; r < 32; r++) sm[pad(tid * 32 + r)] = v[r];
}

// passes 2 and 3 read in[j + 512 r], r = 0..15, for the two butterflies j = tid, tid + 256
TDOA_HD void pass_load16(const float2 *sm, int j, float2 (&u)[16])
{
#pragma unroll
    for (int r = 0; r < 16; r++) u[r] = sm[pad(j + 512 * r)];
}

// pass 2: radix 16, Ns = 32: twiddle W_512^(r k), k = j % 32; out[(j/32)*512 + k + 32 r]
TDOA_HD void pass2_store(float2 (&u)[16], int j, const float2 *tw, float2 *sm)
{
    const int k = j & 31;
    twiddle16(u, tw[16 * k]);  // W_512^k = W_8192^(16 k)
    dft<16>(u);
    const int j0 = (j >> 5) * 512 + k;
#pragma unroll
    for (int r = 0; r < 16; r++) sm[pad(j0 + 32 * r)] = u[r];
}

// pass 2 with the twiddles W_512^(r k) pre-tabulated per lane: tab[r * 32 + k], k = j % 32
// (k is the lane id for both butterflies of a thread, so the reads are conflict free)
TDOA_HD void pass2_store_tab(float2 (&u)[16], int j, const float2 *tab, float2 *sm)
{
    const int k = j & 31;
#pragma unroll
    for (int r = 1; r < 16; r++) u[r] = cmul(u[r], tab[r * 32 + k]);
    dft<16>(u);
    const int j0 = (j >> 5) * 512 + k;
#pragma unroll
    for (int r = 0; r < 16; r++) sm[pad(j0 + 32 * r)] = u[r];
}

// pass 3: radix 16, Ns = 512: twiddle W_8192^(r j); result u[r] = X[j + 512 r]
TDOA_HD void pass3_compute(float2 (&u)[16], int j, const float2 *tw)
{
    twiddle16(u, tw[j]);
    dft<16>(u);
}

}  // namespace fft
}  // namespace tdoa
