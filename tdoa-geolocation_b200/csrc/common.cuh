// common.cuh -- shared device helpers of the B200 TDOA engine (sm_100a only).
//
// Arithmetic conventions: the reference is Go on amd64, which
// never fuses a*b+c and computes complex64 products in f64 with one rounding to f32.
// Every parity-critical expression below therefore uses the explicit round-to-nearest
// intrinsics (__fmul_rn, __fadd_rn, __dmul_rn, ...), which nvcc never contracts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tdoa {

typedef long long i64;

constexpr int kWarp = 32;

// ---------------------------------------------------------------------------
// Where a signal's samples live.  Either a window of a station's raw capture
// (interleaved uint8 I,Q; processor.go:181-201) or two planar f32 arrays.
// A raw window is at most two contiguous runs of the capture, because the
// reference signal is blocks 1 and 3 concatenated (processor.go:208-238).
struct SigSrc {
    const uint8_t *raw;  // capture bytes, or nullptr when the planes are the source
    i64 run0_start;      // raw sample index of signal sample 0
    i64 run0_len;        // signal samples served by run 0
    i64 run1_start;      // raw sample index of signal sample run0_len
    const float *re;     // planar source (raw == nullptr)
    const float *im;
};

__device__ __forceinline__ i64 raw_index(const SigSrc &s, i64 i)
{
    return i < s.run0_len ? s.run0_start + i : s.run1_start + (i - s.run0_len);
}

// processor.go:198-199  (f32(b) - 127.5) / 127.5, true f32 division.
__device__ __forceinline__ float unpack_byte(unsigned b)
{
    return __fdiv_rn(__fsub_rn((float)b, 127.5f), 127.5f);
}

__device__ __forceinline__ float2 load_sample(const SigSrc &s, i64 i)
{
    if (s.raw) {
        const uchar2 v = reinterpret_cast<const uchar2 *>(s.raw)[raw_index(s, i)];
        return make_float2(unpack_byte(v.x), unpack_byte(v.y));
    }
    return make_float2(s.re[i], s.im ? s.im[i] : 0.f);
}

// processor.go:328  f32 re*re + im*im (two f32 products, one f32 add).
__device__ __forceinline__ float mag2_f32(float re, float im)
{
    return __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
}

// ---------------------------------------------------------------------------
// Deterministic block reductions (fixed shuffle tree, fixed warp order).

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block; result valid in thread 0.  `scratch` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double *scratch)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // scratch may still be in use by a previous call
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double r = 0.0;
    if (wid == 0) {
        r = lane < nw ? scratch[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// Grid-wide deterministic sum of K values: every CTA stores its partials (valid in
// its thread 0); the CTA that arrives last adds them in a fixed order and returns
// true in its thread 0 with total[] filled.  `counter` must be zero on entry and is
// zero again on exit.  partials holds K * n_cta doubles.
template <int K>
__device__ __forceinline__ bool grid_sum_last(const double (&part)[K], double *partials, unsigned *counter,
                                              int n_cta, int cta, double *scratch, double (&total)[K])
{
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) partials[(size_t)k * n_cta + cta] = part[k];
        __threadfence();
        const unsigned t = atomicAdd(counter, 1u);
        s_last = (t == (unsigned)n_cta - 1u);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < K; k++) {
        double v = 0.0;
        // fixed order: thread t adds partials t, t+B, t+2B, ... then the block tree
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

}  // namespace tdoa
