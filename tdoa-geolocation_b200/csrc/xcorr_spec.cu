// xcorr_spec.cu -- many stations per window (BASELINE config 4: 16 stations, 120 pairs):
// every station-segment is transformed ONCE, the pairs are formed from the parked spectra.
//
// The 2 x 2 station tiles of xcorr_tile.cu transform a station again in every tile it
// appears in: 16 stations need 36 tiles = 72 transforms per segment where 16 suffice
// (15 template roles + 15 signal roles, two real signals per complex transform).  Here:
//   k_spec_fft : one CTA = one 8192-point transform (fft_tile_core.cuh) of a pair of
//                planes -- template segments zero padded after `seg` samples, signal
//                segments 8192 samples long -- written as a spectrum to global memory
//                ([segment][packed transform][8192]; parked in L2 / HBM)
//   k_spec_acc : one CTA = 64 frequency bins of one (window, lag chunk) for ALL its pairs:
//                per segment it unpacks T_i[k], S_j[k] of every station from Z[k], Z[N-k]
//                into shared memory and adds conj(T_i) S_j into per-thread registers
//                (<= 16 pairs per thread); after the last segment the pair spectra go
//                straight to k_fft_finish (xcorr_fft.cu), no partials, no reduce.
// Because the segment length is a parameter here, a search of 2049..4096 lags runs as ONE
// chunk of 4096-sample segments instead of two chunks of 6144-sample segments.
#include "fft_tile_core.cuh"
#include "kernels.h"
#include "xcorr_fft.h"

namespace tdoa {

using namespace fft2;

namespace {

constexpr int kRowSmem = (kBuf + kTab) * (int)sizeof(float2);
constexpr int kAccBins = kSpecBins;            // 64 bins per CTA
constexpr int kAccThreads = 512;
constexpr int kAccGroups = kAccThreads / kAccBins;   // 8 pair groups
constexpr int kAccPer = kSpecMaxPairs / kAccGroups;  // 16 accumulators per thread

// ---------------------------------------------------------------- transforms
__global__ void __launch_bounds__(kT, 2) k_spec_fft(const SpecFftJob *jobs, const float2 *__restrict__ tw)
{
    extern __shared__ __align__(16) float2 sm[];
    const SpecFftJob &J = jobs[blockIdx.y];
    const int seg = blockIdx.x;
    if (seg >= J.n_seg) return;
    const int t = threadIdx.x;
    float2 *buf = sm, *tab = sm + kBuf;
    for (int idx = t; idx < kTab; idx += kT) tab[idx] = tw[(16 * (idx & 31) * (idx >> 5)) & (kN - 1)];
    const float2 w1a = tw[2 * t], w1b = tw[2 * t + 1];
    {
        // z[m] = x0[i] + i x1[i], i = base + seg * stride + m, for m < seg_len and lo <= i < hi
        const i64 first = J.base + (i64)seg * J.stride;
        const float *__restrict__ x0 = J.x0, *__restrict__ x1 = J.x1;
        float2 v[32];
        if (first >= J.lo && first + J.seg_len <= J.hi && (J.seg_len & 255) == 0) {
            const int rows = J.seg_len >> 8;
#pragma unroll
            for (int r = 0; r < 32; r++) v[r] = r < rows ? make_float2(x0[first + t + 256 * r], x1[first + t + 256 * r]) : make_float2(0.f, 0.f);
        } else {
#pragma unroll
            for (int r = 0; r < 32; r++) {
                const int m = t + 256 * r;
                const i64 i = first + m;
                v[r] = (m < J.seg_len && i >= J.lo && i < J.hi) ? make_float2(x0[i], x1[i]) : make_float2(0.f, 0.f);
            }
        }
        pass1_store(v, t, buf);
    }
    __syncthreads();
    {
        float2 u0[16], u1[16];
        pass_load(buf, t, u0, u1);
        __syncthreads();
        pass2_twiddle(u0, u1, t, tab);
        pass2_store(u0, u1, t, buf);
        __syncthreads();
        pass_load(buf, t, u0, u1);
        __syncthreads();
        pass3_compute(u0, w1a);
        pass3_compute(u1, w1b);
        // X[2t + 512 r], X[2t + 1 + 512 r]: straight to global memory, 16 bytes per thread and row
        float4 *out = reinterpret_cast<float4 *>(J.out + (size_t)seg * J.out_seg_stride) + t;
#pragma unroll
        for (int r = 0; r < 16; r++) out[256 * r] = make_float4(u0[r].x, u0[r].y, u1[r].x, u1[r].y);
    }
}

// ---------------------------------------------------------------- pair accumulation
// Round 1: the next segment's spectrum values fetched into registers one segment ahead, two CTA barriers
// per segment, and per pair two byte loads (which template, which signal), a bounds test and the address
// arithmetic around four FMAs -- 10 000 cycles per segment and CTA.  Round 2:
//   - a thread's <= 16 (template row, signal row) pairs are packed into eight registers before the segment
//     loop (one PRMT per use); pairs beyond the list point at row 0 and their sums are never stored, so the
//     inner loop has no predicate;
//   - the reads run kSpecStages - 1 = 3 segments ahead through cp.async into a shared-memory ring (a thread
//     copies and later unpacks only its own items: no barrier for the ring);
//   - the unpacked station spectra are double buffered: one barrier per segment.
constexpr int kSpecStages = 4;
constexpr int kItems = kSpecMaxPacked * kAccBins / kAccThreads;   // 2 (Z[k], Z[N-k]) pairs per thread and segment
constexpr int kAccSmem = kSpecStages * kItems * 2 * kAccThreads * (int)sizeof(float2);   // 64 KB
constexpr int kRowBytes = kAccBins * (int)sizeof(float2);          // one station's 64 bins: 512 B

__global__ void __launch_bounds__(kAccThreads, 2) k_spec_acc(const SpecAccJob *jobs)
{
    extern __shared__ __align__(16) float2 ring[];           // [stage][item][k / N-k][thread]
    __shared__ __align__(16) float2 s_st[2][2 * kSpecMaxPacked][kAccBins];  // unpacked station spectra, double buffered
    const SpecAccJob &J = jobs[blockIdx.y];
    const int k0 = blockIdx.x * kAccBins;
    const int nb = min(kAccBins, kN / 2 + 1 - k0);
    if (nb <= 0) return;
    const int tid = threadIdx.x, b = tid & (kAccBins - 1), g = tid >> 6;
    const int n_pk = J.n_pk_t + J.n_pk_s;
    const int per = (J.n_pairs + kAccGroups - 1) / kAccGroups;   // pairs per thread group, <= kAccPer
    const int n_pairs = J.n_pairs, sig0 = 2 * J.n_pk_t, n_seg = J.n_seg;
    // this thread's pairs: template rows in rows_t[0..3], signal rows in rows_s[0..3], one byte each
    unsigned rows_t[kAccPer / 4], rows_s[kAccPer / 4];
#pragma unroll
    for (int w = 0; w < kAccPer / 4; w++) {
        rows_t[w] = rows_s[w] = 0u;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int q = 4 * w + c, p = g * per + q;
            const bool ok = q < per && p < n_pairs;
            // twice the row number: one PRMT (byte c into byte 1 of the result) then gives row * 512 bytes
            rows_t[w] |= (ok ? 2u * (unsigned)J.pair_t[p] : 0u) << (8 * c);
            rows_s[w] |= (ok ? 2u * (unsigned)(sig0 + J.pair_s[p]) : 0u) << (8 * c);
        }
    }
    float2 acc[kAccPer];
#pragma unroll
    for (int q = 0; q < kAccPer; q++) acc[q] = make_float2(0.f, 0.f);
    const unsigned ring_base = (unsigned)__cvta_generic_to_shared(ring);
    const unsigned st_base = (unsigned)__cvta_generic_to_shared(&s_st[0][0][0]) + (unsigned)b * (unsigned)sizeof(float2);
    auto slot = [&](int stage, int u, int which) {
        return ring_base + (unsigned)((((stage * kItems + u) * 2 + which) * kAccThreads + tid) * sizeof(float2));
    };
    auto issue = [&](int seg) {
        if (seg < n_seg) {
            const float2 *__restrict__ sp = J.spec + (size_t)seg * n_pk * kN;
            const int stage = seg % kSpecStages;
#pragma unroll
            for (int u = 0; u < kItems; u++) {
                const int item = tid + u * kAccThreads;
                const int m = item >> 6, bb = item & (kAccBins - 1);
                if (m < n_pk && bb < nb) {
                    const int k = k0 + bb;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(slot(stage, u, 0)), "l"(sp + (size_t)m * kN + k) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(slot(stage, u, 1)),
                                 "l"(sp + (size_t)m * kN + ((kN - k) & (kN - 1))) : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto lds2 = [](unsigned addr) {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
        return v;
    };
#pragma unroll
    for (int p = 0; p < kSpecStages - 1; p++) issue(p);
    for (int seg = 0; seg < n_seg; seg++) {
        issue(seg + kSpecStages - 1);   // into the stage whose items this thread unpacked one segment ago
        asm volatile("cp.async.wait_group %0;" ::"n"(kSpecStages - 1) : "memory");
        const int stage = seg % kSpecStages, half = seg & 1;
#pragma unroll
        for (int u = 0; u < kItems; u++) {
            const int item = tid + u * kAccThreads;
            const int m = item >> 6, bb = item & (kAccBins - 1);
            if (m < n_pk && bb < nb) {
                const float2 a = lds2(slot(stage, u, 0)), c = lds2(slot(stage, u, 1));
                // Z = FFT(x0 + i x1): X0[k] = (Z[k] + conj Z[N-k]) / 2, X1[k] = (Z[k] - conj Z[N-k]) / (2i)
                s_st[half][2 * m][bb] = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y - c.y));
                s_st[half][2 * m + 1][bb] = make_float2(0.5f * (a.y + c.y), 0.5f * (c.x - a.x));
            }
        }
        __syncthreads();   // this segment's station spectra are complete; the other buffer was consumed before the previous barrier
        const unsigned base = st_base + (unsigned)half * (unsigned)(2 * kSpecMaxPacked * kRowBytes);
        unsigned last = 0xffffffffu;
        float2 ti = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < kAccPer; q++) {
            // volatile: kept inside the loop (hoisted, the 32 row offsets would not fit the 64 registers)
            unsigned it, js;
            asm volatile("prmt.b32 %0, %1, 0, %2;" : "=r"(it) : "r"(rows_t[q >> 2]), "r"(0x4404 | ((q & 3) << 4)));
            asm volatile("prmt.b32 %0, %1, 0, %2;" : "=r"(js) : "r"(rows_s[q >> 2]), "r"(0x4404 | ((q & 3) << 4)));
            if (it != last) { ti = lds2(base + it); last = it; }   // pairs are listed template by template
            const float2 sj = lds2(base + js);
            // conj(T) S
            acc[q].x = fmaf(ti.x, sj.x, fmaf(ti.y, sj.y, acc[q].x));
            acc[q].y = fmaf(ti.x, sj.y, fmaf(-ti.y, sj.x, acc[q].y));
        }
    }
    if (b < nb) {
#pragma unroll
        for (int q = 0; q < kAccPer; q++) {
            const int p = g * per + q;
            if (q < per && p < n_pairs) J.spectrum[p][k0 + b] = acc[q];
        }
    }
}

}  // namespace

int spec_setup()
{
    return cudaFuncSetAttribute(k_spec_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowSmem) == cudaSuccess &&
                   cudaFuncSetAttribute(k_spec_acc, cudaFuncAttributeMaxDynamicSharedMemorySize, kAccSmem) == cudaSuccess
               ? 0 : -1;
}

void launch_spec_fft(const SpecFftJob *d_jobs, int n_jobs, int max_seg, const float2 *d_tw, cudaStream_t st)
{
    if (n_jobs <= 0 || max_seg <= 0) return;
    k_spec_fft<<<dim3(max_seg, n_jobs), kT, kRowSmem, st>>>(d_jobs, d_tw);
}

void launch_spec_acc(const SpecAccJob *d_jobs, int n_jobs, cudaStream_t st)
{
    if (n_jobs <= 0) return;
    k_spec_acc<<<dim3((kN / 2 + 1 + kAccBins - 1) / kAccBins, n_jobs), kAccThreads, kAccSmem, st>>>(d_jobs);
}

}  // namespace tdoa
