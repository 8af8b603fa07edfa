// seqsum_core.cuh -- the reference's SEQUENTIAL f32 accumulator, s <- RN_f32(s + x[i]) for
// i = 0..n-1 (removeDCBias, processor.go:304-309), evaluated chunk-parallel and still bit
// for bit.
//
// While the running sum s stays in one binade [2^e, 2^(e+1)) (same sign), it is an integer
// multiple m of u = 2^(e-23) and every addition is integer arithmetic:
//     x / u = q + f  (q = floor, 0 <= f < 1)        m' = m + q + r,
//     r = 1 if f > 1/2, 0 if f < 1/2, and on a tie (f = 1/2) whatever makes m' even.
// Only ties look at the state, and only at its parity.  So a chunk of samples, given a
// guess e of the binade, reduces to: the total increment for an even and for an odd start
// (d[0], d[1]), and how far the sum wanders on the way (to prove it never left the
// binade).  chunk_analyse() computes that for every chunk in parallel; chunk_apply() then
// walks the chunks in order with the TRUE sum: where the guess and the excursion check hold
// the chunk costs O(1), elsewhere it is added sample by sample.  The guess comes from exact
// f64 prefix sums, which track the f32 chain to within its accumulated rounding error.
// Everything is __host__ __device__: tests/native/seqsum_emul.cu checks it against the plain
// loop on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace tdoa {
namespace seqsum {

#define TDOA_SQ __host__ __device__ __forceinline__

constexpr int kChunk = 256;   // samples per chunk.  Runs of good chunks cost one scan whatever the size; a bad chunk costs kChunk
                              // dependent additions plus a new scan (worth ~100 additions): 64-sample chunks fall back on 40 % fewer
                              // samples near a binade boundary but rescan 2.4 times as often -- measured on the emulation, no gain

struct ChunkInfo {
    long long d[2];        // total increment of m, start parity even / odd
    long long min_floor;   // min over steps of (m - m_in) + q          (sum before rounding, floor)
    long long max_ceil;    // max over steps of (m - m_in) + q + (f > 0) (sum before rounding, ceiling)
    long long min_after;   // min / max over steps of m' - m_in
    long long max_after;
    int e;                 // binade guessed: 2^e <= |s| < 2^(e+1)
    int valid;             // 0: some sample does not fit the integer picture for this guess
    long long pad;         // 64 bytes: four 16-byte cp.async granules
};
static_assert(sizeof(ChunkInfo) == 64, "ChunkInfo is staged with 16-byte copies");

TDOA_SQ uint32_t f2u(float x)
{
#ifdef __CUDA_ARCH__
    return __float_as_uint(x);
#else
    uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
TDOA_SQ float u2f(uint32_t u)
{
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float x; memcpy(&x, &u, 4); return x;
#endif
}

// binade of a finite normal f32 (floor(log2 |x|)); -1000 for zero / denormal / inf / nan
TDOA_SQ int binade(float x)
{
    const int be = (int)((f2u(x) >> 23) & 0xffu);
    return (be == 0 || be == 255) ? -1000 : be - 127;
}

// one chunk under the guess e
TDOA_SQ ChunkInfo chunk_analyse(const float *x, int count, int e)
{
    ChunkInfo c;
    c.d[0] = c.d[1] = 0;
    c.min_floor = 0; c.max_ceil = 0; c.min_after = 0; c.max_after = 0;
    c.e = e;
    c.pad = 0;
    c.valid = e > -1000 ? 1 : 0;
    if (!c.valid) return c;
    for (int i = 0; i < count; i++) {
        const uint32_t b = f2u(x[i]);
        const int be = (int)((b >> 23) & 0xffu);
        if ((b & 0x7fffffffu) == 0) continue;           // +-0 adds nothing
        if (be == 0 || be == 255) { c.valid = 0; break; }   // denormal / inf / nan: not here
        const int sh = e - (be - 127);                  // x / u = mx * 2^-sh
        if (sh < 0) { c.valid = 0; break; }             // sample above the sum's binade: the sum would leave it
        long long mx = (long long)((b & 0x7fffffu) | 0x800000u);
        if (b >> 31) mx = -mx;
        long long q, frac, half;
        if (sh == 0) { q = mx; frac = 0; half = 1; }
        else if (sh >= 40) { q = mx < 0 ? -1 : 0; frac = 1; half = 4; if (mx < 0) frac = 7; }   // |x/u| < 2^-15: f tiny or 1 - tiny
        else { q = mx >> sh; frac = mx - q * (1ll << sh); half = 1ll << (sh - 1); }   // arithmetic shift = floor
        const long long up = frac > half ? 1 : 0;
        const bool tie = sh > 0 && sh < 40 && frac == half;
        const long long ceil_add = frac != 0 ? 1 : 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const long long before = c.d[h] + q;
            long long r = up;
            if (tie) r = (before + h) & 1;             // parity of m_in + before: round to even
            const long long after = before + r;
            if (before < c.min_floor) c.min_floor = before;
            if (before + ceil_add > c.max_ceil) c.max_ceil = before + ceil_add;
            if (after < c.min_after) c.min_after = after;
            if (after > c.max_after) c.max_after = after;
            c.d[h] = after;
        }
    }
    return c;
}

// Binade guesses of a chunk: the one the exact prefix sum suggests and its neighbours (the f32
// chain drifts from the exact sum by its accumulated rounding error, and the sum lingers around
// powers of two like any random walk).
constexpr int kGuesses = 3;
TDOA_SQ int guess_binade(double prefix, int k)
{
    const int e = binade((float)prefix);
    return e <= -1000 ? e : e + k - 1;
}

// What the in-order walk needs of one (chunk, guess), 32 bytes, 32-bit arithmetic only: with the
// running sum s = m u (m signed, 2^23 <= |m| < 2^24) the chunk may be skipped iff the binade is
// e and  pos_lo <= m <= pos_hi  (s > 0)  or  neg_lo <= m <= neg_hi  (s < 0); then m += d[m & 1].
struct ChunkRule {
    int e;                 // -1000: never applies
    int d[2];
    int pos_lo, pos_hi;
    int neg_lo, neg_hi;
    int pad;
};
static_assert(sizeof(ChunkRule) == 32, "ChunkRule is staged with 16-byte copies");

TDOA_SQ ChunkRule chunk_rule(const ChunkInfo &c)
{
    ChunkRule r;
    r.e = -1000; r.d[0] = r.d[1] = 0; r.pos_lo = 1; r.pos_hi = 0; r.neg_lo = 1; r.neg_hi = 0; r.pad = 0;
    const long long lim = 1ll << 26;   // anything that wanders further cannot stay in the binade anyway
    if (!c.valid || c.min_floor < -lim || c.max_ceil > lim || c.min_after < -lim || c.max_after > lim) return r;
    r.e = c.e;
    r.d[0] = (int)c.d[0]; r.d[1] = (int)c.d[1];
    const long long lo = 1ll << 23, hi = (1ll << 24) - 1;
    r.pos_lo = (int)(lo - c.min_floor); r.pos_hi = (int)(hi - c.max_after);
    r.neg_lo = (int)(-hi - c.min_after); r.neg_hi = (int)(-lo - c.max_ceil);
    return r;
}

// the walk's step on the packed rules; *done = 0 when no rule applied (the caller adds the samples)
TDOA_SQ float rule_apply(float s, const ChunkRule *rules, int *done)
{
    const uint32_t b = f2u(s);
    const int es = (int)((b >> 23) & 0xffu) - 127;
    int m = (int)((b & 0x7fffffu) | 0x800000u);
    const bool neg = (b >> 31) != 0;
    if (neg) m = -m;
    *done = 0;
#pragma unroll
    for (int k = 0; k < kGuesses; k++) {
        const ChunkRule &r = rules[k];
        if (r.e != es) continue;
        const bool ok = neg ? (m >= r.neg_lo && m <= r.neg_hi) : (m >= r.pos_lo && m <= r.pos_hi);
        if (ok) {
            int mo = m + r.d[m & 1];
            if (mo < 0) mo = -mo;
            *done = 1;
            return u2f((b & 0xff800000u) | ((uint32_t)mo & 0x7fffffu));   // same sign, same binade
        }
        break;
    }
    return s;
}

// ---- runs of chunks.  With the binade and the sign of the sum fixed, a chunk rule is a partial
// map on the signed mantissa:  m -> m + d[m & 1]  for  lo[m & 1] <= m <= hi[m & 1]  (a "hop").
// Hops are closed under composition -- the domain of "A then B" is the interval of A's domain
// whose image lies in B's, per start parity -- and composition is associative, so the hops of a
// batch of chunks can be combined by a parallel prefix scan: hop k of the scan takes the sum
// across chunks 0..k at once, and its domain says exactly whether every one of those chunks
// keeps the sum inside the binade.  The walk then costs one scan per run of good chunks instead
// of one dependent step per chunk.
struct Hop {
    int d[2], lo[2], hi[2];
};
constexpr int kHopFar = 1 << 30;   // identity's bounds: |lo|, |hi|, |d| of any non-empty hop stay below 2^29

TDOA_SQ Hop hop_identity()
{
    Hop h;
    h.d[0] = h.d[1] = 0; h.lo[0] = h.lo[1] = -kHopFar; h.hi[0] = h.hi[1] = kHopFar;
    return h;
}
TDOA_SQ Hop hop_empty()
{
    Hop h;
    h.d[0] = h.d[1] = 0; h.lo[0] = h.lo[1] = 1; h.hi[0] = h.hi[1] = 0;
    return h;
}
// the hop of a chunk for a sum in binade es with the given sign (empty if no guess was made for it)
TDOA_SQ Hop hop_from_rules(const ChunkRule *rules, int es, bool neg)
{
    Hop h = hop_empty();
#pragma unroll
    for (int k = 0; k < kGuesses; k++) {
        const ChunkRule &r = rules[k];
        if (r.e != es || r.e <= -1000) continue;
        h.d[0] = r.d[0]; h.d[1] = r.d[1];
        h.lo[0] = h.lo[1] = neg ? r.neg_lo : r.pos_lo;
        h.hi[0] = h.hi[1] = neg ? r.neg_hi : r.pos_hi;
        break;
    }
    return h;
}
// a first, then b
TDOA_SQ Hop hop_compose(const Hop &a, const Hop &b)
{
    Hop c;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const bool odd = ((h + a.d[h]) & 1) != 0;   // parity after a (two's complement: also right for negative d and m)
        const int bd = odd ? b.d[1] : b.d[0], blo = odd ? b.lo[1] : b.lo[0], bhi = odd ? b.hi[1] : b.hi[0];   // selects, not indexing: registers
        const int d = a.d[h] + bd;
        int lo = blo - a.d[h], hi = bhi - a.d[h];
        if (a.lo[h] > lo) lo = a.lo[h];
        if (a.hi[h] < hi) hi = a.hi[h];
        const bool empty = lo > hi || a.lo[h] > a.hi[h] || blo > bhi;
        c.d[h] = empty ? 0 : d;
        c.lo[h] = empty ? 1 : lo;
        c.hi[h] = empty ? 0 : hi;
    }
    return c;
}
// signed mantissa and binade of a normal f32 (es = -127 for zero / denormal: no hop is made for that)
TDOA_SQ int mantissa_signed(float s, int *es, bool *neg)
{
    const uint32_t b = f2u(s);
    *es = (int)((b >> 23) & 0xffu) - 127;
    *neg = (b >> 31) != 0;
    const int m = (int)((b & 0x7fffffu) | 0x800000u);
    return *neg ? -m : m;
}
TDOA_SQ bool hop_admits(const Hop &h, int m)
{
    const bool odd = (m & 1) != 0;
    return m >= (odd ? h.lo[1] : h.lo[0]) && m <= (odd ? h.hi[1] : h.hi[0]);
}
// the sum after a hop that admits m: same sign, same binade
TDOA_SQ float hop_apply(const Hop &h, float s, int m)
{
    int mo = m + ((m & 1) ? h.d[1] : h.d[0]);
    if (mo < 0) mo = -mo;
    return u2f((f2u(s) & 0xff800000u) | ((uint32_t)mo & 0x7fffffu));
}

// s after the chunk, exactly: through the summary whose guess provably applies, else sample by sample
TDOA_SQ float chunk_apply(float s, const ChunkInfo *guesses, const float *x, int count, int *fast)
{
    const int es = binade(s);
#pragma unroll
    for (int k = 0; k < kGuesses; k++) {
        const ChunkInfo &c = guesses[k];
        if (!(c.valid && es == c.e)) continue;
        const uint32_t b = f2u(s);
        long long m = (long long)((b & 0x7fffffu) | 0x800000u);
        const bool neg = (b >> 31) != 0;
        if (neg) m = -m;
        const int h = (int)(m & 1);
        const long long lo = 1ll << 23, hi = (1ll << 24) - 1;
        const bool ok = neg ? (m + c.max_ceil <= -lo && m + c.min_after >= -hi)
                            : (m + c.min_floor >= lo && m + c.max_after <= hi);
        if (ok) {
            long long mo = m + c.d[h];
            const uint32_t sign = mo < 0 ? 0x80000000u : 0u;
            if (mo < 0) mo = -mo;
            if (fast) (*fast)++;
            return u2f(sign | ((uint32_t)(c.e + 127) << 23) | ((uint32_t)mo & 0x7fffffu));
        }
        break;   // the right binade, but the sum leaves it inside the chunk
    }
    for (int i = 0; i < count; i++) {
#ifdef __CUDA_ARCH__
        s = __fadd_rn(s, x[i]);
#else
        volatile float t = s + x[i];
        s = t;
#endif
    }
    return s;
}

}  // namespace seqsum
}  // namespace tdoa
