// fft_tile16_core.cuh -- the 8192-point transform of the tile kernel laid out for 512 threads with 16 points
// each (64 registers per thread, 1024 threads per CTA: eight warps per scheduler instead of the four of
// fft_tile_core.cuh, whose 32-point threads fill the register file with 512).
//
// 8192 = 16 x 16 x 32.  With n = 512 a + b and k = c + 16 d:
//   X[c + 16 d] = sum_b W_512^(b d) { W_8192^(b c) [ sum_a x[512 a + b] W_16^(a c) ] }           pass 1: thread = b
// and for every c a 512-point transform over b = 32 e + f, d = g + 16 h:
//   Y_c[g + 16 h] = sum_f W_32^(f h) { W_512^(f g) [ sum_e z_c[32 e + f] W_16^(e g) ] }          pass 2: thread = (c, f)
// The last factor is a radix-32 step, which does not fit 16 points per thread: a PAIR of adjacent lanes
// (s = bit 4 of the lane) takes it, each a 16-point transform of the samples f = s + 2 f', the final radix-2 stage
//   V[h'] = P_0[h'] + W_32^h' P_1[h'],  V[h' + 16] = P_0[h'] - W_32^h' P_1[h']                   pass 3: thread = (g, c, s)
// after exchanging eight values through one shuffle each way (partner = lane ^ 16): lane s = 0 finishes h' = 0..7, lane s = 1
// h' = 8..15.  To keep that exchange free of selects, lane 1 negates its odd inputs (its outputs come out
// rotated by 8, so both lanes keep their first eight values and send their last eight), and takes its partner's
// -i out of the data (W_32^(8 + i) = -i W_32^i), so both lanes use the compile-time twiddles W_32^i.
//
// Shared memory: rows of 32 points padded by 1 (an odd row stride); every access is a conflict-free 64-bit one:
//   after pass 1: z[c][b]     at row c * 16 + b / 32,  column b % 32
//   after pass 2: u[g][c][f]  at row g * 16 + c,       column f
//   after pass 3: the spectrum in natural order, unpadded (what the cross phase reads)
// Everything is __host__ __device__: tests/native/fft_tile16_emul.cu runs the phases one "thread" at a time on
// the CPU against a direct DFT.
#pragma once
#include "fft_tile_core.cuh"

namespace tdoa {
namespace fft16 {

using namespace fft2;   // kN, kRow, kBuf, kTab, bfly, dft, twiddle16, cmul, cross_accumulate

constexpr int kT16 = 512;   // threads per transform
constexpr int kRow16 = 33;  // float2 per padded row of 32 points: an ODD stride, so that 16 lanes walking down a column hit 16 bank pairs

// pass 1: v[a] = x[512 a + t] (already loaded); w1 = W_8192^t
TDOA_HD2 void pass1(float2 (&v)[16], int t, float2 w1, float2 *buf)
{
    dft<16>(v);
    twiddle16(v, w1);
    float2 *p = buf + (t >> 5) * kRow16 + (t & 31);
#pragma unroll
    for (int c = 0; c < 16; c++) p[c * 16 * kRow16] = v[c];
}

// pass 2: thread t = (c = t / 32, f = t % 32)
TDOA_HD2 void pass2_load(const float2 *buf, int t, float2 (&u)[16])
{
    const float2 *p = buf + (t >> 5) * 16 * kRow16 + (t & 31);
#pragma unroll
    for (int e = 0; e < 16; e++) u[e] = p[e * kRow16];
}
// tab[g * 32 + f] = W_512^(f g)
TDOA_HD2 void pass2_store(float2 (&u)[16], int t, const float2 *tab, float2 *buf)
{
    dft<16>(u);
    const float2 *w = tab + (t & 31);
    float2 *p = buf + (t >> 5) * kRow16 + (t & 31);
    p[0] = u[0];
#pragma unroll
    for (int g = 1; g < 16; g++) p[g * 16 * kRow16] = cmul(u[g], w[g * 32]);
}

// pass 3: thread t = (g = t / 32, s = (t % 32) / 16, c = t % 16): the partner of a lane is lane ^ 16.  First half: the 16-point transform of the
// samples f = s + 2 f'; on return p[0..7] are the values this lane keeps, p[8..15] the ones its partner needs.
TDOA_HD2 void pass3_first(const float2 *buf, int t, float2 (&p)[16])
{
    const int s = (t >> 4) & 1;
    const float2 *q = buf + ((t >> 5) * 16 + (t & 15)) * kRow16 + s;
    const float sgn = s ? -1.f : 1.f;   // lane 1: odd inputs negated = outputs rotated by 8
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float2 x = q[2 * i];
        p[i] = (i & 1) ? make_float2(x.x * sgn, x.y * sgn) : x;
    }
    dft<16>(p);
}

template <int I>
TDOA_HD2 void pass3_bfly(const float2 (&keep)[16], const float2 (&recv)[8], bool s, float2 (&lo)[8], float2 (&hi)[8])
{
    // lane 0: E = P_0[i] (kept), O = P_1[i] (received); lane 1: E = P_0[8 + i] (received), O = -i P_1[8 + i] (kept, rotated)
    float2 E = s ? recv[I] : keep[I];
    float2 O = s ? make_float2(keep[I].y, -keep[I].x) : recv[I];
    bfly<I>(E, O);   // E + W_32^I O, E - W_32^I O
    lo[I] = E; hi[I] = O;
    if constexpr (I + 1 < 8) pass3_bfly<I + 1>(keep, recv, s, lo, hi);
}

// second half: recv[i] = the partner's p[8 + i]; lo[i] = V[8 s + i], hi[i] = V[8 s + i + 16]
TDOA_HD2 void pass3_combine(const float2 (&keep)[16], const float2 (&recv)[8], int t, float2 (&lo)[8], float2 (&hi)[8])
{
    pass3_bfly<0>(keep, recv, (t & 16) != 0, lo, hi);
}

// spectrum in natural order: X[c + 16 g + 256 h]
TDOA_HD2 void spectrum_store16(const float2 (&lo)[8], const float2 (&hi)[8], int t, float2 *Z)
{
    const int g = t >> 5, c = t & 15, s = (t >> 4) & 1;
    float2 *p = Z + c + 16 * g + 256 * 8 * s;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        p[256 * i] = lo[i];
        p[256 * (i + 16)] = hi[i];
    }
}

}  // namespace fft16
}  // namespace tdoa
