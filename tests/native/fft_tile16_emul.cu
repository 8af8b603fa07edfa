// Host emulation of csrc/fft_tile16_core.cuh (512 threads x 16 points, pair-of-lanes radix-32 step): the
// per-thread phases run one "thread" at a time between the barriers -- the shuffle of pass 3 is a swap between the
// two emulated lanes -- and the spectrum is compared with a direct DFT in double precision.  Built and run by
// tests/test_fft_core.py (no GPU).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft_tile16_core.cuh"

using namespace tdoa::fft16;

int main()
{
    std::vector<float2> tw(kN), tab(kTab);
    for (int k = 0; k < kN; k++) {
        const double a = -2.0 * M_PI * k / kN;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int r = 0; r < 16; r++)
        for (int k = 0; k < 32; k++) tab[r * 32 + k] = tw[(16 * k * r) & (kN - 1)];
    srand(99);
    auto rnd = [] { return (float)rand() / RAND_MAX - 0.5f; };
    std::vector<float2> x(kN);
    for (int i = 0; i < kN; i++) x[i] = i < 6144 ? make_float2(rnd(), rnd()) : make_float2(0.f, 0.f);
    std::vector<float2> buf(kBuf + 64, make_float2(0.f, 0.f));
    for (int t = 0; t < kT16; t++) {
        float2 v[16];
        for (int a = 0; a < 16; a++) v[a] = x[512 * a + t];
        pass1(v, t, tw[t], buf.data());
    }
    std::vector<float2> regs(kT16 * 16);
    for (int t = 0; t < kT16; t++) {
        float2 u[16];
        pass2_load(buf.data(), t, u);
        for (int e = 0; e < 16; e++) regs[t * 16 + e] = u[e];
    }
    for (int t = 0; t < kT16; t++) {
        float2 u[16];
        for (int e = 0; e < 16; e++) u[e] = regs[t * 16 + e];
        pass2_store(u, t, tab.data(), buf.data());
    }
    for (int t = 0; t < kT16; t++) {
        float2 p[16];
        pass3_first(buf.data(), t, p);
        for (int i = 0; i < 16; i++) regs[t * 16 + i] = p[i];
    }
    std::vector<float2> Z(kN);
    for (int t = 0; t < kT16; t++) {
        float2 keep[16], recv[8], lo[8], hi[8];
        for (int i = 0; i < 16; i++) keep[i] = regs[t * 16 + i];
        for (int i = 0; i < 8; i++) recv[i] = regs[(t ^ 16) * 16 + 8 + i];   // __shfl_xor(send[i], 16)
        pass3_combine(keep, recv, t, lo, hi);
        spectrum_store16(lo, hi, t, Z.data());
    }
    // direct DFT in double
    std::vector<double> cr(kN), ci(kN);
    for (int k = 0; k < kN; k++) { cr[k] = cos(-2.0 * M_PI * k / kN); ci[k] = sin(-2.0 * M_PI * k / kN); }
    double e2 = 0, r2 = 0;
    for (int k = 0; k < kN; k++) {
        double sr = 0, si = 0;
        for (int n = 0; n < 6144; n++) {
            const int t = (int)(((long long)k * n) & (kN - 1));
            sr += x[n].x * cr[t] - x[n].y * ci[t];
            si += x[n].x * ci[t] + x[n].y * cr[t];
        }
        e2 += pow(Z[k].x - sr, 2) + pow(Z[k].y - si, 2);
        r2 += sr * sr + si * si;
    }
    const double rel = sqrt(e2 / r2);
    printf("rel_rms_err %.3e\n", rel);
    return rel < 1e-6 ? 0 : 1;
}
