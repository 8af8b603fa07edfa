// Host emulation of csrc/fft_tile_core.cuh: the per-thread phases of the 8192-point
// transform run one "thread" at a time between the barriers and are compared with a
// direct DFT in double precision; then the 2 x 2 tile cross-spectrum formulas are checked
// against conj(T_a[k]) S_b[k] formed from four separate real-input DFTs.
// Built and run by tests/test_fft_core.py (no GPU).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft_tile_core.cuh"

using namespace tdoa::fft2;

static std::vector<float2> g_tab(kTab), g_tw(kN);

static void transform(const std::vector<float2> &x, std::vector<float2> &X)
{
    std::vector<float2> buf(kBuf + 64);
    for (int t = 0; t < kT; t++) {
        float2 v[32];
        for (int r = 0; r < 32; r++) v[r] = x[t + 256 * r];
        pass1_store(v, t, buf.data());
    }
    std::vector<float2> regs(kT * 32);
    for (int t = 0; t < kT; t++) {
        float2 u0[16], u1[16];
        pass_load(buf.data(), t, u0, u1);
        pass2_twiddle(u0, u1, t, g_tab.data());
        for (int r = 0; r < 16; r++) { regs[(t * 2) * 16 + r] = u0[r]; regs[(t * 2 + 1) * 16 + r] = u1[r]; }
    }
    for (int t = 0; t < kT; t++) {
        float2 u0[16], u1[16];
        for (int r = 0; r < 16; r++) { u0[r] = regs[(t * 2) * 16 + r]; u1[r] = regs[(t * 2 + 1) * 16 + r]; }
        pass2_store(u0, u1, t, buf.data());
    }
    for (int t = 0; t < kT; t++) {
        float2 u0[16], u1[16];
        pass_load(buf.data(), t, u0, u1);
        for (int r = 0; r < 16; r++) { regs[(t * 2) * 16 + r] = u0[r]; regs[(t * 2 + 1) * 16 + r] = u1[r]; }
    }
    X.assign(kBuf, make_float2(0.f, 0.f));
    for (int t = 0; t < kT; t++) {
        float2 u0[16], u1[16];
        for (int r = 0; r < 16; r++) { u0[r] = regs[(t * 2) * 16 + r]; u1[r] = regs[(t * 2 + 1) * 16 + r]; }
        pass3_compute(u0, g_tw[2 * t]);
        pass3_compute(u1, g_tw[2 * t + 1]);
        spectrum_store(u0, u1, t, X.data());
    }
    X.resize(kN);
}

static void direct(const std::vector<double> &xr, const std::vector<double> &xi, std::vector<double> &Xr, std::vector<double> &Xi)
{
    std::vector<double> cr(kN), ci(kN);
    for (int k = 0; k < kN; k++) { cr[k] = cos(-2.0 * M_PI * k / kN); ci[k] = sin(-2.0 * M_PI * k / kN); }
    Xr.assign(kN, 0); Xi.assign(kN, 0);
    for (int k = 0; k < kN; k++) {
        double sr = 0, si = 0;
        for (int n = 0; n < kN; n++) {
            const int t = (int)(((long long)k * n) & (kN - 1));
            sr += xr[n] * cr[t] - xi[n] * ci[t];
            si += xr[n] * ci[t] + xi[n] * cr[t];
        }
        Xr[k] = sr; Xi[k] = si;
    }
}

int main()
{
    for (int k = 0; k < kN; k++) {
        const double a = -2.0 * M_PI * k / kN;
        g_tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int r = 0; r < 16; r++)
        for (int k = 0; k < 32; k++) g_tab[r * 32 + k] = g_tw[(16 * k * r) & (kN - 1)];
    srand(4321);
    auto rnd = [] { return (float)rand() / RAND_MAX - 0.5f; };
    // A = t0 + i t1 (6144 samples, zero padded), B = s0 + i s1 (8192 samples)
    std::vector<float2> a(kN), b(kN), A, B;
    for (int i = 0; i < kN; i++) {
        a[i] = i < 6144 ? make_float2(rnd(), rnd()) : make_float2(0.f, 0.f);
        b[i] = make_float2(rnd(), rnd());
    }
    transform(a, A);
    transform(b, B);
    std::vector<double> ar(kN), ai(kN), br(kN), bi(kN), zero(kN, 0.0);
    for (int i = 0; i < kN; i++) { ar[i] = a[i].x; ai[i] = a[i].y; br[i] = b[i].x; bi[i] = b[i].y; }
    std::vector<double> Ar, Ai, Br, Bi;
    direct(ar, ai, Ar, Ai);
    direct(br, bi, Br, Bi);
    int rc = 0;
    {
        double e2 = 0, r2 = 0;
        for (int k = 0; k < kN; k++) {
            e2 += pow(A[k].x - Ar[k], 2) + pow(A[k].y - Ai[k], 2) + pow(B[k].x - Br[k], 2) + pow(B[k].y - Bi[k], 2);
            r2 += Ar[k] * Ar[k] + Ai[k] * Ai[k] + Br[k] * Br[k] + Bi[k] * Bi[k];
        }
        const double rel = sqrt(e2 / r2);
        printf("rel_rms_err %.3e\n", rel);
        if (!(rel < 1e-6)) rc = 1;
    }
    // cross-spectra: separate real DFTs in double
    std::vector<double> T0r, T0i, T1r, T1i, S0r, S0i, S1r, S1i;
    direct(ar, zero, T0r, T0i); direct(ai, zero, T1r, T1i); direct(br, zero, S0r, S0i); direct(bi, zero, S1r, S1i);
    const std::vector<double> *Tr[2] = {&T0r, &T1r}, *Ti[2] = {&T0i, &T1i}, *Sr[2] = {&S0r, &S1r}, *Si[2] = {&S0i, &S1i};
    double e2 = 0, r2 = 0;
    for (int k = 0; k <= kN / 2; k++) {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int nk = (kN - k) & (kN - 1);
        cross_accumulate(A[k], A[nk], B[k], B[nk], acc);
        for (int ta = 0; ta < 2; ta++)
            for (int sb = 0; sb < 2; sb++) {
                // conj(T) S
                const double tr = (*Tr[ta])[k], ti = (*Ti[ta])[k], sr = (*Sr[sb])[k], si = (*Si[sb])[k];
                const double cr = tr * sr + ti * si, ci = tr * si - ti * sr;
                const double gr = acc[2 * (2 * ta + sb)] / 4.0, gi = acc[2 * (2 * ta + sb) + 1] / 4.0;
                e2 += (gr - cr) * (gr - cr) + (gi - ci) * (gi - ci);
                r2 += cr * cr + ci * ci;
            }
    }
    const double relc = sqrt(e2 / r2);
    printf("cross_rel_rms_err %.3e\n", relc);
    if (!(relc < 2e-6)) rc = 1;
    return rc;
}
