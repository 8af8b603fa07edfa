// Host emulation of the shared-memory FFT phases in csrc/fft_core.cuh: runs every
// "thread" of a CTA one after another between the barriers and compares the result with
// a direct DFT in double precision.  Built and run by tests/test_fft_core.py (no GPU).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft_core.cuh"

using namespace tdoa::fft;

int main()
{
    std::vector<float2> x(kN), sm(kPad), tw(kN);
    std::vector<double> xr(kN), xi(kN);
    srand(1234);
    for (int i = 0; i < kN; i++) {
        x[i].x = (float)rand() / RAND_MAX - 0.5f;
        x[i].y = (i < 6144) ? (float)rand() / RAND_MAX - 0.5f : 0.f;
        xr[i] = x[i].x; xi[i] = x[i].y;
    }
    for (int k = 0; k < kN; k++) {
        const double a = -2.0 * M_PI * k / kN;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    // phases (variant 0: twiddle product tree in pass 2; variant 1: per-lane table)
    std::vector<float2> tab(512);
    for (int r = 0; r < 16; r++)
        for (int k = 0; k < 32; k++) tab[r * 32 + k] = tw[(16 * k * r) & (kN - 1)];
    int rc = 0;
    for (int variant = 0; variant < 2; variant++) {
    std::vector<float2> X(kN);
    {
        for (int tid = 0; tid < kThreads; tid++) {
            float2 v[32];
            for (int r = 0; r < 32; r++) v[r] = x[tid + 256 * r];
            pass1_store(v, tid, sm.data());
        }
        std::vector<float2> regs(kThreads * 32);
        for (int tid = 0; tid < kThreads; tid++)
            for (int b = 0; b < 2; b++) {
                float2 u[16];
                pass_load16(sm.data(), tid + 256 * b, u);
                for (int r = 0; r < 16; r++) regs[(tid * 2 + b) * 16 + r] = u[r];
            }
        for (int tid = 0; tid < kThreads; tid++)
            for (int b = 0; b < 2; b++) {
                float2 u[16];
                for (int r = 0; r < 16; r++) u[r] = regs[(tid * 2 + b) * 16 + r];
                if (variant) pass2_store_tab(u, tid + 256 * b, tab.data(), sm.data());
                else pass2_store(u, tid + 256 * b, tw.data(), sm.data());
            }
        for (int tid = 0; tid < kThreads; tid++)
            for (int b = 0; b < 2; b++) {
                float2 u[16];
                const int j = tid + 256 * b;
                pass_load16(sm.data(), j, u);
                pass3_compute(u, j, tw.data());
                for (int r = 0; r < 16; r++) X[j + 512 * r] = u[r];
            }
    }
    // direct DFT (double) via the twiddle recurrence-free formula
    std::vector<double> cr(kN), ci(kN);
    for (int k = 0; k < kN; k++) { cr[k] = cos(-2.0 * M_PI * k / kN); ci[k] = sin(-2.0 * M_PI * k / kN); }
    double err2 = 0, ref2 = 0, maxerr = 0;
    for (int k = 0; k < kN; k++) {
        double sr = 0, si = 0;
        for (int n = 0; n < kN; n++) {
            const int t = (int)(((long long)k * n) & (kN - 1));
            sr += xr[n] * cr[t] - xi[n] * ci[t];
            si += xr[n] * ci[t] + xi[n] * cr[t];
        }
        const double dr = X[k].x - sr, di = X[k].y - si;
        err2 += dr * dr + di * di;
        ref2 += sr * sr + si * si;
        maxerr = fmax(maxerr, sqrt(dr * dr + di * di));
    }
    const double rel = sqrt(err2 / ref2);
    printf("rel_rms_err %.3e max_abs_err %.3e rms_ref %.3e\n", rel, maxerr, sqrt(ref2 / kN));
    if (!(rel < 2e-6)) rc = 1;
    }
    return rc;
}
