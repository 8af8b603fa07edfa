// Host check of csrc/atan2_core.cuh against libm atan2 (glibc: < 1 ulp) on discriminator-like
// inputs: f32-valued arguments over all octants and 12 decades of ratio.
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <random>

#include "atan2_core.cuh"

int main()
{
    std::mt19937_64 g(42);
    std::uniform_real_distribution<double> u(-1.0, 1.0), lg(-6.0, 0.0);
    double max_ulp = 0.0;
    long mism32 = 0, n = 0;
    for (int it = 0; it < 4000000; it++) {
        float x = (float)(u(g) * pow(10.0, lg(g))), y = (float)(u(g) * pow(10.0, lg(g)));
        if (it % 7 == 0) y = 0.f;
        if (it % 11 == 0) x = y;
        if (it % 13 == 0) x = -y;
        if (x == 0.f && y == 0.f) continue;
        const double want = atan2((double)y, (double)x);
        const double got = tdoa::atan2_octant((double)y, (double)x, y, x, nullptr);
        const double ulp = want == 0.0 ? (got == 0.0 ? 0.0 : 1e9) : fabs(got - want) / (fabs(want) * 2.220446049250313e-16);
        if (ulp > max_ulp) max_ulp = ulp;
        if ((float)want != (float)got) mism32++;
        n++;
    }
    printf("max_err_ulp %.3f f32_mismatch %ld of %ld\n", max_ulp, mism32, n);
    return (max_ulp <= 2.0 && mism32 <= 2) ? 0 : 1;
}
