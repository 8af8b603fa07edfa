// atan2_core.cuh -- the arctangent of the FM discriminator, sized for the job.
//
// The shipped reference binary evaluates math.Atan2(f64(Im p), f64(Re p)) per sample and
// rounds it to f32 (ELF 0x49d120).  Its arguments are f32 values of moderate magnitude
// (|p|^2 > 1e-10 is gated before the call), never NaN/Inf/denormal, so the general
// libdevice atan2 -- ~100 instructions with all its special cases -- is replaced by:
//   - octant reduction with min/max,
//   - a table step  atan(q) = atan(k/8) + atan((mn - c mx)/(mx + c mn)),  c = k/8, which
//     needs one division and leaves |z| <= ~1/15,
//   - a 7-term odd polynomial (truncation < 2^-60 relative).
// Accuracy ~1 ulp of f64 (checked against libm in tests/native/atan2_check.cu); after
// the rounding to f32 it is the reference's value except where two correctly working
// f64 atan2 implementations may themselves differ (a result within ~1e-16 relative of an
// f32 rounding boundary, probability ~1e-8 per sample).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace tdoa {

#define TDOA_A2_HD __host__ __device__ __forceinline__

// atan(k/8), k = 0..8, correctly rounded doubles
__host__ __device__ constexpr double atan_k8(int k)
{
    constexpr double t[9] = {0.0,
                             0.12435499454676143503,
                             0.24497866312686415417,
                             0.35877067027057222040,
                             0.46364760900080611621,
                             0.55859931534356243597,
                             0.64350110879328438680,
                             0.71882999962162450542,
                             0.78539816339744830962};
    return t[k];
}

#ifdef __CUDACC__
// polynomial of atan(z) = z + z w P(w), w = z^2: -1/3, 1/5, -1/7, 1/9, -1/11, 1/13
// (constant-bank operands: no immediate moves in the inner loop)
__constant__ double kAtanPoly[6] = {-1.0 / 3.0, 1.0 / 5.0, -1.0 / 7.0, 1.0 / 9.0, -1.0 / 11.0, 1.0 / 13.0};
#endif

// num / den for den in [1e-6, 4]: MUFU.RCP64H seed (>= 20 bits), two Newton steps and one
// residual correction -- branch-free, error < 1 ulp.  The host build divides.
TDOA_A2_HD double div_pos(double num, double den)
{
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
    r = fma(r, fma(-den, r, 1.0), r);
    r = fma(r, fma(-den, r, 1.0), r);
    const double q = num * r;
    return fma(r, fma(-den, q, num), q);
#else
    return num / den;
#endif
}

// y, x: finite, not both zero; yf, xf: the same values in f32 (only used to pick the
// table entry).  `table` points at the 9 atan(k/8) values (shared memory on the device;
// nullptr on the host selects the constexpr table).
TDOA_A2_HD double atan2_octant(double y, double x, float yf, float xf, const double *table)
{
    const double ax = fabs(x), ay = fabs(y);
    const double mx = fmax(ax, ay), mn = fmin(ax, ay);
    // k = round(8 mn/mx): an f32 estimate is plenty, the polynomial absorbs the slack
    const float axf = fabsf(xf), ayf = fabsf(yf);
#ifdef __CUDA_ARCH__
    const float qf = __fdividef(fminf(axf, ayf), fmaxf(axf, ayf));
#else
    const float qf = fminf(axf, ayf) / fmaxf(axf, ayf);
#endif
    int k = (int)(qf * 8.0f + 0.5f);
    k = k < 0 ? 0 : (k > 8 ? 8 : k);
    const double c = (double)k * 0.125;
    const double num = fma(-c, mx, mn);  // mn - c mx
    const double den = fma(c, mn, mx);   // mx + c mn
    const double z = div_pos(num, den);
    const double w = z * z;
#ifdef __CUDA_ARCH__
    double p = kAtanPoly[5];
    p = fma(p, w, kAtanPoly[4]);
    p = fma(p, w, kAtanPoly[3]);
    p = fma(p, w, kAtanPoly[2]);
    p = fma(p, w, kAtanPoly[1]);
    p = fma(p, w, kAtanPoly[0]);
    const double base = table[k];
#else
    double p = 1.0 / 13.0;
    p = fma(p, w, -1.0 / 11.0);
    p = fma(p, w, 1.0 / 9.0);
    p = fma(p, w, -1.0 / 7.0);
    p = fma(p, w, 1.0 / 5.0);
    p = fma(p, w, -1.0 / 3.0);
    const double base = table ? table[k] : atan_k8(k);
#endif
    const double az = fma(z * w, p, z);  // atan(z)
    double r = base + az;                               // atan(mn/mx) in [0, pi/4]
    if (ay > ax) r = 1.57079632679489661923 - r;         // pi/2 - r
    if (x < 0.0) r = 3.14159265358979323846 - r;         // pi - r
    return y < 0.0 ? -r : r;
}

}  // namespace tdoa
