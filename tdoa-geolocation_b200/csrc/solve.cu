// solve.cu -- geodesy, the reference's 2x2 damped Newton fix, and the dense lat-lon
// grid multilateration.
//
// Replaces latLonToECEF (processor.go:125-148), distance3D / calculateBaseline
// (:151-163), solveTDOA (:932-1020) and ecefToLatLon (:1023-1045).  All f64, written
// in the reference's operation order.  The grid solve has no reference equivalent
// (oracle: orc_grid_solve).
#include "kernels.h"

namespace tdoa {

namespace {

constexpr double kA = 6378137.0;
constexpr double kF = 1.0 / 298.257223563;
constexpr double kPi = 3.14159265358979323846;

__device__ void llh_to_ecef(double lat, double lon, double elev, double *xyz)
{
    const double e2 = 2 * kF - kF * kF;
    const double lr = lat * kPi / 180, lo = lon * kPi / 180;
    const double sl = sin(lr), cl = cos(lr), so = sin(lo), co = cos(lo);
    const double N = kA / sqrt(1 - e2 * sl * sl);
    xyz[0] = (N + elev) * cl * co;
    xyz[1] = (N + elev) * cl * so;
    xyz[2] = (N * (1 - e2) + elev) * sl;
}

__device__ void ecef_to_llh(double x, double y, double z, double *llh)
{
    const double e2 = 2 * kF - kF * kF;
    const double p = sqrt(x * x + y * y);
    const double lon = atan2(y, x);
    double lat = atan2(z, p * (1 - e2));
    for (int i = 0; i < 5; i++) {
        const double N = kA / sqrt(1 - e2 * sin(lat) * sin(lat));
        const double elev = p / cos(lat) - N;
        lat = atan2(z, p * (1 - e2 * N / (N + elev)));
    }
    const double N = kA / sqrt(1 - e2 * sin(lat) * sin(lat));
    const double elev = p / cos(lat) - N;
    llh[0] = lat * 180.0 / kPi;
    llh[1] = lon * 180.0 / kPi;
    llh[2] = elev;
}

__global__ void k_baselines(const double *llh, int n_st, double *out)
{
    // pair p = (i, j), i < j lexicographic (processor.go:803-809)
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int P = n_st * (n_st - 1) / 2;
    if (p >= P) return;
    int i = 0, rem = p;
    while (rem >= n_st - 1 - i) { rem -= n_st - 1 - i; i++; }
    const int j = i + 1 + rem;
    double a[3], b[3];
    llh_to_ecef(llh[3 * i], llh[3 * i + 1], llh[3 * i + 2], a);
    llh_to_ecef(llh[3 * j], llh[3 * j + 1], llh[3 * j + 2], b);
    const double dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
    out[p] = sqrt(dx * dx + dy * dy + dz * dz);
}

// processor.go:932-1020: stations 0..2 only, rd[0] and rd[1] only, start at the ECEF of
// the mean lat/lon/elev, <= 10 iterations, residual test before the update, Jacobian in
// X,Y, Cramer, |det| < 1e-10 -> singular, step 0.5, Z frozen.  One thread per set.
__global__ void k_solve(const double *llh, const double *rd_all, int n_sets, int rd_stride, double *out_llh,
                        int *status, int *iters)
{
    const int set = blockIdx.x * blockDim.x + threadIdx.x;
    if (set >= n_sets) return;
    const double *rd = rd_all + (size_t)set * rd_stride;
    double s[3][3];
    for (int k = 0; k < 3; k++) llh_to_ecef(llh[3 * k], llh[3 * k + 1], llh[3 * k + 2], s[k]);
    const double clat = (llh[0] + llh[3] + llh[6]) / 3.0;
    const double clon = (llh[1] + llh[4] + llh[7]) / 3.0;
    const double cel = (llh[2] + llh[5] + llh[8]) / 3.0;
    double x[3];
    llh_to_ecef(clat, clon, cel, x);
    int it, st = 0;
    for (it = 0; it < 10; it++) {
        double r[3];
        for (int k = 0; k < 3; k++)
            r[k] = sqrt((x[0] - s[k][0]) * (x[0] - s[k][0]) + (x[1] - s[k][1]) * (x[1] - s[k][1]) +
                        (x[2] - s[k][2]) * (x[2] - s[k][2]));
        const double res1 = (r[1] - r[0]) - rd[0];
        const double res2 = (r[2] - r[0]) - rd[1];
        if (fabs(res1) < 1.0 && fabs(res2) < 1.0) break;
        const double dx1 = (x[0] - s[0][0]) / r[0], dy1 = (x[1] - s[0][1]) / r[0];
        const double dx2 = (x[0] - s[1][0]) / r[1], dy2 = (x[1] - s[1][1]) / r[1];
        const double dx3 = (x[0] - s[2][0]) / r[2], dy3 = (x[1] - s[2][1]) / r[2];
        const double J11 = dx2 - dx1, J12 = dy2 - dy1, J21 = dx3 - dx1, J22 = dy3 - dy1;
        const double det = J11 * J22 - J12 * J21;
        if (fabs(det) < 1e-10) { st = 1; break; }
        const double dx = (-res1 * J22 + res2 * J12) / det;
        const double dy = (res1 * J21 - res2 * J11) / det;
        x[0] += 0.5 * dx;
        x[1] += 0.5 * dy;
    }
    if (iters) iters[set] = it;
    status[set] = st;
    double o[3] = {0.0, 0.0, 0.0};
    if (!st) ecef_to_llh(x[0], x[1], x[2], o);
    out_llh[3 * set] = o[0];
    out_llh[3 * set + 1] = o[1];
    out_llh[3 * set + 2] = o[2];
}

// solveTDOA of the SHIPPED BINARY (ELF 0x4a0360; oracle: orc_solve_binary, same statement, pinned
// by the iteration traces the binary prints): measurements with |rd| > 20400 m are dropped and
// the rest compacted; fewer than two -> status 1; any count but two -> status 2 (the binary solves
// with exactly two); res1 = (r2 - r1) - valid[0], res2 = (r3 - r1) - valid[1]; both < 1 m ->
// converged; |det| < 1e-12 -> 0.1 of a single-equation step in X (status 3 if neither equation
// can be used); else the Newton step times 0.7, a step longer than 1000 m first scaled to 1000 m;
// at most 10 iterations; Z is never updated.
// info[4] = {status, n_valid, n_iter, converged}; trace[10][5] = {det, res1, res2, step, code}
// (code 0 plain step, 1 limited step, 2 / 3 single equation 1 / 2).
__global__ void k_solve_binary(const double *llh, const double *rd, int n_rd, double *out_llh, int *info, double *trace)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double valid[2] = {0.0, 0.0};
    int nv = 0;
    for (int k = 0; k < n_rd; k++)
        if (fabs(rd[k]) <= 20400.0) {
            if (nv < 2) valid[nv] = rd[k];
            nv++;
        }
    int status = 0, n_iter = 0, converged = 0;
    double o[3] = {0.0, 0.0, 0.0};
    if (nv < 2) {
        status = 1;
    } else {
        double s[3][3];
        for (int k = 0; k < 3; k++) llh_to_ecef(llh[3 * k], llh[3 * k + 1], llh[3 * k + 2], s[k]);
        double x[3];
        llh_to_ecef((llh[0] + llh[3] + llh[6]) / 3.0, (llh[1] + llh[4] + llh[7]) / 3.0, (llh[2] + llh[5] + llh[8]) / 3.0, x);
        for (int it = 0; it < 10 && status == 0; it++) {
            double r[3];
            for (int k = 0; k < 3; k++)
                r[k] = sqrt((x[0] - s[k][0]) * (x[0] - s[k][0]) + (x[1] - s[k][1]) * (x[1] - s[k][1]) +
                            (x[2] - s[k][2]) * (x[2] - s[k][2]));
            const double dx1 = (x[0] - s[0][0]) / r[0], dy1 = (x[1] - s[0][1]) / r[0];
            const double dx2 = (x[0] - s[1][0]) / r[1], dy2 = (x[1] - s[1][1]) / r[1];
            const double dx3 = (x[0] - s[2][0]) / r[2], dy3 = (x[1] - s[2][1]) / r[2];
            if (nv != 2) { status = 2; break; }
            const double res1 = (r[1] - r[0]) - valid[0];
            const double res2 = (r[2] - r[0]) - valid[1];
            const double J11 = dx2 - dx1, J12 = dy2 - dy1, J21 = dx3 - dx1, J22 = dy3 - dy1;
            if (fabs(res1) < 1.0 && fabs(res2) < 1.0) { converged = 1; break; }
            const double det = J22 * J11 - J21 * J12;
            n_iter = it + 1;
            double t_step = 0.0, t_code = 0.0;
            // trace may be NULL (tdoa_b200.h): one guarded store per iteration
            auto record = [&]() {
                if (trace) {
                    double *tr = trace + 5 * it;
                    tr[0] = det; tr[1] = res1; tr[2] = res2; tr[3] = t_step; tr[4] = t_code;
                }
            };
            if (fabs(det) < 1e-12) {
                double d = 0.0;
                if (fabs(J11) > fabs(J21) && fabs(J12) > 1e-10) { d = -res1 / J11; t_code = 2.0; }
                else if (fabs(J21) > 1e-10) { d = -res2 / J21; t_code = 3.0; }
                else { status = 3; record(); break; }
                x[0] += d * 0.1;
            } else {
                const double dx = (-res1 * J22 + J12 * res2) / det;
                const double dy = (res1 * J21 - J11 * res2) / det;
                const double step = sqrt(dx * dx + dy * dy);
                double scale = 0.7;
                if (step > 1000.0) { scale = 1000.0 / step * 0.7; t_code = 1.0; }
                t_step = step;
                x[0] += dx * scale;
                x[1] += dy * scale;
            }
            record();
        }
        if (status == 0) ecef_to_llh(x[0], x[1], x[2], o);
    }
    out_llh[0] = o[0]; out_llh[1] = o[1]; out_llh[2] = o[2];
    info[0] = status; info[1] = nv; info[2] = n_iter; info[3] = converged;
}

// ---------------------------------------------------------------- least-squares fix
// SURVEY.md 8f rank 4 (no reference equivalent; oracle: orc_solve_ls, same statement):
// all P = S(S-1)/2 range differences, Levenberg-Marquardt in the local east/north/up
// frame of the current estimate (dims = 2: the elevation of the start point is kept,
// which is what three stations can support; dims = 3: elevation is solved too).
//   f_p(x) = (|x - s_j| - |x - s_i|) - rd_p,  row_p = g_j - g_i,  g_k = ENU components of
//   the unit vector from station k to x;  (J'J + lambda diag(J'J)) d = -J'f;  a step is
//   taken only if it lowers sum f^2 (lambda / 3), else lambda * 4, at most 8 tries;
//   stop after 60 iterations, when no try succeeds, or when |d| < 1e-6 m.
struct LsPoint {
    double lat, lon, h;  // degrees, degrees, metres
};

__device__ double ls_cost(const double *s_st, int n_st, const double *rd, const LsPoint &p, double *r_out)
{
    double x[3];
    llh_to_ecef(p.lat, p.lon, p.h, x);
    for (int k = 0; k < n_st; k++) {
        const double dx = x[0] - s_st[3 * k], dy = x[1] - s_st[3 * k + 1], dz = x[2] - s_st[3 * k + 2];
        r_out[k] = sqrt(dx * dx + dy * dy + dz * dz);
    }
    double c = 0.0;
    int q = 0;
    for (int i = 0; i < n_st; i++)
        for (int j = i + 1; j < n_st; j++, q++) {
            const double f = (r_out[j] - r_out[i]) - rd[q];
            c += f * f;
        }
    return c;
}

// d = -(A + lambda diag A)^-1 b for the leading n x n part (n = 2 or 3); false: singular
__device__ bool ls_step(const double (&A)[3][3], const double (&b)[3], double lambda, int n, double (&d)[3])
{
    double M[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) M[i][j] = A[i][j];
    for (int i = 0; i < 3; i++) M[i][i] = A[i][i] + lambda * A[i][i];
    d[0] = d[1] = d[2] = 0.0;
    if (n == 2) {
        const double det = M[0][0] * M[1][1] - M[0][1] * M[1][0];
        if (!(fabs(det) > 1e-300)) return false;
        d[0] = -(b[0] * M[1][1] - b[1] * M[0][1]) / det;
        d[1] = -(M[0][0] * b[1] - M[1][0] * b[0]) / det;
        return true;
    }
    const double c00 = M[1][1] * M[2][2] - M[1][2] * M[2][1];
    const double c01 = M[1][2] * M[2][0] - M[1][0] * M[2][2];
    const double c02 = M[1][0] * M[2][1] - M[1][1] * M[2][0];
    const double det = M[0][0] * c00 + M[0][1] * c01 + M[0][2] * c02;
    if (!(fabs(det) > 1e-300)) return false;
    const double c10 = M[0][2] * M[2][1] - M[0][1] * M[2][2];
    const double c11 = M[0][0] * M[2][2] - M[0][2] * M[2][0];
    const double c12 = M[0][1] * M[2][0] - M[0][0] * M[2][1];
    const double c20 = M[0][1] * M[1][2] - M[0][2] * M[1][1];
    const double c21 = M[0][2] * M[1][0] - M[0][0] * M[1][2];
    const double c22 = M[0][0] * M[1][1] - M[0][1] * M[1][0];
    d[0] = -(c00 * b[0] + c10 * b[1] + c20 * b[2]) / det;
    d[1] = -(c01 * b[0] + c11 * b[1] + c21 * b[2]) / det;
    d[2] = -(c02 * b[0] + c12 * b[1] + c22 * b[2]) / det;
    return true;
}

constexpr int kLsMaxStations = 64;

__global__ void k_solve_ls(const double *llh, int n_st, const double *rd_all, int n_sets, int rd_stride,
                           const double *init_llh, int dims, double *out_llh, double *out_rms, int *status, int *iters)
{
    extern __shared__ double s_st[];  // station ECEF
    for (int k = threadIdx.x; k < n_st; k += blockDim.x) llh_to_ecef(llh[3 * k], llh[3 * k + 1], llh[3 * k + 2], s_st + 3 * k);
    __syncthreads();
    const int set = blockIdx.x * blockDim.x + threadIdx.x;
    if (set >= n_sets) return;
    const double *rd = rd_all + (size_t)set * rd_stride;
    const int P = n_st * (n_st - 1) / 2;
    const double e2 = 2 * kF - kF * kF;
    LsPoint p;
    if (init_llh) {
        p.lat = init_llh[3 * set]; p.lon = init_llh[3 * set + 1]; p.h = init_llh[3 * set + 2];
    } else {  // mean of the station coordinates
        p.lat = p.lon = p.h = 0.0;
        for (int k = 0; k < n_st; k++) { p.lat += llh[3 * k]; p.lon += llh[3 * k + 1]; p.h += llh[3 * k + 2]; }
        p.lat /= n_st; p.lon /= n_st; p.h /= n_st;
    }
    double r[kLsMaxStations], g[kLsMaxStations][3];
    double lambda = 1e-3;
    double cost = ls_cost(s_st, n_st, rd, p, r);
    int it = 0, st = 0;
    for (; it < 60; it++) {
        const double lr = p.lat * kPi / 180, lo = p.lon * kPi / 180;
        const double sl = sin(lr), cl = cos(lr), so = sin(lo), co = cos(lo);
        const double w = sqrt(1 - e2 * sl * sl);
        const double Nr = kA / w, Mr = kA * (1 - e2) / (w * w * w);
        double x[3];
        llh_to_ecef(p.lat, p.lon, p.h, x);
        for (int k = 0; k < n_st; k++) {
            const double ux = (x[0] - s_st[3 * k]) / r[k], uy = (x[1] - s_st[3 * k + 1]) / r[k], uz = (x[2] - s_st[3 * k + 2]) / r[k];
            g[k][0] = -so * ux + co * uy;
            g[k][1] = -sl * co * ux - sl * so * uy + cl * uz;
            g[k][2] = cl * co * ux + cl * so * uy + sl * uz;
        }
        double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, b[3] = {0, 0, 0};
        int q = 0;
        for (int i = 0; i < n_st; i++)
            for (int j = i + 1; j < n_st; j++, q++) {
                const double f = (r[j] - r[i]) - rd[q];
                const double row[3] = {g[j][0] - g[i][0], g[j][1] - g[i][1], g[j][2] - g[i][2]};
                for (int a = 0; a < 3; a++) {
                    b[a] += row[a] * f;
                    for (int c = 0; c < 3; c++) A[a][c] += row[a] * row[c];
                }
            }
        bool moved = false;
        double d[3] = {0, 0, 0};
        for (int tr = 0; tr < 8 && !moved; tr++) {
            if (!ls_step(A, b, lambda, dims, d)) { lambda *= 10.0; continue; }
            LsPoint c = p;
            c.lat = p.lat + d[1] / (Mr + p.h) * 180.0 / kPi;
            c.lon = p.lon + d[0] / ((Nr + p.h) * cl) * 180.0 / kPi;
            if (dims == 3) c.h = p.h + d[2];
            double rc[kLsMaxStations];
            const double cc = ls_cost(s_st, n_st, rd, c, rc);
            if (cc < cost) {
                p = c; cost = cc; moved = true;
                for (int k = 0; k < n_st; k++) r[k] = rc[k];
                lambda = fmax(lambda / 3.0, 1e-12);
            } else {
                lambda *= 4.0;
            }
        }
        if (!moved) break;
        if (sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) < 1e-6) { it++; break; }
    }
    if (!(cost == cost)) st = 1;  // NaN: degenerate geometry
    out_llh[3 * set] = p.lat;
    out_llh[3 * set + 1] = p.lon;
    out_llh[3 * set + 2] = p.h;
    if (out_rms) out_rms[set] = P > 0 ? sqrt(cost / P) : 0.0;
    status[set] = st;
    if (iters) iters[set] = it;
}

// ---------------------------------------------------------------- grid multilateration
// cost(cell, set) = sum over pairs i<j (lexicographic) of ((r_j - r_i) - rd_ij)^2, f64,
// terms added in pair order (as orc_grid_solve); arg-min, lowest linear index wins ties.
// The S ranges of a cell do not depend on the set, so one thread computes them once
// and sweeps every set; per-CTA minima go to scratch and a second kernel reduces them.
constexpr int kGridThreads = 128;
constexpr int kMaxStations = 16;

struct GridDesc {
    double lat0, lon0, dlat, dlon;
    int nlat, nlon;
    double elev;
};

__device__ __forceinline__ GridDesc read_desc(const double *g)
{
    GridDesc d;
    d.lat0 = g[0]; d.lon0 = g[1]; d.dlat = g[2]; d.dlon = g[3];
    d.nlat = (int)g[4]; d.nlon = (int)g[5]; d.elev = g[6];
    return d;
}

__global__ void __launch_bounds__(kGridThreads) k_grid_cost(const double *llh, int n_st, const double *gdesc,
                                                           const double *rd_all, int n_sets, int rd_stride,
                                                           double *cta_cost, i64 *cta_idx)
{
    extern __shared__ double sm[];  // [n_st*3] station ECEF, then [kGridThreads] cost + idx scratch
    double *s_st = sm;
    double *s_cost = sm + 3 * kMaxStations;
    i64 *s_idx = reinterpret_cast<i64 *>(s_cost + kGridThreads);
    const GridDesc G = read_desc(gdesc);
    if (threadIdx.x < n_st) llh_to_ecef(llh[3 * threadIdx.x], llh[3 * threadIdx.x + 1], llh[3 * threadIdx.x + 2],
                                        s_st + 3 * threadIdx.x);
    __syncthreads();
    const i64 n_cells = (i64)G.nlat * G.nlon;
    const i64 cell = (i64)blockIdx.x * kGridThreads + threadIdx.x;
    double r[kMaxStations];
    const bool live = cell < n_cells;
    if (live) {
        const int a = (int)(cell / G.nlon), b = (int)(cell % G.nlon);
        double x[3];
        llh_to_ecef(G.lat0 + a * G.dlat, G.lon0 + b * G.dlon, G.elev, x);
#pragma unroll
        for (int k = 0; k < kMaxStations; k++) {
            if (k < n_st) {
                const double dx = x[0] - s_st[3 * k], dy = x[1] - s_st[3 * k + 1], dz = x[2] - s_st[3 * k + 2];
                r[k] = sqrt(dx * dx + dy * dy + dz * dz);
            } else {
                r[k] = 0.0;
            }
        }
    }
    for (int set = 0; set < n_sets; set++) {
        const double *rd = rd_all + (size_t)set * rd_stride;
        double cost = 0.0;
        if (live) {
            int p = 0;
#pragma unroll
            for (int i = 0; i < kMaxStations; i++) {
#pragma unroll
                for (int j = i + 1; j < kMaxStations; j++) {
                    if (j < n_st) {
                        const double e = (r[j] - r[i]) - rd[p];
                        cost = __dadd_rn(cost, __dmul_rn(e, e));
                        p++;
                    }
                }
            }
        }
        // block arg-min (lowest index wins ties)
        s_cost[threadIdx.x] = cost;
        s_idx[threadIdx.x] = live ? cell : (i64)-1;
        __syncthreads();
        for (int o = kGridThreads / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const i64 ia = s_idx[threadIdx.x], ib = s_idx[threadIdx.x + o];
                const double ca = s_cost[threadIdx.x], cb = s_cost[threadIdx.x + o];
                const bool take_b = ib >= 0 && (ia < 0 || cb < ca || (cb == ca && ib < ia));
                if (take_b) { s_cost[threadIdx.x] = cb; s_idx[threadIdx.x] = ib; }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            cta_cost[(size_t)set * gridDim.x + blockIdx.x] = s_cost[0];
            cta_idx[(size_t)set * gridDim.x + blockIdx.x] = s_idx[0];
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_grid_reduce(const double *gdesc, const double *cta_cost, const i64 *cta_idx,
                                                    int n_cta, double *best_cost, i64 *best_idx, double *out_llh)
{
    __shared__ double s_cost[256];
    __shared__ i64 s_idx[256];
    const int set = blockIdx.x;
    double c = 0.0;
    i64 ix = -1;
    for (int i = threadIdx.x; i < n_cta; i += 256) {
        const double cc = cta_cost[(size_t)set * n_cta + i];
        const i64 ii = cta_idx[(size_t)set * n_cta + i];
        if (ii >= 0 && (ix < 0 || cc < c || (cc == c && ii < ix))) { c = cc; ix = ii; }
    }
    s_cost[threadIdx.x] = c;
    s_idx[threadIdx.x] = ix;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const i64 ia = s_idx[threadIdx.x], ib = s_idx[threadIdx.x + o];
            const double ca = s_cost[threadIdx.x], cb = s_cost[threadIdx.x + o];
            if (ib >= 0 && (ia < 0 || cb < ca || (cb == ca && ib < ia))) { s_cost[threadIdx.x] = cb; s_idx[threadIdx.x] = ib; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const GridDesc G = read_desc(gdesc);
        const i64 bi = s_idx[0];
        best_cost[set] = s_cost[0];
        best_idx[set] = bi;
        if (bi >= 0) {
            out_llh[3 * set] = G.lat0 + (double)(bi / G.nlon) * G.dlat;
            out_llh[3 * set + 1] = G.lon0 + (double)(bi % G.nlon) * G.dlon;
            out_llh[3 * set + 2] = G.elev;
        }
    }
}

}  // namespace

void launch_baselines(const double *d_llh, int n_st, double *d_out, cudaStream_t st)
{
    const int P = n_st * (n_st - 1) / 2;
    if (P <= 0) return;
    k_baselines<<<(P + 63) / 64, 64, 0, st>>>(d_llh, n_st, d_out);
}

void launch_solve(const double *d_llh, const double *d_rd, int n_sets, int rd_stride, double *d_out_llh,
                  int *d_status, int *d_iters, cudaStream_t st)
{
    if (n_sets <= 0) return;
    k_solve<<<(n_sets + 63) / 64, 64, 0, st>>>(d_llh, d_rd, n_sets, rd_stride, d_out_llh, d_status, d_iters);
}

void launch_solve_binary(const double *d_llh, const double *d_rd, int n_rd, double *d_out_llh, int *d_info, double *d_trace,
                         cudaStream_t st)
{
    k_solve_binary<<<1, 32, 0, st>>>(d_llh, d_rd, n_rd, d_out_llh, d_info, d_trace);
}

// ProcessTDOA between the pair loops and the solver (processor.go:821, :853, :899-903; shipped
// binary: "REFERENCE SIGNAL SYNCHRONIZATION"): dt = delay / fs per pair; SOURCE keeps the
// target differences, the binary's modes subtract the reference differences pair by pair;
// range difference = dt * c.  EXTENDED adds the sub-sample offset to the delay.
__global__ void k_range_diffs(const PeakRec *ref, const PeakRec *tgt, int n_pairs, double fs, int mode, double *td_out,
                              double *rd_out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    double lt = (double)tgt[p].lag, lr = (double)ref[p].lag;
    if (mode == TDOA_MODE_EXTENDED) { lt += (double)tgt[p].frac; lr += (double)ref[p].frac; }
    const double tt = lt / fs, tr = lr / fs;
    const double td = mode == TDOA_MODE_SOURCE ? tt : tt - tr;
    td_out[p] = td;
    rd_out[p] = td * 299792458.0;
}

void launch_range_diffs(const PeakRec *d_ref, const PeakRec *d_tgt, int n_pairs, double fs, int mode, double *d_td,
                        double *d_rd, cudaStream_t st)
{
    if (n_pairs > 0) k_range_diffs<<<(n_pairs + 63) / 64, 64, 0, st>>>(d_ref, d_tgt, n_pairs, fs, mode, d_td, d_rd);
}

int solve_ls_max_stations() { return kLsMaxStations; }

void launch_solve_ls(const double *d_llh, int n_st, const double *d_rd, int n_sets, int rd_stride, const double *d_init,
                     int dims, double *d_out_llh, double *d_rms, int *d_status, int *d_iters, cudaStream_t st)
{
    if (n_sets <= 0) return;
    k_solve_ls<<<(n_sets + 31) / 32, 32, 3 * n_st * sizeof(double), st>>>(d_llh, n_st, d_rd, n_sets, rd_stride, d_init, dims,
                                                                       d_out_llh, d_rms, d_status, d_iters);
}

// ---------------------------------------------------------------- grid multilateration, ranked
// The cost above spends 4 f64 operations per (cell, set, pair): 480 per cell and set at 16 stations.  Expanded
// around q_k = r_k - r_0 (the cost only sees differences of ranges) it separates into a part of the cell, a
// part of the set and ONE dot product that couples them:
//     cost = [S sum q_k^2 - (sum q_k)^2]  -  2 sum_k q_k c_k  +  sum_p rd_p^2 ,
//     c_k  = sum_{i<k} rd_ik - sum_{j>k} rd_kj      (S = stations; c_k and sum rd^2 come from the host, per set)
// i.e. 16 fused multiply-adds per cell and set.  That form only RANKS cells (it cancels ~1e11 m^2 down to the
// cost): its rounding error is bounded by tol = kGridKappa * (2 S sum q^2 + 4 max|q| sum|rd| + sum rd^2), so the
// statement's arg-min x* satisfies f(x*) - tol(x*) <= min_x (f(x) + tol(x)) =: U.  k_grid_rank leaves per CTA and
// set the minima of f - tol and f + tol; k_grid_bound takes U; k_grid_refine revisits only the CTAs that can
// hold a cell under U and lists those cells (normally one per set); k_grid_exact evaluates the listed cells by the
// statement itself -- pair by pair in its order -- and k_grid_pick runs the statement's selection (strict <,
// lowest index on ties) on them.  The
// index and the cost are the statement's, bit for bit; config 5 (10^6 cells x 64 sets) 7.5 ms -> see DESIGN.md.
constexpr int kRankThreads = 128;
constexpr int kRankCells = 2;                    // cells per thread
constexpr int kRankSpan = kRankThreads * kRankCells;
constexpr int kRankTab = 20;                     // doubles per set: c[16], sum rd^2, 2 sum |rd|, pad
constexpr double kGridKappa = 4e-13;             // ~3600 ulp of the largest partial result: far above the ~200 roundings
constexpr int kRankMaxSets = 512;                // per launch (shared memory: 64 B per set)

struct RankCell {
    double q[kMaxStations];
    double a;      // S sum q^2 - (sum q)^2
    double t0;     // 2 S sum q^2
    double qmax;
    bool live;
};

__device__ __forceinline__ void rank_cell(const GridDesc &G, const double *s_st, int n_st, i64 cell, i64 n_cells,
                                          RankCell &c, double *r_out)
{
    c.live = cell < n_cells;
    double r[kMaxStations];
    if (c.live) {
        const int a = (int)(cell / G.nlon), b = (int)(cell % G.nlon);
        double x[3];
        llh_to_ecef(G.lat0 + a * G.dlat, G.lon0 + b * G.dlon, G.elev, x);
#pragma unroll
        for (int k = 0; k < kMaxStations; k++) {
            if (k < n_st) {
                const double dx = x[0] - s_st[3 * k], dy = x[1] - s_st[3 * k + 1], dz = x[2] - s_st[3 * k + 2];
                r[k] = sqrt(dx * dx + dy * dy + dz * dz);
            } else {
                r[k] = 0.0;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < kMaxStations; k++) r[k] = 0.0;
    }
    double sq = 0.0, sl = 0.0, qm = 0.0;
#pragma unroll
    for (int k = 0; k < kMaxStations; k++) {
        const double q = k < n_st ? r[k] - r[0] : 0.0;
        c.q[k] = q;
        sq = __fma_rn(q, q, sq);
        sl += q;
        qm = fmax(qm, fabs(q));
    }
    c.a = (double)n_st * sq - sl * sl;
    c.t0 = 2.0 * (double)n_st * sq;
    c.qmax = qm;
    if (r_out) {
#pragma unroll
        for (int k = 0; k < kMaxStations; k++) r_out[k] = r[k];
    }
}

// f and tol of one cell for one set; tab = the set's row (c[16], B, C1)
__device__ __forceinline__ void rank_eval(const RankCell &c, const double2 *__restrict__ tab, double &lo, double &hi)
{
    double mid = 0.0;
#pragma unroll
    for (int m = 0; m < kMaxStations / 2; m++) {
        const double2 cc = __ldg(tab + m);
        mid = __fma_rn(c.q[2 * m], cc.x, mid);
        mid = __fma_rn(c.q[2 * m + 1], cc.y, mid);
    }
    const double2 bc = __ldg(tab + kMaxStations / 2);   // (sum rd^2, 2 sum |rd|)
    const double f = __fma_rn(-2.0, mid, c.a + bc.x);
    const double tol = kGridKappa * (c.t0 + __fma_rn(2.0 * c.qmax, bc.y, bc.x));
    lo = f - tol;
    hi = f + tol;
}

__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(kRankThreads) k_grid_rank(const double *llh, int n_st, const double *gdesc,
                                                            const double *tab, int n_sets, double *cta_lo, double *cta_hi)
{
    extern __shared__ double sm[];   // [3 * 16] station ECEF, then [n_sets][4 warps] lo, [n_sets][4 warps] hi
    double *s_st = sm;
    double *s_lo = sm + 3 * kMaxStations;
    double *s_hi = s_lo + (size_t)n_sets * (kRankThreads / 32);
    const GridDesc G = read_desc(gdesc);
    if (threadIdx.x < n_st) llh_to_ecef(llh[3 * threadIdx.x], llh[3 * threadIdx.x + 1], llh[3 * threadIdx.x + 2],
                                        s_st + 3 * threadIdx.x);
    __syncthreads();
    const i64 n_cells = (i64)G.nlat * G.nlon;
    const i64 base = (i64)blockIdx.x * kRankSpan + threadIdx.x;
    RankCell c0, c1;
    rank_cell(G, s_st, n_st, base, n_cells, c0, nullptr);
    rank_cell(G, s_st, n_st, base + kRankThreads, n_cells, c1, nullptr);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    for (int set = 0; set < n_sets; set++) {
        const double2 *row = reinterpret_cast<const double2 *>(tab + (size_t)set * kRankTab);
        double lo0, hi0, lo1, hi1;
        rank_eval(c0, row, lo0, hi0);
        rank_eval(c1, row, lo1, hi1);
        double lo = fmin(c0.live ? lo0 : inf, c1.live ? lo1 : inf);
        double hi = fmin(c0.live ? hi0 : inf, c1.live ? hi1 : inf);
        lo = warp_min(lo);
        hi = warp_min(hi);
        if (lane == 0) { s_lo[set * (kRankThreads / 32) + wid] = lo; s_hi[set * (kRankThreads / 32) + wid] = hi; }
    }
    __syncthreads();
    for (int set = threadIdx.x; set < n_sets; set += kRankThreads) {
        double lo = inf, hi = inf;
#pragma unroll
        for (int w = 0; w < kRankThreads / 32; w++) {
            lo = fmin(lo, s_lo[set * (kRankThreads / 32) + w]);
            hi = fmin(hi, s_hi[set * (kRankThreads / 32) + w]);
        }
        cta_lo[(size_t)set * gridDim.x + blockIdx.x] = lo;
        cta_hi[(size_t)set * gridDim.x + blockIdx.x] = hi;
    }
}

__global__ void __launch_bounds__(256) k_grid_bound(const double *cta_hi, int n_cta, double *bound)
{
    __shared__ double s_m[8];
    const int set = blockIdx.x;
    double m = __longlong_as_double(0x7ff0000000000000ll);
    for (int i = threadIdx.x; i < n_cta; i += 256) m = fmin(m, cta_hi[(size_t)set * n_cta + i]);
    m = warp_min(m);
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; w++) m = fmin(m, s_m[w]);
        bound[set] = m;
    }
}

struct GridCand { i64 cell; double cost; int set; int pad; };

__global__ void __launch_bounds__(kRankThreads) k_grid_refine(const double *llh, int n_st, const double *gdesc,
                                                              const double *tab, const double *rd_all, int n_sets,
                                                              int rd_stride, const double *cta_lo, const double *bound,
                                                              GridCand *cands, int cap, int *n_cand)
{
    __shared__ double s_st[3 * kMaxStations];
    __shared__ int s_sets[kRankMaxSets];
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int set = threadIdx.x; set < n_sets; set += kRankThreads)
        if (cta_lo[(size_t)set * gridDim.x + blockIdx.x] <= bound[set]) s_sets[atomicAdd(&s_n, 1)] = set;
    __syncthreads();
    const int ns = s_n;
    if (ns == 0) return;   // nearly every CTA
    const GridDesc G = read_desc(gdesc);
    if (threadIdx.x < n_st) llh_to_ecef(llh[3 * threadIdx.x], llh[3 * threadIdx.x + 1], llh[3 * threadIdx.x + 2],
                                        s_st + 3 * threadIdx.x);
    __syncthreads();
    const i64 n_cells = (i64)G.nlat * G.nlon;
    for (int h = 0; h < kRankCells; h++) {
        const i64 cell = (i64)blockIdx.x * kRankSpan + h * kRankThreads + threadIdx.x;
        RankCell c;
        rank_cell(G, s_st, n_st, cell, n_cells, c, nullptr);
        if (!c.live) continue;
        for (int u = 0; u < ns; u++) {
            const int set = s_sets[u];
            double lo, hi;
            rank_eval(c, reinterpret_cast<const double2 *>(tab + (size_t)set * kRankTab), lo, hi);
            if (!(lo <= bound[set])) continue;
            const int slot = atomicAdd(n_cand, 1);
            if (slot < cap) cands[slot] = GridCand{cell, 0.0, set, 0};
        }
    }
}

// The statement itself (k_grid_cost / orc_grid_solve) on the surviving (cell, set) pairs, one thread each: pair by
// pair, in order, no contraction.  (Inside k_grid_refine the survivors of all sets sit in two or three
// neighbouring cells, i.e. in one warp, and their 120-term chains ran one after the other: 0.41 ms of 0.75.)
__global__ void __launch_bounds__(kRankThreads) k_grid_exact(const double *llh, int n_st, const double *gdesc,
                                                             const double *rd_all, int rd_stride, GridCand *cands, int cap,
                                                             const int *n_cand)
{
    __shared__ double s_st[3 * kMaxStations];
    const int n = min(*n_cand, cap);
    if ((int)(blockIdx.x * kRankThreads) >= n) return;
    const GridDesc G = read_desc(gdesc);
    if (threadIdx.x < n_st) llh_to_ecef(llh[3 * threadIdx.x], llh[3 * threadIdx.x + 1], llh[3 * threadIdx.x + 2],
                                        s_st + 3 * threadIdx.x);
    __syncthreads();
    const int i = blockIdx.x * kRankThreads + threadIdx.x;
    if (i >= n) return;
    const GridCand g = cands[i];
    RankCell c;
    double r[kMaxStations];
    rank_cell(G, s_st, n_st, g.cell, (i64)G.nlat * G.nlon, c, r);
    const double *rd = rd_all + (size_t)g.set * rd_stride;
    double cost = 0.0;
    int p = 0;
#pragma unroll
    for (int a = 0; a < kMaxStations; a++) {
#pragma unroll
        for (int b = a + 1; b < kMaxStations; b++) {
            if (b < n_st) {
                const double e = (r[b] - r[a]) - rd[p];
                cost = __dadd_rn(cost, __dmul_rn(e, e));
                p++;
            }
        }
    }
    cands[i].cost = cost;
}

__global__ void __launch_bounds__(128) k_grid_pick(const double *gdesc, const GridCand *cands, int cap, const int *n_cand,
                                                  double *best_cost, i64 *best_idx, double *out_llh)
{
    __shared__ double s_cost[128];
    __shared__ i64 s_idx[128];
    const int set = blockIdx.x;
    const int n = min(*n_cand, cap);
    double c = 0.0;
    i64 ix = -1;
    for (int i = threadIdx.x; i < n; i += 128) {
        const GridCand g = cands[i];
        if (g.set == set && (ix < 0 || g.cost < c || (g.cost == c && g.cell < ix))) { c = g.cost; ix = g.cell; }
    }
    s_cost[threadIdx.x] = c;
    s_idx[threadIdx.x] = ix;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const i64 ia = s_idx[threadIdx.x], ib = s_idx[threadIdx.x + o];
            const double ca = s_cost[threadIdx.x], cb = s_cost[threadIdx.x + o];
            if (ib >= 0 && (ia < 0 || cb < ca || (cb == ca && ib < ia))) { s_cost[threadIdx.x] = cb; s_idx[threadIdx.x] = ib; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const GridDesc G = read_desc(gdesc);
        const i64 bi = s_idx[0];
        best_cost[set] = s_cost[0];
        best_idx[set] = bi;
        if (bi >= 0) {
            out_llh[3 * set] = G.lat0 + (double)(bi / G.nlon) * G.dlat;
            out_llh[3 * set + 1] = G.lon0 + (double)(bi % G.nlon) * G.dlon;
            out_llh[3 * set + 2] = G.elev;
        }
    }
}

int grid_rank_tab_doubles() { return kRankTab; }
int grid_rank_max_sets() { return kRankMaxSets; }
int grid_rank_cand_cap(int n_sets) { return 64 * n_sets + 1024; }

static int rank_n_cta(int nlat, int nlon)
{
    const i64 cells = (i64)nlat * nlon;
    return (int)((cells + kRankSpan - 1) / kRankSpan);
}

// scratch: [n_sets][n_cta] lo, [n_sets][n_cta] hi, [n_sets] bound, candidates, counter
size_t grid_rank_scratch_bytes(int nlat, int nlon, int n_sets)
{
    const size_t n_cta = (size_t)rank_n_cta(nlat, nlon);
    return (2 * n_cta * (size_t)n_sets + (size_t)n_sets) * sizeof(double) +
           (size_t)grid_rank_cand_cap(n_sets) * sizeof(GridCand) + 64;
}

// Host part of the expansion: the set's row of the table (c_k, sum rd^2, sum |c_k|) from its range differences.
void grid_rank_row(const double *rd, int n_st, double *row)
{
    double c[kMaxStations] = {0.0};
    double b = 0.0;
    int p = 0;
    for (int i = 0; i < n_st; i++)
        for (int j = i + 1; j < n_st; j++, p++) {
            c[j] += rd[p];
            c[i] -= rd[p];
            b += rd[p] * rd[p];
        }
    // the bound takes sum |rd| over the terms of every c_k (not sum |c_k|: a c_k that cancels keeps its rounding)
    double c1 = 0.0;
    p = 0;
    for (int i = 0; i < n_st; i++)
        for (int j = i + 1; j < n_st; j++, p++) c1 += 2.0 * (rd[p] < 0 ? -rd[p] : rd[p]);
    for (int k = 0; k < kMaxStations; k++) row[k] = c[k];
    row[16] = b;
    row[17] = c1;
    row[18] = row[19] = 0.0;
}

// n_sets <= grid_rank_max_sets().  d_count: one int, read back by the caller after the stream has drained --
// more than grid_rank_cand_cap(n_sets) candidates means the table is incomplete (launch_grid_cells then).
void launch_grid_ranked(const double *d_llh, int n_st, const double *d_grid_desc, int nlat, int nlon, const double *d_tab,
                        const double *d_rd, int n_sets, int rd_stride, double *d_best_cost, i64 *d_best_idx,
                        double *d_out_llh, void *d_scratch, int *d_count, cudaStream_t st)
{
    const int n_cta = rank_n_cta(nlat, nlon);
    if (n_cta <= 0 || n_sets <= 0) return;
    double *cta_lo = reinterpret_cast<double *>(d_scratch);
    double *cta_hi = cta_lo + (size_t)n_cta * n_sets;
    double *bound = cta_hi + (size_t)n_cta * n_sets;
    GridCand *cands = reinterpret_cast<GridCand *>(bound + n_sets);
    const int cap = grid_rank_cand_cap(n_sets);
    cudaMemsetAsync(d_count, 0, sizeof(int), st);
    const size_t smem = (3 * kMaxStations + 2 * (size_t)n_sets * (kRankThreads / 32)) * sizeof(double);
    k_grid_rank<<<n_cta, kRankThreads, smem, st>>>(d_llh, n_st, d_grid_desc, d_tab, n_sets, cta_lo, cta_hi);
    k_grid_bound<<<n_sets, 256, 0, st>>>(cta_hi, n_cta, bound);
    k_grid_refine<<<n_cta, kRankThreads, 0, st>>>(d_llh, n_st, d_grid_desc, d_tab, d_rd, n_sets, rd_stride, cta_lo, bound, cands,
                                                cap, d_count);
    k_grid_exact<<<(cap + kRankThreads - 1) / kRankThreads, kRankThreads, 0, st>>>(d_llh, n_st, d_grid_desc, d_rd, rd_stride, cands,
                                                                                   cap, d_count);
    k_grid_pick<<<n_sets, 128, 0, st>>>(d_grid_desc, cands, cap, d_count, d_best_cost, d_best_idx, d_out_llh);
}

static int grid_n_cta(int nlat, int nlon)
{
    const i64 cells = (i64)nlat * nlon;
    return (int)((cells + kGridThreads - 1) / kGridThreads);
}

size_t grid_scratch_bytes(int n_st, int nlat, int nlon, int n_sets)
{
    (void)n_st;
    return (size_t)grid_n_cta(nlat, nlon) * (size_t)n_sets * (sizeof(double) + sizeof(i64));
}

void launch_grid_cells(const double *d_llh, int n_st, const double *d_grid_desc, int nlat, int nlon,
                       const double *d_rd, int n_sets, int rd_stride, double *d_best_cost, i64 *d_best_idx,
                       double *d_out_llh, void *d_scratch, cudaStream_t st)
{
    const int n_cta = grid_n_cta(nlat, nlon);
    if (n_cta <= 0 || n_sets <= 0) return;
    double *cta_cost = reinterpret_cast<double *>(d_scratch);
    i64 *cta_idx = reinterpret_cast<i64 *>(cta_cost + (size_t)n_cta * n_sets);
    const size_t smem = (3 * kMaxStations + kGridThreads) * sizeof(double) + kGridThreads * sizeof(i64);
    k_grid_cost<<<n_cta, kGridThreads, smem, st>>>(d_llh, n_st, d_grid_desc, d_rd, n_sets, rd_stride, cta_cost, cta_idx);
    k_grid_reduce<<<n_sets, 256, 0, st>>>(d_grid_desc, cta_cost, cta_idx, n_cta, d_best_cost, d_best_idx, d_out_llh);
}

}  // namespace tdoa
