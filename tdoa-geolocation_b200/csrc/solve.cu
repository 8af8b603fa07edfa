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
