// Microbenchmark: tensor memory (TMEM) as per-thread accumulator storage.
// 512 threads, each thread owns 64 consecutive 32-bit columns of its TMEM lane
// (warp w uses lanes 32*(w%4).., columns 128*(w/4)..).  Checks that data written with
// tcgen05.st comes back with tcgen05.ld, and times a read-modify-write loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_rw tmem_rw.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16])
{
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(unsigned taddr, const float (&v)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                    "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                    "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])),
                    "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                    "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
                    "r"(__float_as_uint(v[15]))
                 : "memory");
}

__global__ void __launch_bounds__(512, 1) k(float *out, int iters, long long *cycles)
{
    __shared__ unsigned s_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(&s_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned base = s_base;
    // this thread's 64 columns
    const unsigned taddr = base + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)(128 * (warp >> 2));
    float v[16];
    for (int c = 0; c < 4; c++) {
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = (float)(tid * 1000 + c * 16 + i);
        tmem_st16(taddr + 16 * c, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        for (int c = 0; c < 4; c++) {
            tmem_ld16(taddr + 16 * c, v);
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] += 1.0f;
            tmem_st16(taddr + 16 * c, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    const long long t1 = clock64();
    __syncthreads();
    float bad = 0.f;
    for (int c = 0; c < 4; c++) {
        tmem_ld16(taddr + 16 * c, v);
#pragma unroll
        for (int i = 0; i < 16; i++) bad += fabsf(v[i] - (float)(tid * 1000 + c * 16 + i) - (float)iters);
    }
    out[blockIdx.x * 512 + tid] = bad;
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
    (void)lane;
}

int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 512 * sizeof(float)); cudaMalloc(&cyc, 148 * sizeof(long long));
    const int iters = 1000;
    k<<<148, 512>>>(out, iters, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    static float h[148 * 512]; long long hc[148];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double bad = 0; for (float x : h) bad += x;
    printf("sum |error| = %g\n", bad);
    const double bytes = 512.0 * 64 * 4;  // per iteration, read and written
    printf("cycles/iter = %.1f  -> %.1f B/clk read + %.1f B/clk write per SM\n", (double)hc[0] / iters, bytes / ((double)hc[0] / iters),
           bytes / ((double)hc[0] / iters));
    return 0;
}
