// Microbenchmark: per-SM throughput of the FP64 / conversion instructions the parity
// kernels lean on (DADD, DFMA, F2F f32<->f64, FADD for scale).  Build: nvcc -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void bench(double *out, int iters, float seed)
{
    double a[8]; float f[8];
    for (int k = 0; k < 8; k++) { a[k] = threadIdx.x * 1e-3 + k; f[k] = seed + threadIdx.x + k; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (OP == 0) a[k] = __dadd_rn(a[k], 1.0000001);
            if (OP == 1) a[k] = __fma_rn(a[k], 1.0000001, 0.5);
            if (OP == 2) { a[k] = __dadd_rn(a[k], (double)f[k]); f[k] = __fadd_rn(f[k], 1.0f); }      // F2F.F64.F32 + DADD + FADD
            if (OP == 3) { f[k] = __fadd_rn((float)a[k], f[k]); a[k] = __dadd_rn(a[k], 1.0); }          // F2F.F32.F64 + FADD + DADD
            if (OP == 4) f[k] = __fadd_rn(f[k], 1.0000001f);
            if (OP == 5) a[k] = __dmul_rn(a[k], 1.0000001);
        }
    }
    double s = 0; for (int k = 0; k < 8; k++) s += a[k] + f[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char *name, int ops_per_iter)
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double *out; cudaMalloc(&out, sizeof(double) * sms * 4 * 512);
    const int iters = 4096;
    bench<OP><<<sms * 4, 512>>>(out, 16, 1.f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<sms * 4, 512>>>(out, iters, 1.f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)sms * 4 * 512 * iters * 8 * ops_per_iter;
    printf("%-28s %8.3f ms  %8.2f Gop/s  %6.2f op/clk/SM (at %d MHz nominal)\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    cudaFree(out);
}

int main()
{
    run<4>("FADD", 1);
    run<0>("DADD", 1);
    run<1>("DFMA", 1);
    run<5>("DMUL", 1);
    run<2>("F2F.F64.F32+DADD+FADD", 1);
    run<3>("F2F.F32.F64+FADD+DADD", 1);
    return 0;
}
