// Microbenchmark: gather throughput of the texture path against shared memory for the discriminator's
// small tables (per-SM, random indices).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(512) k(cudaTextureObject_t t1, cudaTextureObject_t t2, const float2 *g, unsigned *out, int iters)
{
    __shared__ float2 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = g[i];
    __syncthreads();
    unsigned u = threadIdx.x * 2654435761u + blockIdx.x;
    float acc = 0.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
            u = u * 1664525u + 1013904223u;
            if (MODE == 0) { const float2 v = tex1Dfetch<float2>(t1, (int)(u >> 22)); acc += v.x + v.y; }
            if (MODE == 1) { const float v = tex2D<float>(t2, (float)(u >> 24), (float)((u >> 16) & 255)); acc += v; }
            if (MODE == 2) { const float2 v = sm[u >> 22]; acc += v.x + v.y; }
            if (MODE == 3) { const float2 v = __ldg(g + (u >> 22)); acc += v.x + v.y; }
            if (MODE == 4) { acc += __uint_as_float(u >> 9); }
            if (MODE == 5) { const float2 v = tex1Dfetch<float2>(t1, (int)((u >> 27) + (threadIdx.x & 31))); acc += v.x + v.y; }   // clustered indices
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(acc) + u;
}
template <int MODE>
void run(const char *name, cudaTextureObject_t t1, cudaTextureObject_t t2, const float2 *g, unsigned *out)
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 512;
    k<MODE><<<sms * 4, 512>>>(t1, t2, g, out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms * 4, 512>>>(t1, t2, g, out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_ops = (double)sms * 4 * 16 * iters * 8;
    printf("%-36s %8.3f ms  %6.2f SM-cycles per warp gather (at 1965 MHz)  err=%s\n", name, ms,
           ms * 1e-3 * 1.965e9 / (warp_ops / sms), cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    float2 *g; cudaMalloc(&g, 1024 * sizeof(float2));
    float2 h[1024]; for (int i = 0; i < 1024; i++) h[i] = make_float2(i, -i);
    cudaMemcpy(g, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = g;
    rd.res.linear.desc = cudaCreateChannelDesc<float2>(); rd.res.linear.sizeInBytes = sizeof(h);
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t t1 = 0; cudaCreateTextureObject(&t1, &rd, &td, nullptr);
    cudaArray_t arr; cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
    cudaMallocArray(&arr, &cd, 256, 256);
    float *h2 = new float[65536]; for (int i = 0; i < 65536; i++) h2[i] = i;
    cudaMemcpy2DToArray(arr, 0, 0, h2, 256 * 4, 256 * 4, 256, cudaMemcpyHostToDevice);
    cudaResourceDesc rd2 = {}; rd2.resType = cudaResourceTypeArray; rd2.res.array.array = arr;
    cudaTextureDesc td2 = {}; td2.readMode = cudaReadModeElementType; td2.filterMode = cudaFilterModePoint;
    td2.addressMode[0] = td2.addressMode[1] = cudaAddressModeClamp; td2.normalizedCoords = 0;
    cudaTextureObject_t t2 = 0; cudaCreateTextureObject(&t2, &rd2, &td2, nullptr);
    unsigned *out; cudaMalloc(&out, 148 * 4 * 512 * 4);
    run<4>("baseline (index arithmetic only)", t1, t2, g, out);
    run<0>("tex1Dfetch float2, 1024 random", t1, t2, g, out);
    run<5>("tex1Dfetch float2, clustered", t1, t2, g, out);
    run<1>("tex2D float, 256x256 random", t1, t2, g, out);
    run<2>("LDS.64 random (1024 float2)", t1, t2, g, out);
    run<3>("LDG.64 (__ldg) random, L1 resident", t1, t2, g, out);
    return 0;
}
