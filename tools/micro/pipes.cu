// Microbenchmark: per-SM issue rates of the instruction classes the FM discriminator is
// made of, alone and in pairs, so that its per-sample pipe budget (issue / ALU / FP64 /
// XU) can be written down from measurements rather than from the programming guide.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/pipes.cu -o tools/micro/pipes
#include <cstdio>
#include <cuda_runtime.h>

enum Op { FFMA, DFMA, DADD, LOP, IADD, SHF, FSEL, PRMT, IMAD, RCP32, RCP64H, F2F_DF, F2F_FD, LDS128, LDS64,
          MIX_DFMA_LOP, MIX_DFMA_FFMA, MIX_DFMA_F2F, MIX_LOP_FFMA, MIX_DFMA_RCP64H, I2F,
          FFMA2, FMUL2, FADD2, FMNMX, MIX_FFMA2_LOP, MIX_FFMA2_DFMA, MIX_FFMA2_FFMA, F2F_DF_ONLY, F2F_FD_ONLY, LDS32R, LDS64R, NOPS };

template <int OP>
__global__ void __launch_bounds__(512) bench(unsigned long long *out, int iters, unsigned seed)
{
    __shared__ __align__(16) double sm[512 * 2];
    sm[threadIdx.x * 2] = threadIdx.x; sm[threadIdx.x * 2 + 1] = seed;
    __syncthreads();
    double d[8]; float f[8]; unsigned u[8]; unsigned long long q[8]; const unsigned long long qc = 0x3f8000013f800001ull;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
    for (int k = 0; k < 8; k++) { d[k] = 1.0 + (threadIdx.x + k) * 1e-3; f[k] = 1.0f + (threadIdx.x + k) * 1e-3f; u[k] = seed + threadIdx.x * 8 + k; q[k] = 0x3f8000003f800000ull + u[k]; }
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 15) * 16;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (OP == FFMA || OP == MIX_DFMA_FFMA || OP == MIX_LOP_FFMA) f[k] = fmaf(f[k], 1.0000001f, 0.5f);
            if (OP == DFMA || OP == MIX_DFMA_LOP || OP == MIX_DFMA_FFMA || OP == MIX_DFMA_F2F || OP == MIX_DFMA_RCP64H)
                d[k] = fma(d[k], 1.0000001, 0.5);
            if (OP == DADD) d[k] = __dadd_rn(d[k], 1.0000001);
            if (OP == LOP || OP == MIX_DFMA_LOP || OP == MIX_LOP_FFMA) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[k]) : "r"(u[(k + 1) & 7]), "r"(seed));
            if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[k]) : "r"(u[(k + 1) & 7]));
            if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, 5;" : "+r"(u[k]) : "r"(u[(k + 1) & 7]));
            if (OP == FSEL) asm volatile("{.reg .pred p; setp.lt.u32 p, %1, %2; selp.b32 %0, %0, %1, p;}" : "+r"(u[k]) : "r"(u[(k + 1) & 7]), "r"(seed));
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(u[k]) : "r"(u[(k + 1) & 7]));
            if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[k]) : "r"(u[(k + 1) & 7]), "r"(seed));
            if (OP == RCP32) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[k])); f[k] = __fadd_rn(f[k], 1.5f); }   // MUFU.RCP + FADD
            if (OP == RCP64H || OP == MIX_DFMA_RCP64H) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(d[k]));
            if (OP == F2F_DF) { f[k] = __fadd_rn((float)d[k], f[k]); d[k] = __dadd_rn(d[k], 1.0000001); }   // F2F.F32.F64 + FADD + DADD
            if (OP == F2F_FD || OP == MIX_DFMA_F2F) { d[(k + 1) & 7] = __dadd_rn((double)f[k], d[(k + 1) & 7]); }   // F2F.F64.F32 + DADD
            if (OP == I2F) { f[k] = __fadd_rn((float)(int)u[k], f[k]); u[k] += seed; }
            if (OP == LDS128) { double a, b; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "r"(saddr + (u[k] & 0x1f00))); u[k] += __double2loint(a) + __double2loint(b); }
            if (OP == FFMA2 || OP == MIX_FFMA2_LOP || OP == MIX_FFMA2_DFMA || OP == MIX_FFMA2_FFMA) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[k]) : "l"(q[(k + 1) & 7]), "l"(qc));
            if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(q[k]) : "l"(qc));
            if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[k]) : "l"(qc));
            if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[k]) : "f"(f[(k + 1) & 7]));
            if (OP == MIX_FFMA2_LOP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[k]) : "r"(u[(k + 1) & 7]), "r"(seed));
            if (OP == MIX_FFMA2_DFMA) d[k] = fma(d[k], 1.0000001, 0.5);
            if (OP == MIX_FFMA2_FFMA) f[k] = fmaf(f[k], 1.0000001f, 0.5f);
            if (OP == F2F_DF_ONLY) { float t; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(t) : "d"(d[k])); u[k] ^= __float_as_uint(t); }
            if (OP == F2F_FD_ONLY) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(f[k])); u[k] ^= (unsigned)__double2hiint(t); }
            if (OP == LDS32R) { unsigned a; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a) : "r"(sbase + ((u[k] * 2654435761u >> 20) & 0xffcu))); u[k] += a; }
            if (OP == LDS64R) { double a; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(sbase + ((u[k] * 2654435761u >> 20) & 0xff8u))); u[k] += __double2loint(a); }
            if (OP == LDS64) { double a; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(saddr + (u[k] & 0x1f00))); u[k] += __double2loint(a); }
        }
    }
    unsigned long long s = 0;
    for (int k = 0; k < 8; k++) s += (unsigned long long)__double_as_longlong(d[k]) + __float_as_uint(f[k]) + u[k] + q[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char *name, double ops_per_slot, const char *note = "")
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *out; cudaMalloc(&out, sizeof(*out) * sms * 4 * 512);
    const int iters = 2048;
    bench<OP><<<sms * 4, 512>>>(out, 16, 1u);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<sms * 4, 512>>>(out, iters, 1u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double slots = (double)sms * 4 * 512 * iters * 8;   // per-thread "slots" (one of each op of the mix)
    // 1965 MHz is this pool's max SM clock; the figure is thread-ops per clock per SM at that clock
    printf("%-22s %8.3f ms  %7.2f thread-slots/clk/SM  (%.1f cycles per warp-slot per SMSP)  %s\n", name, ms,
           slots / (ms * 1e-3) / sms / 1.965e9, 32.0 / (slots / (ms * 1e-3) / sms / 1.965e9 / 4.0), note);
    (void)ops_per_slot;
    cudaFree(out);
}

int main()
{
    run<FFMA>("FFMA", 1);
    run<DFMA>("DFMA", 1);
    run<DADD>("DADD", 1);
    run<LOP>("LOP3", 1);
    run<IADD>("IADD", 1);
    run<SHF>("SHF", 1);
    run<FSEL>("ISETP+SEL", 2);
    run<PRMT>("PRMT", 1);
    run<IMAD>("IMAD", 1);
    run<RCP32>("MUFU.RCP+FADD", 2);
    run<RCP64H>("MUFU.RCP64H", 1);
    run<F2F_DF>("F2F.F32.F64+FADD+DADD", 3);
    run<F2F_FD>("F2F.F64.F32+DADD", 2);
    run<I2F>("I2F+FADD", 2);
    run<LDS128>("LDS.128+2 IADD+LOP", 4);
    run<LDS64>("LDS.64+IADD+LOP", 3);
    run<FFMA2>("FFMA2", 1);
    run<FMUL2>("FMUL2", 1);
    run<FADD2>("FADD2", 1);
    run<FMNMX>("FMNMX", 1);
    run<F2F_DF_ONLY>("F2F.F32.F64+LOP", 2);
    run<F2F_FD_ONLY>("F2F.F64.F32+LOP", 2);
    run<LDS32R>("LDS.32 random+IMAD+..", 3);
    run<LDS64R>("LDS.64 random+IMAD+..", 3);
    run<MIX_FFMA2_LOP>("FFMA2+LOP3", 2);
    run<MIX_FFMA2_DFMA>("FFMA2+DFMA", 2);
    run<MIX_FFMA2_FFMA>("FFMA2+FFMA", 2);
    run<MIX_DFMA_LOP>("DFMA+LOP3", 2, "(sum of the two alone = no overlap)");
    run<MIX_DFMA_FFMA>("DFMA+FFMA", 2);
    run<MIX_DFMA_F2F>("DFMA+F2F.F64.F32+DADD", 3);
    run<MIX_LOP_FFMA>("LOP3+FFMA", 2);
    run<MIX_DFMA_RCP64H>("DFMA+RCP64H", 2);
    return 0;
}
