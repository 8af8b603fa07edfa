// seqsum.cu -- removeDCBias's sequential f32 accumulator (processor.go:304-309), bit for bit, in
// four launches instead of one dependent chain of n additions (seqsum_core.cuh has the
// argument): exact f64 chunk sums -> their prefix (a guess of the running sum's binade at
// every chunk) -> per chunk and guess the integer summary -> one warp per chain walks the
// chunks in order with the true sum, O(1) per chunk where the summary provably applies and
// sample by sample elsewhere.  The old kernel (k_seqsum, preprocess.cu) stays for short
// signals and as the reference the tests compare against.
#include "kernels.h"
#include "seqsum_core.cuh"

namespace tdoa {

using namespace seqsum;

namespace {

constexpr int kBatch = 16;                                   // chunks staged per step of the walk (lanes 0..15 of the scan)
constexpr int kInfoBytes = kBatch * kGuesses * (int)sizeof(ChunkRule);   // 1536
constexpr int kSampBytes = kBatch * kChunk * (int)sizeof(float);         // 16384

__global__ void __launch_bounds__(256) k_seq_chunksum(const SeqJob *jobs)
{
    const SeqJob &J = jobs[blockIdx.y];
    const i64 c = (i64)blockIdx.x * 256 + threadIdx.x;
    if (!J.x || c >= J.n_chunks) return;
    const float *__restrict__ x = J.x + c * kChunk;
    const int count = (int)min((i64)kChunk, J.n - c * kChunk);
    double s = 0.0;   // exact: a chunk of f32 values fits a f64 sum without rounding for any realistic range
    for (int i = 0; i < count; i++) s += (double)x[i];
    J.csum[c] = s;
}

// exclusive prefix of the chunk sums (any association: it only feeds the binade guess)
__global__ void __launch_bounds__(1024) k_seq_scan(const SeqJob *jobs)
{
    __shared__ double s_tot[1024];
    const SeqJob &J = jobs[blockIdx.x];
    if (!J.x) return;
    const int t = threadIdx.x;
    const i64 per = (J.n_chunks + 1023) / 1024;
    const i64 c0 = min(J.n_chunks, (i64)t * per), c1 = min(J.n_chunks, c0 + per);
    double loc = 0.0;
    for (i64 c = c0; c < c1; c++) loc += J.csum[c];
    s_tot[t] = loc;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const double v = t >= o ? s_tot[t - o] : 0.0;
        __syncthreads();
        s_tot[t] += v;
        __syncthreads();
    }
    double run = s_tot[t] - loc;   // exclusive
    for (i64 c = c0; c < c1; c++) {
        const double v = J.csum[c];
        J.pre[c] = run;
        run += v;
    }
}

__global__ void __launch_bounds__(256) k_seq_analyse(const SeqJob *jobs)
{
    const SeqJob &J = jobs[blockIdx.y];
    const i64 item = (i64)blockIdx.x * 256 + threadIdx.x;   // chunk * kGuesses + guess
    if (!J.x || item >= J.n_chunks * kGuesses) return;
    const i64 c = item / kGuesses;
    const int k = (int)(item % kGuesses);
    const int count = (int)min((i64)kChunk, J.n - c * kChunk);
    reinterpret_cast<ChunkRule *>(J.infos)[item] = chunk_rule(chunk_analyse(J.x + c * kChunk, count, guess_binade(J.pre[c], k)));
}

// one warp per chain: all lanes stage the next batch with cp.async and take part in the scan of the hops
__global__ void __launch_bounds__(32) k_seq_apply(const SeqJob *jobs)
{
    __shared__ __align__(16) unsigned char s_info[2][kInfoBytes];
    __shared__ __align__(16) float s_x[2][kBatch * kChunk];
    const SeqJob &J = jobs[blockIdx.x];
    const int lane = threadIdx.x;
    if (!J.x) {
        if (lane == 0) *J.out = 0.0;
        return;
    }
    const i64 n = J.n, nc = J.n_chunks;
    const i64 n_batches = (nc + kBatch - 1) / kBatch;
    const ChunkRule *__restrict__ infos = reinterpret_cast<const ChunkRule *>(J.infos);
    auto stage = [&](i64 b, int buf) {
        if (b < n_batches) {
            const i64 c0 = b * kBatch;
            const int chunks = (int)min((i64)kBatch, nc - c0);
            const unsigned char *gi = reinterpret_cast<const unsigned char *>(infos + c0 * kGuesses);
            const int info16 = chunks * kGuesses * (int)sizeof(ChunkRule) / 16;
            for (int j = lane; j < info16; j += 32) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_info[buf][16 * j]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gi + 16 * j) : "memory");
            }
            const i64 first = c0 * kChunk;
            const i64 avail = min((i64)chunks * kChunk, n - first);
            const bool vec = (reinterpret_cast<uintptr_t>(J.x + first) & 15) == 0;
            const int full16 = vec ? (int)(avail / 4) : 0;   // whole 16-byte granules inside the signal
            for (int j = lane; j < full16; j += 32) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_x[buf][4 * j]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(J.x + first + 4 * j) : "memory");
            }
            for (i64 i = 4 * (i64)full16 + lane; i < avail; i += 32) s_x[buf][i] = J.x[first + i];   // ragged tail / unaligned plane
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float s = 0.f;
    stage(0, 0);
    for (i64 b = 0; b < n_batches; b++) {
        const int buf = (int)(b & 1);
        stage(b + 1, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        {
            // Every lane carries the same sum (the whole walk is warp-uniform).  The hops of the
            // batch's remaining chunks, for the binade and sign the sum has now, are combined by an
            // inclusive scan; a ballot finds how far the sum can be taken in one step, the chunk
            // after that (the sum leaves its binade there, or no guess was made for it) is added
            // sample by sample, and the rest of the batch is scanned again from the new sum.
            const i64 c0 = b * kBatch;
            const int chunks = (int)min((i64)kBatch, nc - c0);
            const ChunkRule *ci = reinterpret_cast<const ChunkRule *>(s_info[buf]);
            int pos = 0;
            while (pos < chunks) {
                int es;
                bool neg;
                const int m = mantissa_signed(s, &es, &neg);
                Hop f = lane < pos ? hop_identity() : (lane < chunks ? hop_from_rules(ci + lane * kGuesses, es, neg) : hop_empty());
#pragma unroll
                for (int off = 1; off < kBatch; off <<= 1) {
                    Hop p;
                    p.d[0] = __shfl_up_sync(0xffffffffu, f.d[0], off); p.d[1] = __shfl_up_sync(0xffffffffu, f.d[1], off);
                    p.lo[0] = __shfl_up_sync(0xffffffffu, f.lo[0], off); p.lo[1] = __shfl_up_sync(0xffffffffu, f.lo[1], off);
                    p.hi[0] = __shfl_up_sync(0xffffffffu, f.hi[0], off); p.hi[1] = __shfl_up_sync(0xffffffffu, f.hi[1], off);
                    if (lane >= off) f = hop_compose(p, f);
                }
                const unsigned ok = __ballot_sync(0xffffffffu, lane < pos || hop_admits(f, m));
                const int upto = ok == 0xffffffffu ? chunks : min(chunks, __ffs(~ok) - 1);
                const int d = __shfl_sync(0xffffffffu, (m & 1) ? f.d[1] : f.d[0], max(upto - 1, 0));
                if (upto > pos) {
                    int mo = m + d;
                    if (mo < 0) mo = -mo;
                    s = __uint_as_float((__float_as_uint(s) & 0xff800000u) | ((unsigned)mo & 0x7fffffu));
                    pos = upto;
                }
                if (pos >= chunks) break;
                // chunk pos, sample by sample -- loads first, then the chain
                const int count = (int)min((i64)kChunk, n - (c0 + pos) * kChunk);
                const float4 *xv = reinterpret_cast<const float4 *>(s_x[buf] + pos * kChunk);
                if (count == kChunk) {
#pragma unroll 1
                    for (int q = 0; q < kChunk / 64; q++) {   // 64 samples at a time: 16 loads in flight, then their chain
                        float4 v[16];
#pragma unroll
                        for (int j = 0; j < 16; j++) v[j] = xv[16 * q + j];
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            s = __fadd_rn(s, v[j].x); s = __fadd_rn(s, v[j].y); s = __fadd_rn(s, v[j].z); s = __fadd_rn(s, v[j].w);
                        }
                    }
                } else {
                    for (int i = 0; i < count; i++) s = __fadd_rn(s, s_x[buf][pos * kChunk + i]);
                }
                pos++;
            }
        }
        __syncwarp();   // the buffer is free for the batch after next
    }
    if (lane == 0) *J.out = n > 0 ? (double)__fdiv_rn(s, (float)n) : 0.0;   // processor.go:309
}

}  // namespace

size_t seqsum_scratch_bytes(i64 n)
{
    const i64 nc = (n + kChunk - 1) / kChunk;
    return (size_t)nc * (2 * sizeof(double) + kGuesses * sizeof(ChunkRule)) + 256;
}

i64 seqsum_chunks(i64 n) { return (n + kChunk - 1) / kChunk; }

void seqsum_carve(SeqJob &J, void *scratch)
{
    J.n_chunks = seqsum_chunks(J.n);
    unsigned char *p = static_cast<unsigned char *>(scratch);
    J.infos = p;                                   // 32-byte records first: keeps them 16-byte aligned
    p += (size_t)J.n_chunks * kGuesses * sizeof(ChunkRule);
    J.csum = reinterpret_cast<double *>(p);
    J.pre = J.csum + J.n_chunks;
}

void launch_seqsum_chunked(const SeqJob *d_jobs, int n_jobs, i64 max_n, cudaStream_t st)
{
    if (n_jobs <= 0) return;
    const i64 nc = seqsum_chunks(max_n);
    if (nc > 0) {
        k_seq_chunksum<<<dim3((unsigned)((nc + 255) / 256), n_jobs), 256, 0, st>>>(d_jobs);
        k_seq_scan<<<n_jobs, 1024, 0, st>>>(d_jobs);
        k_seq_analyse<<<dim3((unsigned)((nc * kGuesses + 255) / 256), n_jobs), 256, 0, st>>>(d_jobs);
    }
    k_seq_apply<<<n_jobs, 32, 0, st>>>(d_jobs);
}

}  // namespace tdoa
