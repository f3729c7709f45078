// digest.cu -- order-sensitive 64-bit digests of the resident unique-read table and edge list
// (sage2gpu_digest).  No reference counterpart: it is the parity gate of the measurements.  The same two
// sums are computed with numpy from the reference's own `.reads` / `.graph3` files (tests/digest.py:
// readLoader.cpp:270-287, overlapGraph.cpp:338-369), so a benchmark run proves on every rank that what it
// timed is the reference's result, bit for bit, without formatting a gigabyte of text per step.
//
//   edge at position e of the canonical list:  a = from << 32 | to,  b = type << 48 | delta << 24 | delta_twin
//       h = mix(mix(mix(e) + a) + b)                      edges_digest = sum h + mix(E)
//   unique read id (1-based):  h = mix(id); h = mix(h + (frequency << 16 | length));
//       for each of ceil(length / 32) words (32 bases, first base in the top bits, pad bits 0): h = mix(h + word)
//                                                          reads_digest = sum h + mix(U)
#include "context.h"

namespace sg {

__host__ __device__ __forceinline__ u64 dmix(u64 x)      // splitmix64 finaliser with the golden-ratio increment
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ void block_sum_to(u64 v, unsigned long long *out)
{
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, (unsigned long long)v);
}

__global__ void __launch_bounds__(256) digest_edges_kernel(const u64 *__restrict__ edges, u64 E, const uint16_t *__restrict__ len,
                                                           unsigned long long *__restrict__ out)
{
    u64 acc = 0;
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (u64)gridDim.x * blockDim.x) {
        const ulonglong2 w = reinterpret_cast<const ulonglong2 *>(edges)[e];
        const u64 from = w.x >> 32, to = w.x & 0xFFFFFFFFull;
        const u64 type = (w.y >> 20) & 3ull, delta = w.y & 0xFFFFFull;
        const u64 twin = (u64)len[from - 1] - ((u64)len[to - 1] - delta);      // overlapGraph.cpp:147
        acc += dmix(dmix(dmix(e) + w.x) + ((type << 48) | (delta << 24) | (twin & 0xFFFFFFull)));
    }
    block_sum_to(acc, out);
}

__global__ void __launch_bounds__(256) digest_reads_kernel(const u64 *__restrict__ F, int SW, int SWS, const uint16_t *__restrict__ len,
                                                           const uint16_t *__restrict__ freq, u64 U, unsigned long long *__restrict__ out)
{
    u64 acc = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < U; i += (u64)gridDim.x * blockDim.x) {
        const int l = len[i];
        u64 h = dmix(i + 1);
        h = dmix(h + (((u64)freq[i] << 16) | (u64)l));
        const int nw = (l + 31) >> 5;
        for (int w = 0; w < nw; ++w) {
            u64 x = F[i * SWS + w];
            if (w == SW - 1) x &= ~0xFFFFull;       // the record keeps the length in the low 16 bits of its last word
            h = dmix(h + x);
        }
        acc += h;
    }
    block_sum_to(acc, out);
}

void stage_digest(Context &c, u64 *reads_digest, u64 *edges_digest)
{
    cudaStream_t st = c.stream;
    ArenaScope arena_scope(c.arena, st);
    DevBuf<unsigned long long> d(2, st);
    SG_CUDA(cudaMemsetAsync(d.p, 0, 2 * sizeof(unsigned long long), st));
    const u64 U = c.cnt.unique_reads, E = c.cnt.n_edges;
    if (reads_digest) {
        SG_CHECK(c.have_reads, "no reads loaded");
        if (U) { digest_reads_kernel<<<kSMs * 8, 256, 0, st>>>(c.F.p, c.SW, c.SWS, c.len.p, c.freq.p, U, d.p); SG_LAUNCHED(); }
    }
    if (edges_digest) {
        SG_CHECK(c.have_graph, "overlap graph not built");
        if (E) { digest_edges_kernel<<<kSMs * 8, 256, 0, st>>>(c.edges.p, E, c.len.p, d.p + 1); SG_LAUNCHED(); }
    }
    unsigned long long h[2];
    SG_CUDA(cudaMemcpyAsync(h, d.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    if (reads_digest) *reads_digest = h[0] + dmix(U);
    if (edges_digest) *edges_digest = h[1] + dmix(E);
}

}  // namespace sg
