// synth.cu -- measurement aid (no reference counterpart): synthetic paired-end reads generated ON THE DEVICE, for the
// workloads whose input does not fit a host pipeline (BASELINE.json config #5: 3.1 Gbp, 30x, ~620 M reads = 93 GB of
// characters).  Same shape as sage2_b200/synth.py: a uniform-random genome over {A,C,G,T}, error-free
// fixed-length reads, mates interleaved, fragment start uniform, insert size ~ N(mu, sigma) clipped to >= 2L, mate 2 =
// reverse complement of the fragment end, whole fragments from either strand with equal probability -- but from a
// counter-based generator, so that any rank can produce any slice of the read set without the others: the genome is a
// pure function of (seed, position) and is never stored.
#include "context.h"

namespace sg {

__host__ __device__ __forceinline__ u64 smix(u64 x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// base (0..3) of genome position p: 32 positions share one 64-bit word of the generator
__device__ __forceinline__ int genome_base(u64 seed, u64 p) { return (int)((smix(seed ^ (p >> 5)) >> (2 * (p & 31))) & 3); }

// one thread per pair: reads 2 * pair (mate 1) and 2 * pair + 1 (mate 2), L characters each, into bases[(2 * (pair - first_pair) + m) * L ..]
__global__ void __launch_bounds__(256) synth_pairs_kernel(uint8_t *__restrict__ bases, int64_t *__restrict__ offsets, u64 first_pair, u64 n_pairs,
                                                          u64 genome_bp, int L, float mu, float sigma, u64 seed)
{
    const char ACGT[4] = { 'A', 'C', 'G', 'T' };
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n_pairs; t += (u64)gridDim.x * blockDim.x) {
        const u64 pair = first_pair + t;
        const u64 r0 = smix(seed * 0x100000001B3ull + 3 * pair), r1 = smix(seed * 0x100000001B3ull + 3 * pair + 1), r2 = smix(seed * 0x100000001B3ull + 3 * pair + 2);
        // insert size: mu + sigma * (sum of 12 uniforms - 6), clipped to [2L, genome]
        float g = -6.f;
        for (int q = 0; q < 6; ++q) { g += (float)((r0 >> (10 * q)) & 1023) * (1.f / 1024.f); g += (float)((r1 >> (10 * q)) & 1023) * (1.f / 1024.f); }
        long long ins = (long long)(mu + sigma * g + 0.5f);
        if (ins < 2 * L) ins = 2 * L;
        if ((u64)ins > genome_bp) ins = (long long)genome_bp;
        const u64 start = r2 % (genome_bp - (u64)ins + 1);
        const bool flip = (r1 >> 63) != 0;
        uint8_t *m1 = bases + (2 * t) * (u64)L, *m2 = m1 + L;
        uint8_t *a = flip ? m2 : m1, *b = flip ? m1 : m2;      // a: forward read at the fragment start, b: revcomp of its end
        for (int x = 0; x < L; ++x) {
            a[x] = (uint8_t)ACGT[genome_base(seed, start + (u64)x)];
            b[x] = (uint8_t)ACGT[3 - genome_base(seed, start + (u64)ins - 1 - (u64)x)];
        }
        offsets[2 * t] = (int64_t)((2 * t) * (u64)L);
        offsets[2 * t + 1] = (int64_t)((2 * t + 1) * (u64)L);
        if (t + 1 == n_pairs) offsets[2 * n_pairs] = (int64_t)((2 * n_pairs) * (u64)L);
    }
}

void stage_synth_reads(Context &c, uint8_t *d_bases, int64_t *d_offsets, u64 first_pair, u64 n_pairs, u64 genome_bp, int read_len, float mu, float sigma, u64 seed)
{
    SG_CHECK(d_bases && d_offsets && read_len >= 1 && genome_bp >= (u64)(2 * read_len), "bad synthetic read request");
    if (n_pairs == 0) { const int64_t z = 0; SG_CUDA(cudaMemcpyAsync(d_offsets, &z, sizeof(z), cudaMemcpyHostToDevice, c.stream)); SG_CUDA(cudaStreamSynchronize(c.stream)); return; }
    unsigned g = grid_for(n_pairs, 256, 4);
    if (g > kSMs * 16u) g = kSMs * 16u;
    synth_pairs_kernel<<<g, 256, 0, c.stream>>>(d_bases, d_offsets, first_pair, n_pairs, genome_bp, read_len, mu, sigma, seed);
    SG_LAUNCHED();
    SG_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace sg
