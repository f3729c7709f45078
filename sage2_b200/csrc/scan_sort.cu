// scan_sort.cu -- exclusive scan and stable LSD radix sort (sm_100a, hand written).
//
// Radix sort, one 1..8-bit digit per pass, three launches per pass:
//   rs_hist    : per-tile digit histogram               -> tileHist[digit][tile]
//   scan       : exclusive scan of the flattened array  -> global base of (digit, tile)
//   rs_scatter : stable rank inside the tile (warp match_any ranking, warps own contiguous item
//                ranges so tile order == input order) and scatter of every column.
// HBM traffic per pass: key column read twice, every column read once and written once.
#include "device_utils.cuh"

namespace sg {

// ------------------------------------------------------------------------------------------------
// exclusive scan
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ u32 block_exclusive_scan_256(u32 v, u32 *smem /*[8]*/, u32 &block_total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) smem[warp] = x;
    __syncthreads();
    u32 warp_prefix = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        u32 s = smem[w];
        if (w < warp) warp_prefix += s;
        total += s;
    }
    block_total = total;
    __syncthreads();
    return warp_prefix + x - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const u32 *__restrict__ in, u32 *__restrict__ sums, u64 n)
{
    __shared__ u32 sm[8];
    const u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) s += in[base + i];
    u32 total;
    block_exclusive_scan_256(s, sm, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_kernel(const u32 *in, u32 *out, u64 n, const u32 *__restrict__ block_prefix, u32 *d_total)
{
    __shared__ u32 sm[8];
    const u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    u32 v[SCAN_ITEMS];
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? in[base + i] : 0u; s += v[i]; }
    u32 total;
    u32 ex = block_exclusive_scan_256(s, sm, total);
    if (block_prefix) ex += block_prefix[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
    if (d_total && gridDim.x == 1 && threadIdx.x == 0) *d_total = total;
}

void exclusive_scan_u32(const u32 *in, u32 *out, u64 n, u32 *d_total, cudaStream_t st)
{
    if (n == 0) {
        if (d_total) SG_CUDA(cudaMemsetAsync(d_total, 0, sizeof(u32), st));
        return;
    }
    const u64 nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (nb == 1) {
        scan_tile_kernel<<<1, SCAN_THREADS, 0, st>>>(in, out, n, nullptr, d_total);
        SG_LAUNCHED();
        return;
    }
    DevBuf<u32> sums(nb, st);
    scan_reduce_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, sums.p, n);
    SG_LAUNCHED();
    exclusive_scan_u32(sums.p, sums.p, nb, d_total, st);
    scan_tile_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, out, n, sums.p, nullptr);
    SG_LAUNCHED();
}

// ------------------------------------------------------------------------------------------------
// OR / AND reduction of a key column
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) or_and_kernel(const u64 *__restrict__ keys, u64 n, u64 *out)
{
    u64 o = 0, a = ~0ull;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 k = keys[i];
        o |= k; a &= k;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        o |= __shfl_xor_sync(0xffffffffu, o, s);
        a &= __shfl_xor_sync(0xffffffffu, a, s);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicOr((unsigned long long *)&out[0], (unsigned long long)o);
        atomicAnd((unsigned long long *)&out[1], (unsigned long long)a);
    }
}

void reduce_or_and_u64(const u64 *keys, u64 n, u64 *d_or_and, cudaStream_t st)
{
    const u64 init[2] = { 0ull, ~0ull };
    SG_CUDA(cudaMemcpyAsync(d_or_and, init, sizeof(init), cudaMemcpyHostToDevice, st));
    if (n == 0) return;
    unsigned g = grid_for(n, 256, 8);
    if (g > kSMs * 8) g = kSMs * 8;
    or_and_kernel<<<g, 256, 0, st>>>(keys, n, d_or_and);
    SG_LAUNCHED();
}

// ------------------------------------------------------------------------------------------------
// radix sort
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;                         // per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;       // 4096 records per block
constexpr int RS_WARP_ITEMS = 32 * RS_ITEMS;         // contiguous range owned by one warp

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const u64 *__restrict__ key, u64 n, int shift, u32 mask,
                                                              u32 *__restrict__ tile_hist, u32 num_tiles)
{
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const u64 base = (u64)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int i = 0; i < RS_ITEMS; ++i) {
        const u64 idx = base + (u64)i * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(u32)(key[idx] >> shift) & mask], 1u);
    }
    __syncthreads();
    if (threadIdx.x <= mask) tile_hist[(u64)threadIdx.x * num_tiles + blockIdx.x] = h[threadIdx.x];
}

// One stable pass: the tile's records are first brought into digit order in shared memory, one column at a time, and then
// written out position by position -- the records of one digit leave as contiguous runs (coalesced stores) instead of 4096
// scattered 8-byte stores per tile.
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const u64 *__restrict__ key, u64 n, int shift, u32 mask,
                                                                 const u32 *__restrict__ tile_off, u32 num_tiles,
                                                                 const u64 *__restrict__ a_in, u64 *__restrict__ a_out,
                                                                 const u64 *__restrict__ b_in, u64 *__restrict__ b_out,
                                                                 const u32 *__restrict__ v_in, u32 *__restrict__ v_out)
{
    __shared__ u32 whist[RS_WARPS][256];
    __shared__ u32 dbase[256];           // global position of the tile's first record of digit d
    __shared__ u32 dstart[256];          // position inside the tile of its first record of digit d
    __shared__ u32 wsum[RS_WARPS];
    __shared__ uint8_t sdig[RS_TILE];    // digit of the record at tile position j (after the reordering)
    __shared__ u64 stage[RS_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();

    const u64 tbase = (u64)blockIdx.x * RS_TILE;
    const u64 wbase = tbase + (u64)warp * RS_WARP_ITEMS;
    const u32 in_tile = (u32)((n - tbase) < (u64)RS_TILE ? (n - tbase) : (u64)RS_TILE);
    u32 packed[RS_ITEMS];    // digit | rank-in-warp << 8, then the record's position inside the tile
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const u64 idx = wbase + (u64)r * 32 + lane;
        const bool valid = idx < n;
        const u32 d = valid ? ((u32)(key[idx] >> shift) & mask) : 0xFFFFFFFFu;
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        const unsigned peers = __match_any_sync(0xffffffffu, d) & vm;
        u32 rank = 0;
        if (valid) rank = whist[warp][d] + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
        if (valid && lane == (__ffs(peers) - 1)) whist[warp][d] += __popc(peers);
        __syncwarp();
        packed[r] = valid ? (d | (rank << 8)) : 0xFFFFFFFFu;
    }
    __syncthreads();
    // per digit: exclusive scan over the warps, the tile's total, the global base of (digit, tile)
    u32 total;
    {
        const int d = threadIdx.x;
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { const u32 t = whist[w][d]; whist[w][d] = run; run += t; }
        total = run;
        dbase[d] = ((u32)d <= mask) ? tile_off[(u64)d * num_tiles + blockIdx.x] : 0u;
    }
    // exclusive scan of the totals over the digits -> where each digit starts inside the tile
    u32 inc = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const u32 y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    {
        u32 before = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) if (w < warp) before += wsum[w];
        dstart[threadIdx.x] = before + inc - total;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        if (packed[r] == 0xFFFFFFFFu) continue;
        const u32 d = packed[r] & 0xFF;
        const u32 lpos = dstart[d] + whist[warp][d] + (packed[r] >> 8);
        sdig[lpos] = (uint8_t)d;
        packed[r] = lpos;
    }
    // column a
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r)
        if (packed[r] != 0xFFFFFFFFu) stage[packed[r]] = a_in[wbase + (u64)r * 32 + lane];
    __syncthreads();
    for (u32 j = threadIdx.x; j < in_tile; j += RS_THREADS) { const u32 d = sdig[j]; a_out[(u64)dbase[d] + (j - dstart[d])] = stage[j]; }
    if (b_in) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RS_ITEMS; ++r)
            if (packed[r] != 0xFFFFFFFFu) stage[packed[r]] = b_in[wbase + (u64)r * 32 + lane];
        __syncthreads();
        for (u32 j = threadIdx.x; j < in_tile; j += RS_THREADS) { const u32 d = sdig[j]; b_out[(u64)dbase[d] + (j - dstart[d])] = stage[j]; }
    }
    if (v_in) {
        __syncthreads();
        u32 *stage32 = reinterpret_cast<u32 *>(stage);
#pragma unroll
        for (int r = 0; r < RS_ITEMS; ++r)
            if (packed[r] != 0xFFFFFFFFu) stage32[packed[r]] = v_in[wbase + (u64)r * 32 + lane];
        __syncthreads();
        for (u32 j = threadIdx.x; j < in_tile; j += RS_THREADS) { const u32 d = sdig[j]; v_out[(u64)dbase[d] + (j - dstart[d])] = stage32[j]; }
    }
}

int radix_sort_bits(SortCols &c, int cur, u64 n, bool use_b, int lo, int hi, cudaStream_t st)
{
    if (n <= 1 || hi <= lo) return cur;
    SG_CHECK(n < 0xFFFFFFFFull, "radix sort supports < 2^32 records");
    const int bits = hi - lo;
    const int passes = (bits + 7) / 8;
    const int width = (bits + passes - 1) / passes;
    const u32 num_tiles = (u32)((n + RS_TILE - 1) / RS_TILE);
    DevBuf<u32> hist((u64)256 * num_tiles, st);
    int shift = lo;
    for (int p = 0; p < passes; ++p) {
        const int w = (hi - shift) < width ? (hi - shift) : width;
        const u32 mask = (1u << w) - 1u;
        const u64 *key = use_b ? c.b[cur] : c.a[cur];
        rs_hist_kernel<<<num_tiles, RS_THREADS, 0, st>>>(key, n, shift, mask, hist.p, num_tiles);
        SG_LAUNCHED();
        exclusive_scan_u32(hist.p, hist.p, (u64)(mask + 1) * num_tiles, nullptr, st);
        rs_scatter_kernel<<<num_tiles, RS_THREADS, 0, st>>>(key, n, shift, mask, hist.p, num_tiles,
                                                            c.a[cur], c.a[cur ^ 1], c.b[cur], c.b[cur ^ 1],
                                                            c.v[cur], c.v[cur ^ 1]);
        SG_LAUNCHED();
        cur ^= 1;
        shift += w;
    }
    return cur;
}

int radix_sort_varying(SortCols &c, int cur, u64 n, bool use_b, cudaStream_t st)
{
    if (n <= 1) return cur;
    DevBuf<u64> oa(2, st);
    reduce_or_and_u64(use_b ? c.b[cur] : c.a[cur], n, oa.p, st);
    u64 h[2];
    SG_CUDA(cudaMemcpyAsync(h, oa.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    const u64 varying = h[0] ^ h[1];
    if (varying == 0) return cur;
    const int lo = __builtin_ctzll(varying), hi = 64 - __builtin_clzll(varying);
    return radix_sort_bits(c, cur, n, use_b, lo, hi, st);
}

// ------------------------------------------------------------------------------------------------
// random-sector gather microbenchmark: the denominator SURVEY.md 8(d) asks for.  Every thread reads
// `per_thread` independent, uniformly random, `granule`-byte aligned blocks (32 or 64 bytes) of a
// buffer far larger than L2 and folds them into one word.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_bench_kernel(const uint4 *__restrict__ buf, u64 n_granules, int vec_per_granule,
                                                            int per_thread, u64 seed, u32 *__restrict__ out)
{
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 x = (tid + 1) * 0x9e3779b97f4a7c15ull + seed;
    u32 acc = 0;
    for (int r = 0; r < per_thread; r += 4) {
        uint4 v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
            const u64 g = __umul64hi(x, n_granules);
            const uint4 *p = buf + g * vec_per_granule;
#pragma unroll
            for (int q = 0; q < 4; ++q) if (q < vec_per_granule) v[u][q] = __ldg(p + q);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) if (q < vec_per_granule) acc ^= v[u][q].x ^ v[u][q].y ^ v[u][q].z ^ v[u][q].w;
    }
    out[tid] = acc;
}

__device__ __forceinline__ void ldg256(const void *p, u64 (&v)[4])
{
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}

// mode 1: a granule of 32 / 64 bytes = one / two 256-bit loads of one lane.
// mode 2: a 64-byte granule split over a lane pair, one 256-bit load each (one L1 request, two sectors).
__global__ void __launch_bounds__(256) gather_bench256_kernel(const uint8_t *__restrict__ buf, u64 n_granules, int granule, int mode,
                                                               int per_thread, u64 seed, u32 *__restrict__ out)
{
    const u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 stream_id = mode == 2 ? (tid >> 1) : tid;
    u64 x = (stream_id + 1) * 0x9e3779b97f4a7c15ull + seed;
    u64 acc = 0;
    for (int r = 0; r < per_thread; r += 4) {
        u64 v[4][2][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
            const u64 g = __umul64hi(x, n_granules);
            const uint8_t *p = buf + g * (u64)granule;
            if (mode == 2) { ldg256(p + 32 * (tid & 1), v[u][0]); }
            else {
                ldg256(p, v[u][0]);
                if (granule == 64) ldg256(p + 32, v[u][1]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            acc ^= v[u][0][0] ^ v[u][0][1] ^ v[u][0][2] ^ v[u][0][3];
            if (mode != 2 && granule == 64) acc ^= v[u][1][0] ^ v[u][1][1] ^ v[u][1][2] ^ v[u][1][3];
        }
    }
    out[tid] = (u32)(acc ^ (acc >> 32));
}

float gather_bench(u64 footprint_bytes, int granule_bytes, u64 n_loads, int mode, cudaStream_t st)
{
    SG_CHECK(granule_bytes == 16 || granule_bytes == 32 || granule_bytes == 64, "granule must be 16, 32 or 64 bytes");
    SG_CHECK(mode >= 0 && mode <= 2 && (mode == 0 || granule_bytes >= 32) && (mode != 2 || granule_bytes == 64), "bad gather mode");
    const int vpg = granule_bytes / 16, per_thread = 64;
    const u64 n_granules = footprint_bytes / (u64)granule_bytes;
    SG_CHECK(n_granules > 0, "empty footprint");
    u64 threads = n_loads / per_thread * (mode == 2 ? 2 : 1);
    threads = (threads + 255) / 256 * 256;
    DevBuf<uint4> buf((size_t)(n_granules * vpg), st);
    DevBuf<u32> out((size_t)threads, st);
    SG_CUDA(cudaMemsetAsync(buf.p, 1, buf.bytes(), st));
    cudaEvent_t e0, e1;
    SG_CUDA(cudaEventCreate(&e0)); SG_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        SG_CUDA(cudaEventRecord(e0, st));
        if (mode == 0) gather_bench_kernel<<<(unsigned)(threads / 256), 256, 0, st>>>(buf.p, n_granules, vpg, per_thread, 1234 + it, out.p);
        else gather_bench256_kernel<<<(unsigned)(threads / 256), 256, 0, st>>>((const uint8_t *)buf.p, n_granules, granule_bytes, mode, per_thread, 1234 + it, out.p);
        SG_LAUNCHED();
        SG_CUDA(cudaEventRecord(e1, st));
        SG_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double granules = (double)threads * per_thread / (mode == 2 ? 2 : 1);
    return (float)(granules * granule_bytes / (best * 1e-3) / 1e9);
}

}  // namespace sg
