// table.cu -- step 2 on device: the prefix/suffix table (K3).  Restates
// HashTable::hashPrefixesAndSuffix / hashTableInsert (economyGraph/hashTable.cpp:70-188).
//
// Observable semantics kept exactly: key (h = min(k,64) bases) -> all (readId,type) sharing it, in
// (readId asc, type asc) order; a key with >= 100 entries is invisible to searches.  Layout:
//   entries[4U]  : (readId0<<2 | type), radix-sorted by the exact 128-bit key (stable, so each key's
//                  run is already in bucket order),
//   slots[cap]   : open-addressing index (linear probing over 32-byte sectors of 4 slots, load <= 0.5)
//                  filled with 64-bit atomicCAS, one slot per DISTINCT key: tag | min(count,127) |
//                  (the entry itself when count == 1, else the offset of the key's run) (core.cuh).
#include "context.h"

namespace sg {

__global__ void __launch_bounds__(256) gen_entries_kernel(const u64 *__restrict__ F, const u64 *__restrict__ RC,
                                                           const uint16_t *__restrict__ len, u64 U, int SW, int h,
                                                           u64 *__restrict__ k0, u64 *__restrict__ k1, u32 *__restrict__ val)
{
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < 4 * U; t += (u64)gridDim.x * blockDim.x) {
        const u64 rid = t >> 2;
        const int type = (int)(t & 3);
        const u64 *X = ((type & 2) ? RC : F) + rid * SW;
        const int l = len[rid];
        u64 v0, v1;
        extract_key(X, SW, (type & 1) ? l - h : 0, h, v0, v1);
        k0[t] = v0; k1[t] = v1; val[t] = (u32)t;
    }
}

__global__ void __launch_bounds__(256) key_flag_kernel(const u64 *__restrict__ k0, const u64 *__restrict__ k1, u64 n, u32 *__restrict__ flag)
{
    for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x)
        flag[p] = (p == 0 || k0[p] != k0[p - 1] || k1[p] != k1[p - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) group_start_kernel(const u32 *__restrict__ flag, const u32 *__restrict__ gidx, u64 n, u32 *__restrict__ gstart)
{
    for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (u64)gridDim.x * blockDim.x)
        if (flag[p]) gstart[gidx[p]] = (u32)p;
}

__global__ void __launch_bounds__(256) index_insert_kernel(const u64 *__restrict__ k0, const u64 *__restrict__ k1,
                                                            const u32 *__restrict__ val, const u32 *__restrict__ gstart,
                                                            u64 D, u64 n, u64 *__restrict__ slots, u64 nsec,
                                                            unsigned long long *over)
{
    unsigned long long my_over = 0;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < D; g += (u64)gridDim.x * blockDim.x) {
        const u64 p = gstart[g];
        const u64 e = (g + 1 < D) ? gstart[g + 1] : n;
        const u64 count = e - p;
        if (count >= (u64)kHashThreshold) my_over++;
        const u64 hsh = hash_key(k0[p], k1[p]);
        const u64 v = slot_encode(hsh, count, count == 1 ? (u64)val[p] : p);
        u64 sec = home_sector(hsh, nsec);
        bool done = false;
        while (!done) {
#pragma unroll
            for (int t = 0; t < kSlotsPerSector && !done; ++t)
                done = atomicCAS((unsigned long long *)&slots[kSlotsPerSector * sec + t], 0ull, (unsigned long long)v) == 0ull;
            sec = (sec + 1 == nsec) ? 0 : sec + 1;
        }
    }
    if (my_over) atomicAdd(over, my_over);
}

static unsigned big_grid(u64 n, unsigned block = 256)
{
    unsigned g = grid_for(n, block, 4);
    return g > kSMs * 16u ? kSMs * 16u : g;
}

void stage_build_table(Context &c)
{
    cudaStream_t st = c.stream;
    SG_CHECK(c.have_reads, "organize_reads must run before build_hash_table");
    const u64 U = c.cnt.unique_reads;
    const int SW = c.SW, h = c.h;
    c.cnt.distinct_keys = 0; c.cnt.keys_over_threshold = 0; c.cap = 0; c.cnt.table_capacity = 0;
    c.slots.release(); c.entries.release();
    if (U == 0) { c.have_table = true; return; }
    const u64 n = 4 * U;
    SG_CHECK(n < 0xFFFFFFFFull, "at most 2^30-1 unique reads per context");

    DevBuf<u64> a0(n, st), a1(n, st), b0(n, st), b1(n, st);
    DevBuf<u32> v0(n, st), v1(n, st);
    gen_entries_kernel<<<big_grid(n), 256, 0, st>>>(c.F.p, c.RC.p, c.len.p, U, SW, h, a0.p, b0.p, v0.p);
    SG_LAUNCHED();
    SortCols cols;
    cols.a[0] = a0.p; cols.a[1] = a1.p; cols.b[0] = b0.p; cols.b[1] = b1.p; cols.v[0] = v0.p; cols.v[1] = v1.p;
    int cur = 0;
    // LSD: low key word (last 32 bases) first, then the leading bases (all zero when h <= 32)
    cur = radix_sort_bits(cols, cur, n, true, 0, 2 * (h < 32 ? h : 32), st);
    if (h > 32) cur = radix_sort_bits(cols, cur, n, false, 0, 2 * (h - 32), st);

    DevBuf<u32> flag(n, st), gidx(n, st), d_total(1, st);
    key_flag_kernel<<<big_grid(n), 256, 0, st>>>(cols.a[cur], cols.b[cur], n, flag.p);
    SG_LAUNCHED();
    exclusive_scan_u32(flag.p, gidx.p, n, d_total.p, st);
    u32 D = 0;
    SG_CUDA(cudaMemcpyAsync(&D, d_total.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    DevBuf<u32> gstart(D, st);
    group_start_kernel<<<big_grid(n), 256, 0, st>>>(flag.p, gidx.p, n, gstart.p);
    SG_LAUNCHED();

    u64 nsec = (2 * (u64)D + kSlotsPerSector - 1) / kSlotsPerSector;    // load factor <= 0.5
    if (nsec < 256) nsec = 256;
    const u64 cap = nsec * kSlotsPerSector;
    c.slots.alloc(cap, st);
    SG_CUDA(cudaMemsetAsync(c.slots.p, 0, cap * sizeof(u64), st));
    DevBuf<unsigned long long> d_over(1, st);
    SG_CUDA(cudaMemsetAsync(d_over.p, 0, sizeof(unsigned long long), st));
    index_insert_kernel<<<big_grid(D), 256, 0, st>>>(cols.a[cur], cols.b[cur], cols.v[cur], gstart.p, D, n, c.slots.p, nsec, d_over.p);
    SG_LAUNCHED();
    unsigned long long over = 0;
    SG_CUDA(cudaMemcpyAsync(&over, d_over.p, sizeof(over), cudaMemcpyDeviceToHost, st));

    // keep the sorted entry column
    c.entries.alloc(n, st);
    SG_CUDA(cudaMemcpyAsync(c.entries.p, cols.v[cur], n * sizeof(u32), cudaMemcpyDeviceToDevice, st));
    SG_CUDA(cudaStreamSynchronize(st));
    c.cap = cap;
    c.cnt.table_capacity = cap;
    c.cnt.distinct_keys = D;
    c.cnt.keys_over_threshold = over;
    c.have_table = true;
}

}  // namespace sg
