// table.cu -- step 2 on device: the prefix/suffix table (K3).  Restates
// HashTable::hashPrefixesAndSuffix / hashTableInsert (economyGraph/hashTable.cpp:70-188).
//
// Observable semantics kept exactly: key (h = min(k,64) bases) -> all (readId,type) sharing it, in
// (readId asc, type asc) order; a key with >= 100 entries is invisible to searches.  Layout:
//   slots[cap]   : open-addressing index, one 64-bit slot per DISTINCT key, grouped in 32-byte sectors
//                  of 4 (a probe reads one sector), filled with 64-bit atomicCAS:
//                  tag | min(count,127) | payload (core.cuh).  payload = the key's only entry
//                  (count == 1), a representative entry (count >= 100, masked) or the offset of the
//                  key's run in entries[] (2 <= count <= 99).
//   entries[M]   : (readId0<<2 | type) of every key with 2..99 entries, each run sorted ascending, which
//                  is the reference's bucket order (insertion order, hashTable.cpp:94-109).
// Like the reference (hashTable.cpp:150-160) a slot does not store its 128-bit key: a tag match is
// confirmed by re-extracting the key from the read of the slot's representative entry.
//
// Four passes over the 4U entries, no sort:
//   1 insert  : find-or-claim the key's slot (atomicCAS on an empty slot, saturating CAS increment of
//               the count on a match), remember the slot per entry
//   2 offsets : exclusive scan of the 2..99 counts over the slot array, rewrite those slots' payload
//   3 fill    : every entry of such a key takes the next position of its run (atomic cursor)
//   4 order   : insertion sort of every run (< 100 entries)
#include "context.h"

namespace sg {

constexpr u64 kCountOne = 1ull << 33;
constexpr u32 kNotOwned = 0xFFFFFFFFu;      // where[] of an entry whose key belongs to another shard

__device__ __forceinline__ u64 load_slot(const u64 *p) { return *reinterpret_cast<const volatile u64 *>(p); }

__global__ void __launch_bounds__(256) table_insert_kernel(const u64 *__restrict__ F, const u64 *__restrict__ RC,
                                                            const uint16_t *__restrict__ len, u64 U, int SW, int SWS, int h,
                                                            u64 *__restrict__ slots, u64 nsec, u32 *__restrict__ where,
                                                            int rank, int world, u32 *__restrict__ overflow)
{
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < 4 * U; t += (u64)gridDim.x * blockDim.x) {
        const u64 rid = t >> 2;
        const int type = (int)(t & 3);
        u64 v0, v1;
        entry_key(F + rid * SWS, RC + rid * SWS, SW, len[rid], h, type, v0, v1);
        const u64 hsh = hash_key(v0, v1);
        if (key_owner(hsh, world) != rank) { where[t] = kNotOwned; continue; }      // another shard's key
        const u64 tag = slot_tag(hsh);
        u64 sec = home_sector(hsh, nsec);
        bool done = false;
        u64 tries = 0;
        while (!done) {
            for (int q = 0; q < kSlotsPerSector && !done; ++q) {
                u64 *sp = slots + kSlotsPerSector * sec + q;
                u64 cur = load_slot(sp);
                if (cur == 0) {
                    const u64 old = atomicCAS((unsigned long long *)sp, 0ull, (unsigned long long)slot_encode(hsh, 1, t));
                    if (old == 0) { where[t] = (u32)(kSlotsPerSector * sec + q); done = true; break; }
                    cur = old;
                }
                if (slot_get_tag(cur) != tag) continue;
                // same tag: same key?  (the representative entry never changes once the slot is claimed)
                const u32 rep = (u32)slot_get_payload(cur);
                const u64 r2 = rep >> 2;
                u64 w0, w1;
                entry_key(F + r2 * SWS, RC + r2 * SWS, SW, len[r2], h, (int)(rep & 3), w0, w1);
                if (w0 != v0 || w1 != v1) continue;
                while (slot_get_count(cur) < 127) {            // saturating: >= 100 is all a search needs to know
                    const u64 old = atomicCAS((unsigned long long *)sp, (unsigned long long)cur, (unsigned long long)(cur + kCountOne));
                    if (old == cur) break;
                    cur = old;
                }
                where[t] = (u32)(kSlotsPerSector * sec + q);
                done = true;
            }
            sec = (sec + 1 == nsec) ? 0 : sec + 1;
            if (++tries > nsec) { *overflow = 1; where[t] = kNotOwned; break; }      // index full: reported, never silent
        }
    }
}

// pass 2a: run length of every slot that owns a run in entries[]; distinct / masked key counters
__global__ void __launch_bounds__(256) table_runs_kernel(const u64 *__restrict__ slots, u64 cap, u32 *__restrict__ run, unsigned long long *counters)
{
    unsigned long long distinct = 0, over = 0;
    for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += (u64)gridDim.x * blockDim.x) {
        const u64 v = slots[s];
        const u32 c = v ? slot_get_count(v) : 0u;
        run[s] = (c >= 2 && c < (u32)kHashThreshold) ? c : 0u;
        distinct += v != 0;
        over += c >= (u32)kHashThreshold;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { distinct += __shfl_xor_sync(0xffffffffu, distinct, o); over += __shfl_xor_sync(0xffffffffu, over, o); }
    if ((threadIdx.x & 31) == 0) { if (distinct) atomicAdd(&counters[0], distinct); if (over) atomicAdd(&counters[1], over); }
}

// pass 2b: payload of run-owning slots := offset of the run
__global__ void __launch_bounds__(256) table_offsets_kernel(u64 *__restrict__ slots, u64 cap, const u32 *__restrict__ run, const u32 *__restrict__ off)
{
    for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += (u64)gridDim.x * blockDim.x)
        if (run[s]) slots[s] = (slots[s] & ~0x1FFFFFFFFull) | (u64)off[s];
}

// pass 3: entries of run-owning keys take the next free position of their run
__global__ void __launch_bounds__(256) table_fill_kernel(const u64 *__restrict__ slots, const u32 *__restrict__ where, u64 n,
                                                          u32 *__restrict__ cursor, u32 *__restrict__ entries)
{
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (u64)gridDim.x * blockDim.x) {
        if (where[t] == kNotOwned) continue;
        const u64 v = slots[where[t]];
        const u32 c = slot_get_count(v);
        if (c < 2 || c >= (u32)kHashThreshold) continue;
        const u64 off = slot_get_payload(v);
        entries[off + atomicAdd(&cursor[off], 1u)] = (u32)t;
    }
}

// pass 4: bucket order = (readId asc, type asc) = entry value ascending
__global__ void __launch_bounds__(256) table_order_kernel(const u64 *__restrict__ slots, u64 cap, u32 *__restrict__ entries)
{
    for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += (u64)gridDim.x * blockDim.x) {
        const u64 v = slots[s];
        if (v == 0) continue;
        const u32 c = slot_get_count(v);
        if (c < 2 || c >= (u32)kHashThreshold) continue;
        u32 *e = entries + slot_get_payload(v);
        for (u32 a = 1; a < c; ++a) {
            const u32 x = e[a];
            u32 b = a;
            while (b > 0 && e[b - 1] > x) { e[b] = e[b - 1]; --b; }
            e[b] = x;
        }
    }
}

static unsigned big_grid(u64 n, unsigned block = 256)
{
    unsigned g = grid_for(n, block, 4);
    return g > kSMs * 16u ? kSMs * 16u : g;
}

// rank / world: the key-hash shard this context holds (0 / 1 = the whole table).  A shard indexes only the keys
// with key_owner(hash) == rank; every context still streams all 4U entries (the reads are replicated).
void stage_build_table(Context &c, int rank, int world, bool joint)
{
    cudaStream_t st = c.stream;
    ArenaScope arena_scope(c.arena, st);
    SG_CHECK(c.have_reads, "organize_reads must run before build_hash_table");
    SG_CHECK(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "bad rank / world");
    c.tb_rank = rank; c.tb_world = world; c.tb_shards = 1; c.tb_entries = 0; c.tb_joint = joint && world > 1;
    const u64 U = c.cnt.unique_reads;
    const int SW = c.SW, h = c.h;
    c.cnt.distinct_keys = 0; c.cnt.keys_over_threshold = 0; c.cap = 0; c.cnt.table_capacity = 0;
    if (U == 0) { c.slots.release(); c.entries.release(); c.have_table = true; return; }
    const u64 n = 4 * U;
    SG_CHECK(n < 0xFFFFFFFFull, "at most 2^30-1 unique reads per context");

    // load factor <= 2/3 even if all 4U keys are distinct (typically ~0.5)
    // (a shard expects 1/world of the keys; 10 % head room for the imbalance of the split)
    u64 nsec = (n + n / 2 + kSlotsPerSector - 1) / kSlotsPerSector;
    if (world > 1) nsec = nsec / (u64)world + nsec / (10 * (u64)world) + 1;
    if (nsec < 256) nsec = 256;
    const u64 cap = nsec * kSlotsPerSector;
    SG_CHECK(cap < 0xFFFFFFFFull, "slot index too large for one context");
    // joint: the shard is built in its place inside an array with room for all shards (stage_table_gather_* completes it)
    c.slots.alloc(c.tb_joint ? cap * (u64)world : cap, st);
    u64 *const slots = c.slots.p + (c.tb_joint ? (u64)rank * cap : 0);
    DevBuf<u32> &entries = c.tb_joint ? c.entries_loc : c.entries;
    SG_CUDA(cudaMemsetAsync(slots, 0, cap * sizeof(u64), st));

    DevBuf<u32> where(n, st), d_overflow(1, st);
    SG_CUDA(cudaMemsetAsync(d_overflow.p, 0, sizeof(u32), st));
    table_insert_kernel<<<big_grid(n), 256, 0, st>>>(c.F.p, c.RC.p, c.len.p, U, SW, c.SWS, h, slots, nsec, where.p, rank, world, d_overflow.p);
    SG_LAUNCHED();

    DevBuf<u32> run(cap, st), off(cap, st), d_total(1, st);
    DevBuf<unsigned long long> d_cnt(2, st);
    SG_CUDA(cudaMemsetAsync(d_cnt.p, 0, 2 * sizeof(unsigned long long), st));
    table_runs_kernel<<<big_grid(cap), 256, 0, st>>>(slots, cap, run.p, d_cnt.p);
    SG_LAUNCHED();
    exclusive_scan_u32(run.p, off.p, cap, d_total.p, st);
    u32 M = 0;
    unsigned long long h_cnt[2];
    u32 h_overflow = 0;
    SG_CUDA(cudaMemcpyAsync(&h_overflow, d_overflow.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(&M, d_total.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(h_cnt, d_cnt.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    SG_CHECK(h_overflow == 0, "slot index of this table shard is full (key split too uneven)");

    entries.alloc(M, st);
    c.tb_entries = M;
    if (M) {
        table_offsets_kernel<<<big_grid(cap), 256, 0, st>>>(slots, cap, run.p, off.p);
        SG_LAUNCHED();
        DevBuf<u32> cursor(M, st);
        SG_CUDA(cudaMemsetAsync(cursor.p, 0, (size_t)M * sizeof(u32), st));
        table_fill_kernel<<<big_grid(n), 256, 0, st>>>(slots, where.p, n, cursor.p, entries.p);
        SG_LAUNCHED();
        table_order_kernel<<<big_grid(cap), 256, 0, st>>>(slots, cap, entries.p);
        SG_LAUNCHED();
        SG_CUDA(cudaStreamSynchronize(st));     // cursor / where lifetimes end here
    }
    c.cap = cap;
    c.cnt.table_capacity = cap;
    c.cnt.distinct_keys = h_cnt[0];
    c.cnt.keys_over_threshold = h_cnt[1];
    c.have_table = true;
}

// ---- several GPUs, replicated table: every rank builds one key-hash shard, the shards are all-gathered ---------------
// layout: room for all shards back to back (equal slot counts; entry runs of shard q behind those of shards < q), this
// rank's shard moved to its place; the host all-gathers both arrays; finish: the run offsets stored in the slots of
// shard q become offsets into the joint entries[] array.
__global__ void __launch_bounds__(256) table_rebase_kernel(u64 *__restrict__ slots, u64 cap_shard, u64 ebase)
{
    for (u64 s = (u64)blockIdx.x * blockDim.x + threadIdx.x; s < cap_shard; s += (u64)gridDim.x * blockDim.x) {
        const u64 v = slots[s];
        if (v == 0) continue;
        const u32 c = slot_get_count(v);
        if (c >= 2 && c < (u32)kHashThreshold) slots[s] = (v & ~0x1FFFFFFFFull) | (slot_get_payload(v) + ebase);
    }
}

void stage_table_gather_layout(Context &c, const u64 *entry_counts, void **slots, void **entries, u64 *slots_per_shard, u64 *entries_first)
{
    cudaStream_t st = c.stream;
    SG_CHECK(c.have_table && c.tb_world > 1 && c.tb_shards == 1, "sage2gpu_build_hash_table_shard must run first");
    SG_CHECK(entry_counts && entry_counts[c.tb_rank] == c.tb_entries, "this rank's entry count does not match its shard");
    const int world = c.tb_world;
    const u64 cap_shard = c.cap;
    u64 tot = 0, base = 0;
    for (int q = 0; q < world; ++q) { if (q < c.tb_rank) base += entry_counts[q]; tot += entry_counts[q]; }
    SG_CHECK(tot < 0xFFFFFFFFull && cap_shard * (u64)world < (1ull << 40), "table too large");
    SG_CHECK(c.tb_joint, "the shard must be built with sage2gpu_build_hash_table_part for the gather");
    c.entries.alloc(tot, st);
    if (c.tb_entries) SG_CUDA(cudaMemcpyAsync(c.entries.p + base, c.entries_loc.p, (size_t)c.tb_entries * sizeof(u32), cudaMemcpyDeviceToDevice, st));
    SG_CUDA(cudaStreamSynchronize(st));
    if (slots) *slots = c.slots.p;
    if (entries) *entries = c.entries.p;
    if (slots_per_shard) *slots_per_shard = cap_shard;
    if (entries_first) *entries_first = base;
}

void stage_table_gather_finish(Context &c, const u64 *entry_counts, const u64 *distinct, const u64 *over)
{
    cudaStream_t st = c.stream;
    SG_CHECK(c.have_table && c.tb_world > 1 && c.tb_shards == 1 && entry_counts, "sage2gpu_table_gather_layout must run first");
    const int world = c.tb_world;
    const u64 cap_shard = c.cap;
    u64 ebase = 0, d = 0, o = 0;
    for (int q = 0; q < world; ++q) {
        if (ebase) { table_rebase_kernel<<<big_grid(cap_shard), 256, 0, st>>>(c.slots.p + (size_t)q * cap_shard, cap_shard, ebase); SG_LAUNCHED(); }
        ebase += entry_counts[q];
        if (distinct) d += distinct[q];
        if (over) o += over[q];
    }
    c.tb_entries = ebase;
    c.tb_shards = world; c.tb_world = 1; c.tb_rank = 0;
    c.cap = cap_shard * (u64)world;
    c.cnt.table_capacity = c.cap;
    if (distinct) c.cnt.distinct_keys = d;
    if (over) c.cnt.keys_over_threshold = o;
}

}  // namespace sg
