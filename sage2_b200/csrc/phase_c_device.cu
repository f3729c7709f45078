// phase_c_device.cu -- EconomyGraph::buildOverlapGraphEconomy (economyGraph/economyGraph.cpp:495-707) without the
// sequential walk, for the inputs where the walk's order cannot matter.
//
// What the walk does, read r by read r in breadth-first order: insertAllEdgesOfRead inserts r's overlaps with the reads
// that are still unexplored (both directions, :605-631) and sorts r's list (:634); markTransitiveEdge(r) runs once all
// of r's neighbours have their edges; removeTransitiveEdges(r) once all of r's neighbours are marked.  Hence
//   * nothing is appended to a list after its read was explored, and the list is sorted then: at marking time every
//     list involved is complete and in compareLengthBased order -- the marks are a function of the final lists only
//     (host_phase_c.cpp already computes them after the traversal, checked against the oracle);
//   * the final lists depend on the traversal order only through WHICH endpoint inserted an overlap.  If every
//     candidate (a -> b, type, overhang) has its twin (b -> a, reverse type, twin overhang) in b's own candidate list,
//     either endpoint inserts the same two entries and list[a] = a's own candidates + a's phase-B entries.
// So the lists are a function of the EXPLORATION ORDER alone: overlap {a, b} enters both lists with the candidate record
// of the end point explored first.  If the candidate set is symmetric the order is immaterial and nothing runs on the
// host; if it is not (a read contains another one -- variable read lengths -- or one of the two h-mers of an overlap is a
// masked key) the host does the traversal only (host_phase_c.cpp, run_host_phase_c_order) and hands over the order.  Either
// way every list is built, sorted, marked and filtered here, one warp per read.  A list longer than the per-warp
// capacity sends the whole phase to the host walk.
#include "context.h"

namespace sg {

constexpr int PC_MAXD = 512;        // entries per list handled on the device
constexpr int PC_HCAP = 1024;       // per-warp hash capacity (ids of one list)
constexpr int PC_WARPS = 4;

static unsigned pc_grid(u64 n, unsigned per_block)
{
    u64 g = (n + per_block - 1) / per_block;
    if (g > (u64)kSMs * 16) g = (u64)kSMs * 16;
    if (g == 0) g = 1;
    return (unsigned)g;
}

__device__ __forceinline__ u32 pc_rev(u32 t) { return t == 0 ? 3u : (t == 3 ? 0u : t); }
__device__ __forceinline__ bool pc_rule(u32 t1, u32 t2)      // economyGraph.cpp:661-664
{
    return ((t1 == 0 || t1 == 2) && (t2 == 0 || t2 == 1)) || ((t1 == 1 || t1 == 3) && (t2 == 2 || t2 == 3));
}

// phase-B half edges: every selected record (from,to,type,len) as seen from both endpoints
__global__ void __launch_bounds__(256) pc_half_edges_kernel(const u64 *__restrict__ selB, const u32 *__restrict__ selLen, u64 n,
                                                             u64 *__restrict__ owner, u64 *__restrict__ rec)
{
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x) {
        const u64 w0 = selB[2 * e], w1 = selB[2 * e + 1];
        const u32 a = (u32)(w0 >> 32), b = (u32)w0, type = (u32)(w1 >> 20) & 3u, len = (u32)(w1 & 0xFFFFFu);
        const u32 la = selLen[e] & 0xFFFFu, lb = selLen[e] >> 16;
        owner[2 * e] = a;     rec[2 * e] = ((u64)b << 32) | ((u64)type << 20) | len;
        owner[2 * e + 1] = b; rec[2 * e + 1] = ((u64)a << 32) | ((u64)pc_rev(type) << 20) | ((la - (lb - len)) & 0xFFFFFu);
    }
}

__device__ __forceinline__ void pc_range(const u64 *owner, u64 n, u64 id, u64 &lo, u64 &hi)
{
    u64 a = 0, b = n;
    while (a < b) { const u64 m = (a + b) >> 1; if (owner[m] < id) a = m + 1; else b = m; }
    lo = a; b = n;
    while (a < b) { const u64 m = (a + b) >> 1; if (owner[m] <= id) a = m + 1; else b = m; }
    hi = a;
}

// every candidate must have its twin in the other read's list
__global__ void __launch_bounds__(256) pc_symmetry_kernel(const u32 *__restrict__ s_ids, const u32 *__restrict__ sidx, u64 nS,
                                                           const u32 *__restrict__ counts, const u32 *__restrict__ offs, const u64 *__restrict__ cand,
                                                           const uint16_t *__restrict__ len, u32 *__restrict__ flags)
{
    const int lane = threadIdx.x & 31;
    const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
    bool bad = false;
    for (u64 s = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < nS; s += nwarps) {
        const u32 a = s_ids[s] + 1, la = len[a - 1];
        const u32 base = offs[s], m = counts[s];
        for (u32 x = lane; x < m; x += 32) {
            const u64 cw = cand[base + x];
            const u32 b = (u32)(cw >> 32), t = (u32)(cw >> 20) & 3u;
            u32 d = (u32)(cw & 0xFFFFFu);
            if (d & 0x80000u) d |= 0xFFF00000u;
            const u32 lb = len[b - 1];
            const u64 want = ((u64)a << 32) | ((u64)pc_rev(t) << 20) | ((la - (lb - d)) & 0xFFFFFu);
            const u32 sb = sidx[b - 1], bb = offs[sb], mb = counts[sb];
            bool found = false;
            for (u32 y = 0; y < mb && !found; ++y) found = cand[bb + y] == want;
            bad |= !found;
        }
    }
    if (bad) atomicOr(&flags[0], 1u);
}

struct PcHash {
    u32 *id;
    uint8_t *st;
    __device__ __forceinline__ static u32 h(u32 x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
    __device__ __forceinline__ int find(u32 key) const
    {
        for (u32 s = h(key) & (PC_HCAP - 1);; s = (s + 1) & (PC_HCAP - 1)) {
            const u32 k = id[s];
            if (k == key) return (int)s;
            if (k == 0) return -1;
        }
    }
};

// Half edges of phase C.  An overlap {a, b} of two S reads is inserted by whichever end point is explored first
// (insertAllEdgesOfRead skips partners that were explored already, :605), with THAT read's candidate record and the
// twin computed from it (insertEdgeEconomy, :813-849).  order[s] = position of S read s in the exploration sequence;
// without it (symmetric candidate sets: either end point inserts the same two entries) the S index serves.
__global__ void __launch_bounds__(256) pc_cand_half_edges_kernel(const u32 *__restrict__ s_ids, const u32 *__restrict__ sidx, u64 nS,
                                                                  const u32 *__restrict__ counts, const u32 *__restrict__ offs, const u64 *__restrict__ cand,
                                                                  const uint16_t *__restrict__ len, const u32 *__restrict__ order,
                                                                  u64 *__restrict__ owner, u64 *__restrict__ rec, unsigned long long *__restrict__ n_out)
{
    const int lane = threadIdx.x & 31;
    const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
    for (u64 s = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < nS; s += nwarps) {
        const u32 a = s_ids[s] + 1, la = len[a - 1];
        const u32 base = offs[s], m = counts[s];
        const u32 oa = order ? order[s] : (u32)s;
        for (u32 x0 = 0; x0 < m; x0 += 32) {
            const u32 x = x0 + lane;
            bool keep = false;
            u64 cw = 0;
            u32 b = 0;
            if (x < m) {
                cw = cand[base + x];
                b = (u32)(cw >> 32);
                const u32 sb = sidx[b - 1];
                keep = oa < (order ? order[sb] : sb);
            }
            const unsigned km = __ballot_sync(0xffffffffu, keep);
            if (km == 0) continue;
            unsigned long long pos = 0;
            if (lane == 0) pos = atomicAdd(n_out, 2ull * (unsigned long long)__popc(km));
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (keep) {
                const u64 at = pos + 2ull * (u64)__popc(km & ((1u << lane) - 1u));
                const u32 t = (u32)(cw >> 20) & 3u;
                u32 d = (u32)(cw & 0xFFFFFu);
                if (d & 0x80000u) d |= 0xFFF00000u;
                const u32 lb = len[b - 1];
                owner[at] = a;     rec[at] = cw;
                owner[at + 1] = b; rec[at + 1] = ((u64)a << 32) | ((u64)pc_rev(t) << 20) | ((la - (lb - d)) & 0xFFFFFu);
            }
        }
    }
}

// where the list of every half edge's far end sits in the owner-sorted array (the marking loop of pc_node_kernel walks
// those lists one after the other: the two binary searches per step were its critical path)
__global__ void __launch_bounds__(256) pc_target_range_kernel(const u64 *__restrict__ he_owner, const u64 *__restrict__ he_rec, u64 n_he,
                                                               u64 *__restrict__ tlo, u32 *__restrict__ tcnt)
{
    for (u64 x = (u64)blockIdx.x * blockDim.x + threadIdx.x; x < n_he; x += (u64)gridDim.x * blockDim.x) {
        u64 lo, hi;
        pc_range(he_owner, n_he, he_rec[x] >> 32, lo, hi);
        tlo[x] = lo;
        tcnt[x] = (u32)(hi - lo);
    }
}

// one warp per read of S: list = its half edges (phase C + phase B) out of the owner-sorted array, sorted by
// compareLengthBased (:853-871), marked like markTransitiveEdge, filtered like removeTransitiveEdges; survivors with
// id > read go to `out`.  Sort key: length desc, id desc, type desc == one descending 54-bit number; the low 9 bits
// carry the entry's place in the unsorted list (PC_MAXD = 512), which is where its far end's range was precomputed.
__global__ void __launch_bounds__(PC_WARPS * 32) pc_node_kernel(const u32 *__restrict__ s_ids, u64 nS,
                                                                 const u64 *__restrict__ he_owner, const u64 *__restrict__ he_rec, u64 n_he,
                                                                 const u64 *__restrict__ tlo, const u32 *__restrict__ tcnt,
                                                                 u64 *__restrict__ out, unsigned long long *__restrict__ counters /*[0] out, [1] removed*/,
                                                                 u32 *__restrict__ flags)
{
    __shared__ u64 sKey[PC_WARPS][PC_MAXD];
    __shared__ u32 sHid[PC_WARPS][PC_HCAP];
    __shared__ uint8_t sHst[PC_WARPS][PC_HCAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 *key = sKey[warp];
    PcHash H{ sHid[warp], sHst[warp] };
    const u64 nwarps = (u64)gridDim.x * PC_WARPS;
    unsigned long long removed = 0;
    for (u64 s = (u64)blockIdx.x * PC_WARPS + warp; s < nS; s += nwarps) {
        const u32 n = s_ids[s] + 1;
        u64 plo, phi;
        pc_range(he_owner, n_he, n, plo, phi);
        const u32 d = (u32)(phi - plo);
        if (phi - plo > (u64)PC_MAXD) { if (lane == 0) atomicOr(&flags[0], 2u); continue; }
        if (d == 0) continue;
        u32 P = 1;
        while (P < d) P <<= 1;
        __syncwarp();
        for (u32 x = lane; x < P; x += 32) {
            const u64 cw = x < d ? he_rec[plo + x] : 0ull;
            key[x] = x < d ? (((cw & 0xFFFFFull) << 43) | ((cw >> 32) << 11) | (((cw >> 20) & 3ull) << 9) | (u64)x) : 0ull;
        }
        for (u32 x = lane; x < (u32)PC_HCAP; x += 32) { H.id[x] = 0; H.st[x] = 0; }
        __syncwarp();
        for (u32 k = 2; k <= P; k <<= 1)
            for (u32 j = k >> 1; j > 0; j >>= 1) {
                for (u32 i = lane; i < P; i += 32) {
                    const u32 p = i ^ j;
                    if (p > i) {
                        const u64 a = key[i], b = key[p];
                        const bool desc = (i & k) == 0;
                        if (desc ? a < b : a > b) { key[i] = b; key[p] = a; }
                    }
                }
                __syncwarp();
            }
        for (u32 x = lane; x < d; x += 32) {
            const u32 id = (u32)(key[x] >> 11);
            for (u32 sl = PcHash::h(id) & (PC_HCAP - 1);; sl = (sl + 1) & (PC_HCAP - 1)) {
                const u32 old = atomicCAS(&H.id[sl], 0u, id);
                if (old == 0u || old == id) { H.st[sl] = 1; break; }
            }
        }
        __syncwarp();
        for (u32 x0 = 0; x0 < d; x0 += 32) {
            // 32 entries at a time: their far ends' ranges come in with one load per lane
            const u64 my_key = x0 + lane < d ? key[x0 + lane] : 0ull;
            u64 my_lo = 0;
            u32 my_cnt = 0;
            if (x0 + lane < d) { const u64 at = plo + (my_key & 511ull); my_lo = tlo[at]; my_cnt = tcnt[at]; }
            const u32 m = d - x0 < 32u ? d - x0 : 32u;
            for (u32 xx = 0; xx < m; ++xx) {
                const u64 kx = __shfl_sync(0xffffffffu, my_key, xx);
                const u64 qlo = __shfl_sync(0xffffffffu, my_lo, xx);
                const u32 qn = __shfl_sync(0xffffffffu, my_cnt, xx);
                const u32 ida = (u32)(kx >> 11), t1 = (u32)(kx >> 9) & 3u;
                const int sa = H.find(ida);
                if (H.st[sa] == 1) {
                    for (u32 y = lane; y < qn; y += 32) {
                        const u64 cw = he_rec[qlo + y];
                        const int sf = H.find((u32)(cw >> 32));
                        if (sf >= 0 && H.st[sf] == 1 && pc_rule(t1, (u32)(cw >> 20) & 3u)) H.st[sf] = 2;
                    }
                }
                __syncwarp();
            }
        }
        for (u32 x = lane; x < d; x += 32) {
            const u64 kx = key[x];
            const u32 id = (u32)(kx >> 11);
            if (H.st[H.find(id)] == 2) { removed++; continue; }
            if (id > n) {
                const unsigned long long pos = atomicAdd(&counters[0], 1ull);
                out[2 * pos] = ((u64)n << 32) | id;
                out[2 * pos + 1] = (((kx >> 9) & 3ull) << 20) | (kx >> 43);
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
    if (lane == 0 && removed) atomicAdd(&counters[1], removed);
}

// ---- the lists of the host traversal ---------------------------------------------------------------------------------
// For asymmetric candidate sets the traversal (economyGraph.cpp:513-564) stays on the host, but everything it would sort or
// look up is prepared here: for every S read the entries its list can ever hold -- its own candidates and the twins other
// reads would push into it (insertEdgeEconomy, :813-849) -- in compareLengthBased order (:853-871).  One record per
// potential entry: B = the list's read (S index), A = 55-bit sort key, complemented so that ascending = compareLengthBased:
// (overhang20 << 34 | other read's S index << 2 | type) << 1 | twin.  S indices ascend with the ids, so they order like ids.
constexpr u64 PC_KEYMASK = (1ull << 55) - 1ull;

__global__ void __launch_bounds__(256) pc_list_records_kernel(const u32 *__restrict__ s_ids, const u32 *__restrict__ sidx, u64 nS,
                                                               const u32 *__restrict__ counts, const u32 *__restrict__ offs, const u64 *__restrict__ cand,
                                                               const uint16_t *__restrict__ len, u64 *__restrict__ A, u64 *__restrict__ B)
{
    const int lane = threadIdx.x & 31;
    const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
    for (u64 s = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < nS; s += nwarps) {
        const u32 a = s_ids[s] + 1, la = len[a - 1];
        const u32 base = offs[s], m = counts[s];
        for (u32 x = lane; x < m; x += 32) {
            const u64 q = (u64)base + x, cw = cand[q];
            const u32 b = (u32)(cw >> 32), t = (u32)(cw >> 20) & 3u, d20 = (u32)(cw & 0xFFFFFu);
            const u32 d = (d20 & 0x80000u) ? (d20 | 0xFFF00000u) : d20;
            const u32 sb = sidx[b - 1], lb = len[b - 1];
            const u32 d2 = (la - (lb - d)) & 0xFFFFFu;
            const u64 k_own = ((u64)d20 << 34) | ((u64)sb << 2) | t;
            const u64 k_twin = ((u64)d2 << 34) | ((u64)s << 2) | pc_rev(t);
            A[2 * q] = PC_KEYMASK ^ (k_own << 1);             B[2 * q] = s;
            A[2 * q + 1] = PC_KEYMASK ^ ((k_twin << 1) | 1);  B[2 * q + 1] = sb == (u32)s ? nS : sb;     // a read never inserts an overlap with itself (:605)
        }
    }
}

__global__ void __launch_bounds__(256) pc_list_entries_kernel(const u64 *__restrict__ A, u64 n, u32 *__restrict__ ent)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 k = A[i] ^ PC_KEYMASK;
        ent[i] = ((u32)(k >> 3) << 1) | (u32)(k & 1ull);
    }
}

__global__ void __launch_bounds__(256) pc_list_offsets_kernel(const u64 *__restrict__ B, u64 n, u64 nS, u32 *__restrict__ off)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i <= nS; i += (u64)gridDim.x * blockDim.x) {
        u64 lo = 0, hi = n;
        while (lo < hi) { const u64 m = (lo + hi) >> 1; if (B[m] < i) lo = m + 1; else hi = m; }
        off[i] = (u32)lo;
    }
}

// off[nS+1], ent[2 nC] on the device; false: too many entries for 32-bit offsets (the caller takes the host walk)
bool phase_c_sorted_lists(Context &c, const u32 *s_ids, const u32 *sidx, u64 nS, const u32 *counts, const u32 *offs, const u64 *cand, u64 nC,
                          DevBuf<u32> &off, DevBuf<u32> &ent)
{
    cudaStream_t st = c.stream;
    const u64 n = 2 * nC;
    if (n >= 0xFFFFFFFFull || nS >= 0x7FFFFFFFull) return false;
    off.alloc(nS + 1, st);
    ent.alloc(n + 1, st);
    DevBuf<u64> a0(n + 1, st), a1(n + 1, st), b0(n + 1, st), b1(n + 1, st);
    const u64 *B = b0.p;
    if (n) {
        pc_list_records_kernel<<<pc_grid(nS, 8), 256, 0, st>>>(s_ids, sidx, nS, counts, offs, cand, c.len.p, a0.p, b0.p);
        SG_LAUNCHED();
        SortCols cols;
        cols.a[0] = a0.p; cols.a[1] = a1.p; cols.b[0] = b0.p; cols.b[1] = b1.p; cols.v[0] = cols.v[1] = nullptr;
        int cur = radix_sort_varying(cols, 0, n, false, st);        // the key, then (stable) the list's read
        int nbits = 1;
        while ((nS >> nbits) != 0) ++nbits;
        cur = radix_sort_bits(cols, cur, n, true, 0, nbits, st);
        pc_list_entries_kernel<<<pc_grid(n, 256), 256, 0, st>>>(cols.a[cur], n, ent.p);
        SG_LAUNCHED();
        B = cols.b[cur];
    }
    pc_list_offsets_kernel<<<pc_grid(nS + 1, 256), 256, 0, st>>>(B, n, nS, off.p);
    SG_LAUNCHED();
    return true;
}

// every candidate has its twin in the other read's list?  (then the traversal order cannot matter)
bool phase_c_is_symmetric(Context &c, const u32 *s_ids, const u32 *sidx, u64 nS, const u32 *counts, const u32 *offs, const u64 *cand)
{
    cudaStream_t st = c.stream;
    DevBuf<u32> flags(1, st);
    SG_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(u32), st));
    pc_symmetry_kernel<<<pc_grid(nS, 8), 256, 0, st>>>(s_ids, sidx, nS, counts, offs, cand, c.len.p, flags.p);
    SG_LAUNCHED();
    u32 h_flags = 0;
    SG_CUDA(cudaMemcpyAsync(&h_flags, flags.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    return h_flags == 0;
}

// Lists, marks and filtering of phase C on the device.  d_order: exploration order of the S reads (device array, from the
// host traversal) or null for symmetric candidate sets.  true: `out` holds n_out (w0,w1) records owned by the S reads;
// false: a list exceeds the per-warp capacity, use the host walk.
bool device_phase_c(Context &c, const u32 *s_ids, const u32 *sidx, u64 nS, const u32 *counts, const u32 *offs, const u64 *cand, u64 nC,
                    const u64 *selB, const u32 *selLen, u64 nSel, const u32 *d_order, DevBuf<u64> &out, u64 &n_out, u64 &inserted, u64 &removed)
{
    cudaStream_t st = c.stream;
    const u64 U = c.cnt.unique_reads;
    // half edges: phase B (both end points of every selected record) + phase C (inserting end point + twin), sorted by owner
    const u64 n_pb = 2 * nSel, cap = n_pb + 2 * nC;
    DevBuf<u64> o0(cap + 1, st), o1(cap + 1, st), r0(cap + 1, st), r1(cap + 1, st);
    DevBuf<unsigned long long> cnt(3, st);
    SG_CUDA(cudaMemsetAsync(cnt.p, 0, 3 * sizeof(unsigned long long), st));
    if (n_pb) {
        pc_half_edges_kernel<<<pc_grid(nSel, 256), 256, 0, st>>>(selB, selLen, nSel, o0.p, r0.p);
        SG_LAUNCHED();
    }
    if (nC) {
        pc_cand_half_edges_kernel<<<pc_grid(nS, 8), 256, 0, st>>>(s_ids, sidx, nS, counts, offs, cand, c.len.p, d_order, o0.p + n_pb, r0.p + n_pb, cnt.p + 2);
        SG_LAUNCHED();
    }
    unsigned long long n_c = 0;
    SG_CUDA(cudaMemcpyAsync(&n_c, cnt.p + 2, sizeof(n_c), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    const u64 n_he = n_pb + n_c;
    const u64 *he_owner = o0.p, *he_rec = r0.p;
    if (n_he) {
        SortCols cols;
        cols.a[0] = o0.p; cols.a[1] = o1.p; cols.b[0] = r0.p; cols.b[1] = r1.p; cols.v[0] = cols.v[1] = nullptr;
        int id_bits = 1;
        while ((U >> id_bits) != 0) ++id_bits;
        const int cur = radix_sort_bits(cols, 0, n_he, false, 0, id_bits, st);
        he_owner = cols.a[cur]; he_rec = cols.b[cur];
    }
    out.alloc(2 * n_he + 2, st);
    DevBuf<u32> flags(1, st);
    SG_CUDA(cudaMemsetAsync(flags.p, 0, sizeof(u32), st));
    DevBuf<u64> tlo(n_he + 1, st);
    DevBuf<u32> tcnt(n_he + 1, st);
    if (n_he) {
        pc_target_range_kernel<<<pc_grid(n_he, 256), 256, 0, st>>>(he_owner, he_rec, n_he, tlo.p, tcnt.p);
        SG_LAUNCHED();
    }
    pc_node_kernel<<<pc_grid(nS, PC_WARPS), PC_WARPS * 32, 0, st>>>(s_ids, nS, he_owner, he_rec, n_he, tlo.p, tcnt.p, out.p, cnt.p, flags.p);
    SG_LAUNCHED();
    unsigned long long h_cnt[2];
    u32 h_flags = 0;
    SG_CUDA(cudaMemcpyAsync(h_cnt, cnt.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(&h_flags, flags.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    if (h_flags) return false;           // a list longer than PC_MAXD
    n_out = h_cnt[0];
    removed = h_cnt[1];
    inserted = n_c;
    return true;
}

}  // namespace sg
