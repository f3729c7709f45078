// graph.cu -- step 3 after the search: phase B on device, phase-C candidate generation, the serial
// phase-C walk on the host (host_phase_c.cpp), and the canonical edge list (K6).  Restates
// EconomyGraph::buildInitialOverlapGraph phase B (economyGraph/economyGraph.cpp:455-480),
// sortEconomyGraph (:896-913) and the part of OverlapGraph::convertGraph that decides which entries
// become edges (overlapGraph/overlapGraph.cpp:93-112).
#include <stdlib.h>
#include <algorithm>
#include <stdio.h>
#include "context.h"
#include "host_phase_c.h"

namespace sg {

void launch_phase_c_candidates(Context &c, const u32 *s_ids, u64 nS, u32 *counts, const u32 *offsets, u64 *cand, bool fill);
// phase_c_device.cu
bool phase_c_is_symmetric(Context &c, const u32 *s_ids, const u32 *sidx, u64 nS, const u32 *counts, const u32 *offs, const u64 *cand);
bool phase_c_sorted_lists(Context &c, const u32 *s_ids, const u32 *sidx, u64 nS, const u32 *counts, const u32 *offs, const u64 *cand, u64 nC,
                          DevBuf<u32> &off, DevBuf<u32> &ent);
bool device_phase_c(Context &c, const u32 *s_ids, const u32 *sidx, u64 nS, const u32 *counts, const u32 *offs, const u64 *cand, u64 nC,
                    const u64 *selB, const u32 *selLen, u64 nSel, const u32 *d_order, DevBuf<u64> &out, u64 &n_out, u64 &inserted, u64 &removed);

static unsigned big_grid(u64 n, unsigned block = 256)
{
    unsigned g = grid_for(n, block, 4);
    return g > kSMs * 16u ? kSMs * 16u : g;
}

// B1: which reads have reciprocal unique extensions on both sides (economyGraph.cpp:460)
__global__ void __launch_bounds__(256) phase_b_qualify_kernel(const u64 *__restrict__ extR, const u64 *__restrict__ extL,
                                                               const uint8_t *__restrict__ flag5, const u32 *__restrict__ cont_max,
                                                               u64 U, uint8_t *__restrict__ explored, unsigned long long *counters)
{
    unsigned long long nq = 0, n6 = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < U; i += (u64)gridDim.x * blockDim.x) {
        const u32 id = (u32)i + 1;
        uint8_t s = state_after_a(id, cont_max[i], flag5[i]);
        if (s != 6) {
            const bool q = phase_b_qualifies(extR, extL, i);
            if (q) { s = 4; nq++; }
        } else n6++;
        explored[i] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { nq += __shfl_xor_sync(0xffffffffu, nq, o); n6 += __shfl_xor_sync(0xffffffffu, n6, o); }
    if ((threadIdx.x & 31) == 0) { if (nq) atomicAdd(&counters[0], nq); if (n6) atomicAdd(&counters[1], n6); }
}

// B2: edges.  Read i (state 4) inserts i->L and i->R unless the target was already state 4 when i was
// visited, i.e. unless the target qualifies too and has a smaller id (economyGraph.cpp:462-473).
__global__ void __launch_bounds__(256) phase_b_edges_kernel(const u64 *__restrict__ extR, const u64 *__restrict__ extL,
                                                             const uint8_t *__restrict__ explored, const uint16_t *__restrict__ len,
                                                             u64 U, u64 *__restrict__ edges, unsigned long long *n_edges)
{
    for (u64 i0 = (u64)blockIdx.x * blockDim.x; i0 < U; i0 += (u64)gridDim.x * blockDim.x) {
        const u64 i = i0 + threadIdx.x;
        EdgeRec e[2];
        int n = 0;
        if (i < U) n = phase_b_edges(extR, extL, explored, len, i, e);
        // warp-aggregated append
        const int lane = threadIdx.x & 31;
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        unsigned long long base = 0;
        if (lane == 31 && tot) base = atomicAdd(n_edges, (unsigned long long)tot);
        base = __shfl_sync(0xffffffffu, base, 31);
        const u64 pos = base + (u64)(incl - n);
        for (int t = 0; t < n; ++t) { edges[2 * (pos + t)] = e[t].w0; edges[2 * (pos + t) + 1] = e[t].w1; }
    }
}

__global__ void __launch_bounds__(256) sum_u32_kernel(const u32 *__restrict__ v, u64 n, unsigned long long *__restrict__ out)
{
    unsigned long long acc = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) acc += v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

__global__ void __launch_bounds__(256) flag_state0_kernel(const uint8_t *__restrict__ explored, u64 U, u32 *__restrict__ flag)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < U; i += (u64)gridDim.x * blockDim.x) flag[i] = explored[i] == 0;
}

__global__ void __launch_bounds__(256) compact_ids_kernel(const u32 *__restrict__ flag, const u32 *__restrict__ idx, u64 U, u32 *__restrict__ out)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < U; i += (u64)gridDim.x * blockDim.x)
        if (flag[i]) out[idx[i]] = (u32)i;
}

// Phase-C hand-over: the host walk touches S (state 0 after phase B) and the phase-B neighbours of S,
// whose whole lists markTransitiveEdge scans (economyGraph.cpp:653).  mark_neighbours flags those
// neighbours, flag_needed_edges selects every phase-B record with an endpoint in S or flagged.
__global__ void __launch_bounds__(256) mark_neighbours_kernel(const u64 *__restrict__ edges, u64 n, const uint8_t *__restrict__ explored,
                                                               uint8_t *__restrict__ nbr)
{
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x) {
        const u64 w0 = edges[2 * e];
        const u32 a = (u32)(w0 >> 32) - 1, b = (u32)w0 - 1;
        if (explored[a] == 0) nbr[b] = 1;
        if (explored[b] == 0) nbr[a] = 1;
    }
}

__global__ void __launch_bounds__(256) flag_needed_edges_kernel(const u64 *__restrict__ edges, u64 n, const uint8_t *__restrict__ explored,
                                                                 const uint8_t *__restrict__ nbr, u32 *__restrict__ flag)
{
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x) {
        const u64 w0 = edges[2 * e];
        const u32 a = (u32)(w0 >> 32) - 1, b = (u32)w0 - 1;
        flag[e] = (explored[a] == 0 || explored[b] == 0 || nbr[a] || nbr[b]) ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(256) gather_needed_edges_kernel(const u64 *__restrict__ edges, u64 n, const u32 *__restrict__ flag,
                                                                   const u32 *__restrict__ idx, const uint16_t *__restrict__ len,
                                                                   u64 *__restrict__ out, u32 *__restrict__ out_len)
{
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x) {
        if (!flag[e]) continue;
        const u64 w0 = edges[2 * e], p = idx[e];
        out[2 * p] = w0; out[2 * p + 1] = edges[2 * e + 1];
        out_len[p] = (u32)len[(u32)(w0 >> 32) - 1] | ((u32)len[(u32)w0 - 1] << 16);
    }
}

__global__ void __launch_bounds__(256) gather_len_kernel(const u32 *__restrict__ ids, u64 n, const uint16_t *__restrict__ len, uint16_t *__restrict__ out)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) out[i] = len[ids[i]];
}

// which S reads own phase-B records (their lists are not empty whatever the traversal does, economyGraph.cpp:525)
__global__ void __launch_bounds__(256) has_b_kernel(const u64 *__restrict__ selB, u64 nSel, const uint8_t *__restrict__ explored, const u32 *__restrict__ sidx,
                                                     uint8_t *__restrict__ has_b)
{
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < nSel; e += (u64)gridDim.x * blockDim.x) {
        const u64 w0 = selB[2 * e];
        const u32 a = (u32)(w0 >> 32) - 1, b = (u32)w0 - 1;
        if (explored[a] == 0) has_b[sidx[a]] = 1;
        if (explored[b] == 0) has_b[sidx[b]] = 1;
    }
}

void stage_phase_b(Context &c)
{
    cudaStream_t st = c.stream;
    ArenaScope arena_scope(c.arena, st);
    const u64 U = c.cnt.unique_reads;
    c.explored.alloc(U, st);
    c.edges.alloc(4 * U + 2, st);
    c.cnt.contained_ext = c.cnt.contained_size = 0;
    c.cnt.left_to_explore = 0; c.cnt.edges_phase_b = 0;
    c.have_phase_b = true;
    if (U == 0) return;
    DevBuf<unsigned long long> d_cnt(3, st);
    SG_CUDA(cudaMemsetAsync(d_cnt.p, 0, 3 * sizeof(unsigned long long), st));
    phase_b_qualify_kernel<<<big_grid(U), 256, 0, st>>>(c.extR.p, c.extL.p, c.flag5.p, c.cont_max.p, U, c.explored.p, d_cnt.p);
    SG_LAUNCHED();
    phase_b_edges_kernel<<<big_grid(U), 256, 0, st>>>(c.extR.p, c.extL.p, c.explored.p, c.len.p, U, c.edges.p, d_cnt.p + 2);
    SG_LAUNCHED();
    unsigned long long h[3];
    SG_CUDA(cudaMemcpyAsync(h, d_cnt.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    c.cnt.contained_ext = h[0];
    c.cnt.contained_size = h[1];
    c.cnt.left_to_explore = U - h[0] - h[1];
    c.cnt.edges_phase_b = h[2];
}

// ------------------------------------------------------------------------------------------------
// finalize: drop phase-B records owned by state-0 reads (their lists are rebuilt by the host walk),
// append the host's records, canonical radix sort, (from,to,type) dedupe keeping the smallest
// overhang (overlapGraph.cpp:101).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flag_keep_edges_kernel(const u64 *__restrict__ edges, u64 n, const uint8_t *__restrict__ explored, u32 *__restrict__ flag)
{
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x) {
        const u32 a = (u32)(edges[2 * e] >> 32);
        flag[e] = explored[a - 1] != 0;
    }
}

__global__ void __launch_bounds__(256) split_edges_kernel(const u64 *__restrict__ edges, u64 n, const u32 *__restrict__ flag,
                                                           const u32 *__restrict__ idx, u64 *__restrict__ w0, u64 *__restrict__ w1)
{
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x)
        if (!flag || flag[e]) { const u64 p = flag ? idx[e] : e; w0[p] = edges[2 * e]; w1[p] = edges[2 * e + 1]; }
}

// runs of equal w0 = (from, to): sort the w1 = type << 20 | overhang words ascending (compareIdBased ties)
__global__ void __launch_bounds__(256) order_pair_runs_kernel(const u64 *__restrict__ w0, u64 *__restrict__ w1, u64 n)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u64 k = w0[i];
        if ((i > 0 && w0[i - 1] == k) || i + 1 >= n || w0[i + 1] != k) continue;      // not the head of a run of >= 2
        u64 e = i + 2;
        while (e < n && w0[e] == k) ++e;
        for (u64 a = i + 1; a < e; ++a) {
            const u64 x = w1[a];
            u64 b = a;
            while (b > i && w1[b - 1] > x) { w1[b] = w1[b - 1]; --b; }
            w1[b] = x;
        }
    }
}

__global__ void __launch_bounds__(256) flag_first_edge_kernel(const u64 *__restrict__ w0, const u64 *__restrict__ w1, u64 n, u32 *__restrict__ flag)
{
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x)
        flag[e] = (e == 0 || w0[e] != w0[e - 1] || (w1[e] >> 20) != (w1[e - 1] >> 20)) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) join_edges_kernel(const u64 *__restrict__ w0, const u64 *__restrict__ w1, u64 n,
                                                          const u32 *__restrict__ flag, const u32 *__restrict__ idx, u64 *__restrict__ edges)
{
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (u64)gridDim.x * blockDim.x)
        if (flag[e]) { edges[2 * (u64)idx[e]] = w0[e]; edges[2 * (u64)idx[e] + 1] = w1[e]; }
}

// ---- merge of two sorted record lists without common keys (kept phase-B records, phase-C records) -----------------------
__device__ __forceinline__ bool rec_less2(u64 a0, u64 a1, u64 b0, u64 b1) { return a0 < b0 || (a0 == b0 && a1 < b1); }
// records of X (nx) go to position i + |{y in Y : y < x_i}|
__global__ void __launch_bounds__(256) merge_scatter_kernel(const u64 *__restrict__ x0, const u64 *__restrict__ x1, u64 nx,
                                                             const u64 *__restrict__ y0, const u64 *__restrict__ y1, u64 ny,
                                                             u64 *__restrict__ o0, u64 *__restrict__ o1)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nx; i += (u64)gridDim.x * blockDim.x) {
        const u64 k0 = x0[i], k1 = x1[i];
        u64 lo = 0, hi = ny;
        while (lo < hi) { const u64 m = (lo + hi) >> 1; if (rec_less2(y0[m], y1[m], k0, k1)) lo = m + 1; else hi = m; }
        o0[i + lo] = k0; o1[i + lo] = k1;
    }
}

void stage_phase_c_and_finalize(Context &c)
{
    cudaStream_t st = c.stream;
    ArenaScope arena_scope(c.arena, st);
    const u64 U = c.cnt.unique_reads;
    c.cnt.candidates_c = c.cnt.edges_inserted_c = c.cnt.transitive_removed = 0;
    c.cnt.phase_c_on_device = 0;
    c.cnt.n_edges = 0;
    c.h_edges.clear();
    if (U == 0) { c.have_graph = true; c.rt_for_c = false; return; }

    cudaEvent_t ev0, ev1, ev2;
    SG_CUDA(cudaEventCreate(&ev0)); SG_CUDA(cudaEventCreate(&ev1)); SG_CUDA(cudaEventCreate(&ev2));
    SG_CUDA(cudaEventRecord(ev0, st));

    // ---- state-0 reads and their candidate lists -------------------------------------------------
    DevBuf<u32> flag(U, st), idx(U, st), d_total(1, st);
    flag_state0_kernel<<<big_grid(U), 256, 0, st>>>(c.explored.p, U, flag.p);
    SG_LAUNCHED();
    exclusive_scan_u32(flag.p, idx.p, U, d_total.p, st);
    u32 nS = 0;
    SG_CUDA(cudaMemcpyAsync(&nS, d_total.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));

    u64 nB = c.cnt.edges_phase_b;
    std::vector<u64> host_c_edges;     // records owned by state-0 reads after the walk
    DevBuf<u64> dev_c_edges;           // the same when phase C ran on the device
    u64 n_dev_c = 0;
    bool c_on_device = false;
    float host_ms = 0.f, host_order_ms = 0.f;
    // kept phase-B records, sorted ahead of the merge (sort_kept_b)
    DevBuf<u32> eflag, eidx, d_keep(1, st);
    DevBuf<u64> ka0, ka1, kb0, kb1;
    SortCols kcols;
    int kcur = 0;
    bool b_sorted = false;
    u64 nKeep = nB;
    int id_bits = 1;
    while ((U >> id_bits) != 0) ++id_bits;                   // ids are 1..U
    if (nS > 0) {
        DevBuf<u32> s_ids(nS, st), counts(nS, st), offs(nS, st), d_ctotal(1, st);
        compact_ids_kernel<<<big_grid(U), 256, 0, st>>>(flag.p, idx.p, U, s_ids.p);
        SG_LAUNCHED();
        launch_phase_c_candidates(c, s_ids.p, nS, counts.p, nullptr, nullptr, false);
        // the candidate total is not bounded by 4U (error-rich data): sum it in 64 bits before the 32-bit scan
        DevBuf<unsigned long long> d_c64(1, st);
        SG_CUDA(cudaMemsetAsync(d_c64.p, 0, sizeof(unsigned long long), st));
        sum_u32_kernel<<<big_grid(nS), 256, 0, st>>>(counts.p, nS, d_c64.p);
        SG_LAUNCHED();
        exclusive_scan_u32(counts.p, offs.p, nS, d_ctotal.p, st);
        u32 nC = 0;
        unsigned long long nC64 = 0;
        SG_CUDA(cudaMemcpyAsync(&nC, d_ctotal.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaMemcpyAsync(&nC64, d_c64.p, sizeof(nC64), cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaStreamSynchronize(st));
        SG_CHECK(nC64 < 0xFFFFFFFFull, "more than 2^32 phase-C candidates: not supported (split the input)");
        DevBuf<u64> cand(nC, st);
        if (nC) launch_phase_c_candidates(c, s_ids.p, nS, counts.p, offs.p, cand.p, true);
        c.cnt.candidates_c = nC;

        // ---- host walk (serial by definition, economyGraph.cpp:513-564) -------------------------
        // only the phase-B records the walk can touch leave the device
        DevBuf<uint8_t> nbr(U, st);
        DevBuf<u32> nflag, nidx, d_nsel(1, st), selLen;
        DevBuf<u64> selB;
        DevBuf<uint16_t> sLen(nS, st);
        u32 nSel = 0;
        gather_len_kernel<<<big_grid(nS), 256, 0, st>>>(s_ids.p, nS, c.len.p, sLen.p);
        SG_LAUNCHED();
        if (nB) {
            nflag.alloc(nB, st); nidx.alloc(nB, st);
            SG_CUDA(cudaMemsetAsync(nbr.p, 0, U, st));
            mark_neighbours_kernel<<<big_grid(nB), 256, 0, st>>>(c.edges.p, nB, c.explored.p, nbr.p);
            SG_LAUNCHED();
            flag_needed_edges_kernel<<<big_grid(nB), 256, 0, st>>>(c.edges.p, nB, c.explored.p, nbr.p, nflag.p);
            SG_LAUNCHED();
            exclusive_scan_u32(nflag.p, nidx.p, nB, d_nsel.p, st);
            SG_CUDA(cudaMemcpyAsync(&nSel, d_nsel.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
            SG_CUDA(cudaStreamSynchronize(st));
            selB.alloc(2 * (u64)nSel, st); selLen.alloc(nSel, st);
            if (nSel) {
                gather_needed_edges_kernel<<<big_grid(nB), 256, 0, st>>>(c.edges.p, nB, nflag.p, nidx.p, c.len.p, selB.p, selLen.p);
                SG_LAUNCHED();
            }
        }
        // ---- phase C: lists, marks and filtering on the device (phase_c_device.cu); the host contributes the traversal
        // order when the candidate set is not symmetric, and the whole walk only when a list is too long for a warp
        const bool force_host = getenv("SAGE2GPU_PHASE_C_HOST") != nullptr;       // test knob: always take the walk
        // the walk's input, fetched once (for the traversal order or for the whole walk)
        // the input of the whole walk (only when a list is too long for the device, or on request)
        PhaseCInput in;
        std::vector<u32> h_sids, h_off, h_selLen;
        std::vector<u64> h_cand, h_selB;
        std::vector<uint16_t> h_slen;
        auto fetch_input = [&]() {
            h_sids.resize(nS); h_off.resize((size_t)nS + 1); h_selLen.resize(nSel); h_cand.resize(nC); h_selB.resize(2 * (u64)nSel); h_slen.resize(nS);
            SG_CUDA(cudaMemcpyAsync(h_sids.data(), s_ids.p, nS * sizeof(u32), cudaMemcpyDeviceToHost, st));
            SG_CUDA(cudaMemcpyAsync(h_off.data(), offs.p, nS * sizeof(u32), cudaMemcpyDeviceToHost, st));
            SG_CUDA(cudaMemcpyAsync(h_slen.data(), sLen.p, nS * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
            if (nC) SG_CUDA(cudaMemcpyAsync(h_cand.data(), cand.p, nC * sizeof(u64), cudaMemcpyDeviceToHost, st));
            if (nSel) {
                SG_CUDA(cudaMemcpyAsync(h_selB.data(), selB.p, 2 * (u64)nSel * sizeof(u64), cudaMemcpyDeviceToHost, st));
                SG_CUDA(cudaMemcpyAsync(h_selLen.data(), selLen.p, nSel * sizeof(u32), cudaMemcpyDeviceToHost, st));
            }
            SG_CUDA(cudaStreamSynchronize(st));
            h_off[nS] = nC;
            in.nS = nS; in.s_ids = h_sids.data(); in.s_len = h_slen.data(); in.cand_off = h_off.data();
            in.cand = h_cand.data(); in.nB = nSel; in.edgesB = h_selB.data(); in.edgesB_len = h_selLen.data();
        };
        // the phase-B records that stay (owner not in S) are sorted now, asynchronously: the device works on them while the
        // host computes the traversal order; the few records of phase C are merged in afterwards
        auto sort_kept_b = [&]() {
            if (b_sorted) return;
            b_sorted = true;
            nKeep = nB;
            if (nB == 0) return;
            eflag.alloc(nB, st); eidx.alloc(nB, st);
            flag_keep_edges_kernel<<<big_grid(nB), 256, 0, st>>>(c.edges.p, nB, c.explored.p, eflag.p);
            SG_LAUNCHED();
            exclusive_scan_u32(eflag.p, eidx.p, nB, d_keep.p, st);
            u32 k32 = 0;
            SG_CUDA(cudaMemcpyAsync(&k32, d_keep.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
            SG_CUDA(cudaStreamSynchronize(st));
            nKeep = k32;
            if (nKeep == 0) return;
            ka0.alloc(nKeep, st); ka1.alloc(nKeep, st); kb0.alloc(nKeep, st); kb1.alloc(nKeep, st);
            split_edges_kernel<<<big_grid(nB), 256, 0, st>>>(c.edges.p, nB, eflag.p, eidx.p, ka0.p, kb0.p);
            SG_LAUNCHED();
            kcols.a[0] = ka0.p; kcols.a[1] = ka1.p; kcols.b[0] = kb0.p; kcols.b[1] = kb1.p; kcols.v[0] = kcols.v[1] = nullptr;
            kcur = radix_sort_bits(kcols, 0, nKeep, false, 0, id_bits, st);           // w0: to
            kcur = radix_sort_bits(kcols, kcur, nKeep, false, 32, 32 + id_bits, st);  // w0: from
            order_pair_runs_kernel<<<big_grid(nKeep), 256, 0, st>>>(kcols.a[kcur], kcols.b[kcur], nKeep);
            SG_LAUNCHED();
        };
        u64 removed_dev = 0, inserted_dev = 0;
        if (!force_host) {
            DevBuf<u32> d_order;
            const u32 *order = nullptr;
            bool order_known = true;
            if (!phase_c_is_symmetric(c, s_ids.p, idx.p, nS, counts.p, offs.p, cand.p)) {
                // which end point inserts an overlap depends on the breadth-first traversal (economyGraph.cpp:513-564, :605):
                // that part stays sequential, on the host; everything else follows from its order.  The device hands over
                // every list already sorted (phase_c_sorted_lists), into page-locked memory.
                DevBuf<u32> l_off, l_ent;
                order_known = phase_c_sorted_lists(c, s_ids.p, idx.p, nS, counts.p, offs.p, cand.p, nC, l_off, l_ent);
                if (order_known) {
                    DevBuf<uint8_t> d_hasb(nS, st);
                    SG_CUDA(cudaMemsetAsync(d_hasb.p, 0, nS, st));
                    if (nSel) { has_b_kernel<<<big_grid(nSel), 256, 0, st>>>(selB.p, nSel, c.explored.p, idx.p, d_hasb.p); SG_LAUNCHED(); }
                    const size_t off_bytes = ((size_t)nS + 1) * sizeof(u32), ent_bytes = 2 * (size_t)nC * sizeof(u32);
                    char *stage = (char *)c.pc_stage.ensure(off_bytes + ent_bytes + nS);
                    SG_CUDA(cudaMemcpyAsync(stage, l_off.p, off_bytes, cudaMemcpyDeviceToHost, st));
                    if (nC) SG_CUDA(cudaMemcpyAsync(stage + off_bytes, l_ent.p, ent_bytes, cudaMemcpyDeviceToHost, st));
                    SG_CUDA(cudaMemcpyAsync(stage + off_bytes + ent_bytes, d_hasb.p, nS, cudaMemcpyDeviceToHost, st));
                    SG_CUDA(cudaStreamSynchronize(st));
                    sort_kept_b();          // runs on the device while the host walks
                    PhaseCLists lists;
                    lists.nS = nS; lists.off = (const u32 *)stage; lists.ent = (const u32 *)(stage + off_bytes);
                    lists.has_b = (const uint8_t *)(stage + off_bytes + ent_bytes);
                    std::vector<u32> h_order;
                    host_order_ms = run_host_phase_c_order_lists(lists, h_order);
                    d_order.alloc(nS, st);
                    SG_CUDA(cudaMemcpyAsync(d_order.p, h_order.data(), nS * sizeof(u32), cudaMemcpyHostToDevice, st));
                    SG_CUDA(cudaStreamSynchronize(st));
                    order = d_order.p;
                    c.cnt.phase_c_on_device = 2;
                }
            } else c.cnt.phase_c_on_device = 1;
            if (order_known)
                c_on_device = device_phase_c(c, s_ids.p, idx.p, nS, counts.p, offs.p, cand.p, nC, selB.p, selLen.p, nSel, order, dev_c_edges, n_dev_c,
                                             inserted_dev, removed_dev);
        }
        if (c_on_device) {
            c.cnt.edges_inserted_c = inserted_dev;
            c.cnt.transitive_removed = removed_dev;
            SG_CUDA(cudaEventRecord(ev1, st));
        } else {
        c.cnt.phase_c_on_device = 0;
        fetch_input();
        sort_kept_b();
        SG_CUDA(cudaEventRecord(ev1, st));
        PhaseCOutput out;
        host_ms += run_host_phase_c(in, out);
        if (const char *dump = getenv("SAGE2GPU_DUMP_PHASE_C")) {      // test knob: the walk's input and output, for tools/phase_c_bench
            if (FILE *f = fopen(dump, "wb")) {
                const u64 hdr[6] = { nS, nC, nSel, out.edges.size(), out.inserted, out.removed };
                fwrite(hdr, sizeof(u64), 6, f);
                fwrite(h_sids.data(), sizeof(u32), nS, f); fwrite(h_slen.data(), sizeof(uint16_t), nS, f);
                fwrite(h_off.data(), sizeof(u32), (size_t)nS + 1, f); fwrite(h_cand.data(), sizeof(u64), nC, f);
                fwrite(h_selB.data(), sizeof(u64), 2 * (size_t)nSel, f); fwrite(h_selLen.data(), sizeof(u32), nSel, f);
                fwrite(out.edges.data(), sizeof(u64), out.edges.size(), f);
                fclose(f);
            }
        }
        c.cnt.edges_inserted_c = out.inserted;
        c.cnt.transitive_removed = out.removed;
        host_c_edges.swap(out.edges);
        }
    } else {
        SG_CUDA(cudaEventRecord(ev1, st));
    }

    if (c.opt_low_memory) {     // the table is not needed after the candidate scan: its memory serves the edge sort
        SG_CUDA(cudaStreamSynchronize(st));
        c.slots.release(); c.entries.release(); c.entries_loc.release();
        c.have_table = false;
        trim_default_pool(c.device, st);
    }
    // ---- assemble the final record set on device --------------------------------------------------
    // kept phase-B records (sorted above, or now) + the records of phase C (sorted here, few) -> one sorted list
    const u64 nH = c_on_device ? n_dev_c : host_c_edges.size() / 2;
    if (nS == 0) nKeep = nB;        // no read reached phase C: every phase-B record stays
    else if (!b_sorted) {           // (sort_kept_b is a lambda of the nS > 0 block: same steps when phase C never went to the host)
        nKeep = nB;
        if (nB) {
            eflag.alloc(nB, st); eidx.alloc(nB, st);
            flag_keep_edges_kernel<<<big_grid(nB), 256, 0, st>>>(c.edges.p, nB, c.explored.p, eflag.p);
            SG_LAUNCHED();
            exclusive_scan_u32(eflag.p, eidx.p, nB, d_keep.p, st);
            u32 k32 = 0;
            SG_CUDA(cudaMemcpyAsync(&k32, d_keep.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
            SG_CUDA(cudaStreamSynchronize(st));
            nKeep = k32;
        }
    }
    const u64 nAll = nKeep + nH;
    if (nAll == 0) {
        SG_CUDA(cudaEventRecord(ev2, st));
        SG_CUDA(cudaEventSynchronize(ev2));
        c.have_graph = true; c.rt_for_c = false;
        c.tm.phase_c_host = host_ms + host_order_ms;
        cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
        return;
    }
    const u64 *f0 = nullptr, *f1 = nullptr;     // the merged, sorted records
    DevBuf<u64> a0, a1, b0, b1, m0, m1;
    if (b_sorted) {
        // phase C's records: sort them alone, then merge the two sorted lists (no key occurs in both: a kept phase-B record
        // is owned by a read outside S, a phase-C record by a read of S)
        a0.alloc(nH + 1, st); a1.alloc(nH + 1, st); b0.alloc(nH + 1, st); b1.alloc(nH + 1, st);
        DevBuf<u64> tmp;
        if (nH && c_on_device) {
            split_edges_kernel<<<big_grid(nH), 256, 0, st>>>(dev_c_edges.p, nH, nullptr, nullptr, a0.p, b0.p);
            SG_LAUNCHED();
        } else if (nH) {
            tmp.alloc(2 * nH, st);
            SG_CUDA(cudaMemcpyAsync(tmp.p, host_c_edges.data(), 2 * nH * sizeof(u64), cudaMemcpyHostToDevice, st));
            split_edges_kernel<<<big_grid(nH), 256, 0, st>>>(tmp.p, nH, nullptr, nullptr, a0.p, b0.p);
            SG_LAUNCHED();
            SG_CUDA(cudaStreamSynchronize(st));   // host_c_edges / tmp lifetime
        }
        SortCols cc;
        cc.a[0] = a0.p; cc.a[1] = a1.p; cc.b[0] = b0.p; cc.b[1] = b1.p; cc.v[0] = cc.v[1] = nullptr;
        int ccur = 0;
        if (nH) {
            ccur = radix_sort_bits(cc, ccur, nH, false, 0, id_bits, st);
            ccur = radix_sort_bits(cc, ccur, nH, false, 32, 32 + id_bits, st);
            order_pair_runs_kernel<<<big_grid(nH), 256, 0, st>>>(cc.a[ccur], cc.b[ccur], nH);
            SG_LAUNCHED();
        }
        if (nH == 0) { f0 = kcols.a[kcur]; f1 = kcols.b[kcur]; }
        else if (nKeep == 0) { f0 = cc.a[ccur]; f1 = cc.b[ccur]; }
        else {
            m0.alloc(nAll, st); m1.alloc(nAll, st);
            merge_scatter_kernel<<<big_grid(nKeep), 256, 0, st>>>(kcols.a[kcur], kcols.b[kcur], nKeep, cc.a[ccur], cc.b[ccur], nH, m0.p, m1.p);
            SG_LAUNCHED();
            merge_scatter_kernel<<<big_grid(nH), 256, 0, st>>>(cc.a[ccur], cc.b[ccur], nH, kcols.a[kcur], kcols.b[kcur], nKeep, m0.p, m1.p);
            SG_LAUNCHED();
            f0 = m0.p; f1 = m1.p;
        }
        if (!f0) { f0 = kcols.a[kcur]; f1 = kcols.b[kcur]; }
    } else {
        a0.alloc(nAll, st); a1.alloc(nAll, st); b0.alloc(nAll, st); b1.alloc(nAll, st);
        if (nB) {
            split_edges_kernel<<<big_grid(nB), 256, 0, st>>>(c.edges.p, nB, nS > 0 ? eflag.p : nullptr, nS > 0 ? eidx.p : nullptr, a0.p, b0.p);
            SG_LAUNCHED();
        }
        if (nH && c_on_device) {
            split_edges_kernel<<<big_grid(nH), 256, 0, st>>>(dev_c_edges.p, nH, nullptr, nullptr, a0.p + nKeep, b0.p + nKeep);
            SG_LAUNCHED();
        } else if (nH) {
            DevBuf<u64> tmp(2 * nH, st);
            SG_CUDA(cudaMemcpyAsync(tmp.p, host_c_edges.data(), 2 * nH * sizeof(u64), cudaMemcpyHostToDevice, st));
            split_edges_kernel<<<big_grid(nH), 256, 0, st>>>(tmp.p, nH, nullptr, nullptr, a0.p + nKeep, b0.p + nKeep);
            SG_LAUNCHED();
            SG_CUDA(cudaStreamSynchronize(st));   // host_c_edges / tmp lifetime
        }
        SortCols cols;
        cols.a[0] = a0.p; cols.a[1] = a1.p; cols.b[0] = b0.p; cols.b[1] = b1.p; cols.v[0] = cols.v[1] = nullptr;
        int cur = 0;
        // stable LSD passes, least significant field first; the bit ranges are known, no reduction / host round trip
        cur = radix_sort_bits(cols, cur, nAll, false, 0, id_bits, st);           // w0: to
        cur = radix_sort_bits(cols, cur, nAll, false, 32, 32 + id_bits, st);     // w0: from
        // several edges between one pair of reads are rare (tandem repeats, palindromes): order their
        // (type, overhang) words in place instead of spending three more passes on every edge
        order_pair_runs_kernel<<<big_grid(nAll), 256, 0, st>>>(cols.a[cur], cols.b[cur], nAll);
        SG_LAUNCHED();
        f0 = cols.a[cur]; f1 = cols.b[cur];
    }
    DevBuf<u32> fflag(nAll, st), fidx(nAll, st), d_ne(1, st);
    flag_first_edge_kernel<<<big_grid(nAll), 256, 0, st>>>(f0, f1, nAll, fflag.p);
    SG_LAUNCHED();
    exclusive_scan_u32(fflag.p, fidx.p, nAll, d_ne.p, st);
    u32 nE = 0;
    SG_CUDA(cudaMemcpyAsync(&nE, d_ne.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    c.edges.alloc(2 * (u64)nE, st);
    join_edges_kernel<<<big_grid(nAll), 256, 0, st>>>(f0, f1, nAll, fflag.p, fidx.p, c.edges.p);
    SG_LAUNCHED();
    SG_CUDA(cudaEventRecord(ev2, st));
    SG_CUDA(cudaEventSynchronize(ev2));
    c.cnt.n_edges = nE;
    float ms01 = 0, ms12 = 0;
    cudaEventElapsedTime(&ms01, ev0, ev1);
    cudaEventElapsedTime(&ms12, ev1, ev2);
    c.tm.phase_c_dev = ms01 - host_order_ms > 0 ? ms01 - host_order_ms : 0;
    c.tm.phase_c_host = host_ms + host_order_ms;
    c.tm.sort_edges = ms12 - host_ms > 0 ? ms12 - host_ms : 0;
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaEventDestroy(ev2);
    c.have_graph = true; c.rt_for_c = false;
}

}  // namespace sg
