// search.cu -- step 3 search kernels (K4 phase A, K5 phase-C candidates).  Restates
// EconomyGraph::buildInitialOverlapGraph phase A (economyGraph/economyGraph.cpp:64-452),
// insertAllEdgesOfRead (:580-638), HashTable::hashTableSearch (hashTable.cpp:193-231) and
// compareStringInBytes[Previous] (economyGraph.cpp:712-808).
//
// One warp per read.  Lanes probe 32 consecutive windows at once (key extraction by funnel shifts from
// the read's words in shared memory, one slot sector per probe); every found window's bucket is then
// expanded 32 entries at a time: each lane fetches one partner record in the orientation the entry
// type asks for and verifies the whole overlap by XOR under a mask (which also re-verifies the key).
// Accepted hits are fed, in the reference's order (window ascending, bucket order), to the
// unique-extension state machine of core.cuh.
#include "context.h"

namespace sg {

constexpr int SR_WARPS = 8;

struct SearchParams {
    const u64 *F, *RC;
    const u64 *slots;
    const u32 *entries;
    u64 cap, U;
    int h, k;
};

template <int SW>
__device__ __forceinline__ u64 t_window32(const u64 *X, int s)
{
    const int i = s >> 5, sh = (s & 31) * 2;
    const u64 a = i < SW ? X[i] : 0ull;
    const u64 b = (i + 1) < SW ? X[i + 1] : 0ull;
    return sh == 0 ? a : ((a << sh) | (b >> (64 - sh)));
}

template <int SW>
__device__ __forceinline__ void t_extract_key(const u64 *X, int j, int h, u64 &v0, u64 &v1)
{
    if (h <= 32) { v0 = 0; v1 = t_window32<SW>(X, j) >> (64 - 2 * h); }
    else { v0 = t_window32<SW>(X, j) >> (64 - 2 * (h - 32)); v1 = t_window32<SW>(X, j + h - 32); }
}

// X (shared / global pointer, dynamic word index) against Y (registers, static index)
template <int SW>
__device__ __forceinline__ bool t_overlap_equal(const u64 *X, int lenX, int start, const u64 (&Y)[SW], int lenY, bool &contained)
{
    const int rem = lenX - start;
    contained = lenY <= rem;
    const int ov = contained ? lenY : rem;
    u64 acc = 0;
#pragma unroll
    for (int w = 0; w < SW; ++w) {
        const int nb = ov - 32 * w;
        if (nb > 0) {
            const u64 m = nb >= 32 ? ~0ull : ~(~0ull >> (2 * nb));
            acc |= (t_window32<SW>(X, start + 32 * w) ^ Y[w]) & m;
        }
    }
    return acc == 0;
}

// hashTableSearch: linear probe; a slot whose tag matches is confirmed by re-extracting the key from
// its first entry's read (hashTable.cpp:203-220); masked keys (>= 100 entries) read as absent.
template <int SW>
__device__ __forceinline__ bool probe_key(const SearchParams &P, u64 v0, u64 v1, u32 &off, u32 &cnt)
{
    const u64 hsh = hash_key(v0, v1);
    const u64 tag = slot_tag(hsh);
    u64 s = slot_home(hsh, P.cap);
    for (;;) {
        const u64 slot = __ldg(&P.slots[s]);
        if (slot == 0) return false;
        if (slot_get_tag(slot) == tag) {
            const u64 o = slot_get_offset(slot);
            const u32 ent = __ldg(&P.entries[o]);
            const u64 rid = ent >> 2;
            const int type = (int)(ent & 3);
            const u64 *X = ((type & 2) ? P.RC : P.F) + rid * SW;
            const int l = (int)(__ldg(&X[SW - 1]) & 0xFFFF);
            u64 w0, w1;
            t_extract_key<SW>(X, (type & 1) ? l - P.h : 0, P.h, w0, w1);
            if (w0 == v0 && w1 == v1) {
                const u32 c = slot_get_count(slot);
                if (c >= (u32)kHashThreshold) return false;
                off = (u32)o; cnt = c;
                return true;
            }
        }
        s = (s + 1 == P.cap) ? 0 : s + 1;
    }
}

template <int SW>
__device__ __forceinline__ void load_record(const u64 *src, u64 (&q)[SW])
{
#pragma unroll
    for (int w = 0; w < SW; ++w) q[w] = __ldg(&src[w]);
}

// ------------------------------------------------------------------------------------------------
// K4: phase A, exact sequential chain (reference order).
// ------------------------------------------------------------------------------------------------
template <int SW>
__global__ void __launch_bounds__(SR_WARPS * 32)
phase_a_kernel(SearchParams P, u64 *__restrict__ extR, u64 *__restrict__ extL, uint8_t *__restrict__ flag5,
               u32 *__restrict__ cont_max, unsigned long long *__restrict__ counters)
{
    __shared__ u64 sXf[SR_WARPS][SW], sXr[SR_WARPS][SW], sPrevR[SR_WARPS][SW], sPrevL[SR_WARPS][SW], sQ[SR_WARPS][SW];
    __shared__ ExtState sState[SR_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 *Xf = sXf[warp], *Xr = sXr[warp];
    ExtState &st = sState[warp];
    const u64 nwarps = (u64)gridDim.x * SR_WARPS;
    unsigned long long calls = 0, probes = 0;

    for (u64 i = (u64)blockIdx.x * SR_WARPS + warp; i < P.U; i += nwarps) {
        if (lane < SW) { Xf[lane] = P.F[i * SW + lane]; Xr[lane] = P.RC[i * SW + lane]; }
        if (lane == 0) ext_init(st);
        __syncwarp();
        const int len1 = (int)(Xf[SW - 1] & 0xFFFF);
        const int W = len1 - P.h + 1;
        for (int base = 0; base < W; base += 32) {
            const int j = base + lane;
            bool found = false;
            u32 off = 0, cnt = 0;
            if (j < W) {
                u64 v0, v1;
                t_extract_key<SW>(Xf, j, P.h, v0, v1);
                found = probe_key<SW>(P, v0, v1, off, cnt);
                probes++;
            }
            unsigned fm = __ballot_sync(0xffffffffu, found);
            while (fm) {
                const int l = __ffs(fm) - 1;
                fm &= fm - 1;
                const int jj = base + l;
                const u32 o = __shfl_sync(0xffffffffu, off, l), c = __shfl_sync(0xffffffffu, cnt, l);
                if (lane == 0) ext_new_window(st);
                __syncwarp();
                const bool gateR = gate_right(jj, len1, P.k), gateL = gate_left(jj, P.k, P.h);
                for (u32 e0 = 0; e0 < c; e0 += 32) {
                    const u32 e = e0 + lane;
                    bool hit = false, right = false;
                    u32 rid2 = 0;
                    int len2 = 0, type = 0;
                    u64 q[SW];
                    if (e < c) {
                        const u32 ent = __ldg(&P.entries[o + e]);
                        rid2 = ent >> 2; type = (int)(ent & 3);
                        right = !(type & 1);
                        if (rid2 != (u32)i && (right ? gateR : gateL)) {
                            const bool use_rc = partner_uses_rc(type);
                            load_record<SW>((use_rc ? P.RC : P.F) + (u64)rid2 * SW, q);
                            len2 = (int)(q[SW - 1] & 0xFFFF);
                            bool contained;
                            calls++;
                            const bool ok = t_overlap_equal<SW>(right ? Xf : Xr, len1, right ? jj : len1 - jj - P.h, q, len2, contained);
                            if (ok && contained) atomicMax(&cont_max[rid2], (u32)(i + 1));   // economyGraph.cpp:735
                            hit = ok && !contained;
                        }
                    }
                    unsigned hm = __ballot_sync(0xffffffffu, hit);
                    while (hm) {
                        const int b = __ffs(hm) - 1;
                        hm &= hm - 1;
                        if (lane == b) {
                            u64 *Q = sQ[warp];
#pragma unroll
                            for (int w = 0; w < SW; ++w) Q[w] = q[w];
                            if (right) ext_right_hit(st, sPrevR[warp], Q, SW, rid2 + 1, type >> 1, jj, len1, len2);
                            else ext_left_hit(st, sPrevL[warp], Q, SW, rid2 + 1, type >> 1, jj, P.h, len1, len2);
                        }
                        __syncwarp();
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            flag5[i] = st.connections > kConnectionsLimit ? 1 : 0;              // :443
            const bool amb = st.itsAmbigR == 1 || st.itsAmbigL == 1;           // :446-450
            extR[i] = ext_pack(st.Rid, st.Rtype, amb ? 0u : st.Rlen);
            extL[i] = ext_pack(st.Lid, st.Ltype, amb ? 0u : st.Llen);
        }
        __syncwarp();
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        calls += __shfl_xor_sync(0xffffffffu, calls, s);
        probes += __shfl_xor_sync(0xffffffffu, probes, s);
    }
    if (lane == 0) { atomicAdd(&counters[0], calls); atomicAdd(&counters[1], probes); }
}

// ------------------------------------------------------------------------------------------------
// K5: phase-C candidates.  For every read r1 still unexplored after phase B (state 0) list, in the
// reference's order, every (read2, edge type, overhang) insertAllEdgesOfRead would test positive
// (economyGraph.cpp:599-631), restricted to read2 that are themselves state 0 (any other state is
// skipped at :605 and never returns to 0).  Pass 1 counts, pass 2 fills.
// candidate = read2(1-based) << 32 | edgeType << 20 | (overhang & 0xFFFFF)
// ------------------------------------------------------------------------------------------------
template <int SW, bool FILL>
__global__ void __launch_bounds__(SR_WARPS * 32)
phase_c_kernel(SearchParams P, const u32 *__restrict__ s_ids, u64 nS, const uint8_t *__restrict__ explored,
               u32 *__restrict__ counts, const u32 *__restrict__ offsets, u64 *__restrict__ cand)
{
    __shared__ u64 sXf[SR_WARPS][SW], sXr[SR_WARPS][SW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 *Xf = sXf[warp], *Xr = sXr[warp];
    const u64 nwarps = (u64)gridDim.x * SR_WARPS;
    for (u64 s = (u64)blockIdx.x * SR_WARPS + warp; s < nS; s += nwarps) {
        const u64 i = s_ids[s];      // 0-based
        if (lane < SW) { Xf[lane] = P.F[i * SW + lane]; Xr[lane] = P.RC[i * SW + lane]; }
        __syncwarp();
        const int len1 = (int)(Xf[SW - 1] & 0xFFFF);
        const int W = len1 - P.h + 1;
        u32 total = 0;
        const u64 out_base = FILL ? offsets[s] : 0;
        for (int base = 0; base < W; base += 32) {
            const int j = base + lane;
            bool found = false;
            u32 off = 0, cnt = 0;
            if (j < W) {
                u64 v0, v1;
                t_extract_key<SW>(Xf, j, P.h, v0, v1);
                found = probe_key<SW>(P, v0, v1, off, cnt);
            }
            unsigned fm = __ballot_sync(0xffffffffu, found);
            while (fm) {
                const int l = __ffs(fm) - 1;
                fm &= fm - 1;
                const int jj = base + l;
                const u32 o = __shfl_sync(0xffffffffu, off, l), c = __shfl_sync(0xffffffffu, cnt, l);
                const bool gateR = gate_right(jj, len1, P.k), gateL = gate_left(jj, P.k, P.h);
                for (u32 e0 = 0; e0 < c; e0 += 32) {
                    const u32 e = e0 + lane;
                    bool ok = false;
                    u64 rec = 0;
                    if (e < c) {
                        const u32 ent = __ldg(&P.entries[o + e]);
                        const u32 rid2 = ent >> 2;
                        const int type = (int)(ent & 3);
                        const bool right = !(type & 1);
                        if (rid2 != (u32)i && explored[rid2] == 0 && (right ? gateR : gateL)) {
                            const bool use_rc = partner_uses_rc(type);
                            u64 q[SW];
                            load_record<SW>((use_rc ? P.RC : P.F) + (u64)rid2 * SW, q);
                            const int len2 = (int)(q[SW - 1] & 0xFFFF);
                            bool contained;
                            ok = t_overlap_equal<SW>(right ? Xf : Xr, len1, right ? jj : len1 - jj - P.h, q, len2, contained);
                            rec = candidate_record(type, jj, P.h, len1, len2, rid2);
                        }
                    }
                    const unsigned om = __ballot_sync(0xffffffffu, ok);
                    if (FILL && ok) cand[out_base + total + __popc(om & ((1u << lane) - 1u))] = rec;
                    total += __popc(om);
                }
            }
        }
        if (!FILL && lane == 0) counts[s] = total;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static unsigned search_grid(u64 n_reads)
{
    u64 g = (n_reads + SR_WARPS - 1) / SR_WARPS;
    const u64 cap = (u64)kSMs * 8;
    if (g > cap) g = cap;
    if (g == 0) g = 1;
    return (unsigned)g;
}

template <int SW>
static void launch_phase_a(Context &c, const SearchParams &P, unsigned long long *d_counters)
{
    phase_a_kernel<SW><<<search_grid(P.U), SR_WARPS * 32, 0, c.stream>>>(P, c.extR.p, c.extL.p, c.flag5.p, c.cont_max.p, d_counters);
}

template <int SW>
static void launch_phase_c(Context &c, const SearchParams &P, const u32 *s_ids, u64 nS, u32 *counts, const u32 *offsets, u64 *cand, bool fill)
{
    if (fill) phase_c_kernel<SW, true><<<search_grid(nS), SR_WARPS * 32, 0, c.stream>>>(P, s_ids, nS, c.explored.p, counts, offsets, cand);
    else phase_c_kernel<SW, false><<<search_grid(nS), SR_WARPS * 32, 0, c.stream>>>(P, s_ids, nS, c.explored.p, counts, offsets, cand);
}

#define SG_DISPATCH_SW(SWV, CALL)                                                       \
    switch (SWV) {                                                                      \
        case 2: { constexpr int SWC = 2; CALL; } break;                                 \
        case 3: { constexpr int SWC = 3; CALL; } break;                                 \
        case 4: { constexpr int SWC = 4; CALL; } break;                                 \
        case 5: { constexpr int SWC = 5; CALL; } break;                                 \
        case 6: { constexpr int SWC = 6; CALL; } break;                                 \
        case 8: { constexpr int SWC = 8; CALL; } break;                                 \
        case 12: { constexpr int SWC = 12; CALL; } break;                               \
        case 16: { constexpr int SWC = 16; CALL; } break;                               \
        case 32: { constexpr int SWC = 32; CALL; } break;                               \
        default: throw CudaError("unsupported record stride");                         \
    }

static SearchParams make_params(const Context &c)
{
    SearchParams P;
    P.F = c.F.p; P.RC = c.RC.p; P.slots = c.slots.p; P.entries = c.entries.p;
    P.cap = c.cap; P.U = c.cnt.unique_reads; P.h = c.h; P.k = c.min_overlap;
    return P;
}

void stage_phase_a(Context &c)
{
    cudaStream_t st = c.stream;
    SG_CHECK(c.have_table, "build_hash_table must run before the overlap search");
    const u64 U = c.cnt.unique_reads;
    c.extR.alloc(U, st); c.extL.alloc(U, st); c.flag5.alloc(U, st); c.cont_max.alloc(U, st);
    c.cnt.compare_calls = 0; c.cnt.window_probes = 0;
    if (U == 0) return;
    SG_CUDA(cudaMemsetAsync(c.cont_max.p, 0, U * sizeof(u32), st));
    DevBuf<unsigned long long> d_counters(2, st);
    SG_CUDA(cudaMemsetAsync(d_counters.p, 0, 2 * sizeof(unsigned long long), st));
    const SearchParams P = make_params(c);
    cudaEvent_t e0, e1;
    SG_CUDA(cudaEventCreate(&e0)); SG_CUDA(cudaEventCreate(&e1));
    SG_CUDA(cudaEventRecord(e0, st));
    SG_DISPATCH_SW(c.SW, launch_phase_a<SWC>(c, P, d_counters.p));
    SG_LAUNCHED();
    SG_CUDA(cudaEventRecord(e1, st));
    unsigned long long h[2];
    SG_CUDA(cudaMemcpyAsync(h, d_counters.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    c.cnt.compare_calls = h[0];
    c.cnt.window_probes = h[1];
    cudaEventElapsedTime(&c.tm.phase_a_kernel, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

// used by graph.cu
void launch_phase_c_candidates(Context &c, const u32 *s_ids, u64 nS, u32 *counts, const u32 *offsets, u64 *cand, bool fill)
{
    const SearchParams P = make_params(c);
    SG_DISPATCH_SW(c.SW, launch_phase_c<SWC>(c, P, s_ids, nS, counts, offsets, cand, fill));
    SG_LAUNCHED();
}

}  // namespace sg
