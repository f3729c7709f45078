// mapids.cu -- step-6 mapping of reads to read ids: ReadLoader::getIdOfRead
// (inputReader/readLoader.cpp:319-353) for a whole batch of reads, behind the isGoodRead gate of
// MatePair::processMatePairs (matePair/matePair.cpp:176-179).
//
// Pass 1, one warp per query read: the ASCII bases are read coalesced and packed in registers (lane w holds word w of
// the forward and of the reverse-complement record; core.cuh layout, so an unsigned word-wise compare is
// Read::operator< / stringCompareInBytes).  The orientation the reference would look up is the smaller one (ties ->
// reverse complement, flag -1, readLoader.cpp:325-334).  Pass 2, one thread per query: the binary search over the
// sorted unique reads F (readLoader.cpp:335-348), 32 independent random lines in flight per warp.
#include "context.h"

namespace sg {

constexpr int MP_WARPS = 8;

__device__ __forceinline__ int base_code(uint8_t c) { return ((c >> 1) ^ (c >> 2)) & 3; }       // A0 C1 G2 T3 (either case)
__device__ __forceinline__ bool base_valid(uint8_t c) { c &= 0xDF; return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// Packing of one read by ONE thread, four characters per step: two aligned 32-bit words funnel-shifted to the read's
// byte phase, the codes of four bytes at once, validity by rebuilding the upper-case character each code stands for, the
// four 2-bit codes gathered into one byte by a multiply (the same arithmetic as reads.cu's pack_thread_kernel).
// f[] receives the forward record without its length.  Returns false on a non-ACGT character.
template <int SW>
__device__ __forceinline__ bool pack_read_simd(const uint8_t *__restrict__ s, int len, u64 (&f)[SW])
{
    const u32 *Bw = reinterpret_cast<const u32 *>(reinterpret_cast<uintptr_t>(s) & ~(uintptr_t)3);
    const unsigned ph = (unsigned)(reinterpret_cast<uintptr_t>(s) & 3) * 8;
    u32 lo = __ldg(Bw), bad_bits = 0;
#pragma unroll
    for (int w = 0; w < SW; ++w) {
        u64 acc = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c4 = 8 * w + q;
            if (4 * c4 < len) {
                const int left = len - 4 * c4;
                // the second word is needed only when the four characters reach into it (never read past the read's last byte's word)
                const u32 hi = (ph != 0 && (int)(ph >> 3) + left > 4) || (ph == 0 && left > 4) ? __ldg(Bw + c4 + 1) : 0u;
                u32 x = __funnelshift_r(lo, hi, ph);
                lo = hi;
                if (left < 4) { const u32 keep = 0xFFFFFFFFu >> (8 * (4 - left)); x = (x & keep) | (0x41414141u & ~keep); }
                const u32 c = ((x >> 1) ^ (x >> 2)) & 0x03030303u;
                const u32 c0 = c & 0x01010101u, c1 = (c >> 1) & 0x01010101u;
                const u32 expect = 0x41414141u + (c0 & ~c1) * 2u + (c1 & ~c0) * 6u + (c1 & c0) * 0x13u;
                bad_bits |= expect ^ (x & 0xDFDFDFDFu);
                acc |= (u64)((c * 0x40100401u) >> 24) << (56 - 8 * q);
            }
        }
        f[w] = acc;
    }
    return bad_bits == 0;
}

// One thread per query: pack (above), reverse complement, pick the orientation the reference looks up, binary search of
// readLoader.cpp:335-348 over the sorted unique reads -- a warp keeps 32 independent random lines in flight; words after
// the first differing one are never read.
template <int SW>
__global__ void __launch_bounds__(256) map_reads_thread_kernel(const uint8_t *__restrict__ bases, const int64_t *__restrict__ off, u64 n, int k,
                                                               const u64 *__restrict__ F, u64 U, int SWS,
                                                               long long *__restrict__ ids, uint8_t *__restrict__ good)
{
    for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (u64)gridDim.x * blockDim.x) {
        const int64_t o0 = off[q];
        const int64_t len64 = off[q + 1] - o0;
        const uint8_t *s = bases + o0;
        const bool fits = len64 <= (int64_t)(32 * SW - 8);
        bool ok = len64 > (int64_t)k;                       // utils.cpp:146
        long long id = 0;
        if (ok && fits) {
            const int len = (int)len64;
            u64 f[SW], r[SW];
            ok = pack_read_simd<SW>(s, len, f);
            if (ok && U > 0) {
                revcomp_record(f, r, SW, len);              // r carries the length
                f[SW - 1] |= (u64)len;
                bool fwd_smaller = false;                   // read.compare(read_r) < 0 (readLoader.cpp:325); ties look the revcomp up
#pragma unroll
                for (int w = SW - 1; w >= 0; --w) if (f[w] != r[w]) fwd_smaller = f[w] < r[w];
#pragma unroll
                for (int w = 0; w < SW; ++w) f[w] = fwd_smaller ? f[w] : r[w];
                long long lb = 0, ub = (long long)U - 1;
                while (lb <= ub) {
                    const long long mid = (lb + ub) >> 1;
                    const u64 *X = F + (u64)mid * SWS;
                    u64 x = __ldg(&X[0]), y = f[0];
#pragma unroll
                    for (int w = 1; w < SW; ++w) if (x == y) { x = __ldg(&X[w]); y = f[w]; }
                    if (x == y) { id = fwd_smaller ? mid + 1 : -(mid + 1); break; }
                    if (y > x) lb = mid + 1; else ub = mid - 1;
                }
            }
        } else if (ok) {                                    // longer than any read of the set: good or not, never present
            for (int64_t p = 0; p < len64; ++p) ok = ok && base_valid(s[p]);
        }
        ids[q] = id;
        if (good) good[q] = ok ? 1 : 0;
    }
}

// records of more than 8 words (reads longer than 248 bases): one warp per query packs it (coalesced byte loads) and writes
// the record to look up + its flags
//   meta bit 0: isGoodRead, bit 1: the read itself is the smaller orientation (flag +1), bit 2: worth searching
__global__ void __launch_bounds__(MP_WARPS * 32) map_pack_kernel(const uint8_t *__restrict__ bases, const int64_t *__restrict__ off, u64 n, int k,
                                                                 u64 U, int SW, u64 *__restrict__ qrec, uint8_t *__restrict__ meta)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const u64 warp = (u64)blockIdx.x * MP_WARPS + (threadIdx.x >> 5), nwarps = (u64)gridDim.x * MP_WARPS;
    for (u64 q = warp; q < n; q += nwarps) {
        const int64_t o0 = off[q];
        const int64_t len64 = off[q + 1] - o0;
        const uint8_t *s = bases + o0;
        const bool fits = len64 <= (int64_t)(32 * SW - 8);
        const int len = fits ? (int)len64 : 0;
        bool ok = len64 > (int64_t)k;                       // utils.cpp:146
        u64 myF = 0, myR = 0;
        if (ok && fits) {
            const int nw = (len + 31) >> 5;
            for (int w = 0; w < nw; ++w) {
                const int p = 32 * w + lane;
                uint8_t cf = 'A', cr = 'T';
                if (p < len) { cf = s[p]; cr = s[len - 1 - p]; }
                ok = ok && base_valid(cf);
                const unsigned f = p < len ? (unsigned)base_code(cf) : 0u, r = p < len ? 3u - (unsigned)base_code(cr) : 0u;
                const int sh = 30 - 2 * (lane & 15);
                const unsigned fh = __reduce_or_sync(FULL, lane < 16 ? f << sh : 0u), fl = __reduce_or_sync(FULL, lane >= 16 ? f << sh : 0u);
                const unsigned rh = __reduce_or_sync(FULL, lane < 16 ? r << sh : 0u), rl = __reduce_or_sync(FULL, lane >= 16 ? r << sh : 0u);
                if (lane == w) { myF = ((u64)fh << 32) | fl; myR = ((u64)rh << 32) | rl; }
            }
            if (lane == SW - 1) { myF |= (u64)len; myR |= (u64)len; }
        } else if (ok) {                                    // longer than any read of the set: good or not, never present
            for (int64_t p = lane; p < len64; p += 32) ok = ok && base_valid(s[p]);
        }
        ok = __all_sync(FULL, ok);
        // read.compare(read_r) < 0 (readLoader.cpp:325): the codes order like the characters; ties look the revcomp up
        const unsigned dm = __ballot_sync(FULL, lane < SW && myF != myR);
        bool fwd_smaller = false;
        if (dm) {
            const int d = __ffs(dm) - 1;
            fwd_smaller = __shfl_sync(FULL, myF, d) < __shfl_sync(FULL, myR, d);
        }
        if (lane < SW) qrec[q * SW + lane] = fwd_smaller ? myF : myR;
        if (lane == 0) meta[q] = (uint8_t)((ok ? 1 : 0) | (fwd_smaller ? 2 : 0) | ((ok && fits && U > 0) ? 4 : 0));
    }
}

// pass 2: one THREAD per query walks the binary search of readLoader.cpp:335-348 over the sorted unique reads: a warp
// keeps 32 independent random lines in flight; words after the first differing one are never read
__global__ void __launch_bounds__(256) map_search_kernel(const u64 *__restrict__ qrec, const uint8_t *__restrict__ meta, u64 n,
                                                         const u64 *__restrict__ F, u64 U, int SW, int SWS,
                                                         long long *__restrict__ ids, uint8_t *__restrict__ good)
{
    for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (u64)gridDim.x * blockDim.x) {
        const uint8_t m = meta[q];
        long long id = 0;
        if (m & 4) {
            const u64 *Q = qrec + q * SW;
            const u64 q0 = Q[0];
            long long lb = 0, ub = (long long)U - 1;
            while (lb <= ub) {
                const long long mid = (lb + ub) >> 1;
                const u64 *X = F + (u64)mid * SWS;
                u64 x = __ldg(&X[0]), y = q0;
                for (int w = 1; x == y && w < SW; ++w) { x = __ldg(&X[w]); y = Q[w]; }
                if (x == y) { id = (m & 2) ? mid + 1 : -(mid + 1); break; }
                if (y > x) lb = mid + 1; else ub = mid - 1;
            }
        }
        ids[q] = id;
        if (good) good[q] = m & 1;
    }
}

// bases / offsets: host (uploaded here) or device buffers; ids / good: host buffers.  Returns the kernel milliseconds.
float stage_map_reads(Context &c, const uint8_t *bases, const int64_t *offsets, int64_t n_reads, bool device_resident,
                      int64_t *ids, uint8_t *good)
{
    cudaStream_t st = c.stream;
    ArenaScope arena_scope(c.arena, st);
    SG_CHECK(c.have_reads, "organize_reads must run first");
    SG_CHECK(n_reads >= 0 && ids != nullptr, "bad arguments");
    if (n_reads == 0) return 0.f;
    SG_CHECK(bases != nullptr && offsets != nullptr, "null input");
    const u64 n = (u64)n_reads;
    DevBuf<int64_t> d_off;
    DevBuf<uint8_t> d_bases, d_good(n, st);
    DevBuf<long long> d_ids(n, st);
    const uint8_t *pb = bases;
    const int64_t *po = offsets;
    if (!device_resident) {
        const int64_t first = offsets[0], total = offsets[n] - first;      // offsets are relative to `bases`
        SG_CHECK(total >= 0, "offsets must ascend");
        d_off.alloc(n + 1, st);
        d_bases.alloc((size_t)total + 1, st);
        SG_CUDA(cudaMemcpyAsync(d_off.p, offsets, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        if (total) SG_CUDA(cudaMemcpyAsync(d_bases.p, bases + first, (size_t)total, cudaMemcpyHostToDevice, st));
        pb = d_bases.p - first; po = d_off.p;
    }
    cudaEvent_t e0, e1;
    SG_CUDA(cudaEventCreate(&e0)); SG_CUDA(cudaEventCreate(&e1));
    SG_CUDA(cudaEventRecord(e0, st));
    u64 g = (n + 255) / 256;
    if (g > (u64)kSMs * 8) g = (u64)kSMs * 8;
    const u64 U = c.cnt.unique_reads;
    if (c.SW <= 8) {      // one fused thread-per-read kernel
        switch (c.SW) {
            case 2: map_reads_thread_kernel<2><<<(unsigned)g, 256, 0, st>>>(pb, po, n, c.min_overlap, c.F.p, U, c.SWS, d_ids.p, d_good.p); break;
            case 3: map_reads_thread_kernel<3><<<(unsigned)g, 256, 0, st>>>(pb, po, n, c.min_overlap, c.F.p, U, c.SWS, d_ids.p, d_good.p); break;
            case 4: map_reads_thread_kernel<4><<<(unsigned)g, 256, 0, st>>>(pb, po, n, c.min_overlap, c.F.p, U, c.SWS, d_ids.p, d_good.p); break;
            case 5: map_reads_thread_kernel<5><<<(unsigned)g, 256, 0, st>>>(pb, po, n, c.min_overlap, c.F.p, U, c.SWS, d_ids.p, d_good.p); break;
            case 6: map_reads_thread_kernel<6><<<(unsigned)g, 256, 0, st>>>(pb, po, n, c.min_overlap, c.F.p, U, c.SWS, d_ids.p, d_good.p); break;
            default: map_reads_thread_kernel<8><<<(unsigned)g, 256, 0, st>>>(pb, po, n, c.min_overlap, c.F.p, U, c.SWS, d_ids.p, d_good.p); break;
        }
        SG_LAUNCHED();
    } else {              // long reads: warp-per-read pack, then the thread-per-read search
        DevBuf<u64> qrec(n * (u64)c.SW, st);
        DevBuf<uint8_t> meta(n, st);
        u64 gp = (n + MP_WARPS - 1) / MP_WARPS;
        if (gp > (u64)kSMs * 16) gp = (u64)kSMs * 16;
        map_pack_kernel<<<(unsigned)gp, MP_WARPS * 32, 0, st>>>(pb, po, n, c.min_overlap, U, c.SW, qrec.p, meta.p);
        SG_LAUNCHED();
        map_search_kernel<<<(unsigned)g, 256, 0, st>>>(qrec.p, meta.p, n, c.F.p, U, c.SW, c.SWS, d_ids.p, d_good.p);
        SG_LAUNCHED();
    }
    SG_CUDA(cudaEventRecord(e1, st));
    SG_CUDA(cudaMemcpyAsync(ids, d_ids.p, n * sizeof(long long), cudaMemcpyDeviceToHost, st));
    if (good) SG_CUDA(cudaMemcpyAsync(good, d_good.p, n, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

}  // namespace sg
