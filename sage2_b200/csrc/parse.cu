// parse.cu -- FASTA / FASTQ record splitting on the device: the host only moves raw text.
// Restates what the reference does per record through kseq (inputReader/fastAQReader.cpp:16-45,
// readLoader.cpp:146-160) for the REGULAR layouts -- 4-line FASTQ (@name / sequence / +... / quality of the
// same length) and 2-line FASTA (>name / sequence).  Anything else (multi-line records, blank lines, a
// sequence line that starts with '>', '@' or '+', an empty sequence, a truncated tail) makes the call report
// "irregular" without appending anything, and the caller hands that text to its sequential parser.
//
// Per chunk: newline flags -> exclusive scan -> line starts; per record a structure check and the sequence
// length -> exclusive scan -> sequences copied (one warp per record) behind what was uploaded before.
#include "context.h"

namespace sg {

__global__ void __launch_bounds__(256) nl_flag_kernel(const uint8_t *__restrict__ text, u64 n, u32 *__restrict__ flag)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) flag[i] = text[i] == '\n';
}

__global__ void __launch_bounds__(256) line_start_kernel(const uint8_t *__restrict__ text, u64 n, const u32 *__restrict__ idx, u32 *__restrict__ ls)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        if (i == 0) ls[0] = 0;
        if (text[i] == '\n') ls[idx[i] + 1] = (u32)(i + 1);
    }
}

// out[0] = bytes consumed by R records, out[1] |= 1 on any irregular record
__global__ void __launch_bounds__(256) record_check_kernel(const uint8_t *__restrict__ text, const u32 *__restrict__ ls, u64 R, int lpr, int marker,
                                                            u32 *__restrict__ seq_len, u32 *__restrict__ out)
{
    bool bad = false;
    for (u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x; r < R; r += (u64)gridDim.x * blockDim.x) {
        const u32 h0 = ls[lpr * r], s0 = ls[lpr * r + 1], s1 = ls[lpr * r + 2];
        bad |= text[h0] != (uint8_t)marker;
        u32 sl = s1 - 1 - s0;                                      // without the '\n'
        if (sl && text[s0 + sl - 1] == '\r') --sl;
        const uint8_t c0 = sl ? text[s0] : (uint8_t)'>';
        bad |= sl == 0 || c0 == '>' || c0 == '@' || c0 == '+';
        if (lpr == 4) {
            const u32 q0 = ls[4 * r + 3], q1 = ls[4 * r + 4];
            u32 ql = q1 - 1 - q0;
            if (ql && text[q0 + ql - 1] == '\r') --ql;
            bad |= text[s1] != '+' || ql != sl;
        }
        seq_len[r] = sl;
        if (r + 1 == R) out[0] = ls[lpr * R];
    }
    if (bad) atomicOr(&out[1], 1u);
}

__global__ void __launch_bounds__(256) record_copy_kernel(const uint8_t *__restrict__ text, const u32 *__restrict__ ls, const u32 *__restrict__ seq_len,
                                                           const u32 *__restrict__ seq_off, u64 R, int lpr, u64 base_off,
                                                           uint8_t *__restrict__ bases, int64_t *__restrict__ offsets)
{
    const int lane = threadIdx.x & 31;
    const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
    for (u64 r = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < R; r += nwarps) {
        const u32 s0 = ls[lpr * r + 1], sl = seq_len[r];
        const u64 o = base_off + seq_off[r];
        for (u32 t = lane; t < sl; t += 32) bases[o + t] = text[s0 + t];
        if (lane == 0) { offsets[r] = (int64_t)o; if (r + 1 == R) offsets[R] = (int64_t)(o + sl); }
    }
}

template <typename T>
static void grow_persistent(DevBuf<T> &b, size_t used, size_t need, cudaStream_t st)
{
    if (need <= b.cap) { b.n = need; return; }
    size_t cap = b.cap ? b.cap : (size_t)1 << 20;
    while (cap < need) cap += cap / 2 + 1;
    DevBuf<T> nb;
    nb.persistent = true;
    nb.alloc(cap, st);
    if (used) SG_CUDA(cudaMemcpyAsync(nb.p, b.p, used * sizeof(T), cudaMemcpyDeviceToDevice, st));
    b = std::move(nb);
}

static unsigned pgrid(u64 n, unsigned per_thread = 4)
{
    unsigned g = grid_for(n, 256, per_thread);
    return g > kSMs * 16u ? kSMs * 16u : g;
}

bool stage_parse_text_chunk(Context &c, const uint8_t *text, u64 n_bytes, bool final, int &marker, u64 max_records, u64 &consumed, u64 &n_records)
{
    cudaStream_t st = c.stream;
    ArenaScope arena_scope(c.arena, st);
    consumed = 0; n_records = 0;
    if (n_bytes == 0 || max_records == 0) return true;
    SG_CHECK(n_bytes < 0x7FFFFFF0ull, "text chunks must be smaller than 2 GiB");
    if (marker == 0) {
        if (text[0] != '@' && text[0] != '>') return false;
        marker = text[0];
    }
    const int lpr = marker == '@' ? 4 : 2;
    const bool add_nl = final && text[n_bytes - 1] != '\n';
    const u64 n = n_bytes + (add_nl ? 1 : 0);
    DevBuf<uint8_t> d_text(n, st);
    SG_CUDA(cudaMemcpyAsync(d_text.p, text, n_bytes, cudaMemcpyHostToDevice, st));
    if (add_nl) SG_CUDA(cudaMemsetAsync(d_text.p + n_bytes, '\n', 1, st));
    DevBuf<u32> flag(n, st), idx(n, st), d_lines(1, st);
    nl_flag_kernel<<<pgrid(n), 256, 0, st>>>(d_text.p, n, flag.p);
    SG_LAUNCHED();
    exclusive_scan_u32(flag.p, idx.p, n, d_lines.p, st);
    u32 L = 0;
    SG_CUDA(cudaMemcpyAsync(&L, d_lines.p, sizeof(u32), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    u64 R = (u64)L / (u64)lpr;
    if (R > max_records) R = max_records;
    if (R == 0) return !final;                       // no complete record in this chunk (a final stub is irregular)
    DevBuf<u32> ls((size_t)L + 1, st);
    line_start_kernel<<<pgrid(n), 256, 0, st>>>(d_text.p, n, idx.p, ls.p);
    SG_LAUNCHED();
    DevBuf<u32> seq_len(R, st), seq_off(R, st), d_out(3, st);
    SG_CUDA(cudaMemsetAsync(d_out.p, 0, 3 * sizeof(u32), st));
    record_check_kernel<<<pgrid(R, 1), 256, 0, st>>>(d_text.p, ls.p, R, lpr, marker, seq_len.p, d_out.p);
    SG_LAUNCHED();
    exclusive_scan_u32(seq_len.p, seq_off.p, R, d_out.p + 2, st);
    u32 h_out[3];
    SG_CUDA(cudaMemcpyAsync(h_out, d_out.p, sizeof(h_out), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    if (h_out[1]) return false;
    const u64 used = h_out[0], nb = h_out[2];
    if (final && R < max_records && used != n) return false;      // a truncated last record: let the sequential parser decide
    SG_CHECK(c.up_reads + R < 0x3FFFFFFFull, "at most 2^30-1 reads per context");
    grow_persistent(c.up_d_bases, (size_t)c.up_bases, (size_t)(c.up_bases + nb), st);
    grow_persistent(c.up_d_offsets, (size_t)(c.up_reads ? c.up_reads + 1 : 0), (size_t)(c.up_reads + R + 1), st);
    unsigned g = grid_for(R, 8, 1);
    if (g > kSMs * 16u) g = kSMs * 16u;
    record_copy_kernel<<<g, 256, 0, st>>>(d_text.p, ls.p, seq_len.p, seq_off.p, R, lpr, c.up_bases, c.up_d_bases.p, c.up_d_offsets.p + c.up_reads);
    SG_LAUNCHED();
    SG_CUDA(cudaStreamSynchronize(st));              // the caller may overwrite `text` now
    c.up_reads += R;
    c.up_bases += nb;
    consumed = used > n_bytes ? n_bytes : used;
    n_records = R;
    return true;
}

// drop the uploaded reads [first, first + count)
__global__ void __launch_bounds__(256) shift_offsets_kernel(const int64_t *__restrict__ in, u64 n, int64_t delta, int64_t *__restrict__ out)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) out[i] = in[i] - delta;
}

void stage_remove_uploaded(Context &c, u64 first, u64 count)
{
    cudaStream_t st = c.stream;
    ArenaScope arena_scope(c.arena, st);
    SG_CHECK(first + count <= c.up_reads, "range outside the uploaded reads");
    if (count == 0) return;
    int64_t o[2];
    SG_CUDA(cudaMemcpyAsync(&o[0], c.up_d_offsets.p + first, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(&o[1], c.up_d_offsets.p + first + count, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    const u64 tail_reads = c.up_reads - (first + count), tail_bases = c.up_bases - (u64)o[1];
    if (tail_reads) {
        DevBuf<uint8_t> tb(tail_bases, st);
        DevBuf<int64_t> to(tail_reads + 1, st);
        SG_CUDA(cudaMemcpyAsync(tb.p, c.up_d_bases.p + o[1], tail_bases, cudaMemcpyDeviceToDevice, st));
        SG_CUDA(cudaMemcpyAsync(c.up_d_bases.p + o[0], tb.p, tail_bases, cudaMemcpyDeviceToDevice, st));
        shift_offsets_kernel<<<pgrid(tail_reads + 1), 256, 0, st>>>(c.up_d_offsets.p + first + count, tail_reads + 1, o[1] - o[0], to.p);
        SG_LAUNCHED();
        SG_CUDA(cudaMemcpyAsync(c.up_d_offsets.p + first, to.p, (tail_reads + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
        SG_CUDA(cudaStreamSynchronize(st));
    }
    c.up_reads -= count;
    c.up_bases -= (u64)(o[1] - o[0]);
}

}  // namespace sg
