// device_utils.cuh -- error handling, stream-ordered allocation, exclusive scan and LSD radix sort.
// Hand-written for sm_100a (no CUB/Thrust on the product path).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdexcept>
#include <string>

namespace sg {

typedef uint64_t u64;
typedef uint32_t u32;

struct CudaError : public std::runtime_error {
    explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};

#define SG_CUDA(call)                                                                           \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            throw sg::CudaError(std::string(#call) + " failed: " + cudaGetErrorString(e__) +    \
                                " at " + __FILE__ + ":" + std::to_string(__LINE__));            \
    } while (0)

#define SG_CHECK(cond, msg)                                                                     \
    do {                                                                                        \
        if (!(cond)) throw sg::CudaError(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" + \
                                         std::to_string(__LINE__));                             \
    } while (0)

// Stream-ordered device buffer (cudaMallocAsync pool: after warm-up an allocation is a pointer bump).
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0, cap = 0;      // n = elements in use, cap = elements allocated
    cudaStream_t s = nullptr;
    DevBuf() {}
    DevBuf(size_t count, cudaStream_t st) { alloc(count, st); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), cap(o.cap), s(o.s) { o.p = nullptr; o.n = 0; o.cap = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if (this != &o) { release(); p = o.p; n = o.n; cap = o.cap; s = o.s; o.p = nullptr; o.n = 0; o.cap = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    // Grow-only: a buffer that is already large enough is kept (the long-lived buffers of a context are
    // re-used run after run instead of cycling hundreds of MB through the pool).  Contents are undefined.
    void alloc(size_t count, cudaStream_t st)
    {
        if (p && cap >= count && s == st) { n = count; return; }
        release();
        s = st; n = count; cap = count ? count : 1;
        SG_CUDA(cudaMallocAsync((void **)&p, cap * sizeof(T), st));
    }
    void release()
    {
        if (p) { cudaFreeAsync(p, s); p = nullptr; n = 0; cap = 0; }
    }
    size_t bytes() const { return n * sizeof(T); }
};

constexpr int kSMs = 148;   // B200

// every kernel launch of the library is followed by SG_LAUNCHED(): error check + launch count
// (bench.py reports the count as `gpu_launches`).
inline unsigned long long &launch_counter() { static unsigned long long n = 0; return n; }
#define SG_LAUNCHED()                     \
    do {                                  \
        ++sg::launch_counter();           \
        SG_CUDA(cudaGetLastError());      \
    } while (0)

inline unsigned grid_for(u64 n, unsigned block, unsigned per_thread = 1)
{
    u64 g = (n + (u64)block * per_thread - 1) / ((u64)block * per_thread);
    if (g == 0) g = 1;
    SG_CHECK(g < 0x7FFFFFFFull, "grid too large");
    return (unsigned)g;
}

// exclusive prefix sum of n u32 values (in may alias out); returns the grand total through
// d_total (device pointer, may be null).  Sums must fit in 32 bits.
void exclusive_scan_u32(const u32 *in, u32 *out, u64 n, u32 *d_total, cudaStream_t st);

// OR / AND of n 64-bit keys -> d_or_and[0], d_or_and[1] (device); used to skip constant digits.
void reduce_or_and_u64(const u64 *keys, u64 n, u64 *d_or_and, cudaStream_t st);

// Stable LSD radix sort of n records by bits [lo,hi) of keyA (use_b=false) or keyB (use_b=true).
// A record is (keyA[, keyB][, val]); null pointers mean "that column does not exist".  Buffers
// ping-pong between (a0,b0,v0) and (a1,b1,v1); returns 0 if the result is in set 0, else 1.
struct SortCols {
    u64 *a[2];
    u64 *b[2];
    u32 *v[2];
};
int radix_sort_bits(SortCols &c, int cur, u64 n, bool use_b, int lo, int hi, cudaStream_t st);

// Convenience: sort by all VARYING bits of the chosen key column (one host sync to read OR/AND).
int radix_sort_varying(SortCols &c, int cur, u64 n, bool use_b, cudaStream_t st);

// GB/s of uniformly random `granule_bytes` gathers over `footprint_bytes` of HBM (best of 3 after warm-up).
float gather_bench(u64 footprint_bytes, int granule_bytes, u64 n_loads, int mode, cudaStream_t st);

}  // namespace sg
