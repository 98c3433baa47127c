// common.cuh -- shared helpers for libbwtk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bwtk.h"

namespace bwtk {

void set_error(const char *fmt, ...);
void count_launch(int k = 1);
// Zero-fill by a kernel.  cudaMemsetAsync may be served by a copy engine, where it queues behind the
// bulk downloads of other streams (streaming.IndexPipeline) and stalls the build for milliseconds.
cudaError_t zero_async(void *p, size_t bytes, cudaStream_t st);
// dst <- bytes at d_src via mapped pinned memory (no copy engine); synchronises st
int read_back(void *dst, const void *d_src, size_t bytes, cudaStream_t st);

#define BWTK_CUDA(expr)                                                              \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            bwtk::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,            \
                            cudaGetErrorString(_e));                                 \
            return BWTK_ECUDA;                                                       \
        }                                                                            \
    } while (0)

#define BWTK_LAUNCH_CHECK()                                                          \
    do {                                                                             \
        bwtk::count_launch();                                                        \
        cudaError_t _e = cudaGetLastError();                                         \
        if (_e != cudaSuccess) {                                                     \
            bwtk::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,        \
                            cudaGetErrorString(_e));                                 \
            return BWTK_ECUDA;                                                       \
        }                                                                            \
    } while (0)

#define BWTK_REQUIRE(cond, msg)                                                      \
    do {                                                                             \
        if (!(cond)) {                                                               \
            bwtk::set_error("%s:%d: %s", __FILE__, __LINE__, msg);                   \
            return BWTK_EINVAL;                                                      \
        }                                                                            \
    } while (0)

// Optional per-kernel timing (bwtk_profile_enable): CUDA events recorded on the
// launching stream around a named launch, with its algorithmic byte count.
namespace prof {
bool enabled();
void begin(const char *name, int64_t algo_bytes, cudaStream_t st);
void end(cudaStream_t st);
struct Scope {
    cudaStream_t st;
    bool on;
    Scope(const char *name, int64_t algo_bytes, cudaStream_t s) : st(s), on(enabled())
    {
        if (on) begin(name, algo_bytes, st);
    }
    ~Scope()
    {
        if (on) end(st);
    }
};
}  // namespace prof

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Carves aligned sub-buffers out of a caller-provided workspace.
struct Carver {
    char *base;
    int64_t off, cap;
    Carver(void *p, int64_t bytes) : base((char *)p), off(0), cap(bytes) {}
    template <typename T> T *take(int64_t count)
    {
        off = align_up(off, 256);
        T *r = (T *)(base + off);
        off += count * (int64_t)sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

static const int NUM_SMS = 148;  // B200

// ---- bit-packed text ---------------------------------------------------------
// Symbols are stored with `bits` bits each, MSB first, in a stream of 32-bit
// words: symbol i occupies stream bits [i*bits, (i+1)*bits).  Positions past the
// end read as zero.  The stream is padded with 3 zero words.
__device__ __forceinline__ uint32_t window32(const uint32_t *__restrict__ p, int64_t bitoff)
{
    int64_t w = bitoff >> 5;
    uint32_t r = (uint32_t)bitoff & 31u;
    uint32_t a = __ldg(p + w), b = __ldg(p + w + 1);
    return __funnelshift_l(b, a, r);
}

__device__ __forceinline__ uint64_t window64(const uint32_t *__restrict__ p, int64_t bitoff)
{
    int64_t w = bitoff >> 5;
    uint32_t r = (uint32_t)bitoff & 31u;
    uint32_t a = __ldg(p + w), b = __ldg(p + w + 1), c = __ldg(p + w + 2);
    uint32_t hi = __funnelshift_l(b, a, r), lo = __funnelshift_l(c, b, r);
    return ((uint64_t)hi << 32) | lo;
}

// Warp-parallel decoupled look-back over 64-bit tile status words
// ([63:62] flag: 1 = tile aggregate, 2 = inclusive prefix; [61:0] payload).
// All 32 lanes of one warp call it for tile > 0; 32 predecessors are inspected per
// round trip.  `comb` must be associative and commutative with identity 0.
// Returns the exclusive prefix of `tile`.
template <typename Comb>
__device__ __forceinline__ unsigned long long warp_lookback(volatile unsigned long long *st, int64_t tile, Comb comb,
                                                            int *err, int err_code, int spin_limit)
{
    const unsigned long long FL = 3ull << 62, INCL = 2ull << 62;
    const unsigned lane = threadIdx.x & 31u;
    unsigned long long excl = 0;
    int64_t t = tile - 1;
    int spins = 0;
    while (true) {
        int64_t mine = t - (int64_t)lane;
        unsigned long long s = mine >= 0 ? st[mine] : INCL;   // before tile 0: inclusive identity
        unsigned ready = __ballot_sync(0xffffffffu, (s & FL) != 0ull);
        unsigned incl = __ballot_sync(0xffffffffu, (s & FL) == INCL);
        int fi = incl ? (__ffs(incl) - 1) : 32;
        unsigned need = fi < 31 ? ((2u << fi) - 1u) : 0xffffffffu;
        if ((ready & need) != need) {
            if (++spins > spin_limit) { if (lane == 0) *err = err_code; return excl; }
            __nanosleep(20);
            continue;
        }
        unsigned long long v = ((need >> lane) & 1u) ? (s & ~FL) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = comb(v, __shfl_xor_sync(0xffffffffu, v, o));
        excl = comb(excl, v);
        if (fi < 32) return excl;
        t -= 32;
    }
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

}  // namespace bwtk
