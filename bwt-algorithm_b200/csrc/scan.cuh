// scan.cuh -- single-pass exclusive prefix sum with a per-element callback
// ("scan + compact" skeleton).  Tiles take dynamic ids (atomic counter) and
// chain through a decoupled look-back over 64-bit status words, so input and
// output are each touched once.
//
//   Count  : __device__ uint64_t operator()(int64_t i)            (items at i)
//   Emit   : __device__ void operator()(int64_t i, uint64_t excl, uint64_t cnt)
// Counts are summed in 62 bits, so two 31-bit counters may be packed in one value.
// A Count that declares a nested type `Ctx` gets one default-constructed Ctx per thread,
// shared by its ITEMS consecutive elements and by both phases:
//   Count  : operator()(int64_t i, int k, Ctx &)      Emit : operator()(i, k, excl, cnt, Ctx &)
// (k = 0..ITEMS-1, the element's slot in the thread), so that what the count phase computed
// can be reused by the next element and by the emit phase.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace bwtk {
namespace scan {

constexpr int THREADS = 256;
constexpr int ITEMS = 8;
constexpr int TILE = THREADS * ITEMS;
constexpr unsigned long long AGG = 1ull << 62;
constexpr unsigned long long INCL = 2ull << 62;
constexpr unsigned long long FLAGS = 3ull << 62;
constexpr int SPIN_LIMIT = 1 << 22;

struct Workspace {
    unsigned long long *status;  // [tiles]
    unsigned *counter;           // dynamic tile id
    unsigned long long *total;   // grand total
    int *err;
};

static inline int64_t tiles_for(int64_t n) { return ceil_div(n, TILE); }
static inline int64_t workspace_bytes(int64_t n) { return align_up(tiles_for(n) * 8 + 64, 256) + 1024; }
static inline Workspace carve(Carver &c, int64_t n)
{
    Workspace w;
    w.status = c.take<unsigned long long>(tiles_for(n) + 8);
    w.counter = c.take<unsigned>(4);
    w.total = c.take<unsigned long long>(2);
    w.err = c.take<int>(4);
    return w;
}

struct NoCtx {};
template <typename F, typename = void> struct CtxOf {
    using type = NoCtx;
    static constexpr bool has = false;
};
template <typename F> struct CtxOf<F, std::void_t<typename F::Ctx>> {
    using type = typename F::Ctx;
    static constexpr bool has = true;
};

struct SumComb {
    __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const
    {
        return a + b;
    }
};

template <typename Count, typename Emit>
__global__ void __launch_bounds__(THREADS)
    scan_kernel(int64_t n, Count count, Emit emit, unsigned long long *status, unsigned *counter,
                unsigned long long *total, int *err)
{
    __shared__ unsigned s_tile;
    __shared__ unsigned long long s_warp[THREADS / 32];
    __shared__ unsigned long long s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(counter, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t i0 = tile * TILE + (int64_t)tid * ITEMS;
    unsigned long long c[ITEMS];
    unsigned long long sum = 0;
    typename CtxOf<Count>::type ctx;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int64_t i = i0 + k;
        c[k] = 0ull;
        if (i < n) {
            if constexpr (CtxOf<Count>::has) c[k] = (unsigned long long)count(i, k, ctx);
            else c[k] = (unsigned long long)count(i);
        }
        sum += c[k];
    }
    unsigned long long inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned long long wpre = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++)
        if (w < warp) wpre += s_warp[w];
    if (warp == THREADS / 32 - 1) {
        const unsigned long long tile_agg = __shfl_sync(0xffffffffu, wpre + inc, 31);
        volatile unsigned long long *st = status;
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 31) st[0] = INCL | tile_agg;
        } else {
            if (lane == 31) st[tile] = AGG | tile_agg;
            excl = warp_lookback(st, tile, SumComb{}, err, 3, SPIN_LIMIT);
            if (lane == 31) st[tile] = INCL | (excl + tile_agg);
        }
        if (lane == 31) {
            s_prefix = excl;
            if ((tile + 1) * (int64_t)TILE >= n) *total = excl + tile_agg;
        }
    }
    __syncthreads();
    unsigned long long run = s_prefix + wpre + inc - sum;
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int64_t i = i0 + k;
        if (i < n) {
            if constexpr (CtxOf<Count>::has) emit(i, k, run, c[k], ctx);
            else emit(i, run, c[k]);
        }
        run += c[k];
    }
}

// Launches the scan over [0, n).  The grand total lands in ws.total[0] (device).
template <typename Count, typename Emit>
int run(int64_t n, Count count, Emit emit, const Workspace &ws, cudaStream_t st)
{
    BWTK_CUDA(bwtk::zero_async(ws.total, 8, st));
    if (n <= 0) return BWTK_OK;
    int64_t tiles = tiles_for(n);
    BWTK_CUDA(bwtk::zero_async(ws.status, (size_t)tiles * 8, st));
    BWTK_CUDA(bwtk::zero_async(ws.counter, sizeof(unsigned), st));
    scan_kernel<Count, Emit><<<(unsigned)tiles, THREADS, 0, st>>>(n, count, emit, ws.status, ws.counter,
                                                                 ws.total, ws.err);
    BWTK_LAUNCH_CHECK();
    return BWTK_OK;
}

}  // namespace scan
}  // namespace bwtk
