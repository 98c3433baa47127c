// radix_sort.cuh -- LSD radix sort of (key, uint32 value) pairs, one global
// read + one global write per pass ("onesweep": per-tile digit offsets come from
// a decoupled look-back over tile status words, global digit bases from one
// up-front histogram of every pass).  Ranking inside a tile uses warp-level
// match masks; scatters are staged through shared memory so global stores are
// contiguous per digit.  Stable.  Keys are uint32_t or uint64_t.
#pragma once
#include "common.cuh"

namespace bwtk {
namespace rsort {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int ITEMS = 16;
constexpr int TILE = THREADS * ITEMS;
constexpr int MAX_PASSES = 8;
static_assert(THREADS == RADIX, "one thread per digit in the look-back");

constexpr uint32_t FLAG_AGG = 1u << 30;
constexpr uint32_t FLAG_INCL = 2u << 30;
constexpr uint32_t FLAG_MASK = 3u << 30;
constexpr uint32_t VAL_MASK = ~FLAG_MASK;
constexpr int SPIN_LIMIT = 1 << 22;

struct Plan {
    int passes;
    int shift[MAX_PASSES];
    int bits[MAX_PASSES];
};

// Split [bit_lo, bit_hi) into the fewest passes of <= 8 bits, balanced.
static inline Plan make_plan(int bit_lo, int bit_hi)
{
    Plan p;
    int total = bit_hi - bit_lo;
    if (total < 1) total = 1;
    p.passes = (total + RADIX_BITS - 1) / RADIX_BITS;
    int pos = bit_lo, rem = total;
    for (int i = 0; i < p.passes; i++) {
        int b = (rem + (p.passes - i) - 1) / (p.passes - i);
        p.shift[i] = pos;
        p.bits[i] = b;
        pos += b;
        rem -= b;
    }
    return p;
}

struct Workspace {
    uint32_t *ghist;     // [MAX_PASSES][RADIX]
    uint32_t *status;    // [tiles][RADIX]
    unsigned *counters;  // [MAX_PASSES] dynamic tile ids
    int *err;            // look-back spin overflow flag
    int64_t max_tiles;
};

static inline int64_t tiles_for(int64_t n) { return ceil_div(n, TILE); }
static inline int64_t workspace_bytes(int64_t n)
{
    return align_up(MAX_PASSES * RADIX * 4, 256) + align_up(tiles_for(n) * RADIX * 4 + 256, 256) +
           512;
}
static inline Workspace carve(Carver &c, int64_t n)
{
    Workspace w;
    w.ghist = c.take<uint32_t>(MAX_PASSES * RADIX);
    w.max_tiles = tiles_for(n) + 1;
    w.status = c.take<uint32_t>(w.max_tiles * RADIX);
    w.counters = c.take<unsigned>(MAX_PASSES);
    w.err = c.take<int>(4);
    return w;
}

template <typename KeyT>
__global__ void __launch_bounds__(512) hist_kernel(const KeyT *__restrict__ keys, int64_t n, Plan plan,
                                                   uint32_t *__restrict__ ghist)
{
    __shared__ uint32_t s_hist[MAX_PASSES * RADIX];
    for (int i = threadIdx.x; i < plan.passes * RADIX; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // per-thread run-length cache: high digits of nearly-sorted keys repeat, and
    // same-address shared atomics from a whole warp would serialise.
    uint32_t last_d[MAX_PASSES], run[MAX_PASSES];
#pragma unroll
    for (int p = 0; p < MAX_PASSES; p++) { last_d[p] = 0; run[p] = 0; }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        KeyT k = keys[i];
#pragma unroll
        for (int p = 0; p < MAX_PASSES; p++) {
            if (p < plan.passes) {
                uint32_t d = (uint32_t)(k >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u);
                if (d == last_d[p]) {
                    run[p]++;
                } else {
                    if (run[p]) atomicAdd(&s_hist[p * RADIX + last_d[p]], run[p]);
                    last_d[p] = d;
                    run[p] = 1;
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < MAX_PASSES; p++)
        if (p < plan.passes && run[p]) atomicAdd(&s_hist[p * RADIX + last_d[p]], run[p]);
    __syncthreads();
    for (int i = threadIdx.x; i < plan.passes * RADIX; i += blockDim.x) {
        uint32_t v = s_hist[i];
        if (v) atomicAdd(&ghist[i], v);
    }
}

// exclusive scan of each pass's 256 bins, in place; one block per pass
static __global__ void scan_hist_kernel(uint32_t *ghist)
{
    __shared__ uint32_t s[RADIX];
    uint32_t *h = ghist + blockIdx.x * RADIX;
    s[threadIdx.x] = h[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int i = 0; i < RADIX; i++) {
            uint32_t v = s[i];
            s[i] = acc;
            acc += v;
        }
    }
    __syncthreads();
    h[threadIdx.x] = s[threadIdx.x];
}

template <typename KeyT>
__global__ void __launch_bounds__(THREADS)
    onesweep_kernel(const KeyT *__restrict__ kin, const uint32_t *__restrict__ vin,
                    KeyT *__restrict__ kout, uint32_t *__restrict__ vout, int64_t n, int shift,
                    uint32_t digit_mask, const uint32_t *__restrict__ gbase,
                    uint32_t *status, unsigned *tile_counter, int *err)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KeyT *s_keys = reinterpret_cast<KeyT *>(smem_raw);                            // TILE keys
    uint32_t *s_vals = reinterpret_cast<uint32_t *>(smem_raw + TILE * sizeof(KeyT));  // TILE values
    uint32_t *s_whist = s_vals + TILE;                                            // [WARPS][RADIX]
    uint32_t *s_gbase = s_whist + WARPS * RADIX;                                  // [RADIX]
    uint32_t *s_dstart = s_gbase + RADIX;                                         // [RADIX]
    uint32_t *s_wsum = s_dstart + RADIX;                                          // [WARPS]
    __shared__ unsigned s_tile;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < WARPS * RADIX; i += THREADS) s_whist[i] = 0;
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t base = tile * TILE;
    const int tile_n = (int)((n - base) < TILE ? (n - base) : TILE);

    KeyT key[ITEMS];
    uint32_t val[ITEMS];
    uint32_t rpos[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = warp * (ITEMS * 32) + k * 32 + lane;
        bool valid = idx < tile_n;
        key[k] = valid ? kin[base + idx] : (KeyT)0;
        val[k] = valid ? vin[base + idx] : 0u;
    }
    uint32_t *my_hist = s_whist + warp * RADIX;
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = warp * (ITEMS * 32) + k * 32 + lane;
        bool valid = idx < tile_n;
        uint32_t d = valid ? ((uint32_t)(key[k] >> shift) & digit_mask) : (uint32_t)RADIX;
        unsigned m = __match_any_sync(0xffffffffu, d);
        uint32_t prev = valid ? my_hist[d] : 0u;
        __syncwarp();
        if (valid && (lane == (__ffs(m) - 1))) my_hist[d] = prev + __popc(m);
        __syncwarp();
        rpos[k] = prev + __popc(m & lt);
    }
    __syncthreads();

    // thread d: exclusive scan over warps of digit d, tile count of digit d
    uint32_t cnt = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) {
        uint32_t t = s_whist[w * RADIX + tid];
        s_whist[w * RADIX + tid] = cnt;
        cnt += t;
    }
    // exclusive scan of cnt over the 256 digits
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < WARPS; w++)
        if (w < warp) woff += s_wsum[w];
    const uint32_t dstart = woff + incl - cnt;

    // decoupled look-back for digit `tid`
    uint32_t excl = 0;
    volatile uint32_t *st = status;
    if (tile == 0) {
        st[tid] = FLAG_INCL | cnt;
    } else {
        st[tile * RADIX + tid] = FLAG_AGG | cnt;
        int64_t t = tile - 1;
        while (true) {
            uint32_t s = st[t * RADIX + tid];
            int spins = 0;
            while ((s & FLAG_MASK) == 0u) {
                if (++spins > SPIN_LIMIT) {
                    *err = 1;
                    s = FLAG_INCL;
                    break;
                }
                __nanosleep(32);
                s = st[t * RADIX + tid];
            }
            excl += s & VAL_MASK;
            if (s & FLAG_INCL) break;
            t--;
        }
        st[tile * RADIX + tid] = FLAG_INCL | (excl + cnt);
    }
    s_dstart[tid] = dstart;
    s_gbase[tid] = gbase[tid] + excl - dstart;
    __syncthreads();

#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = warp * (ITEMS * 32) + k * 32 + lane;
        if (idx < tile_n) {
            uint32_t d = (uint32_t)(key[k] >> shift) & digit_mask;
            uint32_t p = s_dstart[d] + my_hist[d] + rpos[k];
            s_keys[p] = key[k];
            s_vals[p] = val[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = tid + k * THREADS;
        if (idx < tile_n) {
            KeyT kk = s_keys[idx];
            uint32_t d = (uint32_t)(kk >> shift) & digit_mask;
            uint32_t g = s_gbase[d] + (uint32_t)idx;
            kout[g] = kk;
            vout[g] = s_vals[idx];
        }
    }
}

template <typename KeyT> constexpr size_t onesweep_smem()
{
    return TILE * sizeof(KeyT) + TILE * 4 + WARPS * RADIX * 4 + RADIX * 4 * 2 + WARPS * 4 + 64;
}

// Sorts n pairs by key bits [bit_lo, bit_hi).  Buffers ping-pong between
// (k0,v0) and (k1,v1); *in_first is set to 1 when the result is in (k0,v0).
template <typename KeyT>
int sort_pairs(KeyT *k0, uint32_t *v0, KeyT *k1, uint32_t *v1, int64_t n, int bit_lo, int bit_hi,
               const Workspace &ws, cudaStream_t st, int *in_first, int64_t *passes_out)
{
    *in_first = 1;
    if (n <= 1) return BWTK_OK;
    static bool attr_set = false;
    if (!attr_set) {
        BWTK_CUDA(cudaFuncSetAttribute(onesweep_kernel<KeyT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)onesweep_smem<KeyT>()));
        attr_set = true;
    }
    Plan plan = make_plan(bit_lo, bit_hi);
    int64_t tiles = tiles_for(n);
    if (tiles > ws.max_tiles) {
        set_error("radix sort workspace too small (%lld tiles > %lld)", (long long)tiles,
                  (long long)ws.max_tiles);
        return BWTK_EWORKSPACE;
    }
    BWTK_CUDA(cudaMemsetAsync(ws.ghist, 0, MAX_PASSES * RADIX * 4, st));
    BWTK_CUDA(cudaMemsetAsync(ws.counters, 0, MAX_PASSES * sizeof(unsigned), st));
    int hgrid = (int)(ceil_div(n, 512 * 16) < NUM_SMS * 4 ? ceil_div(n, 512 * 16) : NUM_SMS * 4);
    { prof::Scope ps(sizeof(KeyT) == 4 ? "radix_hist_u32" : "radix_hist_u64", n * (int64_t)sizeof(KeyT), st);
    hist_kernel<KeyT><<<hgrid, 512, 0, st>>>(k0, n, plan, ws.ghist); }
    BWTK_LAUNCH_CHECK();
    scan_hist_kernel<<<plan.passes, RADIX, 0, st>>>(ws.ghist);
    BWTK_LAUNCH_CHECK();
    KeyT *kin = k0, *kout = k1;
    uint32_t *vin = v0, *vout = v1;
    for (int p = 0; p < plan.passes; p++) {
        BWTK_CUDA(cudaMemsetAsync(ws.status, 0, (size_t)tiles * RADIX * 4, st));
        { prof::Scope ps(sizeof(KeyT) == 4 ? "onesweep_u32" : "onesweep_u64", 2 * n * (int64_t)(sizeof(KeyT) + 4), st);
        onesweep_kernel<KeyT><<<(unsigned)tiles, THREADS, onesweep_smem<KeyT>(), st>>>(
            kin, vin, kout, vout, n, plan.shift[p], (1u << plan.bits[p]) - 1u,
            ws.ghist + p * RADIX, ws.status, ws.counters + p, ws.err); }
        BWTK_LAUNCH_CHECK();
        KeyT *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    *in_first = (kin == k0) ? 1 : 0;
    if (passes_out) *passes_out += plan.passes;
    return BWTK_OK;
}

}  // namespace rsort
}  // namespace bwtk
