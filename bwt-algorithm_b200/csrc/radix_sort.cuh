// radix_sort.cuh -- LSD radix sort of (key, uint32 value) pairs, one global
// read + one global write per pass ("onesweep": per-tile digit offsets come from
// a decoupled look-back over tile status words, global digit bases from one
// up-front histogram of every pass).  Ranking inside a tile: the peer mask of
// equal digits comes from one ballot per digit bit (PTX, see refine_peers;
// match.any is bound by the address-divergence unit), four keys in flight for
// ILP, and one shared-memory atomic per distinct digit per warp step; scatters
// are staged through shared memory so global stores are contiguous per digit.
// Stable.  Keys are uint32_t or uint64_t.  The first pass can read its pairs
// from a generator (Source) instead of arrays, e.g. suffix keys straight from
// the bit-packed text, whose digit histograms are all one chunk histogram.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace bwtk {
namespace rsort {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
#ifndef BWTK_RS_THREADS
#define BWTK_RS_THREADS 512
#endif
#ifndef BWTK_RS_ITEMS
#define BWTK_RS_ITEMS 16
#endif
#ifndef BWTK_RS_WINDOW
#define BWTK_RS_WINDOW 4
#endif
#ifndef BWTK_RS_MINB32
#define BWTK_RS_MINB32 2
#endif
#ifndef BWTK_RS_MINB64
#define BWTK_RS_MINB64 1
#endif
constexpr int THREADS = BWTK_RS_THREADS;
constexpr int WARPS = THREADS / 32;
constexpr int ITEMS = BWTK_RS_ITEMS;
constexpr int TILE = THREADS * ITEMS;
constexpr int MAX_PASSES = 8;
constexpr int LOOKBACK_WINDOW = BWTK_RS_WINDOW;
static_assert(THREADS >= RADIX && RADIX % 32 == 0, "the first RADIX threads own one digit each");

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
    uint32_t *status;    // [MAX_PASSES][tiles][RADIX]
    unsigned *counters;  // [MAX_PASSES] dynamic tile ids
    int *err;            // look-back spin overflow flag
    int64_t max_tiles;
};

static inline int64_t tiles_for(int64_t n) { return ceil_div(n, TILE); }
static inline int64_t workspace_bytes(int64_t n)
{
    return align_up(MAX_PASSES * RADIX * 4, 256) +
           align_up(MAX_PASSES * (tiles_for(n) + 1) * RADIX * 4 + 256, 256) + 512;
}
static inline Workspace carve(Carver &c, int64_t n)
{
    Workspace w;
    w.ghist = c.take<uint32_t>(MAX_PASSES * RADIX);
    w.max_tiles = tiles_for(n) + 1;
    w.status = c.take<uint32_t>(MAX_PASSES * w.max_tiles * RADIX);
    w.counters = c.take<unsigned>(MAX_PASSES);
    w.err = c.take<int>(4);
    return w;
}

// ---- pair sources -------------------------------------------------------------
template <typename KeyT> struct ArraySource {
    const KeyT *k;
    const uint32_t *v;
    __device__ __forceinline__ void load(int64_t t, KeyT &key, uint32_t &val) const
    {
        key = k[t];
        val = v[t];
    }
};

// Suffix keys of round 0 (sa.cu): slot t holds suffix i = n-1-t (decreasing index
// order); key = its first S symbols, read as one window of the packed text.
struct PackedSuffixSource {
    const uint32_t *packed;
    int64_t n;
    int bits;
    int used;  // S * bits
    __device__ __forceinline__ void load(int64_t t, uint32_t &key, uint32_t &val) const
    {
        int64_t i = n - 1 - t;
        uint32_t w = window32(packed, i * bits);
        key = used == 32 ? w : (w >> (32 - used));
        val = (uint32_t)i;
    }
};

template <typename KeyT, typename Source>
__global__ void __launch_bounds__(512) hist_kernel(Source src, int64_t n, const unsigned *__restrict__ d_n, Plan plan,
                                                   uint32_t *__restrict__ ghist)
{
    __shared__ uint32_t s_hist[MAX_PASSES * RADIX];
    if (d_n) n = *d_n;   // device-resident count (<= the host bound the grid was sized for)
    for (int i = threadIdx.x; i < plan.passes * RADIX; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // per-thread run-length cache: high digits of nearly-sorted keys repeat, and
    // same-address shared atomics from a whole warp would serialise.
    uint32_t last_d[MAX_PASSES], run[MAX_PASSES];
#pragma unroll
    for (int p = 0; p < MAX_PASSES; p++) { last_d[p] = 0; run[p] = 0; }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        KeyT k;
        uint32_t v;
        src.load(i, k, v);
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

// Histogram of the 8-bit chunks that start at every symbol of a bit-packed text (symbol i
// at stream bit i*bits): the top digit of every round-0 suffix key.  One 32-bit word (32/bits
// symbol starts) per thread step, per-warp shared histograms.
static __global__ void __launch_bounds__(512)
    chunk_hist_kernel(const uint32_t *__restrict__ packed, int64_t n, int bits, uint32_t *__restrict__ out)
{
    __shared__ uint32_t s_h[16][RADIX];
    for (int i = threadIdx.x; i < 16 * RADIX; i += blockDim.x) (&s_h[0][0])[i] = 0;
    __syncthreads();
    const int per = 32 / bits;
    const int64_t nwords = (n + per - 1) / per;
    uint32_t *mine = s_h[threadIdx.x >> 5];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += stride) {
        const uint32_t a = __ldg(packed + w), b = __ldg(packed + w + 1);
        const int64_t i0 = w * per;
        const int cnt = (n - i0) < per ? (int)(n - i0) : per;
        for (int q = 0; q < cnt; q++) atomicAdd(&mine[__funnelshift_l(b, a, q * bits) >> 24], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RADIX; i += blockDim.x) {
        uint32_t v = 0;
#pragma unroll
        for (int w = 0; w < 16; w++) v += s_h[w][i];
        if (v) atomicAdd(&out[i], v);
    }
}

// Digit p of key(i) is the chunk at symbol i + k_p, k_p = (24 - 8p)/bits, and chunks that
// start at or after the end of the text are zero: hist_p = H - (chunks of the first k_p
// symbols) + k_p * [v == 0].  `h0` (the chunk histogram) may alias no row of `ghist` in use.
static __global__ void shift_hist_kernel(const uint32_t *__restrict__ packed, int64_t n, int bits,
                                         const uint32_t *__restrict__ h0, uint32_t *__restrict__ ghist)
{
    const uint32_t v = threadIdx.x;
    const uint32_t base = h0[v];
    for (int p = 0; p < 4; p++) {
        const int k = (24 - 8 * p) / bits;
        uint32_t c = base;
        for (int i = 0; i < k && i < n; i++) {
            if ((window32(packed, (int64_t)i * bits) >> 24) == v) c--;   // symbol i leaves the range ...
            if (v == 0) c++;                                              // ... and slot n + i (all zero) enters
        }
        ghist[p * RADIX + v] = c;
    }
}

static inline int64_t packed_hist_bytes(int64_t n, int bits) { return n * bits / 8 + 1; }

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

// One step of the peer-mask refinement: keeps in `pm` the lanes whose digit agrees with
// this lane's in bit B.  Written in PTX so that the bit test feeds both the ballot and the
// mask selection (4 instructions per bit; the C++ form compiled to 6-9).
template <int B> __device__ __forceinline__ unsigned refine_peers(unsigned pm, uint32_t d)
{
    unsigned out;
    asm("{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 t, bm;\n\t"
        "and.b32 t, %2, %3;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 bm, p, 0xffffffff;\n\t"
        "@!p not.b32 bm, bm;\n\t"
        "and.b32 %0, %1, bm;\n\t"
        "}"
        : "=r"(out)
        : "r"(pm), "r"(d), "n"(1u << B));
    return out;
}

// lanes of the warp whose 8-bit digit equals this lane's (all 32 lanes must call)
__device__ __forceinline__ unsigned peers_of_digit(uint32_t d)
{
    unsigned pm = 0xffffffffu;
    pm = refine_peers<0>(pm, d);
    pm = refine_peers<1>(pm, d);
    pm = refine_peers<2>(pm, d);
    pm = refine_peers<3>(pm, d);
    pm = refine_peers<4>(pm, d);
    pm = refine_peers<5>(pm, d);
    pm = refine_peers<6>(pm, d);
    pm = refine_peers<7>(pm, d);
    return pm;
}

template <typename KeyT, typename Source, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
    onesweep_kernel(Source src, KeyT *__restrict__ kout, uint32_t *__restrict__ vout, int64_t n,
                    const unsigned *__restrict__ d_n, int shift, int nbits,
                    const uint32_t *__restrict__ gbase, uint32_t *status, unsigned *tile_counter, int *err)
{
    if (d_n) n = *d_n;
    const uint32_t digit_mask = (1u << nbits) - 1u;
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
    if (base >= n) return;   // grid sized for an upper bound of the device-resident count
    const int tile_n = (int)((n - base) < TILE ? (n - base) : TILE);
    const int warp_base = warp * (ITEMS * 32) + lane;

    KeyT key[ITEMS];
    uint32_t val[ITEMS];
    uint32_t rpos[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = warp_base + k * 32;
        key[k] = (KeyT)0;
        val[k] = 0u;
        if (idx < tile_n) src.load(base + idx, key[k], val[k]);
    }
    uint32_t *my_hist = s_whist + warp * RADIX;
    const unsigned lt = lanemask_lt();
    // rank of every key among the keys of its warp with the same digit, in index order
#pragma unroll
    for (int k0 = 0; k0 < ITEMS; k0 += 4) {
        uint32_t d[4];
        unsigned m[4];
        // Slots past the end of the last tile take the top digit: they follow every real key
        // in index order, so they rank behind the real keys of that digit and only inflate its
        // count, which is corrected below.
#pragma unroll
        for (int j = 0; j < 4; j++) {
            bool valid = (warp_base + (k0 + j) * 32) < tile_n;
            d[j] = valid ? ((uint32_t)(key[k0 + j] >> shift) & digit_mask) : (uint32_t)(RADIX - 1);
        }
        // peer mask of equal digits from one ballot per digit bit (MATCH.ANY runs on the
        // address-divergence unit at ~70 cycles per warp and was the kernel's bottleneck);
        // bits above nbits are zero in every lane and leave the mask unchanged
#pragma unroll
        for (int j = 0; j < 4; j++) {
            m[j] = peers_of_digit(d[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int leader = __ffs(m[j]) - 1;
            uint32_t before = 0;
            if (lane == leader) before = atomicAdd(&my_hist[d[j]], (uint32_t)__popc(m[j]));
            before = __shfl_sync(0xffffffffu, before, leader);
            rpos[k0 + j] = before + __popc(m[j] & lt);
        }
    }
    __syncthreads();

    // thread d (< RADIX): exclusive scan over warps of digit d, tile count of digit d
    uint32_t cnt = 0, incl = 0;
    if (tid < RADIX) {
#pragma unroll
        for (int w = 0; w < WARPS; w++) {
            uint32_t t = s_whist[w * RADIX + tid];
            s_whist[w * RADIX + tid] = cnt;
            cnt += t;
        }
        if (tid == RADIX - 1) cnt -= (uint32_t)(TILE - tile_n);   // the padding slots of the last tile
        // exclusive scan of cnt over the 256 digits
        incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[warp] = incl;
    }
    __syncthreads();
    if (tid < RADIX) {
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < RADIX / 32; w++)
            if (w < warp) woff += s_wsum[w];
        const uint32_t dstart = woff + incl - cnt;

        // decoupled look-back for digit `tid`, LOOKBACK_WINDOW predecessors per round trip
        uint32_t excl = 0;
        volatile uint32_t *st = status;
        if (tile == 0) {
            st[tid] = FLAG_INCL | cnt;
        } else {
            st[tile * RADIX + tid] = FLAG_AGG | cnt;
            int64_t t = tile - 1;
            int spins = 0;
            bool done = false;
            while (!done) {
                uint32_t s[LOOKBACK_WINDOW];
#pragma unroll
                for (int j = 0; j < LOOKBACK_WINDOW; j++)
                    s[j] = (t - j >= 0) ? (uint32_t)st[(t - j) * RADIX + tid] : (uint32_t)(2u << 30);  // before tile 0: inclusive 0
                int used = LOOKBACK_WINDOW;
#pragma unroll
                for (int j = 0; j < LOOKBACK_WINDOW; j++) {
                    if (done || used != LOOKBACK_WINDOW) continue;
                    if ((s[j] & FLAG_MASK) == 0u) { used = j; continue; }   // not published yet: poll again from here
                    excl += s[j] & VAL_MASK;
                    if (s[j] & FLAG_INCL) done = true;
                }
                if (!done) {
                    t -= used;
                    if (used != LOOKBACK_WINDOW) {
                        if (++spins > SPIN_LIMIT) { *err = 1; done = true; }
                        __nanosleep(20);
                    }
                }
            }
            st[tile * RADIX + tid] = FLAG_INCL | (excl + cnt);
        }
        s_dstart[tid] = dstart;
        s_gbase[tid] = gbase[tid] + excl - dstart;
    }
    __syncthreads();

#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        int idx = warp_base + k * 32;
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
template <typename KeyT> constexpr int min_blocks() { return sizeof(KeyT) == 4 ? BWTK_RS_MINB32 : BWTK_RS_MINB64; }

template <typename KeyT, typename Source>
static int launch_pass(Source src, KeyT *kout, uint32_t *vout, int64_t n, const unsigned *d_n, const Plan &plan,
                       int p, const Workspace &ws, int64_t tiles, cudaStream_t st)
{
    auto kern = onesweep_kernel<KeyT, Source, min_blocks<KeyT>()>;
    static bool attr_set = false;
    if (!attr_set) {
        BWTK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)onesweep_smem<KeyT>()));
        attr_set = true;
    }
    // algorithmic bytes of the launch: every pair written once; read once from arrays, or -- the generator
    // pass of round 0 -- only the bit-packed text the keys are windows of
    int64_t read_bytes = n * (int64_t)(sizeof(KeyT) + 4);
    if constexpr (std::is_same<Source, PackedSuffixSource>::value) read_bytes = packed_hist_bytes(n, src.bits);
    prof::Scope ps(sizeof(KeyT) == 4 ? (std::is_same<Source, PackedSuffixSource>::value ? "onesweep_u32_gen" : "onesweep_u32")
                                     : "onesweep_u64",
                   read_bytes + n * (int64_t)(sizeof(KeyT) + 4), st);
    kern<<<(unsigned)tiles, THREADS, onesweep_smem<KeyT>(), st>>>(
        src, kout, vout, n, d_n, plan.shift[p], plan.bits[p], ws.ghist + p * RADIX,
        ws.status + (int64_t)p * ws.max_tiles * RADIX, ws.counters + p, ws.err);
    return BWTK_OK;
}

// Sorts n pairs by key bits [bit_lo, bit_hi).  The first pass reads from `first`
// (a Source); later passes ping-pong between (k0,v0) and (k1,v1), starting by
// writing (k0,v0) when `first` is a generator, or (k1,v1) when it wraps (k0,v0).
// *in_first is set to 1 when the result is in (k0,v0).  When d_n is not null the
// element count is read from device memory (n is then the bound the grids are
// sized for) and the pass count / ping-pong parity stay those of the plan.
template <typename KeyT, typename Source>
int sort_pairs_from(Source first, bool first_is_k0, KeyT *k0, uint32_t *v0, KeyT *k1, uint32_t *v1, int64_t n,
                    int bit_lo, int bit_hi, const Workspace &ws, cudaStream_t st, int *in_first,
                    int64_t *passes_out, const unsigned *d_n = nullptr)
{
    *in_first = 1;
    if (n <= 1 && first_is_k0 && !d_n) return BWTK_OK;
    if (n < 1) n = 1;
    Plan plan = make_plan(bit_lo, bit_hi);
    int64_t tiles = tiles_for(n);
    if (tiles > ws.max_tiles) {
        set_error("radix sort workspace too small (%lld tiles > %lld)", (long long)tiles, (long long)ws.max_tiles);
        return BWTK_EWORKSPACE;
    }
    BWTK_CUDA(bwtk::zero_async(ws.ghist, MAX_PASSES * RADIX * 4, st));
    BWTK_CUDA(bwtk::zero_async(ws.counters, MAX_PASSES * sizeof(unsigned), st));
    // status words of pass p live at ws.status + p*max_tiles*RADIX; only `tiles` of them are used
    BWTK_CUDA(bwtk::zero_async(ws.status, (size_t)((plan.passes - 1) * ws.max_tiles + tiles) * RADIX * 4, st));
    int hgrid = (int)(ceil_div(n, 512 * 16) < NUM_SMS * 4 ? ceil_div(n, 512 * 16) : NUM_SMS * 4);
    if (hgrid < 1) hgrid = 1;
    bool hist_done = false;
    if constexpr (std::is_same<Source, PackedSuffixSource>::value) {
        if (!d_n && bit_lo == 0 && bit_hi == 32 && first.used == 32 && plan.passes == 4) {
            // suffix keys are windows of one bit stream: one chunk histogram serves all four digits
            prof::Scope ps("radix_hist_u32", packed_hist_bytes(n, first.bits), st);
            int cgrid = (int)(ceil_div(n, 512 * 64) < NUM_SMS * 4 ? ceil_div(n, 512 * 64) : NUM_SMS * 4);
            chunk_hist_kernel<<<cgrid < 1 ? 1 : cgrid, 512, 0, st>>>(first.packed, n, first.bits,
                                                                    ws.ghist + (MAX_PASSES - 1) * RADIX);
            BWTK_LAUNCH_CHECK();
            shift_hist_kernel<<<1, RADIX, 0, st>>>(first.packed, n, first.bits, ws.ghist + (MAX_PASSES - 1) * RADIX,
                                                  ws.ghist);
            hist_done = true;
        }
    }
    if (!hist_done) {
        prof::Scope ps(sizeof(KeyT) == 4 ? "radix_hist_u32" : "radix_hist_u64", n * (int64_t)sizeof(KeyT), st);
        hist_kernel<KeyT, Source><<<hgrid, 512, 0, st>>>(first, n, d_n, plan, ws.ghist);
    }
    BWTK_LAUNCH_CHECK();
    scan_hist_kernel<<<plan.passes, RADIX, 0, st>>>(ws.ghist);
    BWTK_LAUNCH_CHECK();
    KeyT *kout = first_is_k0 ? k1 : k0;
    uint32_t *vout = first_is_k0 ? v1 : v0;
    int rc = launch_pass<KeyT, Source>(first, kout, vout, n, d_n, plan, 0, ws, tiles, st);
    if (rc) return rc;
    BWTK_LAUNCH_CHECK();
    for (int p = 1; p < plan.passes; p++) {
        ArraySource<KeyT> a{kout, vout};
        KeyT *nk = kout == k0 ? k1 : k0;
        uint32_t *nv = vout == v0 ? v1 : v0;
        rc = launch_pass<KeyT, ArraySource<KeyT>>(a, nk, nv, n, d_n, plan, p, ws, tiles, st);
        if (rc) return rc;
        BWTK_LAUNCH_CHECK();
        kout = nk;
        vout = nv;
    }
    *in_first = (kout == k0) ? 1 : 0;
    if (passes_out) *passes_out += plan.passes;
    return BWTK_OK;
}

template <typename KeyT>
int sort_pairs(KeyT *k0, uint32_t *v0, KeyT *k1, uint32_t *v1, int64_t n, int bit_lo, int bit_hi,
               const Workspace &ws, cudaStream_t st, int *in_first, int64_t *passes_out,
               const unsigned *d_n = nullptr)
{
    ArraySource<KeyT> a{k0, v0};
    return sort_pairs_from<KeyT, ArraySource<KeyT>>(a, true, k0, v0, k1, v1, n, bit_lo, bit_hi, ws, st, in_first,
                                                    passes_out, d_n);
}

}  // namespace rsort
}  // namespace bwtk
