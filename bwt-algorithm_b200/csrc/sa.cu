// sa.cu -- suffix-array construction (replaces BWTCore._build_suffix_array,
// reference bwt.py:212-264) by prefix doubling over a bit-packed text:
//
//   round 0 : 32-bit key = the first S symbols of every suffix (S = 32/bits,
//             16 bases for the ACGT$ fast path), read straight from the packed
//             text by the histogram and the first radix pass; (key, index)
//             pairs are fed in DECREASING index order, so that the stable sort
//             puts a suffix that runs off the end of the text before longer ones
//             with the same zero-padded key ("shorter suffix first", the
//             reference's key2 = -1 rule);
//   regroup : head flags -> group-head position (max-scan) -> rank[] scatter,
//             SA write, compaction of suffixes whose group is not yet a
//             singleton (one single-pass decoupled look-back scan);
//   round r : for the active suffixes only, every group is ordered by
//             rank[i+h]+1: an LSD radix sort of (group head, rank) while the
//             active list is long, a shared-memory segmented sort (whole groups
//             per CTA, see segsort_kernel) once it is short enough for the radix
//             passes to be launch-bound; then regroup; h doubles.  Stops when
//             nothing is active.
//
// The host reads the active count once per round (4 bytes) to size the grids and
// to pick the sort.  The result is the unique suffix array, so it is bit-identical
// to the reference's regardless of the refinement path taken.
#include <stdlib.h>

#include "bucket_sort.cuh"
#include "radix_sort.cuh"

namespace bwtk {

int64_t packed_words(int64_t n, int bits);

namespace sa {

#ifndef BWTK_RG_THREADS
#define BWTK_RG_THREADS 256
#endif
#ifndef BWTK_RG_ITEMS
#define BWTK_RG_ITEMS 16
#endif
constexpr int RG_THREADS = BWTK_RG_THREADS;
constexpr int RG_ITEMS = BWTK_RG_ITEMS;   // a multiple of 4 (16-byte vector loads)
constexpr int RG_TILE = RG_THREADS * RG_ITEMS;
constexpr int MAX_ROUNDS = 40;       // 2^40 symbols of common prefix: unreachable for n < 2^30

// 64-bit look-back status: [63:62] flag, [61:31] (max head position + 1), [30:0] active count
constexpr unsigned long long RG_AGG = 1ull << 62;
constexpr unsigned long long RG_INCL = 2ull << 62;
constexpr unsigned long long RG_FLAGS = 3ull << 62;

__device__ __forceinline__ unsigned long long rg_pack(uint32_t mx, uint32_t sum)
{
    return ((unsigned long long)mx << 31) | (unsigned long long)sum;
}
__device__ __forceinline__ uint32_t rg_max(unsigned long long v) { return (uint32_t)((v >> 31) & 0x7fffffffu); }
__device__ __forceinline__ uint32_t rg_sum(unsigned long long v) { return (uint32_t)(v & 0x7fffffffu); }
__device__ __forceinline__ unsigned long long rg_combine(unsigned long long a, unsigned long long b)
{
    uint32_t ma = rg_max(a), mb = rg_max(b);
    return rg_pack(ma > mb ? ma : mb, rg_sum(a) + rg_sum(b));
}

struct RgComb {
    __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const
    {
        return rg_combine(a, b);
    }
};

// Round-0 singletons never get their rank scattered (that scatter is 46.7 M random
// 4-byte writes for a chr21-sized contig, ~3 GB of DRAM traffic, while the doubling
// rounds only ever read the ranks of ~6 % of the suffixes).  Their rank is their
// SA position, recovered on demand: 16-bit prefix table -> binary search in the
// sorted round-0 keys -> step over the (<= 15) short suffixes of that key.
struct LazyRank {
    const uint32_t *abits;   // bit x set: suffix x was active after round 0, rank[x] is maintained
    const int32_t *rank;
    const uint32_t *skey;    // round-0 keys in sorted order
    const int32_t *ptab;     // [65537] first sorted position of every 16-bit key prefix
    const int32_t *sa;       // singleton positions of SA never change after round 0
    const uint32_t *packed;
    int bits;
    __device__ __forceinline__ int32_t operator()(int64_t x) const
    {
        if ((__ldg(abits + (x >> 5)) >> (x & 31)) & 1u) return __ldg(rank + x);
        uint32_t key = window32(packed, x * bits);
        int32_t lo = __ldg(ptab + (key >> 16)), hi = __ldg(ptab + (key >> 16) + 1);
        // lower bound of key in skey[lo, hi).  On a chromosome-sized contig the range is thousands of keys and
        // every probe of a plain binary search is its own DRAM sector: start from the interpolated place (the
        // low 16 key bits are close to uniform inside one prefix) and gallop out, so that all but the first
        // probes fall into the same two or three sectors.
        if (hi - lo > 32) {
            const int32_t p = lo + (int32_t)(((uint64_t)(key & 0xffffu) * (uint32_t)(hi - lo)) >> 16);
            int32_t step = 16;
            if (__ldg(skey + p) < key) {
                lo = p + 1;
                while (true) {
                    const int32_t q = lo + step;
                    if (q < hi && __ldg(skey + q) < key) { lo = q + 1; step <<= 1; }
                    else { hi = q < hi ? q : hi; break; }
                }
            } else {
                hi = p;
                while (true) {
                    const int32_t q = hi - step;
                    if (q <= lo) break;
                    if (__ldg(skey + q) >= key) { hi = q; step <<= 1; }
                    else { lo = q + 1; break; }
                }
            }
        }
        while (lo < hi) {
            int32_t mid = (lo + hi) >> 1;
            if (__ldg(skey + mid) < key) lo = mid + 1; else hi = mid;
        }
        while (__ldg(sa + lo) != (int32_t)x) lo++;
        return lo;
    }
};

__global__ void build_keys_kernel(const uint32_t *__restrict__ suf, const int32_t *__restrict__ grp, LazyRank lr,
                                  const unsigned *__restrict__ d_m, int64_t n, int64_t h, int kbits,
                                  uint64_t *__restrict__ key)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)*d_m) return;
    int64_t s = (int64_t)suf[t] + h;
    uint64_t k2 = s < n ? (uint64_t)(lr(s) + 1) : 0ull;
    key[t] = ((uint64_t)(uint32_t)grp[t] << kbits) | k2;
}

// ---- segmented sort of one doubling round -------------------------------------------
// The active list is grouped (elements of one group are adjacent, grp[] = the group's head
// SA position, increasing along the list), and a round only has to order every group by
// k2 = rank[suffix + h] + 1.  After round 0 almost all groups are tiny, so instead of
// seven global radix passes over (grp, k2) each CTA sorts a window of whole groups in
// shared memory with one bitonic network over (grp, k2, original slot):
//   * CTA b owns the groups whose first element lies in [b*SS_T, (b+1)*SS_T); its window
//     runs from its first group head to the end of the group that contains its last
//     nominal element (<= SS_W elements);
//   * if that last group makes the window overflow it is deferred: groups of up to
//     SS_MID_MAX elements go on a list sorted by midsort_kernel, one CTA per group;
//     larger ones raise ctl->huge and the host redoes this round with the radix path.
// Output: key_out = (grp << kbits) | k2 and suf_out in sorted order at the same list slots
// (what regroup_kernel consumes).
// k2[t] = rank[suffix + h] + 1 of every active element, one thread each: the lazy-rank lookup
// is a chain of dependent loads, hidden only by running all elements at once (inside the
// sort CTAs, 8 lookups per thread in sequence cost more than the sort itself)
__global__ void build_k2_kernel(const uint32_t *__restrict__ suf, LazyRank lr, const unsigned *__restrict__ d_m,
                                int64_t n, int64_t h, uint32_t *__restrict__ k2)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)*d_m) return;
    int64_t s = (int64_t)suf[t] + h;
    k2[t] = s < n ? (uint32_t)(lr(s) + 1) : 0u;
}

constexpr int SS_THREADS = 1024;
constexpr int SS_T = 1024;
constexpr int SS_W = 2048;
constexpr int SS_MID_MAX = 16384;
constexpr int SS_MID_CAP = 4096;
constexpr int SS_MID_THREADS = 1024;
#ifndef BWTK_SEG_MAX
#define BWTK_SEG_MAX 1000000
#endif
constexpr int64_t SEG_MAX = BWTK_SEG_MAX;   // longest active list handled by the segmented sort

struct SegCtl {
    unsigned mid_count;
    unsigned huge;
    unsigned mid_start[SS_MID_CAP];
    unsigned mid_size[SS_MID_CAP];
};

// bitonic sort of N (power of two) (hi, lo) pairs in shared memory, ascending by (hi, lo)
template <int THREADS, bool HAS_HI>
__device__ __forceinline__ void bitonic_sort(uint32_t *hi, unsigned long long *lo, int N)
{
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < (N >> 1); i += THREADS) {
                const int a = 2 * i - (i & (j - 1)), b = a + j;
                const bool up = (a & k) == 0;
                const unsigned long long la = lo[a], lb = lo[b];
                bool gt;
                if (HAS_HI) {
                    const uint32_t ha = hi[a], hb = hi[b];
                    gt = ha > hb || (ha == hb && la > lb);
                    if (gt == up) { hi[a] = hb; hi[b] = ha; }
                } else {
                    gt = la > lb;
                }
                if (gt == up) { lo[a] = lb; lo[b] = la; }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(SS_THREADS)
    segsort_kernel(const uint32_t *__restrict__ suf, const int32_t *__restrict__ grp,
                   const uint32_t *__restrict__ k2in, const unsigned *__restrict__ d_m, int kbits,
                   uint64_t *__restrict__ key_out, uint32_t *__restrict__ suf_out, SegCtl *ctl)
{
    __shared__ uint32_t s_hi[SS_W];
    __shared__ unsigned long long s_lo[SS_W];
    __shared__ uint32_t s_suf[SS_W];
    __shared__ int s_first, s_lasthead, s_end;
    const int64_t m = *d_m;
    const int64_t t0 = (int64_t)blockIdx.x * SS_T;
    if (t0 >= m) return;
    const int64_t t1 = t0 + SS_T < m ? t0 + SS_T : m;
    const int tid = threadIdx.x;
    if (tid == 0) { s_first = 0x7fffffff; s_lasthead = -1; s_end = 0x7fffffff; }
    __syncthreads();
    // group heads inside the nominal range
    for (int e = tid; e < (int)(t1 - t0); e += SS_THREADS) {
        const int64_t t = t0 + e;
        if (t == 0 || grp[t] != grp[t - 1]) { atomicMin(&s_first, e); atomicMax(&s_lasthead, e); }
    }
    __syncthreads();
    if (s_first == 0x7fffffff) return;   // the whole range belongs to a group that started earlier
    const int64_t w0 = t0 + s_first;     // first owned element
    const int64_t gl = t0 + s_lasthead;  // head of the group holding the last nominal element
    const int32_t g_last = grp[gl];
    // end of that group: first t >= t1 with another grp (bounded: beyond SS_MID_MAX it is "huge")
    bool huge = false;
    for (int64_t off = 0;; off += SS_THREADS) {
        const int64_t t = t1 + off + tid;
        const bool same = t < m && grp[t] == g_last;
        if (!same) atomicMin(&s_end, (int)(off + tid));
        __syncthreads();
        const int found = s_end;
        __syncthreads();   // nobody updates s_end for the next chunk before everyone has read it
        if (found != 0x7fffffff) break;
        if (t1 + off + SS_THREADS - gl > SS_MID_MAX) { huge = true; break; }
    }
    int64_t wend = t1 + (huge ? 0 : s_end);   // end of the last group (if known)
    if (huge || wend - w0 > SS_W) {
        // defer the last group; the window keeps the groups before it
        if (tid == 0) {
            if (huge || wend - gl > SS_MID_MAX) {
                atomicExch(&ctl->huge, 1u);
            } else {
                unsigned slot = atomicAdd(&ctl->mid_count, 1u);
                if (slot < (unsigned)SS_MID_CAP) { ctl->mid_start[slot] = (unsigned)gl; ctl->mid_size[slot] = (unsigned)(wend - gl); }
                else atomicExch(&ctl->huge, 1u);
            }
        }
        wend = gl;
    }
    const int len = (int)(wend - w0);
    if (len <= 0) return;
    int N = 32;
    while (N < len) N <<= 1;
    for (int e = tid; e < N; e += SS_THREADS) {
        if (e < len) {
            const int64_t t = w0 + e;
            const unsigned long long k2 = k2in[t];
            s_suf[e] = suf[t];
            s_hi[e] = (uint32_t)grp[t];
            s_lo[e] = (k2 << 11) | (unsigned long long)e;
        } else {
            s_hi[e] = 0xffffffffu;
            s_lo[e] = ~0ull;
        }
    }
    __syncthreads();
    bitonic_sort<SS_THREADS, true>(s_hi, s_lo, N);
    for (int e = tid; e < len; e += SS_THREADS) {
        const unsigned long long lo = s_lo[e];
        key_out[w0 + e] = ((uint64_t)s_hi[e] << kbits) | (uint64_t)(lo >> 11);
        suf_out[w0 + e] = s_suf[(int)(lo & 2047ull)];
    }
}

// one CTA per deferred group (SS_W < window, size <= SS_MID_MAX): all of it in shared memory
__global__ void __launch_bounds__(SS_MID_THREADS)
    midsort_kernel(const uint32_t *__restrict__ suf, const int32_t *__restrict__ grp,
                   const uint32_t *__restrict__ k2in, int kbits, uint64_t *__restrict__ key_out,
                   uint32_t *__restrict__ suf_out, const SegCtl *__restrict__ ctl)
{
    extern __shared__ __align__(16) unsigned char ss_raw[];
    unsigned long long *s_lo = reinterpret_cast<unsigned long long *>(ss_raw);
    uint32_t *s_suf = reinterpret_cast<uint32_t *>(ss_raw + (size_t)SS_MID_MAX * 8);
    unsigned count = ctl->mid_count;
    if (count > (unsigned)SS_MID_CAP) count = SS_MID_CAP;
    for (unsigned g = blockIdx.x; g < count; g += gridDim.x) {
        const int64_t w0 = ctl->mid_start[g];
        const int len = (int)ctl->mid_size[g];
        const uint64_t ghead = (uint64_t)(uint32_t)grp[w0];
        int N = 32;
        while (N < len) N <<= 1;
        for (int e = threadIdx.x; e < N; e += SS_MID_THREADS) {
            if (e < len) {
                const unsigned long long k2 = k2in[w0 + e];
                s_suf[e] = suf[w0 + e];
                s_lo[e] = (k2 << 14) | (unsigned long long)e;
            } else {
                s_lo[e] = ~0ull;
            }
        }
        __syncthreads();
        bitonic_sort<SS_MID_THREADS, false>(nullptr, s_lo, N);
        for (int e = threadIdx.x; e < len; e += SS_MID_THREADS) {
            const unsigned long long lo = s_lo[e];
            key_out[w0 + e] = (ghead << kbits) | (uint64_t)(lo >> 14);
            suf_out[w0 + e] = s_suf[(int)(lo & 16383ull)];
        }
        __syncthreads();
    }
}

// isa[sa[j]] = j (only when the caller asks for the inverse suffix array)
__global__ void invert_kernel(const int32_t *__restrict__ sa, int64_t n, int32_t *__restrict__ isa)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) isa[sa[j]] = (int32_t)j;
}

__device__ __forceinline__ void unpack16(const uint4 &q, uint32_t *dst)
{
    dst[0] = q.x; dst[1] = q.y; dst[2] = q.z; dst[3] = q.w;
}
__device__ __forceinline__ void unpack16(const uint4 &q, uint64_t *dst)
{
    dst[0] = ((uint64_t)q.y << 32) | q.x;
    dst[1] = ((uint64_t)q.w << 32) | q.z;
}

// One pass over the sorted (key, suffix) list: see file header.  The element
// count is m_host when d_m is null, else *d_m (grid sized for an upper bound).
template <typename KeyT, bool FIRST>
__global__ void __launch_bounds__(RG_THREADS)
    regroup_kernel(const KeyT *__restrict__ skey, const uint32_t *__restrict__ ssuf,
                   const int32_t *__restrict__ pos, int64_t m, const unsigned *__restrict__ d_m,
                   int64_t short_from, int32_t *__restrict__ sa, int32_t *__restrict__ rank,
                   int32_t *__restrict__ npos, uint32_t *__restrict__ nsuf, int32_t *__restrict__ ngrp,
                   uint32_t *__restrict__ abits, int32_t *__restrict__ ptab,
                   unsigned long long *status, unsigned *tile_counter, unsigned *out_count, int *err)
{
    __shared__ unsigned s_tile;
    __shared__ unsigned long long s_warp[RG_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    if (d_m) m = *d_m;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t tile0 = tile * RG_TILE;
    if (tile0 >= m) return;
    const int64_t t0 = tile0 + (int64_t)tid * RG_ITEMS;

    // Blocked loads straight into registers: RG_ITEMS consecutive elements per thread as
    // 16-byte vectors (the buffers are 256-byte aligned and t0 is a multiple of RG_ITEMS),
    // plus the two neighbours t0-1 and t0+RG_ITEMS, which hit the sectors the adjacent
    // threads fetch anyway.  k[j+1], sf[j+1] <-> element t0+j.
    KeyT k[RG_ITEMS + 2];
    uint32_t sf[RG_ITEMS + 2];
    constexpr int KV = 16 / (int)sizeof(KeyT);   // keys per 16-byte vector
    if (t0 + RG_ITEMS <= m) {
#pragma unroll
        for (int v = 0; v < RG_ITEMS / KV; v++) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(skey + t0) + v);
            unpack16(q, &k[1 + v * KV]);
        }
#pragma unroll
        for (int v = 0; v < RG_ITEMS / 4; v++) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(ssuf + t0) + v);
            sf[1 + v * 4 + 0] = q.x; sf[1 + v * 4 + 1] = q.y; sf[1 + v * 4 + 2] = q.z; sf[1 + v * 4 + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < RG_ITEMS; j++) {
            const bool ok = t0 + j < m;
            k[1 + j] = ok ? skey[t0 + j] : (KeyT)0;
            sf[1 + j] = ok ? ssuf[t0 + j] : 0u;
        }
    }
    {
        const bool okp = t0 >= 1 && t0 - 1 < m, okn = t0 + RG_ITEMS < m;
        k[0] = okp ? __ldg(skey + t0 - 1) : (KeyT)0;
        sf[0] = okp ? __ldg(ssuf + t0 - 1) : 0u;
        k[RG_ITEMS + 1] = okn ? __ldg(skey + t0 + RG_ITEMS) : (KeyT)0;
        sf[RG_ITEMS + 1] = okn ? __ldg(ssuf + t0 + RG_ITEMS) : 0u;
    }
    if (FIRST) {
        // SA[t] = sorted suffix: every thread stores its RG_ITEMS consecutive entries
        if (t0 + RG_ITEMS <= m) {
#pragma unroll
            for (int v = 0; v < RG_ITEMS / 4; v++)
                reinterpret_cast<int4 *>(sa + t0)[v] = make_int4((int)sf[1 + v * 4], (int)sf[2 + v * 4],
                                                                 (int)sf[3 + v * 4], (int)sf[4 + v * 4]);
        } else {
#pragma unroll
            for (int j = 0; j < RG_ITEMS; j++)
                if (t0 + j < m) sa[t0 + j] = (int32_t)sf[1 + j];
        }
    }
    bool head[RG_ITEMS + 1];
#pragma unroll
    for (int j = 0; j <= RG_ITEMS; j++) {
        int64_t t = t0 + j;
        bool h = (t <= 0) || (t >= m) || (k[j + 1] != k[j]);
        if (FIRST) h = h || ((int64_t)sf[j + 1] >= short_from) || ((int64_t)sf[j] >= short_from);
        head[j] = h;
    }
    // thread-local aggregate
    uint32_t lmax = 0, lsum = 0;
    int32_t p[RG_ITEMS];
#pragma unroll
    for (int j = 0; j < RG_ITEMS; j++) {
        int64_t t = t0 + j;
        p[j] = 0;
        if (t < m) {
            p[j] = FIRST ? (int32_t)t : pos[t];
            if (head[j]) lmax = (uint32_t)p[j] + 1u;  // positions increase with t
            if (!(head[j] && head[j + 1])) lsum++;
        }
    }
    unsigned long long inc = rg_pack(lmax, lsum);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = rg_combine(t, inc);
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned long long wpre = 0;
#pragma unroll
    for (int w = 0; w < RG_THREADS / 32; w++)
        if (w < warp) wpre = rg_combine(wpre, s_warp[w]);
    unsigned long long excl_in_tile = rg_combine(wpre, __shfl_up_sync(0xffffffffu, inc, 1));
    if (lane == 0) excl_in_tile = wpre;

    if (warp == RG_THREADS / 32 - 1) {
        // the last warp chains the tile: lane 31 holds the tile aggregate, all 32 lanes
        // inspect 32 predecessor tiles per round trip
        const unsigned long long tile_agg = __shfl_sync(0xffffffffu, rg_combine(wpre, inc), 31);
        volatile unsigned long long *st = status;
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 31) st[0] = RG_INCL | tile_agg;
        } else {
            if (lane == 31) st[tile] = RG_AGG | tile_agg;
            excl = warp_lookback(st, tile, RgComb{}, err, 2, rsort::SPIN_LIMIT);
            if (lane == 31) st[tile] = RG_INCL | rg_combine(excl, tile_agg);
        }
        if (lane == 31) {
            s_prefix = excl;
            if ((tile + 1) * (int64_t)RG_TILE >= m) *out_count = rg_sum(rg_combine(excl, tile_agg));
        }
    }
    __syncthreads();
    unsigned long long run = rg_combine(s_prefix, excl_in_tile);
    uint32_t cur_max = rg_max(run), cur_sum = rg_sum(run);
#pragma unroll
    for (int j = 0; j < RG_ITEMS; j++) {
        int64_t t = t0 + j;
        if (t < m) {
            if (head[j]) cur_max = (uint32_t)p[j] + 1u;
            int32_t hp = (int32_t)cur_max - 1;
            uint32_t s = sf[j + 1];
            if (!FIRST) sa[p[j]] = (int32_t)s;   // round 0 wrote SA from the staged tile
            const bool active = !(head[j] && head[j + 1]);
            // round 0: only still-ambiguous suffixes get a rank entry (+ their bit);
            // later rounds: every element was active once, so its entry is maintained
            if (!FIRST || active) rank[s] = hp;
            if (FIRST) {
                if (active) atomicOr(abits + (s >> 5), 1u << (s & 31));
                // first sorted position of every 16-bit key prefix
                uint32_t pc = (uint32_t)(k[j + 1] >> 16), pp = (uint32_t)(k[j] >> 16);
                if (t == 0) { for (uint32_t q = 0; q <= pc; q++) ptab[q] = 0; }
                else if (pc != pp) { for (uint32_t q = pp + 1; q <= pc; q++) ptab[q] = (int32_t)t; }
                if (t == m - 1) { for (uint32_t q = pc + 1; q <= 65536u; q++) ptab[q] = (int32_t)m; }
            }
            if (active) {
                npos[cur_sum] = p[j];
                nsuf[cur_sum] = s;
                ngrp[cur_sum] = hp;
                cur_sum++;
            }
        }
    }
}

static int bits_for(int64_t v)  // bits needed to represent values 0..v
{
    int b = 1;
    while ((1ll << b) <= v) b++;
    return b;
}

}  // namespace sa

// Round 0 goes through the MSD bucket sort for 2-bit texts from BWTK_MSD_MIN_N symbols on (below that the
// 8192 sort CTAs are mostly empty and the four LSD passes are cheaper).  BWTK_MSD=0 / BWTK_MSD_FUSE=0 switch
// the path / the fused regroup off (tuning and tests).
static bool msd_enabled(int64_t n)
{
    const char *e = getenv("BWTK_MSD");
    if (e && atoi(e) == 0) return false;
    const char *m = getenv("BWTK_MSD_MIN_N");
    const int64_t min_n = m ? atoll(m) : (1ll << 24);
    return n >= min_n;
}
static bool msd_fuse_enabled()
{
    const char *e = getenv("BWTK_MSD_FUSE");
    return !(e && atoi(e) == 0);
}

// Workspace of the doubling rounds (the packed text is provided by the caller).
int64_t sa_core_workspace_bytes(int64_t n)
{
    if (n < 1) n = 1;
    int64_t b = 0;
    b += align_up(n * 4, 256);                        // rank of the suffixes that stay ambiguous after round 0
    b += align_up(n * 4, 256);                        // round-0 keys in sorted order (lazy singleton ranks)
    b += align_up((n / 32 + 2) * 4, 256);             // "was active" bitmap
    b += align_up(65538 * 4, 256);                    // 16-bit key-prefix table
    b += 2 * align_up(n * 8, 256);                    // key buffers A, B
    b += 2 * align_up(n * 4, 256);                    // value buffers
    b += 2 * align_up(n * 4, 256);                    // position buffers
    b += align_up(n * 4, 256);                        // group heads
    b += align_up(ceil_div(n, sa::RG_TILE) * 8 + 256, 256);  // regroup status
    b += rsort::workspace_bytes(n);
    b += 8192 + align_up((int64_t)sizeof(sa::SegCtl), 256);  // counters, segmented-sort control block
    b += msd::workspace_bytes(n) + 1024;   // round-0 bucket sort (bucket_sort.cuh)
    return b;
}

// Suffix array of the text whose packed form (bits per symbol, `fast` = ACGT$
// layout) is already on the device.  Synchronises the stream.
int sa_build_core(const uint32_t *packed, int64_t n, int bits, bool fast, int32_t *d_sa, int32_t *d_isa_out,
                  void *d_ws, int64_t ws_bytes, int64_t *h_stats, cudaStream_t st, const uint32_t **d_skey0_out,
                  void **d_free_out, int64_t *free_bytes_out)
{
    if (h_stats) memset(h_stats, 0, 8 * sizeof(int64_t));
    if (d_skey0_out) *d_skey0_out = nullptr;
    if (d_free_out) { *d_free_out = nullptr; *free_bytes_out = 0; }
    BWTK_REQUIRE((((uintptr_t)d_sa | (uintptr_t)d_ws) & 15) == 0, "d_sa and the workspace must be 16-byte aligned");
    if (ws_bytes < sa_core_workspace_bytes(n)) {
        set_error("sa workspace: need %lld bytes, got %lld", (long long)sa_core_workspace_bytes(n),
                  (long long)ws_bytes);
        return BWTK_EWORKSPACE;
    }
    if (n == 1) {
        BWTK_CUDA(bwtk::zero_async(d_sa, 4, st));
        if (d_isa_out) BWTK_CUDA(bwtk::zero_async(d_isa_out, 4, st));
        BWTK_CUDA(cudaStreamSynchronize(st));
        if (h_stats) { h_stats[0] = 1; h_stats[1] = bits; h_stats[2] = 32 / bits; h_stats[6] = fast ? 1 : 0; }
        return BWTK_OK;
    }
    Carver c(d_ws, ws_bytes);
    int32_t *rank = c.take<int32_t>(n);
    uint32_t *skeep = c.take<uint32_t>(n);
    uint32_t *abits = c.take<uint32_t>(n / 32 + 2);
    int32_t *ptab = c.take<int32_t>(65538);
    uint64_t *keyA = c.take<uint64_t>(n);
    uint64_t *keyB = c.take<uint64_t>(n);
    uint32_t *val0 = c.take<uint32_t>(n);
    uint32_t *val1 = c.take<uint32_t>(n);
    int32_t *pos0 = c.take<int32_t>(n);
    int32_t *pos1 = c.take<int32_t>(n);
    int32_t *grp = c.take<int32_t>(n);
    int64_t rg_tiles_max = ceil_div(n, sa::RG_TILE);
    unsigned long long *rg_status = c.take<unsigned long long>(rg_tiles_max + 8);
    rsort::Workspace rws = rsort::carve(c, n);
    unsigned *counters = c.take<unsigned>(sa::MAX_ROUNDS + 8);  // [0] regroup tile id, [1+r] active count entering round r
    sa::SegCtl *ctl = c.take<sa::SegCtl>(1);
    msd::Workspace mws = msd::carve(c, n);
    if (!c.ok()) {
        set_error("sa workspace carve overflow");
        return BWTK_EWORKSPACE;
    }
    unsigned *d_counts = counters + 1;
    const int S = 32 / bits;

    BWTK_CUDA(bwtk::zero_async(rws.err, sizeof(int), st));
    BWTK_CUDA(bwtk::zero_async(counters, (sa::MAX_ROUNDS + 8) * sizeof(unsigned), st));
    BWTK_CUDA(bwtk::zero_async(abits, (size_t)(n / 32 + 2) * 4, st));
    int64_t passes = 0;
    int in_first = 1;
    uint32_t *skey32 = nullptr, *sval = nullptr, *suf_other = nullptr;
    int32_t *pos_in = pos0, *pos_out = pos1;
    bool msd_done = false, msd_fused = false;
    if (bits == 2 && msd_enabled(n)) {
        // round 0 as an MSD bucket sort that only moves positions (bucket_sort.cuh); with no oversize
        // bucket it also does the first regroup pass
        // scratch of the oversize pass: pos0 is free either way; fused, the active list goes to (pos_out, val1, grp)
        // and val0 is free; unfused, val0 receives the sorted suffixes and val1 is free
        const bool fuse = msd_fuse_enabled();
        msd::Regroup rg{d_sa, rank, pos_out, val1, grp, abits, ptab, nullptr, d_counts + 1, rws.err, nullptr};
        int rc = msd::round0_sort(packed, n, mws, rws, reinterpret_cast<uint2 *>(keyA), reinterpret_cast<uint2 *>(keyB),
                                  skeep, val0, reinterpret_cast<uint32_t *>(pos0), fuse ? val0 : val1, fuse ? &rg : nullptr,
                                  ptab, n - S + 1, st, &msd_done, &msd_fused, &passes);
        if (rc) return rc;
        if (msd_done) { skey32 = skeep; sval = val0; suf_other = val1; }
    }
    if (!msd_done) {
        // round-0 ping-pong buffers chosen so that the sorted keys land in `skeep`
        // (the generator pass writes k0; an odd number of passes ends in k0)
        const bool odd = (rsort::make_plan(0, S * bits).passes & 1) != 0;
        uint32_t *key32a = odd ? skeep : (uint32_t *)keyA, *key32b = odd ? (uint32_t *)keyA : skeep;
        // the histogram and the first radix pass read the suffix keys
        // straight from the packed text (no key/value arrays are materialised)
        rsort::PackedSuffixSource src{packed, n, bits, S * bits};
        int rc = rsort::sort_pairs_from<uint32_t, rsort::PackedSuffixSource>(src, false, key32a, val0, key32b, val1, n,
                                                                            0, S * bits, rws, st, &in_first, &passes);
        if (rc) return rc;
        skey32 = in_first ? key32a : key32b;
        sval = in_first ? val0 : val1;
        suf_other = in_first ? val1 : val0;  // free value buffer receives the active suffixes
    }
    // key of the suffix at every SA position (later rounds only permute suffixes inside a
    // group of equal keys); stays valid until the caller reuses the workspace
    if (d_skey0_out) *d_skey0_out = skey32;
    if (d_free_out) {
        // everything from the second key buffer on is scratch for the caller once this function returns
        // (28 n bytes + the radix workspace: enough for the LCP's deep-pair pass, lcp.cu)
        *d_free_out = keyB;
        *free_bytes_out = (int64_t)((char *)d_ws + ws_bytes - (char *)keyB);
    }
    if (!msd_fused) {
        int64_t tiles = ceil_div(n, sa::RG_TILE);
        BWTK_CUDA(bwtk::zero_async(rg_status, (size_t)tiles * 8, st));
        {
            prof::Scope ps("regroup_first", n * 16, st);
            sa::regroup_kernel<uint32_t, true><<<(unsigned)tiles, sa::RG_THREADS, 0, st>>>(
                skey32, sval, nullptr, n, nullptr, n - S + 1, d_sa, rank, pos_out, suf_other, grp, abits, ptab,
                rg_status, counters, d_counts + 1, rws.err);
        }
        BWTK_LAUNCH_CHECK();
    }
    unsigned h_count = 0;
    { int rc = read_back(&h_count, d_counts + 1, sizeof(unsigned), st); if (rc) return rc; }
    const int64_t active0 = h_count;
    // after round 0 the active list lives in (pos_out, suf_other, grp)
    uint32_t *suf_in = suf_other;
    uint32_t *suf_free = sval;
    { int32_t *t = pos_in; pos_in = pos_out; pos_out = t; }
    const int kbits = sa::bits_for(n);
    const int gbits = sa::bits_for(n - 1);
    const sa::LazyRank lr{abits, rank, skey32, ptab, d_sa, packed, bits};
    int64_t h = S;
    int round = 1;   // next round index; its count is d_counts[round]
    static bool mid_attr = false;
    const size_t mid_smem = (size_t)sa::SS_MID_MAX * 12;
    if (!mid_attr) {
        BWTK_CUDA(cudaFuncSetAttribute(sa::midsort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mid_smem));
        mid_attr = true;
    }
    while (true) {
        const unsigned *d_m = d_counts + round;
        if (round > 1) {   // round 1's count was read after the first regroup
            int rc = read_back(&h_count, d_m, sizeof(unsigned), st);
            if (rc) return rc;
        }
        if (h_count == 0) break;
        if (round >= sa::MAX_ROUNDS || h > 4 * n) {
            set_error("suffix array refinement did not converge");
            return BWTK_EINTERNAL;
        }
        const int64_t bound = h_count;
        // Every group is ordered by rank[suffix + h].  Long lists go through the radix sort
        // (bandwidth-bound, ~7 passes); below SEG_MAX elements the passes are launch- and
        // latency-bound and the shared-memory segmented sort (segsort_kernel) is cheaper.
        const bool seg = bound <= sa::SEG_MAX;
        unsigned h_ctl[2] = {0, 1};   // [1] != 0: use the radix path
        if (seg) {
            BWTK_CUDA(bwtk::zero_async(ctl, 8, st));
            uint32_t *k2buf = reinterpret_cast<uint32_t *>(keyA);   // keyA is only used by the radix path
            {
                prof::Scope ps("build_k2_kernel", bound * 12, st);
                sa::build_k2_kernel<<<(unsigned)ceil_div(bound, 256), 256, 0, st>>>(suf_in, lr, d_m, n, h, k2buf);
            }
            BWTK_LAUNCH_CHECK();
            {
                prof::Scope ps("segsort_kernel", bound * 28, st);
                sa::segsort_kernel<<<(unsigned)ceil_div(bound, sa::SS_T), sa::SS_THREADS, 0, st>>>(
                    suf_in, grp, k2buf, d_m, kbits, keyB, suf_free, ctl);
            }
            BWTK_LAUNCH_CHECK();
            {
                prof::Scope ps("midsort_kernel", 0, st);
                sa::midsort_kernel<<<64, sa::SS_MID_THREADS, mid_smem, st>>>(suf_in, grp, k2buf, kbits, keyB, suf_free,
                                                                           ctl);
            }
            BWTK_LAUNCH_CHECK();
            int rc = read_back(h_ctl, ctl, 8, st);
            if (rc) return rc;
        }
        uint64_t *sk = keyB;
        uint32_t *ss = suf_free, *sn = suf_in;
        if (h_ctl[1]) {
            // long list, or a group larger than one CTA's shared memory: radix sort of (group, k2)
            {
                prof::Scope ps("build_keys_kernel", (int64_t)h_count * 20, st);
                sa::build_keys_kernel<<<(unsigned)ceil_div((int64_t)h_count, 256), 256, 0, st>>>(suf_in, grp, lr, d_m, n, h,
                                                                                                kbits, keyA);
            }
            BWTK_LAUNCH_CHECK();
            int rc = rsort::sort_pairs<uint64_t>(keyA, suf_in, keyB, suf_free, h_count, 0, kbits + gbits, rws, st,
                                                 &in_first, &passes, d_m);
            if (rc) return rc;
            sk = in_first ? keyA : keyB;
            ss = in_first ? suf_in : suf_free;
            sn = in_first ? suf_free : suf_in;  // the other value buffer takes the next list
        }
        int64_t tiles = ceil_div((int64_t)h_count, sa::RG_TILE);
        BWTK_CUDA(bwtk::zero_async(rg_status, (size_t)tiles * 8, st));
        BWTK_CUDA(bwtk::zero_async(counters, sizeof(unsigned), st));
        {
            prof::Scope ps("regroup_round", (int64_t)h_count * 36, st);
            sa::regroup_kernel<uint64_t, false><<<(unsigned)tiles, sa::RG_THREADS, 0, st>>>(
                sk, ss, pos_in, h_count, d_m, 0, d_sa, rank, pos_out, sn, grp, nullptr, nullptr, rg_status, counters,
                d_counts + round + 1, rws.err);
        }
        BWTK_LAUNCH_CHECK();
        suf_in = sn;
        suf_free = ss;
        { int32_t *t = pos_in; pos_in = pos_out; pos_out = t; }
        h <<= 1;
        round++;
    }
    if (d_isa_out) {
        prof::Scope ps("invert_kernel", n * 8, st);
        sa::invert_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(d_sa, n, d_isa_out);
        BWTK_LAUNCH_CHECK();
    }
    int h_err = 0;
    unsigned h_counts[sa::MAX_ROUNDS + 1];
    { int rc = read_back(&h_err, rws.err, sizeof(int), st); if (rc) return rc; }
    { int rc = read_back(h_counts, d_counts, sizeof(h_counts), st); if (rc) return rc; }
    if (h_err) {
        set_error("look-back spin limit hit (code %d)", h_err);
        return BWTK_EINTERNAL;
    }
    if (h_stats) {
        int64_t sum_active = 0, rounds = 1;
        for (int r = 1; r < sa::MAX_ROUNDS; r++) {
            if (h_counts[r] == 0) break;
            sum_active += h_counts[r];
            rounds++;
        }
        h_stats[0] = rounds; h_stats[1] = bits; h_stats[2] = S; h_stats[3] = active0;
        h_stats[4] = sum_active; h_stats[5] = passes;
        h_stats[6] = (fast ? 1 : 0) | (msd_done ? 2 : 0) | (msd_fused ? 4 : 0);   // bit 1: round 0 by bucket sort, bit 2: fused regroup
    }
    return BWTK_OK;
}

}  // namespace bwtk

using namespace bwtk;

namespace bwtk {
int prepare_text(const uint8_t *d_text, int64_t n, uint32_t *d_packed, unsigned long long *d_hist_scratch,
                 int64_t *h_totals, int *bits_out, bool *fast_out, cudaStream_t st);
}

extern "C" int64_t bwtk_sa_workspace_bytes(int64_t n)
{
    if (n < 1) n = 1;
    return sa_core_workspace_bytes(n) + align_up(packed_words(n, 8) * 4, 256) + 4096;
}

extern "C" int32_t bwtk_sa_build(const uint8_t *d_text, int64_t n, int32_t *d_sa, int32_t *d_isa_out,
                                 void *d_ws, int64_t ws_bytes, int64_t *h_stats, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (h_stats) memset(h_stats, 0, 8 * sizeof(int64_t));
    BWTK_REQUIRE(n >= 0 && n < (1ll << 30), "n must be in [0, 2^30)");
    if (n == 0) return BWTK_OK;
    BWTK_REQUIRE(d_text && d_sa && d_ws, "null pointer");
    if (ws_bytes < bwtk_sa_workspace_bytes(n)) {
        set_error("sa workspace: need %lld bytes, got %lld", (long long)bwtk_sa_workspace_bytes(n),
                  (long long)ws_bytes);
        return BWTK_EWORKSPACE;
    }
    Carver c(d_ws, ws_bytes);
    uint32_t *packed = c.take<uint32_t>(packed_words(n, 8));
    unsigned long long *d_hist = c.take<unsigned long long>(260);
    int64_t totals[256];
    int bits;
    bool fast;
    int rc = prepare_text(d_text, n, packed, d_hist, totals, &bits, &fast, st);
    if (rc) return rc;
    c.off = align_up(c.off, 256);
    return sa_build_core(packed, n, bits, fast, d_sa, d_isa_out, (char *)d_ws + c.off, ws_bytes - c.off, h_stats, st,
                         nullptr, nullptr, nullptr);
}
