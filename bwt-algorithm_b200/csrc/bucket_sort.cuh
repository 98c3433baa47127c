// bucket_sort.cuh -- round 0 of the suffix sort (sa.cu) for 2-bit texts as an MSD bucket sort.
//
// The round-0 key of suffix i is the 32-bit window of the packed text at symbol i.  Per-key scattered
// global accesses are bound by the L2 request rate (~2.4 cycles per key and SM on a B200,
// tools/microbench/lsu_bench.cu), ballot ranking costs ~1 cycle per key and pass, shared-memory atomics
// 0.15 -- so every level below ranks with shared atomics (an MSD sort needs no stable passes) and only
// ever writes global memory in runs:
//
//   coarse_scatter  tile of 32 768 text positions staged in shared memory, grouped by the top 8 key bits
//                   (256 shared counters, one global claim per bin and tile), (position, key) pairs out in
//                   runs of ~128;
//   fine_hist       sizes of the NB-bit buckets (NB = 13..18, 3-6 k suffixes each) from those pairs;
//   bucket_scan     bucket starts, oversize buckets;
//   fine_scatter    tiles of 8 192 pairs of one coarse bucket, grouped by the next NB-8 bits the same way,
//                   (position, low key bits) pairs out;
//   bucket_sort     one CTA per bucket: a counting sort on the next 12 key bits with shared atomics, the
//                   remaining <= 7 bits by counting the smaller peers inside the sub-bucket (elements are
//                   32-bit words (low key bits, slot) -- one unsigned compare orders them; the <= 15 suffixes
//                   that run off the end of the text take the lowest slots, shortest first, which is the
//                   reference's "shorter suffix first" rule); sorted keys + suffixes out, coalesced.
//                   FUSED: also everything sa::regroup_kernel<FIRST> did (SA, group heads, active list,
//                   "was active" bits, rank of active suffixes, 16-bit prefix table), the active-list
//                   offsets chained bucket to bucket by a decoupled look-back.
//
// A bucket with more than CAP_EFF suffixes (low-complexity sequence) goes through the LSD radix sort
// instead -- 64-bit keys (key, ~position) of the oversize buckets only; fused, all of its suffixes enter the
// active list (ovf_regroup_kernel).  If most of the text is in such buckets the caller keeps the plain LSD path.
#pragma once
#include <stdlib.h>

#include "radix_sort.cuh"
#include "scan.cuh"

namespace bwtk {
namespace msd {

constexpr int COARSE_BITS = 8;
constexpr int COARSE = 1 << COARSE_BITS;
constexpr int NB_MIN = 13, NB_MAX = 18;
constexpr int MAX_SUB = 1 << (NB_MAX - COARSE_BITS);   // fine buckets per coarse bucket, at most
constexpr int CAP = 8192;                       // shared-memory slots of one sort CTA
constexpr int SHORT_SLOTS = 16;                 // slots reserved for the suffixes that run off the text
constexpr int CAP_EFF = CAP - 128;              // largest bucket one sort CTA takes (two CTAs per SM: 2 x 113 KB)
constexpr int IDX_BITS = 13;                    // log2(CAP)
constexpr int SUB_BITS = 12;                    // digit of the in-bucket counting sort
constexpr int SUB = 1 << SUB_BITS;
#ifndef BWTK_MSD_TARGET
#define BWTK_MSD_TARGET 5800
#endif
constexpr int TARGET = BWTK_MSD_TARGET;         // mean bucket size aimed for
constexpr int SORT_THREADS = 512;
constexpr int SORT_ITEMS = CAP / SORT_THREADS;  // 16
#ifndef BWTK_MSD_BIN_PEER
#define BWTK_MSD_BIN_PEER 32
#endif
constexpr int BIN_PEER = BWTK_MSD_BIN_PEER;     // sub-buckets up to this size: every element counts its smaller peers
constexpr int PREFETCH_AHEAD = 2 * 148 + 64;          // sort CTAs resident at a time (two per SM) and a few more
constexpr int BIG_LIST = 256;                   // larger ones (at most CAP_EFF / (BIN_PEER + 1) < 256 of them): counting sort by the CTA
static_assert(CAP_EFF / (BIN_PEER + 1) < BIG_LIST, "big-bin list");
constexpr int CA_THREADS = 512;
constexpr int CA_WPT = 4;                       // packed words per thread
constexpr int CA_WORDS = CA_THREADS * CA_WPT;   // 2048 words = 32 768 positions at 2 bits
constexpr int CA_TILE = CA_WORDS * 16;
constexpr int FI_THREADS = 512;
constexpr int FI_ITEMS = 16;
constexpr int FI_TILE = FI_THREADS * FI_ITEMS;  // 8192 pairs
constexpr int FH_CHUNK = 65536;                 // pairs per fine_hist CTA
static_assert(SUB == SORT_THREADS * 8, "the sub-bucket scan gives every thread 8 bins");
static_assert((1 << IDX_BITS) == CAP, "element index bits");
static_assert(CA_TILE <= 65536, "local positions are 16 bits");
static_assert(MAX_SUB == 2 * FI_THREADS, "the fine scan gives every thread 2 bins");

static inline int nb_bits_for(int64_t n)
{
    int nb = NB_MIN;
    while (nb < NB_MAX && (n >> nb) > TARGET) nb++;
    return nb;
}

struct Info {
    unsigned ovf_buckets;   // buckets with more than CAP_EFF suffixes
    unsigned ovf_elems;     // suffixes in them
    unsigned max_bucket;
    unsigned pad;
};

struct Workspace {
    uint32_t *chist;       // [256]      coarse sizes
    uint32_t *cstart;      // [257]      coarse starts
    uint32_t *cfill;       // [256]      claim cursors of the coarse scatter
    uint32_t *ftile0;      // [257]      first fine-scatter tile of every coarse bucket
    uint32_t *htile0;      // [257]      first fine-hist chunk of every coarse bucket
    uint32_t *hist;        // [NBK]      bucket sizes
    uint32_t *bstart;      // [NBK + 1]  bucket starts
    uint32_t *fill;        // [NBK]      claim cursors of the fine scatter
    uint32_t *ovf_prefix;  // [NBK]      oversize suffixes in earlier buckets
    uint32_t *ovf_list;    // [NBK]      oversize bucket ids
    unsigned long long *lb_status;  // [NBK] look-back over the active counts (fused regroup)
    unsigned *ticket;      // dynamic bucket ids of the sort kernel
    Info *info;
};

static inline int64_t workspace_bytes(int64_t n)
{
    const int64_t nbk = 1ll << nb_bits_for(n);
    return 5 * align_up((nbk + 2) * 4, 256) + align_up(nbk * 8 + 64, 256) + 5 * align_up(260 * 4, 256) + 2048;
}

static inline Workspace carve(Carver &c, int64_t n)
{
    const int64_t nbk = 1ll << nb_bits_for(n);
    Workspace w;
    w.chist = c.take<uint32_t>(260);
    w.cstart = c.take<uint32_t>(260);
    w.cfill = c.take<uint32_t>(260);
    w.ftile0 = c.take<uint32_t>(260);
    w.htile0 = c.take<uint32_t>(260);
    w.hist = c.take<uint32_t>(nbk + 2);
    w.bstart = c.take<uint32_t>(nbk + 2);
    w.fill = c.take<uint32_t>(nbk + 2);
    w.ovf_prefix = c.take<uint32_t>(nbk + 2);
    w.ovf_list = c.take<uint32_t>(nbk + 2);
    w.lb_status = c.take<unsigned long long>(nbk + 8);
    w.ticket = c.take<unsigned>(4);
    w.info = c.take<Info>(1);
    return w;
}

// ---- coarse level ---------------------------------------------------------------------------------
// chist = sizes of the 256 coarse buckets (rsort::chunk_hist_kernel) -> starts, claim cursors, tile tables
__global__ void __launch_bounds__(COARSE) coarse_scan_kernel(const uint32_t *__restrict__ chist, uint32_t *cstart,
                                                             uint32_t *cfill, uint32_t *ftile0, uint32_t *htile0)
{
    __shared__ uint32_t s_c[COARSE];
    const int t = threadIdx.x;
    s_c[t] = chist[t];
    __syncthreads();
    if (t == 0) {
        uint32_t a = 0, ft = 0, ht = 0;
        for (int i = 0; i < COARSE; i++) {
            const uint32_t c = s_c[i];
            cstart[i] = a; cfill[i] = a; ftile0[i] = ft; htile0[i] = ht;
            a += c;
            ft += (c + FI_TILE - 1) / FI_TILE;
            ht += (c + FH_CHUNK - 1) / FH_CHUNK;
        }
        cstart[COARSE] = a; ftile0[COARSE] = ft; htile0[COARSE] = ht;
    }
}

// window at local position lp of the staged tile (2-bit symbols)
__device__ __forceinline__ uint32_t tile_window(const uint32_t *s_text, uint32_t lp)
{
    const uint32_t w = lp >> 4;
    return __funnelshift_l(s_text[w + 1], s_text[w], (lp & 15u) * 2u);
}

constexpr size_t coarse_smem() { return (size_t)(CA_WORDS + 8) * 4 + (size_t)CA_TILE * 2; }

__global__ void __launch_bounds__(CA_THREADS, 2)
    coarse_scatter_kernel(const uint32_t *__restrict__ packed, int64_t n, uint32_t *__restrict__ cfill,
                          uint2 *__restrict__ out)
{
    extern __shared__ __align__(16) uint32_t sm_ca[];
    uint32_t *s_text = sm_ca;                                                    // [CA_WORDS + 8]
    unsigned short *s_lp = reinterpret_cast<unsigned short *>(sm_ca + CA_WORDS + 8);   // [CA_TILE] local positions by digit
    __shared__ uint32_t s_cnt[COARSE];      // counts -> placement cursors
    __shared__ uint32_t s_delta[COARSE];    // global claim - local start
    __shared__ uint32_t s_wtot[COARSE / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t w0 = (int64_t)blockIdx.x * CA_WORDS;
    const int64_t p0 = w0 * 16;
    const int tile_n = (int)((n - p0) < CA_TILE ? (n - p0) : CA_TILE);
    const int64_t nwords = (n + 15) / 16;
    for (int i = tid; i < CA_WORDS + 2; i += CA_THREADS) s_text[i] = (w0 + i < nwords + 3) ? __ldg(packed + w0 + i) : 0u;
    if (tid < COARSE) s_cnt[tid] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CA_WPT; k++) {
        const int w = tid + k * CA_THREADS;
        const uint32_t a = s_text[w], b = s_text[w + 1];
        const int lim = tile_n - w * 16;
        if (lim >= 16) {
#pragma unroll
            for (int q = 0; q < 16; q++) atomicAdd(&s_cnt[__funnelshift_l(b, a, q * 2) >> 24], 1u);
        } else {
            for (int q = 0; q < lim; q++) atomicAdd(&s_cnt[__funnelshift_l(b, a, q * 2) >> 24], 1u);
        }
    }
    __syncthreads();
    uint32_t c = 0, inc = 0;
    if (tid < COARSE) {
        c = s_cnt[tid];
        inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_wtot[warp] = inc;
    }
    __syncthreads();
    if (tid < COARSE) {
        uint32_t start = inc - c;
        for (int w = 0; w < warp; w++) start += s_wtot[w];
        const uint32_t claim = c ? atomicAdd(&cfill[tid], c) : 0u;
        s_cnt[tid] = start;
        s_delta[tid] = claim - start;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CA_WPT; k++) {
        const int w = tid + k * CA_THREADS;
        const uint32_t a = s_text[w], b = s_text[w + 1];
        const int lim = tile_n - w * 16;
        if (lim >= 16) {
#pragma unroll
            for (int q = 0; q < 16; q++)
                s_lp[atomicAdd(&s_cnt[__funnelshift_l(b, a, q * 2) >> 24], 1u)] = (unsigned short)(w * 16 + q);
        } else {
            for (int q = 0; q < lim; q++)
                s_lp[atomicAdd(&s_cnt[__funnelshift_l(b, a, q * 2) >> 24], 1u)] = (unsigned short)(w * 16 + q);
        }
    }
    __syncthreads();
    for (int j = tid; j < tile_n; j += CA_THREADS) {
        const uint32_t lp = s_lp[j];
        const uint32_t key = tile_window(s_text, lp);
        out[(uint32_t)j + s_delta[key >> 24]] = make_uint2((uint32_t)(p0 + lp), key);
    }
}

// ---- fine level -----------------------------------------------------------------------------------
// coarse bucket of tile / chunk `t`: largest c with tab[c] <= t (257-entry table of first tiles)
__device__ __forceinline__ int coarse_of_tile(const uint32_t *s_tab, uint32_t t)
{
    int lo = 0, hi = COARSE;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_tab[mid] <= t) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(512)
    fine_hist_kernel(const uint2 *__restrict__ pairs, const uint32_t *__restrict__ cstart,
                     const uint32_t *__restrict__ htile0, int nb_bits, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t s_tab[COARSE + 1];
    __shared__ uint32_t s_h[MAX_SUB];
    const int tid = threadIdx.x;
    for (int i = tid; i <= COARSE; i += 512) s_tab[i] = htile0[i];
    const int sb = nb_bits - COARSE_BITS, nsub = 1 << sb;
    for (int i = tid; i < nsub; i += 512) s_h[i] = 0;
    __syncthreads();
    if (blockIdx.x >= s_tab[COARSE]) return;
    const int c = coarse_of_tile(s_tab, blockIdx.x);
    const uint32_t lo = cstart[c] + (blockIdx.x - s_tab[c]) * FH_CHUNK;
    const uint32_t end = cstart[c + 1];
    const uint32_t hi = lo + FH_CHUNK < end ? lo + FH_CHUNK : end;
    const int sh = 32 - nb_bits;
    for (uint32_t i = lo + tid; i < hi; i += 512) atomicAdd(&s_h[(__ldg(&pairs[i].y) >> sh) & (nsub - 1)], 1u);
    __syncthreads();
    for (int i = tid; i < nsub; i += 512) {
        const uint32_t v = s_h[i];
        if (v) atomicAdd(&hist[((uint32_t)c << sb) + i], v);
    }
}

// ---- bucket starts (one CTA) --------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
    bucket_scan_kernel(const uint32_t *__restrict__ hist, int64_t nbk, uint32_t n, uint32_t *__restrict__ bstart,
                       uint32_t *__restrict__ fill, uint32_t *__restrict__ ovf_prefix, uint32_t *__restrict__ ovf_list,
                       int32_t *__restrict__ ptab, int nb_bits, Info *info)
{
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned s_nb, s_mx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_nb = 0; s_mx = 0; }
    __syncthreads();
    const int64_t ipt = (nbk + 1023) / 1024;
    const int64_t lo = (int64_t)tid * ipt, hi = lo + ipt < nbk ? lo + ipt : nbk;
    unsigned long long sum = 0;                      // [63:32] oversize elements, [31:0] elements
    unsigned mx = 0;
    for (int64_t i = lo; i < hi; i++) {
        const uint32_t c = hist[i];
        sum += c;
        if (c > (uint32_t)CAP_EFF) {
            sum += (unsigned long long)c << 32;
            ovf_list[atomicAdd(&s_nb, 1u)] = (uint32_t)i;
        }
        mx = c > mx ? c : mx;
    }
    unsigned long long inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[warp] = inc;
    atomicMax(&s_mx, mx);
    __syncthreads();
    unsigned long long run = inc - sum;
    for (int w = 0; w < warp; w++) run += s_w[w];
    for (int64_t i = lo; i < hi; i++) {
        const uint32_t c = hist[i];
        const uint32_t st = (uint32_t)run;
        bstart[i] = st;
        fill[i] = st;
        ovf_prefix[i] = (uint32_t)(run >> 32);
        run += c;
        if (c > (uint32_t)CAP_EFF) run += (unsigned long long)c << 32;
    }
    if (tid == 1023) {
        bstart[nbk] = (uint32_t)run;   // == n
        info->ovf_elems = (uint32_t)(run >> 32);
    }
    __syncthreads();
    if (tid == 0) {
        info->ovf_buckets = s_nb;
        info->max_bucket = s_mx;
        info->pad = n;
    }
    // with >= 16 bucket bits the 16-bit prefix table (sa::LazyRank) is a sub-sampling of the bucket starts;
    // with fewer the sort CTAs fill it (the caller's regroup pass does when the sort runs unfused)
    if (ptab != nullptr && nb_bits >= 16) {
        const int sh = nb_bits - 16;
        for (int q = tid; q <= 65536; q += 1024) ptab[q] = q == 65536 ? (int32_t)n : (int32_t)bstart[(int64_t)q << sh];
    }
}

constexpr size_t fine_smem() { return (size_t)FI_TILE * 8; }

__global__ void __launch_bounds__(FI_THREADS, 2)
    fine_scatter_kernel(const uint2 *__restrict__ in, const uint32_t *__restrict__ cstart,
                        const uint32_t *__restrict__ ftile0, int nb_bits, uint32_t *__restrict__ fill,
                        uint2 *__restrict__ out)
{
    extern __shared__ __align__(16) uint2 s_stage[];   // [FI_TILE]
    __shared__ uint32_t s_tab[COARSE + 1];
    __shared__ uint32_t s_cnt[MAX_SUB];
    __shared__ uint32_t s_delta[MAX_SUB];
    __shared__ uint32_t s_wsum[FI_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i <= COARSE; i += FI_THREADS) s_tab[i] = ftile0[i];
    const int sb = nb_bits - COARSE_BITS, nsub = 1 << sb;
    for (int i = tid; i < nsub; i += FI_THREADS) s_cnt[i] = 0;
    __syncthreads();
    if (blockIdx.x >= s_tab[COARSE]) return;
    const int c = coarse_of_tile(s_tab, blockIdx.x);
    const uint32_t lo = cstart[c] + (blockIdx.x - s_tab[c]) * FI_TILE;
    const uint32_t end = cstart[c + 1];
    const int tile_n = (int)(end - lo < (uint32_t)FI_TILE ? end - lo : (uint32_t)FI_TILE);
    const int sh = 32 - nb_bits;
    const uint32_t lowmask = (1u << sh) - 1u;
    uint2 e[FI_ITEMS];
#pragma unroll
    for (int k = 0; k < FI_ITEMS; k++) {
        const int i = tid + k * FI_THREADS;
        e[k] = make_uint2(0u, 0u);
        if (i < tile_n) e[k] = __ldg(in + lo + i);
    }
#pragma unroll
    for (int k = 0; k < FI_ITEMS; k++) {
        const int i = tid + k * FI_THREADS;
        if (i < tile_n) atomicAdd(&s_cnt[(e[k].y >> sh) & (nsub - 1)], 1u);
    }
    __syncthreads();
    // exclusive scan of nsub (<= 1024) counts: two bins per thread
    {
        const int b0 = tid * 2;
        const uint32_t c0 = b0 < nsub ? s_cnt[b0] : 0u, c1 = b0 + 1 < nsub ? s_cnt[b0 + 1] : 0u;
        uint32_t inc = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_wsum[warp] = inc;
        __syncthreads();
        uint32_t start = inc - (c0 + c1);
        for (int w = 0; w < warp; w++) start += s_wsum[w];
        if (b0 < nsub) {
            const uint32_t claim = c0 ? atomicAdd(&fill[((uint32_t)c << sb) + b0], c0) : 0u;
            s_cnt[b0] = start;
            s_delta[b0] = claim - start;
        }
        if (b0 + 1 < nsub) {
            const uint32_t claim = c1 ? atomicAdd(&fill[((uint32_t)c << sb) + b0 + 1], c1) : 0u;
            s_cnt[b0 + 1] = start + c0;
            s_delta[b0 + 1] = claim - (start + c0);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < FI_ITEMS; k++) {
        const int i = tid + k * FI_THREADS;
        if (i < tile_n) s_stage[atomicAdd(&s_cnt[(e[k].y >> sh) & (nsub - 1)], 1u)] = e[k];
    }
    __syncthreads();
    for (int j = tid; j < tile_n; j += FI_THREADS) {
        const uint2 v = s_stage[j];
        out[(uint32_t)j + s_delta[(v.y >> sh) & (nsub - 1)]] = make_uint2(v.x, v.y & lowmask);
    }
}

// ---- in-bucket sort -----------------------------------------------------------------------------
#ifdef BWTK_MSD_TIMING
__device__ unsigned long long g_phase_cycles[16];
#define PHASE_MARK(k) do { if (tid == 0) { const long long now_ = clock64(); atomicAdd(&g_phase_cycles[k], (unsigned long long)(now_ - t_prev_)); t_prev_ = now_; } } while (0)
#define PHASE_INIT() long long t_prev_ = clock64()
#else
#define PHASE_MARK(k) do { } while (0)
#define PHASE_INIT() do { } while (0)
#endif

// What the fused form needs besides the sorted pairs (see sa::regroup_kernel<.., true>).
struct Regroup {
    int32_t *sa;
    int32_t *rank;
    int32_t *npos;
    uint32_t *nsuf;
    int32_t *ngrp;
    uint32_t *abits;
    int32_t *ptab;
    unsigned long long *status;   // look-back over active counts, one word per bucket
    unsigned *out_count;
    int *err;
    uint32_t *ovf_abase;          // [NBK] first active-list slot of an oversize bucket
};

struct SumComb64 {
    __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const
    {
        return a + b;
    }
};

// short_from: suffixes from this position on run off the end of the text (singletons by rule; they sort
// before longer suffixes with the same zero-padded key, shortest first)
template <bool FUSED>
__global__ void __launch_bounds__(SORT_THREADS, 2)
    bucket_sort_kernel(const uint2 *__restrict__ pairs, const uint32_t *__restrict__ packed, int64_t n, int nb_bits,
                       int64_t short_from, const uint32_t *__restrict__ bstart, uint32_t *__restrict__ skey,
                       uint32_t *__restrict__ sval, unsigned *ticket, Regroup rg, int refine)
{
    extern __shared__ __align__(16) uint32_t sm[];
    uint32_t *s_pos = sm;                                      // [CAP_EFF + 16] position of the element in slot i
    uint32_t *s_a = sm + CAP_EFF + SHORT_SLOTS;                // [CAP_EFF] sorted elements: (low key bits << 13) | slot
    uint32_t *s_b = sm + 2 * CAP_EFF + SHORT_SLOTS;            // [CAP_EFF] elements grouped by sub-bucket
    uint32_t *s_h = sm + 3 * CAP_EFF + SHORT_SLOTS;            // [SUB + 4] sub-bucket counts -> starts
    __shared__ uint32_t s_wsum[SORT_THREADS / 32];
    __shared__ uint32_t s_c[256];                  // counters of the big-bin counting sort
    __shared__ unsigned short s_list[BIG_LIST];    // bins above BIN_PEER
    __shared__ unsigned short s_sidx[SHORT_SLOTS]; // where short suffixes of big bins landed
    __shared__ uint32_t s_nbig, s_nshort;
    __shared__ unsigned s_ticket;
    __shared__ unsigned long long s_excl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    PHASE_INIT();
    // dynamic bucket ids: a CTA only ever waits (look-back) for buckets whose CTAs already run
    if (tid == 0) s_ticket = FUSED ? atomicAdd(ticket, 1u) : blockIdx.x;
    for (int i = tid; i < SUB + 4; i += SORT_THREADS) s_h[i] = 0;
    if (tid == 0) { s_nbig = 0; s_nshort = 0; }
    __syncthreads();
    const uint32_t b = s_ticket;
    const uint32_t b0 = __ldg(bstart + b);
    const uint32_t cnt = __ldg(bstart + b + 1) - b0;
    const int LB = 32 - nb_bits, DS = LB - SUB_BITS;
    const uint32_t prefix = b << LB;
    const bool sorted_here = cnt > 0 && cnt <= (uint32_t)CAP_EFF;
    {
        // the pairs of the bucket that will run when the CTAs now resident are done: into L2 ahead of time
        const uint32_t pb = b + PREFETCH_AHEAD;
        if (pb < (1u << nb_bits)) {
            const uint32_t p0 = __ldg(bstart + pb), p1 = __ldg(bstart + pb + 1);
            const uint32_t lines = (p1 - p0 + 15u) / 16u;                 // 128-byte lines of 16 pairs
            if ((uint32_t)tid < lines && lines <= (uint32_t)SORT_THREADS)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pairs + p0 + (uint32_t)tid * 16u));
        }
    }

    if (sorted_here) {
        uint32_t e[SORT_ITEMS];
        uint32_t rk[SORT_ITEMS / 2];
#pragma unroll
        for (int k = 0; k < SORT_ITEMS / 2; k++) rk[k] = 0;
        {
            uint2 in[SORT_ITEMS];
#pragma unroll
            for (int k = 0; k < SORT_ITEMS; k++) {
                const uint32_t i = tid + k * SORT_THREADS;
                in[k] = make_uint2(0u, 0u);
                if (i < cnt) in[k] = __ldg(pairs + b0 + i);
            }
#pragma unroll
            for (int k = 0; k < SORT_ITEMS; k++) {
                const uint32_t i = tid + k * SORT_THREADS;
                e[k] = 0xffffffffu;
                if (i < cnt) {
                    const uint32_t pos = in[k].x, low = in[k].y;
                    // slot: load order + 16; a suffix of L < 16 symbols takes slot L - 1
                    const uint32_t slot = (int64_t)pos >= short_from ? (uint32_t)(n - 1 - (int64_t)pos) : i + SHORT_SLOTS;
                    s_pos[slot] = pos;
                    e[k] = (low << IDX_BITS) | slot;
                    const uint32_t r = atomicAdd(&s_h[low >> DS], 1u);   // arrival order inside the sub-bucket
                    rk[k >> 1] |= r << ((k & 1) * 16);
                }
            }
        }
        __syncthreads();
        PHASE_MARK(0);
        // exclusive scan of the SUB counts, 8 consecutive bins per thread; bins above BIN_PEER go on a list
        {
            uint32_t c[8], o[8];
            const uint4 q0 = *reinterpret_cast<const uint4 *>(s_h + tid * 8);
            const uint4 q1 = *reinterpret_cast<const uint4 *>(s_h + tid * 8 + 4);
            c[0] = q0.x; c[1] = q0.y; c[2] = q0.z; c[3] = q0.w; c[4] = q1.x; c[5] = q1.y; c[6] = q1.z; c[7] = q1.w;
            uint32_t sum = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                sum += c[k];
                if (c[k] > (uint32_t)BIN_PEER) s_list[atomicAdd(&s_nbig, 1u)] = (unsigned short)(tid * 8 + k);
            }
            uint32_t inc = sum;
#pragma unroll
            for (int of = 1; of < 32; of <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, inc, of);
                if (lane >= of) inc += t;
            }
            if (lane == 31) s_wsum[warp] = inc;
            __syncthreads();
            uint32_t run = inc - sum;
#pragma unroll
            for (int w = 0; w < SORT_THREADS / 32; w++)
                if (w < warp) run += s_wsum[w];
#pragma unroll
            for (int k = 0; k < 8; k++) { o[k] = run; run += c[k]; }
            *reinterpret_cast<uint4 *>(s_h + tid * 8) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4 *>(s_h + tid * 8 + 4) = make_uint4(o[4], o[5], o[6], o[7]);
            if (tid == SORT_THREADS - 1) s_h[SUB] = run;   // == cnt
        }
        __syncthreads();
        PHASE_MARK(1);
        // group by sub-bucket
#pragma unroll
        for (int k = 0; k < SORT_ITEMS; k++) {
            const uint32_t i = tid + k * SORT_THREADS;
            if (i < cnt) s_b[s_h[e[k] >> (IDX_BITS + DS)] + ((rk[k >> 1] >> ((k & 1) * 16)) & 0xffffu)] = e[k];
        }
        __syncthreads();
        PHASE_MARK(2);
        // sub-buckets of up to BIN_PEER elements: every element counts the peers that sort before it
        for (uint32_t r = tid; r < cnt; r += SORT_THREADS) {
            const uint32_t x = s_b[r];
            const uint32_t d = x >> (IDX_BITS + DS);
            const uint32_t lo = s_h[d], hi = s_h[d + 1];
            const uint32_t sz = hi - lo;
            if (sz <= 4u) {
                // the common case (mean sub-bucket: 1.4 elements) without a loop; s_b has a readable tail
                const uint32_t a0 = s_b[lo], a1 = s_b[lo + 1], a2 = s_b[lo + 2], a3 = s_b[lo + 3];
                uint32_t f = lo + (a0 < x ? 1u : 0u);
                f += (sz > 1u && a1 < x) ? 1u : 0u;
                f += (sz > 2u && a2 < x) ? 1u : 0u;
                f += (sz > 3u && a3 < x) ? 1u : 0u;
                s_a[f] = x;
            } else if (sz <= (uint32_t)BIN_PEER) {
                uint32_t f = lo;
                for (uint32_t j = lo; j < hi; j++) f += s_b[j] < x ? 1u : 0u;
                s_a[f] = x;
            }
        }
        PHASE_MARK(3);
        // larger ones (runs of one 16-mer: equal keys may stay in any order): counting sort on the remaining
        // <= 7 key bits + "not a short suffix", one bin at a time by the whole CTA
        const uint32_t nbig = s_nbig;
        for (uint32_t li = 0; li < nbig; li++) {
            const uint32_t bin = s_list[li];
            const uint32_t lo = s_h[bin], g = s_h[bin + 1] - lo;
            if (tid < 256) s_c[tid] = 0;
            __syncthreads();
            for (uint32_t i = tid; i < g; i += SORT_THREADS) {
                const uint32_t x = s_b[lo + i];
                atomicAdd(&s_c[(((x >> IDX_BITS) & ((1u << DS) - 1u)) << 1) | ((x & (CAP - 1)) >= (uint32_t)SHORT_SLOTS ? 1u : 0u)], 1u);
            }
            __syncthreads();
            if (warp == 0) {
                uint32_t v[8], sum = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) { v[k] = s_c[lane * 8 + k]; sum += v[k]; }
                uint32_t inc = sum;
#pragma unroll
                for (int of = 1; of < 32; of <<= 1) {
                    uint32_t t = __shfl_up_sync(0xffffffffu, inc, of);
                    if (lane >= of) inc += t;
                }
                uint32_t run = inc - sum;
#pragma unroll
                for (int k = 0; k < 8; k++) { s_c[lane * 8 + k] = run; run += v[k]; }
            }
            __syncthreads();
            for (uint32_t i = tid; i < g; i += SORT_THREADS) {
                const uint32_t x = s_b[lo + i];
                const bool is_long = (x & (CAP - 1)) >= (uint32_t)SHORT_SLOTS;
                const uint32_t r = atomicAdd(&s_c[(((x >> IDX_BITS) & ((1u << DS) - 1u)) << 1) | (is_long ? 1u : 0u)], 1u);
                s_a[lo + r] = x;
                if (!is_long) s_sidx[atomicAdd(&s_nshort, 1u)] = (unsigned short)(lo + r);
            }
            __syncthreads();
        }
        __syncthreads();
        if (tid == 0 && s_nshort >= 2) {
            // short suffixes with one key sit next to each other, before the long ones, but in arrival order:
            // the set of their places is right, so sort places and values separately
            const int ns = (int)s_nshort;
            uint32_t idx[SHORT_SLOTS], val[SHORT_SLOTS];
            for (int i = 0; i < ns; i++) { idx[i] = s_sidx[i]; val[i] = s_a[idx[i]]; }
            for (int i = 1; i < ns; i++) {
                const uint32_t vi = idx[i], vv = val[i];
                int j = i;
                while (j > 0 && idx[j - 1] > vi) { idx[j] = idx[j - 1]; j--; }
                idx[j] = vi;
                j = i;
                while (j > 0 && val[j - 1] > vv) { val[j] = val[j - 1]; j--; }
                val[j] = vv;
            }
            for (int i = 0; i < ns; i++) s_a[idx[i]] = val[i];
        }
        __syncthreads();
        PHASE_MARK(4);
        if (!FUSED) {
            for (uint32_t j = tid; j < cnt; j += SORT_THREADS) {
                const uint32_t x = s_a[j];
                skey[b0 + j] = prefix | (x >> IDX_BITS);
                sval[b0 + j] = s_pos[x & (CAP - 1)];
            }
            PHASE_MARK(5);
        }
    }
    if (!FUSED) return;

    // ---- fused regroup: heads, active list, ranks, prefix table -----------------------------------
    // s_h is free again: [0, 256) head bits of the 16-symbol groups, [256, 512) active bits, [512, 768) last head at
    // or before the start of each 32-element word, [768, 1024) active elements before each word, [1024, 1280) head
    // bits after the in-bucket refinement (what everything below uses); s_b holds the second keys.
    __syncthreads();
    uint32_t *s_hb0 = s_h, *s_ab = s_h + 256, *s_wh = s_h + 512, *s_wa = s_h + 768, *s_hb = s_h + 1024;
    uint32_t n_active = 0;
    if (sorted_here) {
        const int words = (int)((cnt + 31) >> 5);
        // head flags: element j is a head when its key differs from its predecessor's or either of them is a
        // short suffix (slots past the end count as heads: the bucket's last element is followed by one);
        // the same compare finds where the 16-bit key prefix changes (prefix table, < 16 bucket bits only)
        const int psub = 16 - nb_bits;
        for (uint32_t j0 = warp * 32; j0 < cnt; j0 += SORT_THREADS) {
            const uint32_t j = j0 + lane;
            bool head = true;
            if (j < cnt) {
                const uint32_t x = s_a[j];
                const uint32_t p = j > 0 ? s_a[j - 1] : ~x;
                head = j == 0 || (p >> IDX_BITS) != (x >> IDX_BITS) || (x & (CAP - 1)) < (uint32_t)SHORT_SLOTS ||
                       (p & (CAP - 1)) < (uint32_t)SHORT_SLOTS;
                if (psub > 0) {
                    const int32_t lp = (int32_t)((x >> IDX_BITS) >> (LB - psub));
                    const int32_t pp = j > 0 ? (int32_t)((p >> IDX_BITS) >> (LB - psub)) : -1;
                    for (int32_t q = pp + 1; q <= lp; q++) rg.ptab[(b << psub) + q] = (int32_t)(b0 + j);
                    if (j + 1 == cnt)
                        for (int32_t q = lp + 1; q < (1 << psub); q++) rg.ptab[(b << psub) + q] = (int32_t)(b0 + cnt);
                }
            }
            const unsigned hb = __ballot_sync(0xffffffffu, head);
            if (lane == 0) { s_hb0[j0 >> 5] = hb; s_hb[j0 >> 5] = hb; }
        }
        __syncthreads();
        if (refine) {
            // Round 1 of the doubling (h = 16) for the small groups, here: a group of 2..32 suffixes with equal keys
            // that lies inside one 32-element window of the sorted bucket is ordered by the NEXT 16 symbols (the
            // round-0 key of suffix + 16, which is what rank[suffix + 16] orders by) and split where they differ.
            // Members that become singletons are final and never enter the active list; groups that straddle a
            // window, larger groups and groups that touch the end of the text stay 16-symbol groups.  The global
            // rounds that follow are unchanged -- a group that is already finer than the round's h is only sorted
            // again by ranks that are finer still.  Everything is warp-local: one bitmap word = one window.
            for (uint32_t j0 = warp * 32; j0 < cnt; j0 += SORT_THREADS) {
                const uint32_t w = j0 >> 5;
                const uint32_t hbw = s_hb0[w];                                   // padding lanes of the last window are heads
                const bool next_head = (w + 1 >= (uint32_t)words) || (s_hb0[w + 1] & 1u);
                if (hbw == 0xffffffffu && next_head) {                           // singletons only
                    if (lane == 0) s_hb[w] = hbw;
                    continue;
                }
                const uint32_t below = hbw & (0xffffffffu >> (31 - lane));       // heads at or before this lane
                const uint32_t above = lane < 31 ? (hbw >> (lane + 1)) : 0u;      // heads after it
                const bool inside = below != 0u && (above != 0u || next_head);  // the whole group lies in this window
                const uint32_t lo = below ? 31u - (uint32_t)__clz(below) : 0u;
                const uint32_t hi = above ? lane + 1u + (uint32_t)(__ffs(above) - 1) : 32u;
                const uint32_t size = hi - lo;
                const uint32_t j = j0 + lane;
                const uint32_t x = j < cnt ? s_a[j] : 0u;
                // second key of the members of whole groups of >= 2: the round-0 key of suffix + 16 (all ones also
                // stands for "touches the end of the text")
                uint32_t k2 = 0u;
                if (j < cnt && inside && size >= 2u) {
                    const int64_t q = (int64_t)s_pos[x & (CAP - 1)] + 16;
                    k2 = q < short_from ? window32(packed, q * 2) : 0xffffffffu;
                }
                // a group is refined when it is whole, has >= 2 members, and none of them touches the end of the text
                // (a true second key of all ones cannot tie with a smaller one, so treating it as "touches" only
                // leaves that group to the global round)
                const unsigned touch = __ballot_sync(0xffffffffu, j < cnt && k2 == 0xffffffffu);
                const uint32_t gmask = size >= 32u ? 0xffffffffu : (((1u << size) - 1u) << lo);
                const bool refine_me = j < cnt && inside && size >= 2u && (touch & gmask) == 0u;
                const unsigned any = __ballot_sync(0xffffffffu, refine_me);
                uint32_t newpos = lane;
                bool head2 = (hbw >> lane) & 1u;
                if (any) {
                    const uint32_t maxsize = __reduce_max_sync(0xffffffffu, refine_me ? size : 0u);
                    uint32_t r = 0, eqbefore = 0;
                    for (uint32_t d = 0; d < maxsize; d++) {
                        const uint32_t xl = (lo + d) & 31u;
                        const uint32_t kx = __shfl_sync(0xffffffffu, k2, xl);
                        const bool valid = refine_me && d < size;
                        r += (valid && (kx < k2 || (kx == k2 && xl < (uint32_t)lane))) ? 1u : 0u;
                        eqbefore += (valid && kx == k2 && xl < (uint32_t)lane) ? 1u : 0u;
                    }
                    if (refine_me) {
                        newpos = lo + r;
                        head2 = eqbefore == 0u;          // first of its (key, second key) class
                    }
                }
                const uint32_t hb2 = __reduce_or_sync(0xffffffffu, head2 ? (1u << newpos) : 0u);
                if (any) {
                    __syncwarp();
                    if (refine_me) s_a[j0 + newpos] = x;     // a permutation inside the group: every lane has read its x
                }
                if (lane == 0) s_hb[w] = hb2;
            }
            __syncthreads();
        }
        // per-word carries: 256 words at most, 8 per lane of warp 0 (max-scan of heads, sum-scan of actives);
        // an element is active unless it and its successor are both heads
        if (warp == 0) {
            uint32_t lh[8], la[8];
            uint32_t mh = 0, sa_ = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int w = lane * 8 + k;
                uint32_t hbw = 0u, abw = 0u;
                if (w < words) {
                    hbw = s_hb[w];
                    const uint32_t nxt = w + 1 < words ? s_hb[w + 1] : 0xffffffffu;
                    const uint32_t valid = (uint32_t)(w + 1) * 32u <= cnt ? 0xffffffffu : ((1u << (cnt & 31u)) - 1u);
                    abw = ~(hbw & ((hbw >> 1) | (nxt << 31))) & valid;
                    s_ab[w] = abw;
                    hbw &= valid;
                }
                lh[k] = hbw ? (uint32_t)(w * 32 + 31 - __clz(hbw)) + 1u : 0u;   // last head in the word (+1), 0: none
                la[k] = (uint32_t)__popc(abw);
                mh = lh[k] > mh ? lh[k] : mh;
                sa_ += la[k];
            }
            uint32_t imh = mh, isa = sa_;
#pragma unroll
            for (int of = 1; of < 32; of <<= 1) {
                const uint32_t t1 = __shfl_up_sync(0xffffffffu, imh, of), t2 = __shfl_up_sync(0xffffffffu, isa, of);
                if (lane >= of) { imh = t1 > imh ? t1 : imh; isa += t2; }
            }
            uint32_t ch = __shfl_up_sync(0xffffffffu, imh, 1), ca = isa - sa_;
            if (lane == 0) ch = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int w = lane * 8 + k;
                if (w < words) { s_wh[w] = ch; s_wa[w] = ca; }
                ch = lh[k] > ch ? lh[k] : ch;
                ca += la[k];
            }
            n_active = __shfl_sync(0xffffffffu, isa, 31);
        }
    } else if (cnt > (uint32_t)CAP_EFF) {
        // oversize bucket: sorted by the LSD pass afterwards; all of its suffixes go on the active list
        // (ovf_regroup_kernel fills the slots reserved here)
        n_active = cnt;
    }
    // chain the active counts bucket to bucket (every CTA publishes, also for empty buckets): the aggregate goes
    // out first, then everything that does not need the bucket's offset in the active list is written, and only
    // then does warp 0 look back (its predecessors have had that time to publish)
    __syncthreads();   // n_active is warp 0's; the bitmaps are complete
    volatile unsigned long long *lb_st = rg.status;
    const unsigned long long F_AGG = 1ull << 62, F_INCL = 2ull << 62;
    if (tid == 0) lb_st[b] = (b == 0 ? F_INCL : F_AGG) | n_active;
    // 16-bit prefix table (fewer than 16 bucket bits only): an empty bucket's prefixes all start where the bucket would
    // (sorted buckets wrote theirs with the head flags, oversize ones do in ovf_regroup_kernel)
    if (nb_bits < 16 && cnt == 0) {
        const int sub = 16 - nb_bits;
        for (uint32_t q = tid; q < (1u << sub); q += SORT_THREADS) rg.ptab[(b << sub) + q] = (int32_t)b0;
    }
    if (nb_bits < 16 && b + 1 == (1u << nb_bits) && tid == 0) rg.ptab[65536] = (int32_t)n;
    if (sorted_here) {
        for (uint32_t j = tid; j < cnt; j += SORT_THREADS) {
            const uint32_t x = s_a[j];
            const uint32_t s = s_pos[x & (CAP - 1)];
            const uint32_t t = b0 + j;
            skey[t] = prefix | (x >> IDX_BITS);
            rg.sa[t] = (int32_t)s;
            const uint32_t w = j >> 5, bit = j & 31;
            if ((s_ab[w] >> bit) & 1u) {
                const uint32_t hm = s_hb[w] & (0xffffffffu >> (31 - bit));
                const uint32_t hl = hm ? (w * 32 + 31 - __clz(hm)) : (s_wh[w] - 1u);   // a head exists at or before j
                rg.rank[s] = (int32_t)(b0 + hl);
                atomicOr(rg.abits + (s >> 5), 1u << (s & 31));
            }
        }
    }
    PHASE_MARK(6);
    if (warp == 0) {
        unsigned long long excl = 0;
        if (b != 0) {
            excl = warp_lookback(lb_st, (int64_t)b, SumComb64{}, rg.err, 4, rsort::SPIN_LIMIT);
            if (lane == 0) lb_st[b] = F_INCL | (excl + n_active);
        }
        if (lane == 0) {
            s_excl = excl;
            if (cnt > (uint32_t)CAP_EFF) rg.ovf_abase[b] = (uint32_t)excl;
            if (b + 1 == (1u << nb_bits)) *rg.out_count = (unsigned)(excl + n_active);
        }
    }
    __syncthreads();
    PHASE_MARK(7);
    if (!sorted_here) return;
    const uint32_t abase = (uint32_t)s_excl;
    // the active list: only the words of the active bitmap that have a bit set
    const int words = (int)((cnt + 31) >> 5);
    for (int w = warp; w < words; w += SORT_THREADS / 32) {
        const uint32_t abw = s_ab[w];
        if (!((abw >> lane) & 1u)) continue;
        const uint32_t j = (uint32_t)w * 32 + lane;
        const uint32_t x = s_a[j];
        const uint32_t hm = s_hb[w] & (0xffffffffu >> (31 - lane));
        const uint32_t hl = hm ? ((uint32_t)w * 32 + 31 - __clz(hm)) : (s_wh[w] - 1u);
        const uint32_t cidx = abase + s_wa[w] + __popc(abw & ((1u << lane) - 1u));
        rg.npos[cidx] = (int32_t)(b0 + j);
        rg.nsuf[cidx] = s_pos[x & (CAP - 1)];
        rg.ngrp[cidx] = (int32_t)(b0 + hl);
    }
    PHASE_MARK(8);
}

constexpr size_t sort_smem() { return (size_t)(3 * CAP_EFF + SHORT_SLOTS + SUB + 4) * 4; }

// ---- oversize buckets through the LSD sort --------------------------------------------------------
// 37-bit keys (key << 5 | tie) of the suffixes of the oversize buckets, in bucket order; tie = L - 1 for a suffix of
// L < 16 symbols (it sorts before longer suffixes with the same zero-padded key, shortest first), 31 otherwise
// (equal keys of long suffixes may stay in any order: the doubling rounds sort them)
constexpr int OVF_TIE_BITS = 5;
__global__ void ovf_gather_kernel(const uint2 *__restrict__ pairs, const uint32_t *__restrict__ ovf_list, unsigned nlist,
                                  const uint32_t *__restrict__ bstart, const uint32_t *__restrict__ ovf_prefix, int nb_bits,
                                  int64_t n, int64_t short_from, unsigned long long *__restrict__ k, uint32_t *__restrict__ v)
{
    for (unsigned li = blockIdx.y; li < nlist; li += gridDim.y) {
        const uint32_t b = ovf_list[li];
        const uint32_t b0 = bstart[b], cnt = bstart[b + 1] - b0, o = ovf_prefix[b];
        const uint32_t prefix = b << (32 - nb_bits);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
            const uint2 e = __ldg(pairs + b0 + i);
            const unsigned long long tie = (int64_t)e.x >= short_from ? (unsigned long long)(n - 1 - (int64_t)e.x) : 31ull;
            k[o + i] = ((unsigned long long)(prefix | e.y) << OVF_TIE_BITS) | tie;
            v[o + i] = e.x;
        }
    }
}

__global__ void ovf_place_kernel(const unsigned long long *__restrict__ sk, const uint32_t *__restrict__ sv, int64_t m,
                                 int shift, const uint32_t *__restrict__ bstart, const uint32_t *__restrict__ ovf_prefix,
                                 uint32_t *__restrict__ skey, uint32_t *__restrict__ sval)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t key = (uint32_t)(sk[j] >> OVF_TIE_BITS);
    const uint32_t b = key >> shift;
    const uint32_t dst = __ldg(bstart + b) + ((uint32_t)j - __ldg(ovf_prefix + b));
    skey[dst] = key;
    sval[dst] = sv[j];
}

// Fused form of the oversize pass: the sorted (key, ~position) list of the oversize buckets straight into SA, the
// sorted keys, and the active-list slots bucket_sort_kernel<true> reserved -- every suffix of an oversize bucket is
// listed (singletons as groups of one; the first doubling round drops them again).  Group head = first entry of
// the run of equal keys (binary search), after the suffixes that run off the text, which are singletons.
__global__ void ovf_regroup_kernel(const unsigned long long *__restrict__ sk, const uint32_t *__restrict__ sv, int64_t m,
                                   int nb_bits, const uint32_t *__restrict__ bstart, const uint32_t *__restrict__ ovf_prefix,
                                   int64_t short_from, uint32_t *__restrict__ skey, Regroup rg)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int shift = 32 - nb_bits;
    const uint32_t key = (uint32_t)(sk[j] >> OVF_TIE_BITS), pos = sv[j];
    const uint32_t b = key >> shift;
    const uint32_t o = __ldg(ovf_prefix + b), b0 = __ldg(bstart + b), cnt = __ldg(bstart + b + 1) - b0;
    const uint32_t i = (uint32_t)j - o, t = b0 + i;
    skey[t] = key;
    rg.sa[t] = (int32_t)pos;
    int64_t lo = o, hi = j;
    const unsigned long long target = (unsigned long long)key << OVF_TIE_BITS;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sk[mid] < target) lo = mid + 1; else hi = mid;
    }
    int ns = 0;
    while (ns < SHORT_SLOTS && lo + ns < m && (uint32_t)(sk[lo + ns] >> OVF_TIE_BITS) == key && (int64_t)sv[lo + ns] >= short_from) ns++;
    const int64_t hj = (int64_t)pos >= short_from ? j : lo + ns;
    const int32_t hp = (int32_t)(b0 + (uint32_t)(hj - o));
    const uint32_t c = rg.ovf_abase[b] + i;
    rg.npos[c] = (int32_t)t;
    rg.nsuf[c] = pos;
    rg.ngrp[c] = hp;
    rg.rank[pos] = hp;
    atomicOr(rg.abits + (pos >> 5), 1u << (pos & 31));
    if (nb_bits < 16) {
        const int sub = 16 - nb_bits;
        const int32_t per = 1 << sub;
        const int32_t lp = (int32_t)((key >> (shift - sub)) & (uint32_t)(per - 1));
        const int32_t pp = i > 0 ? (int32_t)(((uint32_t)(sk[j - 1] >> OVF_TIE_BITS) >> (shift - sub)) & (uint32_t)(per - 1)) : -1;
        for (int32_t q = pp + 1; q <= lp; q++) rg.ptab[(b << sub) + q] = (int32_t)t;
        if (i + 1 == cnt)
            for (int32_t q = lp + 1; q < per; q++) rg.ptab[(b << sub) + q] = (int32_t)(b0 + cnt);
    }
}

// BWTK_MSD_REFINE=0 switches the in-bucket refinement of small groups off (tuning and tests)
static bool refine_enabled()
{
    const char *e = getenv("BWTK_MSD_REFINE");
    return !(e && atoi(e) == 0);
}

// Sorts the n suffixes of a 2-bit packed text by their 32-bit round-0 key (ties: suffixes that run off the
// text first, shortest first; the rest in no particular order).
//   fused == nullptr : sorted keys -> skey, suffixes -> sval (what the LSD sort delivers);
//   fused != nullptr : as above for skey, and the whole first regroup pass (sval is not written); *did_fuse tells.
// Returns BWTK_OK with *done = false when the text is too skewed for this path (nothing written).
// pairs_a / pairs_b: n (position, key) pairs each; tmp_v0 / tmp_v1: n-element scratch for the oversize pass
// (its 64-bit keys reuse pairs_a and pairs_b).
static int round0_sort(const uint32_t *packed, int64_t n, const Workspace &ws, const rsort::Workspace &rws,
                       uint2 *pairs_a, uint2 *pairs_b, uint32_t *skey, uint32_t *sval, uint32_t *tmp_v0,
                       uint32_t *tmp_v1, const Regroup *fused, int32_t *ptab, int64_t short_from, cudaStream_t st,
                       bool *done, bool *did_fuse, int64_t *passes_out)
{
    *done = false;
    *did_fuse = false;
    const int nb = nb_bits_for(n);
    const int64_t nbk = 1ll << nb;
    const int64_t nwords = ceil_div(n, 16);
    static bool attr = false;
    if (!attr) {
        BWTK_CUDA(cudaFuncSetAttribute(bucket_sort_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem()));
        BWTK_CUDA(cudaFuncSetAttribute(bucket_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem()));
        BWTK_CUDA(cudaFuncSetAttribute(coarse_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)coarse_smem()));
        BWTK_CUDA(cudaFuncSetAttribute(fine_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fine_smem()));
        attr = true;
    }
    BWTK_CUDA(bwtk::zero_async(ws.chist, 260 * 4, st));
    BWTK_CUDA(bwtk::zero_async(ws.hist, (size_t)(nbk + 2) * 4, st));
    {
        prof::Scope ps("coarse_hist", rsort::packed_hist_bytes(n, 2), st);
        int cgrid = (int)(ceil_div(n, 512 * 64) < NUM_SMS * 4 ? ceil_div(n, 512 * 64) : NUM_SMS * 4);
        rsort::chunk_hist_kernel<<<cgrid < 1 ? 1 : cgrid, 512, 0, st>>>(packed, n, 2, ws.chist);
        BWTK_LAUNCH_CHECK();
    }
    coarse_scan_kernel<<<1, COARSE, 0, st>>>(ws.chist, ws.cstart, ws.cfill, ws.ftile0, ws.htile0);
    BWTK_LAUNCH_CHECK();
    {
        prof::Scope ps("coarse_scatter", rsort::packed_hist_bytes(n, 2) + n * 8, st);
        coarse_scatter_kernel<<<(unsigned)ceil_div(nwords, CA_WORDS), CA_THREADS, coarse_smem(), st>>>(packed, n, ws.cfill,
                                                                                                      pairs_a);
        BWTK_LAUNCH_CHECK();
    }
    {
        prof::Scope ps("fine_hist", n * 4, st);
        fine_hist_kernel<<<(unsigned)(ceil_div(n, FH_CHUNK) + COARSE), 512, 0, st>>>(pairs_a, ws.cstart, ws.htile0, nb, ws.hist);
        BWTK_LAUNCH_CHECK();
    }
    bucket_scan_kernel<<<1, 1024, 0, st>>>(ws.hist, nbk, (uint32_t)n, ws.bstart, ws.fill, ws.ovf_prefix, ws.ovf_list,
                                          ptab, nb, ws.info);
    BWTK_LAUNCH_CHECK();
    Info info;
    { int rc = read_back(&info, ws.info, sizeof(Info), st); if (rc) return rc; }
    if ((int64_t)info.ovf_elems * 2 > n) return BWTK_OK;   // mostly low-complexity: the plain LSD sort is the better path
    {
        prof::Scope ps("fine_scatter", n * 16, st);
        fine_scatter_kernel<<<(unsigned)(ceil_div(n, FI_TILE) + COARSE), FI_THREADS, fine_smem(), st>>>(
            pairs_a, ws.cstart, ws.ftile0, nb, ws.fill, pairs_b);
        BWTK_LAUNCH_CHECK();
    }
    const int64_t m = info.ovf_elems;
    unsigned long long *ok0 = reinterpret_cast<unsigned long long *>(pairs_a);   // free once fine_scatter has run
    if (info.ovf_buckets) {
        prof::Scope ps("ovf_gather", m * 20, st);
        dim3 grid((unsigned)(ceil_div(info.max_bucket, 256) < 64 ? ceil_div(info.max_bucket, 256) : 64),
                  info.ovf_buckets < 1024 ? info.ovf_buckets : 1024);
        ovf_gather_kernel<<<grid, 256, 0, st>>>(pairs_b, ws.ovf_list, info.ovf_buckets, ws.bstart, ws.ovf_prefix, nb, n,
                                                short_from, ok0, tmp_v0);
        BWTK_LAUNCH_CHECK();
    }
    const bool fuse = fused != nullptr;
    Regroup rgf{};
    if (fuse) {
        BWTK_CUDA(bwtk::zero_async(ws.lb_status, (size_t)nbk * 8, st));
        BWTK_CUDA(bwtk::zero_async(ws.ticket, sizeof(unsigned), st));
        Regroup rg = *fused;
        rg.status = ws.lb_status;
        rg.ovf_abase = ws.fill;    // the claim cursors are dead once fine_scatter has run
        rgf = rg;
        prof::Scope ps("bucket_sort_regroup", n * 16, st);
        bucket_sort_kernel<true><<<(unsigned)nbk, SORT_THREADS, sort_smem(), st>>>(pairs_b, packed, n, nb, short_from,
                                                                                   ws.bstart, skey, sval, ws.ticket, rg,
                                                                                   refine_enabled() ? 1 : 0);
        BWTK_LAUNCH_CHECK();
        *did_fuse = true;
    } else {
        Regroup rg{};
        prof::Scope ps("bucket_sort", n * 16, st);
        bucket_sort_kernel<false><<<(unsigned)nbk, SORT_THREADS, sort_smem(), st>>>(pairs_b, packed, n, nb, short_from,
                                                                                    ws.bstart, skey, sval, ws.ticket, rg, 0);
        BWTK_LAUNCH_CHECK();
    }
    if (info.ovf_buckets) {
        unsigned long long *ok1 = reinterpret_cast<unsigned long long *>(pairs_b);   // free once the buckets are sorted
        int in_first = 1;
        int rc = rsort::sort_pairs<uint64_t>(reinterpret_cast<uint64_t *>(ok0), tmp_v0, reinterpret_cast<uint64_t *>(ok1),
                                             tmp_v1, m, 0, 32 + OVF_TIE_BITS, rws, st, &in_first, passes_out);
        if (rc) return rc;
        if (fuse) {
            prof::Scope ps("ovf_regroup", m * 36, st);
            ovf_regroup_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(in_first ? ok0 : ok1, in_first ? tmp_v0 : tmp_v1, m, nb,
                                                                          ws.bstart, ws.ovf_prefix, short_from, skey, rgf);
            BWTK_LAUNCH_CHECK();
        } else {
            prof::Scope ps("ovf_place", m * 20, st);
            ovf_place_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(in_first ? ok0 : ok1, in_first ? tmp_v0 : tmp_v1, m,
                                                                        32 - nb, ws.bstart, ws.ovf_prefix, skey, sval);
            BWTK_LAUNCH_CHECK();
        }
    }
#ifdef BWTK_MSD_TIMING
    {
        unsigned long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_phase_cycles, sizeof(h));
        fprintf(stderr, "bucket_sort phase cycles (thread 0, summed over CTAs): load %llu scan %llu place %llu peer %llu bigbins %llu out %llu | fused: flags %llu lookback %llu out %llu\n",
                h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8]);
        memset(h, 0, sizeof(h));
        cudaMemcpyToSymbol(g_phase_cycles, h, sizeof(h));
    }
#endif
    *done = true;
    return BWTK_OK;
}

}  // namespace msd
}  // namespace bwtk
