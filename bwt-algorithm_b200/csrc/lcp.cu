// lcp.cu -- LCP array (reference bwt.py:55-72 `_kasai_lcp_uint8`, 2108-2116):
// lcp[0] = 0, lcp[j] = LCP(suffix sa[j-1], suffix sa[j]).
//
// Two stages.
//  1. Every pair is compared directly on the bit-packed text, 64 stream bits per step.  In the
//     fused build the sorted round-0 keys are at hand, so wherever two neighbouring keys differ
//     (94 % of a genome) the answer is the common bit prefix of two coalesced reads.  A pair gets a
//     budget of DEEP_STEPS steps (256 bases at 2 bits); what it has verified by then is stored as
//     a lower bound and the pair goes on a list.
//  2. The listed ("deep") pairs are finished in TEXT order, which is where Kasai's invariant
//     lives: with Phi(i) = sa[isa[i]-1] and PLCP[i] = LCP(i, Phi(i)), PLCP[i+k] >= PLCP[i] - k.
//     The list is sorted by text position (radix sort of (sa[j], j)), cut into chunks, one warp per
//     chunk (one CTA per chunk; its first entry is finished by all threads together, 8192 bases per
//     step), 32 entries per batch: every lane starts from max(own lower bound, previous exact value
//     - distance) and extends.  A lane that is still not done after another DEEP_STEPS is finished
//     by the WHOLE warp (32 windows = 1024 bases per step), in list order, so that the entries
//     behind it inherit its exact value; with that a jump of the PLCP by L costs L/1024 steps once
//     instead of L/32 steps per pair.  Total work is O(n) windows as in Kasai's algorithm; the
//     unbounded case of stage 1 alone (a megabase N block or satellite array: run^2/32) is gone.
#include "common.cuh"
#include "radix_sort.cuh"

namespace bwtk {

namespace {

constexpr int DEEP_STEPS = 8;      // direct-compare budget of a pair, in 64-bit windows
constexpr int DEEP_CHUNK = 8192;   // list entries per CTA in stage 2
constexpr int PLCP_THREADS = 256;
constexpr int LCPK_THREADS = 256;
constexpr int LCPK_ITEMS = 8;

struct DeepList {
    unsigned *count;
    uint32_t *j;       // SA positions of the unfinished pairs, any order
};

// Extends the common prefix of suffixes a and b from h for at most `budget` windows.
// Returns true when the LCP is final (mismatch found or limit reached); h is updated either way.
__device__ __forceinline__ bool extend(const uint32_t *__restrict__ packed, int64_t a, int64_t b, int64_t limit,
                                       int bits, int budget, int64_t &h)
{
    const int per = 64 / bits;
    for (int s = 0; s < budget; s++) {
        if (h >= limit) { h = limit; return true; }
        const uint64_t x = window64(packed, (a + h) * bits) ^ window64(packed, (b + h) * bits);
        if (x) {
            h += __clzll((long long)x) / bits;
            if (h > limit) h = limit;
            return true;
        }
        h += per;
    }
    if (h >= limit) { h = limit; return true; }
    return false;
}

// the same, by all 32 lanes of a warp on one pair (all arguments warp-uniform); always final
__device__ __forceinline__ int64_t extend_warp(const uint32_t *__restrict__ packed, int64_t a, int64_t b,
                                               int64_t limit, int bits, int64_t h)
{
    const int per = 64 / bits;
    const unsigned lane = lane_id();
    while (true) {
        const int64_t off = h + (int64_t)lane * per;
        uint64_t x = 0;
        if (off < limit) x = window64(packed, (a + off) * bits) ^ window64(packed, (b + off) * bits);
        const bool hit = off >= limit || x != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal) {
            const int f = __ffs(bal) - 1;
            int64_t mine = off >= limit ? limit : off + __clzll((long long)x) / bits;
            if (mine > limit) mine = limit;
            return __shfl_sync(0xffffffffu, mine, f);
        }
        h += 32 * (int64_t)per;
    }
}

__device__ __forceinline__ void push_deep(const DeepList &dl, bool deep, int64_t j)
{
    const unsigned act = __activemask();
    const unsigned bal = __ballot_sync(act, deep);
    if (!bal) return;
    const unsigned lane = lane_id();
    const int leader = __ffs(bal) - 1;
    unsigned at = 0;
    if ((int)lane == leader) at = atomicAdd(dl.count, (unsigned)__popc(bal));
    at = __shfl_sync(act, at, leader);
    if (deep) dl.j[at + __popc(bal & lanemask_lt())] = (uint32_t)j;
}

// stage 1 without keys (stand-alone bwtk_lcp_build)
__global__ void __launch_bounds__(256)
    lcp_kernel(const uint32_t *__restrict__ packed, const int32_t *__restrict__ sa, int64_t n, int bits,
               int slack, int32_t *__restrict__ lcp, DeepList dl)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool deep = false;
    if (j < n) {
        if (j == 0) lcp[0] = 0;
        else {
            const int64_t a = __ldg(sa + j - 1), b = __ldg(sa + j);
            const int64_t limit = n - slack - (a > b ? a : b);
            int64_t h = 0;
            deep = !extend(packed, a, b, limit, bits, DEEP_STEPS, h);
            lcp[j] = (int32_t)h;
        }
    }
    push_deep(dl, deep, j);
}

// stage 1 with the sorted round-0 keys (fused index build): skey[j] is the first S symbols of
// suffix sa[j].  Pairs with equal keys and the few suffixes whose window touches the end of the
// text are compacted into shared memory and compared window by window by all threads of the CTA.
__global__ void __launch_bounds__(LCPK_THREADS)
    lcp_keys_kernel(const uint32_t *__restrict__ packed, const int32_t *__restrict__ sa,
                    const uint32_t *__restrict__ skey, int64_t n, int bits, int slack, int32_t *__restrict__ lcp,
                    DeepList dl)
{
    __shared__ int s_list[LCPK_THREADS * LCPK_ITEMS];
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * (LCPK_THREADS * LCPK_ITEMS);
    const int S = 32 / bits;
#pragma unroll
    for (int it = 0; it < LCPK_ITEMS; it++) {
        const int local = it * LCPK_THREADS + threadIdx.x;
        const int64_t j = base + local;
        bool slow = false;
        if (j < n) {
            if (j == 0) {
                lcp[0] = 0;
            } else {
                const uint32_t x = __ldg(skey + j - 1) ^ __ldg(skey + j);
                const int64_t a = __ldg(sa + j - 1), b = __ldg(sa + j);
                if (x != 0u && (a > b ? a : b) < n - S) lcp[j] = __clz((int)x) / bits;
                else slow = true;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, slow);
        if (bal) {
            int at = 0;
            if ((threadIdx.x & 31) == 0) at = atomicAdd(&s_cnt, __popc(bal));
            at = __shfl_sync(0xffffffffu, at, 0);
            if (slow) s_list[at + __popc(bal & lanemask_lt())] = local;
        }
    }
    __syncthreads();
    const int cnt = s_cnt;
    for (int q0 = 0; q0 < cnt; q0 += LCPK_THREADS) {
        const int q = q0 + threadIdx.x;
        bool deep = false;
        int64_t j = 0;
        if (q < cnt) {
            j = base + s_list[q];
            const int64_t a = __ldg(sa + j - 1), b = __ldg(sa + j);
            const int64_t limit = n - slack - (a > b ? a : b);
            int64_t h = 0;
            deep = !extend(packed, a, b, limit, bits, DEEP_STEPS, h);
            lcp[j] = (int32_t)h;
        }
        push_deep(dl, deep, j);
    }
}

// stage 2 ---------------------------------------------------------------------------
__global__ void deep_keys_kernel(const int32_t *__restrict__ sa, const uint32_t *__restrict__ dj, int64_t m,
                                 uint32_t *__restrict__ key, uint32_t *__restrict__ val)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const uint32_t j = dj[k];
    key[k] = (uint32_t)__ldg(sa + j);
    val[k] = j;
}

// the same by a whole CTA (blockDim.x windows per step); every thread returns the exact value
__device__ __forceinline__ int64_t extend_cta(const uint32_t *__restrict__ packed, int64_t a, int64_t b, int64_t limit,
                                              int bits, int64_t h, unsigned long long *s_best)
{
    const int per = 64 / bits;
    const unsigned long long NONE = ~0ull;
    while (true) {
        if (threadIdx.x == 0) *s_best = NONE;
        __syncthreads();
        const int64_t off = h + (int64_t)threadIdx.x * per;
        uint64_t x = 0;
        if (off < limit) x = window64(packed, (a + off) * bits) ^ window64(packed, (b + off) * bits);
        if (off >= limit || x != 0) {
            int64_t mine = off >= limit ? limit : off + __clzll((long long)x) / bits;
            if (mine > limit) mine = limit;
            atomicMin(s_best, (unsigned long long)mine);
        }
        __syncthreads();
        const unsigned long long best = *s_best;
        __syncthreads();
        if (best != NONE) return (int64_t)best;
        h += (int64_t)blockDim.x * per;
    }
}

// pos[] ascending text positions i = sa[j] of the deep pairs, jdx[] their SA positions j.
// One CTA per DEEP_CHUNK list entries: all warps finish the chunk's first entry together (a start from
// scratch can be megabases deep), then every warp walks its own slice, 32 entries per batch, carrying
// the previous exact value; the chunk's first value bounds every slice's first batch.
__global__ void __launch_bounds__(PLCP_THREADS)
    plcp_kernel(const uint32_t *__restrict__ packed, const int32_t *__restrict__ sa, const uint32_t *__restrict__ pos,
                const uint32_t *__restrict__ jdx, int64_t m, int64_t n, int bits, int slack,
                int32_t *__restrict__ lcp)
{
    __shared__ unsigned long long s_best;
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * DEEP_CHUNK;
    if (c0 >= m) return;
    int64_t carry_i, carry_h;
    {
        const int64_t i = __ldg(pos + c0), j = __ldg(jdx + c0);
        const int64_t a = __ldg(sa + j - 1);
        const int64_t limit = n - slack - (a > i ? a : i);
        carry_i = i;
        carry_h = extend_cta(packed, a, i, limit, bits, lcp[j], &s_best);
        if (threadIdx.x == 0) lcp[j] = (int32_t)carry_h;
    }
    const int64_t e0 = c0 + (int64_t)warp * (DEEP_CHUNK / (PLCP_THREADS / 32));
    const int64_t cend = c0 + DEEP_CHUNK < m ? c0 + DEEP_CHUNK : m;
    const int64_t e1 = e0 + DEEP_CHUNK / (PLCP_THREADS / 32) < cend ? e0 + DEEP_CHUNK / (PLCP_THREADS / 32) : cend;
    for (int64_t base = e0; base < e1; base += 32) {
        const int64_t e = base + lane;
        const bool valid = e < e1;
        int64_t i = 0, j = 0, a = 0, limit = 0, h = 0;
        if (valid) {
            i = __ldg(pos + e);
            j = __ldg(jdx + e);
            a = __ldg(sa + j - 1);                       // Phi(i)
            limit = n - slack - (a > i ? a : i);
            h = lcp[j];                                  // what stage 1 verified (or the chunk's first entry, exact)
            const int64_t lb = carry_h - (i - carry_i);  // PLCP[i] >= PLCP[i'] - (i - i') for i' <= i
            if (lb > h) h = lb;
        }
        bool done = !valid;
        if (valid) done = extend(packed, a, i, limit, bits, DEEP_STEPS, h);
        unsigned open = __ballot_sync(0xffffffffu, !done);
        while (open) {
            const int l = __ffs(open) - 1;
            open &= open - 1;
            // bound from the entry before it (exact by now: finished above or by an earlier turn of this loop)
            int64_t hl = __shfl_sync(0xffffffffu, h, l);
            const int64_t il = __shfl_sync(0xffffffffu, i, l);
            const int64_t al = __shfl_sync(0xffffffffu, a, l);
            const int64_t liml = __shfl_sync(0xffffffffu, limit, l);
            const int lp = l > 0 ? l - 1 : 0;
            int64_t ph = __shfl_sync(0xffffffffu, h, lp), pi = __shfl_sync(0xffffffffu, i, lp);
            if (l == 0) { ph = carry_h; pi = carry_i; }
            if (ph - (il - pi) > hl) hl = ph - (il - pi);
            const int64_t exact = extend_warp(packed, al, il, liml, bits, hl);
            if ((int)lane == l) { h = exact; done = true; }
        }
        if (valid) lcp[j] = (int32_t)h;
        const int last = (int)((e1 - base < 32 ? e1 - base : 32) - 1);
        carry_h = __shfl_sync(0xffffffffu, h, last);
        carry_i = __shfl_sync(0xffffffffu, i, last);
    }
}

static int bits_for(int64_t v)
{
    int b = 1;
    while ((1ll << b) <= v) b++;
    return b;
}

}  // namespace

int64_t lcp_scratch_bytes(int64_t n)
{
    if (n < 1) n = 1;
    return 5 * align_up(n * 4, 256) + rsort::workspace_bytes(n) + 1024;
}

int launch_lcp(const uint32_t *packed, const int32_t *d_sa, const uint32_t *d_skey0, int64_t n, int bits, bool fast,
               int32_t *d_lcp, void *d_scratch, int64_t scratch_bytes, cudaStream_t st)
{
    if (n == 1) return bwtk::zero_async(d_lcp, 4, st) == cudaSuccess ? BWTK_OK : BWTK_ECUDA;
    if (!d_scratch || scratch_bytes < lcp_scratch_bytes(n)) {
        set_error("lcp scratch: need %lld bytes, got %lld", (long long)lcp_scratch_bytes(n), (long long)scratch_bytes);
        return BWTK_EWORKSPACE;
    }
    Carver c(d_scratch, scratch_bytes);
    unsigned *d_count = c.take<unsigned>(4);
    uint32_t *dj = c.take<uint32_t>(n);
    uint32_t *key0 = c.take<uint32_t>(n);
    uint32_t *key1 = c.take<uint32_t>(n);
    uint32_t *val0 = c.take<uint32_t>(n);
    uint32_t *val1 = c.take<uint32_t>(n);
    rsort::Workspace rws = rsort::carve(c, n);
    if (!c.ok()) { set_error("lcp scratch carve overflow"); return BWTK_EWORKSPACE; }
    BWTK_CUDA(bwtk::zero_async(d_count, 16, st));
    DeepList dl{d_count, dj};
    const int slack = fast ? 1 : 0;
    {
        prof::Scope ps("lcp_kernel", n * 8 + (d_skey0 ? n * 4 : 0), st);
        if (d_skey0)
            lcp_keys_kernel<<<(unsigned)ceil_div(n, LCPK_THREADS * LCPK_ITEMS), LCPK_THREADS, 0, st>>>(
                packed, d_sa, d_skey0, n, bits, slack, d_lcp, dl);
        else
            lcp_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(packed, d_sa, n, bits, slack, d_lcp, dl);
        BWTK_LAUNCH_CHECK();
    }
    unsigned h_count = 0;
    int rc = read_back(&h_count, d_count, sizeof(unsigned), st);
    if (rc) return rc;
    const int64_t m = h_count;
    if (m == 0) return BWTK_OK;
    // the deep pairs in text order
    deep_keys_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(d_sa, dj, m, key0, val0);
    BWTK_LAUNCH_CHECK();
    BWTK_CUDA(bwtk::zero_async(rws.err, sizeof(int), st));
    int in_first = 1;
    rc = rsort::sort_pairs<uint32_t>(key0, val0, key1, val1, m, 0, bits_for(n), rws, st, &in_first, nullptr);
    if (rc) return rc;
    const uint32_t *pos = in_first ? key0 : key1, *jdx = in_first ? val0 : val1;
    {
        prof::Scope ps("plcp_kernel", m * 16, st);
        plcp_kernel<<<(unsigned)ceil_div(m, DEEP_CHUNK), PLCP_THREADS, 0, st>>>(packed, d_sa, pos, jdx, m, n, bits, slack,
                                                                             d_lcp);
        BWTK_LAUNCH_CHECK();
    }
    int h_err = 0;
    rc = read_back(&h_err, rws.err, sizeof(int), st);
    if (rc) return rc;
    if (h_err) { set_error("look-back spin limit hit in the LCP deep-pair sort"); return BWTK_EINTERNAL; }
    return BWTK_OK;
}

}  // namespace bwtk
