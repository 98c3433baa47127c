// scans.cu -- the perfect-repeat detectors as parallel scans + ordered resolves.
//
//  * strict adjacency scan  (Tier2LCPFinder.find_long_unit_repeats_strict,
//    reference bwt.py:1891-2001; the only detector the CLI runs)
//  * Tier 1 sliding window  (Tier1STRFinder._find_simple_tandems_kmer,
//    reference bwt.py:1426-1531)
//  * LCP plateaus           (Tier2LCPFinder._detect_lcp_plateaus and
//    _analyze_sa_interval_for_tandems, reference bwt.py:2118-2145, 2500-2560)
//
// The reference walks the text greedily (emit, jump past the array, otherwise
// step).  Here every kernel first computes the position-independent facts in
// parallel (maximal runs of text[j]==text[j+u]; "would emit if visited"
// candidates), and the greedy order is then replayed over the sparse candidate
// list only: a short dependency-chain walk for the strict scan, pointer jumping
// over the candidate successor graph for Tier 1 (which also reproduces the
// reference's adaptive position_step sampling exactly).
#include "radix_sort.cuh"
#include "scan.cuh"
#include "tma.cuh"

namespace bwtk {

// ===========================================================================
// strict adjacency scan
// ===========================================================================
namespace strict {

constexpr int MAX_DYN_SMEM = 160 * 1024;   // dynamic shared memory the run finders may ask for (units up to 60 000: 116 KB)
constexpr int TP = 8192;       // positions per CTA tile
constexpr int THREADS = 256;   // 8 warps x 1024 positions

struct CandOut {
    unsigned long long *key;   // ((umax - u) << abits) | run_start
    uint32_t *val;             // run_end
    unsigned long long *count;
    int64_t cap;
    int abits;
    int64_t umax;
};

__device__ __forceinline__ void push_cand(const CandOut &o, int64_t u, int64_t a, int64_t b)
{
    unsigned long long slot = atomicAdd(o.count, 1ull);
    if ((int64_t)slot < o.cap) {
        o.key[slot] = ((unsigned long long)(o.umax - u) << o.abits) | (unsigned long long)a;
        o.val[slot] = (uint32_t)b;
    }
}

// A warp step found matching groups: extend every streak that can still reach the minimum run
// length to its maximal run and push it.  Kept out of line so that the hot loop stays small.
__device__ __noinline__ void run_candidates(const uint8_t *__restrict__ text, int64_t n, int u, int64_t L,
                                            int64_t pos, bool ok, unsigned B, const CandOut &out)
{
    const int lane = threadIdx.x & 31;
    const int64_t lim = n - u;  // E_u[j] defined for j < lim
    bool prev_ok = __shfl_up_sync(0xffffffffu, ok, 1);
    if (lane == 0) {
        prev_ok = false;
        if (pos >= 4) {
            prev_ok = true;
            for (int q = 1; q <= 4; q++)
                if (__ldg(text + pos - q) != __ldg(text + pos - q + u)) { prev_ok = false; break; }
        }
    }
    if (ok && !prev_ok) {
        // streak of full groups starting at this lane
        unsigned rest = ~(B >> lane);
        int g = rest ? (__ffs(rest) - 1) : 32;
        if (g > 32 - lane) g = 32 - lane;
        bool ends_here = (lane + g) < 32;
        if (!ends_here || (int64_t)4 * g + 6 >= L) {
            int64_t ra = pos;
            while (ra > 0 && __ldg(text + ra - 1) == __ldg(text + ra - 1 + u)) ra--;
            int64_t rb = pos + 4 * g;
            while (rb < lim && __ldg(text + rb) == __ldg(text + rb + u)) rb++;
            if (rb - ra >= L) push_cand(out, u, ra, rb);
        }
    }
}

// Unit lengths [u_from, u_to] of one tile.  INTERIOR: every group of the tile lies before
// n - u - 3, no bounds test.  PAIRS: a run of >= 11 matches contains two adjacent aligned groups
// of 4 (or reaches the next chunk through lane 31), so single matching groups -- one warp step
// in eight on random sequence -- are not looked at.
template <bool INTERIOR, bool PAIRS>
__device__ __forceinline__ void scan_units(const uint8_t *__restrict__ text, int64_t n, int64_t t0, int u_from,
                                           int u_to, int mc, const uint32_t *wbase, const uint32_t (&base)[8],
                                           const CandOut &out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int local = warp * 1024 + 4 * lane;
    for (int u = u_from; u <= u_to; u++) {
        const uint32_t *p = wbase + (u >> 2);
        const int sh = (u & 3) * 8;
        const int last_ok = INTERIOR ? 0 : (int)((n - u - 3 - t0 < (int64_t)TP + 8) ? (n - u - 3 - t0) : (int64_t)TP + 8);
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const uint32_t w = __funnelshift_r(p[c * 32], p[c * 32 + 1], sh);
            bool ok = w == base[c];
            if (!INTERIOR) ok = ok && (local + c * 128 < last_ok);
            const unsigned B = __ballot_sync(0xffffffffu, ok);
            if (PAIRS ? ((B & (B >> 1)) == 0u && (int)B >= 0) : (B == 0u)) continue;
            run_candidates(text, n, u, (int64_t)(mc - 1) * u, t0 + local + c * 128, ok, B, out);
        }
    }
}

// Units whose minimum run (mc-1)*u is >= 8 positions: a qualifying run then
// contains at least one aligned group of 4 matching positions, so one 4-byte
// compare per lane + a ballot filters 128 positions per warp step.
__global__ void __launch_bounds__(THREADS)
    find_runs_kernel(const uint8_t *__restrict__ text, int64_t n, int64_t u_lo, int64_t u_hi,
                     int64_t u_per_block, int64_t mc, CandOut out)
{
    extern __shared__ __align__(16) uint8_t s_text[];
    const int64_t t0 = (int64_t)blockIdx.x * TP;
    const int64_t ua = u_lo + (int64_t)blockIdx.y * u_per_block;
    int64_t ub = ua + u_per_block - 1;
    if (ub > u_hi) ub = u_hi;
    if (ua > ub) return;
    const int span = TP + (int)ub + 8;
    // The tile (+ halo of the longest unit) arrives as ONE 1-D bulk copy through the TMA unit: no byte
    // loads, no registers, one elected thread.  Only what the copy cannot carry -- the last < 16 bytes
    // and the zero padding past the end of the text -- is written by ordinary stores.
    __shared__ __align__(8) uint64_t s_bar;
    int bulk = 0;
    if ((((uintptr_t)(text + t0)) & 15) == 0) {
        const int64_t avail = n - t0;
        bulk = (int)((span < avail ? (int64_t)span : avail) & ~15ll);
    }
    if (bulk > 0) {
        if (threadIdx.x == 0) tma::mbar_init(&s_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            tma::expect_tx(&s_bar, (uint32_t)bulk);
            tma::bulk_load(s_text, text + t0, (uint32_t)bulk, &s_bar);
        }
    }
    for (int i = bulk + threadIdx.x * 4; i < span; i += THREADS * 4) {
        uint32_t w = 0;
        int64_t g = t0 + i;
        for (int q = 0; q < 4; q++)
            if (g + q < n) w |= (uint32_t)__ldg(text + g + q) << (8 * q);
        *reinterpret_cast<uint32_t *>(s_text + i) = w;
    }
    if (bulk > 0) tma::wait(&s_bar, 0);
    __syncthreads();
    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(s_text);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t base[8];
#pragma unroll
    for (int c = 0; c < 8; c++) base[c] = s32[(warp * 1024 + c * 128 + 4 * lane) >> 2];
    const uint32_t *wbase = s32 + warp * 256 + lane;   // word of this lane's group in chunk 0
    // first unit length whose minimum run (mc-1)*u reaches 11
    int64_t u_pairs = (10 + (mc - 1)) / (mc - 1);
    if (u_pairs < ua) u_pairs = ua;
    if (u_pairs > ub + 1) u_pairs = ub + 1;
    if (t0 + TP + 3 + ub < n) {
        scan_units<true, false>(text, n, t0, (int)ua, (int)u_pairs - 1, (int)mc, wbase, base, out);
        scan_units<true, true>(text, n, t0, (int)u_pairs, (int)ub, (int)mc, wbase, base, out);
    } else {
        scan_units<false, false>(text, n, t0, (int)ua, (int)u_pairs - 1, (int)mc, wbase, base, out);
        scan_units<false, true>(text, n, t0, (int)u_pairs, (int)ub, (int)mc, wbase, base, out);
    }
}

// ---- the same search on 2-bit codes: 16 positions per lane and step -----------------------
// Units whose minimum run (mc-1)*u is >= 31 positions: a qualifying run contains an aligned group
// of SIXTEEN matching positions, so the tile is re-coded to 2 bits per symbol in shared memory
// (code = (byte >> 1) & 3: equal bytes give equal codes, so the filter has no false negatives)
// and one 32-bit compare per lane + a ballot filters 512 positions per warp step -- a quarter of
// the byte kernel's steps for the same text.  On random sequence a group matches with probability
// 4^-16, so the ballot is zero practically always and the loop body is 2 LDS + SHF + compare +
// vote + branch.  Bytes other than A/C/G/T can alias (N/G, '$'/T): tiles that hold any are marked
// dirty and every matching group of such a tile is re-checked on the bytes before it counts.
constexpr int TPK = 32768;      // positions per CTA tile: 8 warps x 8 chunks x 512
constexpr int GK = 16;          // positions per group

__device__ __noinline__ void run_candidates16(const uint8_t *__restrict__ text, const uint8_t *s_bytes, int64_t n,
                                              int64_t t0, int u, int64_t L, int64_t pos, bool ok, bool dirty,
                                              const CandOut &out)
{
    const int lane = threadIdx.x & 31;
    const int64_t lim = n - u;
    if (dirty && ok) {
        const uint8_t *a = s_bytes + (pos - t0);
#pragma unroll 4
        for (int q = 0; q < GK; q++) ok = ok && a[q] == a[q + u];
    }
    const unsigned B = __ballot_sync(0xffffffffu, ok);
    if (B == 0u) return;
    bool prev_ok = __shfl_up_sync(0xffffffffu, ok, 1);
    if (lane == 0) {
        prev_ok = false;
        if (pos >= GK) {
            prev_ok = true;
            for (int q = 1; q <= GK; q++)
                if (__ldg(text + pos - q) != __ldg(text + pos - q + u)) { prev_ok = false; break; }
        }
    }
    if (ok && !prev_ok) {
        unsigned rest = ~(B >> lane);
        int g = rest ? (__ffs(rest) - 1) : 32;
        if (g > 32 - lane) g = 32 - lane;
        const bool ends_here = (lane + g) < 32;
        if (!ends_here || (int64_t)GK * g + 2 * (GK - 1) >= L) {
            int64_t ra = pos;
            while (ra > 0 && __ldg(text + ra - 1) == __ldg(text + ra - 1 + u)) ra--;
            int64_t rb = pos + (int64_t)GK * g;
            while (rb < lim && __ldg(text + rb) == __ldg(text + rb + u)) rb++;
            if (rb - ra >= L) push_cand(out, u, ra, rb);
        }
    }
}

template <bool INTERIOR>
__device__ __forceinline__ void scan_units16(const uint8_t *__restrict__ text, const uint8_t *s_bytes, int64_t n,
                                             int64_t t0, int u_from, int u_to, int mc, const uint32_t *wbase,
                                             const uint32_t (&base)[8], bool dirty, const CandOut &out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int local = warp * (TPK / 8) + GK * lane;
    for (int u = u_from; u <= u_to; u++) {
        const uint32_t *p = wbase + (u >> 4);
        const int sh = (u & 15) * 2;
        // groups whose 16 positions all lie before n - u
        const int64_t lim_local = INTERIOR ? 0 : (n - u - t0);
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const uint32_t w = __funnelshift_r(p[c * 32], p[c * 32 + 1], sh);
            bool ok = w == base[c];
            if (!INTERIOR) ok = ok && (local + c * 512 + GK <= lim_local);
            if (!__any_sync(0xffffffffu, ok)) continue;
            run_candidates16(text, s_bytes, n, t0, u, (int64_t)(mc - 1) * u, t0 + local + c * 512, ok, dirty, out);
        }
    }
}

__global__ void __launch_bounds__(THREADS)
    find_runs16_kernel(const uint8_t *__restrict__ text, int64_t n, int64_t u_lo, int64_t u_hi,
                       int64_t u_per_block, int64_t mc, CandOut out)
{
    extern __shared__ __align__(16) uint8_t s_text[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_dirty;
    const int64_t t0 = (int64_t)blockIdx.x * TPK;
    const int64_t ua = u_lo + (int64_t)blockIdx.y * u_per_block;
    int64_t ub = ua + u_per_block - 1;
    if (ub > u_hi) ub = u_hi;
    if (ua > ub) return;
    const int span = TPK + (int)ub + 2 * GK;            // bytes the compares can touch (+ one spare word)
    const int span16 = (span + 15) & ~15;
    uint32_t *s_codes = reinterpret_cast<uint32_t *>(s_text + span16);
    if (threadIdx.x == 0) s_dirty = 0;
    // byte tile: one bulk copy through the TMA unit + the unaligned / past-the-end remainder by stores
    int bulk = 0;
    if ((((uintptr_t)(text + t0)) & 15) == 0) {
        const int64_t avail = n - t0;
        bulk = (int)((span < avail ? (int64_t)span : avail) & ~15ll);
    }
    if (bulk > 0) {
        if (threadIdx.x == 0) tma::mbar_init(&s_bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            tma::expect_tx(&s_bar, (uint32_t)bulk);
            tma::bulk_load(s_text, text + t0, (uint32_t)bulk, &s_bar);
        }
    }
    for (int i = bulk + threadIdx.x * 4; i < span16; i += THREADS * 4) {
        uint32_t w = 0;
        const int64_t g = t0 + i;
        for (int q = 0; q < 4; q++)
            if (g + q < n) w |= (uint32_t)__ldg(text + g + q) << (8 * q);
        *reinterpret_cast<uint32_t *>(s_text + i) = w;
    }
    if (bulk > 0) tma::wait(&s_bar, 0);
    __syncthreads();
    // 2-bit codes, 16 symbols per word, symbol j at bits [2j, 2j+1]; bytes past the end are zero
    bool other = false;
    const int valid = (int)((n - t0) < (int64_t)span16 ? (n - t0) : (int64_t)span16);
    for (int w = threadIdx.x; w < span16 / 16; w += THREADS) {
        const uint4 v = *reinterpret_cast<const uint4 *>(s_text + w * 16);
        const uint32_t q[4] = {v.x, v.y, v.z, v.w};
        uint32_t code = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const uint32_t ch = (q[k] >> (8 * b)) & 0xffu;
                code |= ((ch >> 1) & 3u) << (2 * (4 * k + b));
                const bool acgt = ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T';
                other = other || (!acgt && (w * 16 + 4 * k + b) < valid);
            }
        }
        s_codes[w] = code;
    }
    if (other) s_dirty = 1;
    __syncthreads();
    const bool dirty = s_dirty != 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t base[8];
#pragma unroll
    for (int c = 0; c < 8; c++) base[c] = s_codes[warp * (TPK / 8 / GK) + c * 32 + lane];
    const uint32_t *wbase = s_codes + warp * (TPK / 8 / GK) + lane;
    if (t0 + TPK + GK + ub < n)
        scan_units16<true>(text, s_text, n, t0, (int)ua, (int)ub, (int)mc, wbase, base, dirty, out);
    else
        scan_units16<false>(text, s_text, n, t0, (int)ua, (int)ub, (int)mc, wbase, base, dirty, out);
}

// ---- the 16-position search restricted to groups whose 16-mer occurs more than once -----------------------
// A group of 16 aligned positions p can only match its copy at p + u if the 16-mer at p occurs a second time in
// the text, i.e. if suffix p shares >= 16 symbols with a neighbour in suffix order.  With the index at hand
// (bwtk_repeat_hint: one pass over the LCP array sets a bit for both suffixes of every neighbouring pair with
// LCP >= 16) only those groups -- 6 % on planted random sequence -- are compared at all: they are listed, and a
// warp per listed group stages the group's next ~1000 bases as 2-bit codes in shared memory and tries every
// unit length against them, 32 at a time.  Same runs as find_runs16_kernel (matches are re-checked on the bytes;
// a run is reported by the first matching group of its streak), in any order.
__global__ void __launch_bounds__(256)
    hint_list_kernel(const uint32_t *__restrict__ hint, int64_t n, uint32_t *__restrict__ list, unsigned *__restrict__ count,
                     unsigned cap)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // aligned group index
    const int64_t p = g * GK;
    bool on = false;
    if (p + GK <= n) on = (__ldg(hint + (p >> 5)) >> (p & 31)) & 1u;
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    if (bal == 0u) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (on) {
        const unsigned at = base + (unsigned)__popc(bal & ((1u << lane) - 1u));
        if (at < cap) list[at] = (uint32_t)p;
    }
}

constexpr int HL_WARPS = 8;          // groups per CTA
constexpr int HL_WORDS = 72;         // 2-bit words staged per group: 16 * 72 = 1152 positions >= 16 + 1000 + 16

__global__ void __launch_bounds__(HL_WARPS * 32)
    find_runs16_list_kernel(const uint8_t *__restrict__ text, int64_t n, const uint32_t *__restrict__ list,
                            const unsigned *__restrict__ count, unsigned cap, int64_t u_lo, int64_t u_hi, int64_t mc,
                            CandOut out)
{
    __shared__ uint32_t s_codes[HL_WARPS][HL_WORDS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned total = *count;
    if (total > cap) total = cap;
    uint32_t *codes = s_codes[warp];
    for (unsigned gi = blockIdx.x * HL_WARPS + warp; gi < total; gi += gridDim.x * HL_WARPS) {
        const int64_t p = list[gi];
        // 2-bit codes of text[p, p + 16 * HL_WORDS): code = (byte >> 1) & 3, zero past the end
        for (int w = lane; w < HL_WORDS; w += 32) {
            const int64_t q0 = p + (int64_t)w * 16;
            uint32_t code = 0;
            if (q0 + 16 <= n) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(text + q0));     // text and p are 16-byte aligned
                const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; k++)
#pragma unroll
                    for (int b = 0; b < 4; b++) code |= (((q[k] >> (8 * b)) >> 1) & 3u) << (2 * (4 * k + b));
            } else {
                for (int j = 0; j < 16 && q0 + j < n; j++) code |= (((uint32_t)__ldg(text + q0 + j) >> 1) & 3u) << (2 * j);
            }
            codes[w] = code;
        }
        __syncwarp();
        const uint32_t base = codes[0];
        for (int64_t u0 = u_lo; u0 <= u_hi; u0 += 32) {
            const int64_t u = u0 + lane;
            bool ok = false;
            if (u <= u_hi && p + GK <= n - u) {
                const uint32_t w = (uint32_t)(u >> 4), sh = (uint32_t)(u & 15) * 2u;
                ok = __funnelshift_r(codes[w], codes[w + 1], sh) == base;
            }
            if (!__any_sync(0xffffffffu, ok)) continue;
            if (ok) {
                // on the bytes: the group really matches, and the aligned group before it does not (else that one reports)
                const uint8_t *a = text + p;
                for (int q = 0; q < GK && ok; q++) ok = __ldg(a + q) == __ldg(a + q + u);
                if (ok && p >= GK) {
                    bool prev = true;
                    for (int q = 1; q <= GK && prev; q++) prev = __ldg(a - q) == __ldg(a - q + u);
                    if (prev) {
                        // the previous group matches too; it is listed (its 16-mer repeats as well) and reports the run
                        ok = false;
                    }
                }
                if (ok) {
                    const int64_t lim = n - u;
                    int64_t ra = p;
                    while (ra > 0 && __ldg(text + ra - 1) == __ldg(text + ra - 1 + u)) ra--;
                    int64_t rb = p + GK;
                    while (rb < lim && __ldg(text + rb) == __ldg(text + rb + u)) rb++;
                    if (rb - ra >= (mc - 1) * u) push_cand(out, u, ra, rb);
                }
            }
        }
        __syncwarp();
    }
}

// Units with (mc-1)*u < 8: one thread per position, run starts found directly.
__global__ void __launch_bounds__(256)
    find_runs_small_kernel(const uint8_t *__restrict__ text, int64_t n, int64_t u_lo, int64_t u_hi,
                           int64_t mc, int64_t min_run_u1, CandOut out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t ci = __ldg(text + i);
    for (int64_t u = u_lo; u <= u_hi; u++) {
        int64_t lim = n - u;
        if (i >= lim) break;
        if (ci != __ldg(text + i + u)) continue;
        if (i > 0 && __ldg(text + i - 1) == __ldg(text + i - 1 + u)) continue;  // not a run start
        int64_t b = i + 1;
        while (b < lim && __ldg(text + b) == __ldg(text + b + u)) b++;
        if (b - i >= (u == 1 && min_run_u1 > mc - 1 ? min_run_u1 : (mc - 1) * u)) push_cand(out, u, i, b);
    }
}

// ---- small units, ordered: the maximal runs of ONE unit length u (<= 7 positions) in text order ------
// The per-position kernel above pushes every run through one global counter and leaves the list to be
// sorted; homopolymer runs alone are ~5 % of the positions.  Here a scan (scan.cuh) over 16-position
// chunks counts the qualifying run starts of a chunk from one match mask (two 16-byte loads, byte-wise
// compares on 64-bit words), and the emit phase writes them at their place in the sorted list: no
// atomics, no sort, and the list of the large units that does get sorted stays short.
template <int U> struct SmallRuns {
    const uint8_t *text;    // 16-byte aligned
    int64_t n;
    int64_t L;              // minimum run length
    // bit j: text[p0 + j] == text[p0 + j + U] (and p0 + j < n - U), j = 0..23 -- the chunk's 16 positions and 8 of
    // look-ahead, enough to see whether a run of <= 9 starts at any of them; *prev: the same for p0 - 1
    // carry (may be null): the 16 bytes at p0, loaded as the look-ahead half of the chunk before; on return the
    // 16 bytes at p0 + 16.  prev_known >= 0: the match bit of p0 - 1 is already known (bit 15 of the chunk before).
    __device__ __forceinline__ uint32_t mask(int64_t p0, bool *prev, uint4 *carry = nullptr, bool have_carry = false,
                                             int prev_known = -1) const
    {
        const int64_t lim = n - U;
        uint32_t e = 0;
        if (p0 + 32 <= n) {
            const uint4 a = have_carry ? *carry : __ldg(reinterpret_cast<const uint4 *>(text + p0));
            const uint4 b = __ldg(reinterpret_cast<const uint4 *>(text + p0 + 16));
            if (carry) *carry = b;
            const unsigned long long w0 = ((unsigned long long)a.y << 32) | a.x, w1 = ((unsigned long long)a.w << 32) | a.z,
                                     w2 = ((unsigned long long)b.y << 32) | b.x, w3 = ((unsigned long long)b.w << 32) | b.z;
            // bytes j + U of the three 8-byte words
            const unsigned long long s0 = (w0 >> (8 * U)) | (w1 << (64 - 8 * U)), s1 = (w1 >> (8 * U)) | (w2 << (64 - 8 * U)),
                                     s2 = (w2 >> (8 * U)) | (w3 << (64 - 8 * U));
            unsigned long long x0 = w0 ^ s0, x1 = w1 ^ s1, x2 = w2 ^ s2;
            // zero bytes -> bit 7 of the byte (exact form: no borrow across bytes)
            const unsigned long long K = 0x7f7f7f7f7f7f7f7full;
            x0 = ~(((x0 & K) + K) | x0 | K);
            x1 = ~(((x1 & K) + K) | x1 | K);
            x2 = ~(((x2 & K) + K) | x2 | K);
            // bit 7 of every byte -> one byte (the multiply adds the eight bits up in its top byte)
            e = (uint32_t)(((x0 >> 7) * 0x0102040810204080ull) >> 56) |
                ((uint32_t)(((x1 >> 7) * 0x0102040810204080ull) >> 56) << 8) |
                ((uint32_t)(((x2 >> 7) * 0x0102040810204080ull) >> 56) << 16);
            if (p0 + 24 > lim) e &= lim > p0 ? ((1u << (int)(lim - p0)) - 1u) : 0u;
        } else {
            for (int j = 0; j < 24; j++) {
                const int64_t q = p0 + j;
                if (q < lim && __ldg(text + q) == __ldg(text + q + U)) e |= 1u << j;
            }
        }
        *prev = prev_known >= 0 ? prev_known != 0
                                : (p0 > 0 && p0 - 1 < lim && __ldg(text + p0 - 1) == __ldg(text + p0 - 1 + U));
        return e;
    }
    // bit j (j < 16): a maximal run of >= L matches starts at p0 + j.  L <= 9: the 24-bit mask decides.
    __device__ __forceinline__ uint32_t qualifying_starts(uint32_t e, bool prev) const
    {
        uint32_t r = e;
        for (int k = 1; k < (int)L; k++) r &= e >> k;
        return e & ~((e << 1) | (prev ? 1u : 0u)) & r & 0xffffu;
    }
    // calls f(start, end) for every maximal run of >= L matches that starts in the chunk with match mask e
    template <typename F> __device__ __forceinline__ void for_runs(int64_t c, uint32_t e, bool prev, F f) const
    {
        const int64_t p0 = c * 16;
        uint32_t starts = qualifying_starts(e, prev);
        const int64_t lim = n - U;
        while (starts) {
            const int j = __ffs(starts) - 1;
            starts &= starts - 1;
            const uint32_t rest = ~(e >> j);                 // first mismatch after j
            const int t = __ffs(rest) - 1;                    // e has 24 bits: a zero exists at or below bit 24 - j
            int64_t b = p0 + j + t;
            if (j + t >= 24) {                                // the run leaves what the mask covers
                b = p0 + 24;
                while (b < lim && __ldg(text + b) == __ldg(text + b + U)) b++;
            }
            f(p0 + j, b);
        }
    }
};
// the match masks of a thread's chunks are computed once, in the count phase, and kept for the emit phase
struct SmallRunsCtx {
    uint32_t e[scan::ITEMS];
    uint32_t prev;
    uint4 carry;            // the 16 bytes after the chunk handled last (a thread's chunks are consecutive)
};
template <int U> struct SmallRunsCount {
    using Ctx = SmallRunsCtx;
    SmallRuns<U> sr;
    __device__ unsigned long long operator()(int64_t c, int k, Ctx &ctx) const
    {
        if (k == 0) ctx.prev = 0;
        ctx.e[k] = 0;
        if (c * 16 >= sr.n - U) return 0;
        // chunk c - 1 (slot k - 1 of this thread) left its look-ahead bytes and, in bit 15 of its mask, the match
        // bit of the position before this chunk: 9 instead of 16 vector loads and no byte loads per thread
        bool prev;
        const uint32_t e = k > 0 ? sr.mask(c * 16, &prev, &ctx.carry, true, (int)((ctx.e[k - 1] >> 15) & 1u))
                                 : sr.mask(c * 16, &prev, &ctx.carry, false, -1);
        ctx.e[k] = e;
        if (prev) ctx.prev |= 1u << k;
        return (unsigned long long)__popc(sr.qualifying_starts(e, prev));
    }
};
template <int U> struct SmallRunsEmit {
    SmallRuns<U> sr;
    CandOut out;
    const unsigned long long *base;   // device: runs already in the list
    __device__ void operator()(int64_t c, int k, unsigned long long excl, unsigned long long cnt, SmallRunsCtx &ctx) const
    {
        if (!cnt) return;
        unsigned long long at = *base + excl;
        sr.for_runs(c, ctx.e[k], (ctx.prev >> k) & 1u, [&](int64_t a, int64_t b) {
            if ((int64_t)at < out.cap) {
                out.key[at] = ((unsigned long long)(out.umax - U) << out.abits) | (unsigned long long)a;
                out.val[at] = (uint32_t)b;
            }
            at++;
        });
    }
};
static __global__ void add_total_kernel(unsigned long long *base, const unsigned long long *total) { *base += *total; }

template <int U>
static int small_runs_append(const uint8_t *d_text, int64_t n, int64_t L, const CandOut &out, unsigned long long *d_base,
                             const scan::Workspace &sws, cudaStream_t st)
{
    SmallRuns<U> sr{d_text, n, L};
    SmallRunsCount<U> cf{sr};
    SmallRunsEmit<U> ef{sr, out, d_base};
    int rc = scan::run(ceil_div(n, 16), cf, ef, sws, st);
    if (rc) return rc;
    add_total_kernel<<<1, 1, 0, st>>>(d_base, sws.total);
    BWTK_LAUNCH_CHECK();
    return BWTK_OK;
}

struct Resolved {
    uint8_t *emit;       // per candidate
    int32_t *row;        // per candidate, REC_W ints
};

// Replays the greedy walk over the sorted candidate runs of one unit length:
// entry e = max(run start, end of the previous emission); emit when
// e <= run_end - (mc-1)*u; the array then ends at e + u*(1 + (run_end-e)/u).
// A run only depends on the run before it when that one's array can reach past its start
// (start < previous run end + u); runs linked that way form a chain.  The thread of the chain's FIRST run walks
// the chain forward once and leaves (entry, copies) for every member -- O(chain) in total, where every
// member walking back to the head on its own would be O(chain^2) on texts that are one long chain (a period-2
// text with a mismatch every few bases).  Chains of ordinary sequence have one or two members.
__global__ void __launch_bounds__(256)
    resolve_walk_kernel(const unsigned long long *__restrict__ key, const uint32_t *__restrict__ val, int64_t m,
                        int abits, int64_t umax, int64_t mc, Resolved res)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const unsigned long long amask = (1ull << abits) - 1ull;
    const int64_t urev = (int64_t)(key[t] >> abits);
    const int64_t u = umax - urev;
    const int64_t L = (mc - 1) * u;
    if (t > 0 && (int64_t)(key[t - 1] >> abits) == urev && (int64_t)(key[t] & amask) < (int64_t)val[t - 1] + u)
        return;                                  // not the first run of its chain
    int64_t prev_end = -1, prev_b = 0;
    for (int64_t r = t; r < m; r++) {
        const unsigned long long kr = key[r];
        const int64_t a = (int64_t)(kr & amask), b = (int64_t)val[r];
        if (r > t && ((int64_t)(kr >> abits) != urev || a >= prev_b + u)) break;      // the chain ends
        const int64_t e = a > prev_end ? a : prev_end;
        const bool emitted = e <= b - L;
        res.emit[r] = emitted ? 1 : 0;
        if (emitted) {
            const int64_t cnt = 1 + (b - e) / u;
            prev_end = e + cnt * u;
            int32_t *row = res.row + r * BWTK_REC_W;
            row[0] = (int32_t)e;
            row[3] = (int32_t)cnt;
        }
        prev_b = b;
    }
}

// The record of every emitted run: primitive period of the first unit (MotifUtils.smallest_period_str,
// bwt.py:1124-1133), span, copies.  One thread per run; (entry, copies) come from resolve_walk_kernel.
__global__ void __launch_bounds__(256)
    resolve_rows_kernel(const uint8_t *__restrict__ text, const unsigned long long *__restrict__ key, int64_t m,
                        int abits, int64_t umax, Resolved res)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m || !res.emit[t]) return;
    const int64_t u = umax - (int64_t)(key[t] >> abits);
    int32_t *row = res.row + t * BWTK_REC_W;
    const int64_t e = row[0], cnt = row[3];
    int64_t prim = u;
    for (int64_t p = 1; p <= u / 2; p++) {
        if (u % p) continue;
        bool ok = true;
        for (int64_t j = p; j < u; j++)
            if (__ldg(text + e + j) != __ldg(text + e + j - p)) { ok = false; break; }
        if (ok) { prim = p; break; }
    }
    const int64_t end = e + cnt * u;
    row[1] = (int32_t)end; row[2] = (int32_t)prim;
    row[3] = (int32_t)(prim < u ? (end - e) / prim : cnt);
    row[4] = 0; row[5] = 0; row[6] = (int32_t)u; row[7] = 0;
}

struct CountEmit {
    const uint8_t *emit;
    __device__ uint64_t operator()(int64_t i) const { return emit[i]; }
};
struct WriteRows {
    const int32_t *row;
    int32_t *out;
    int64_t cap;
    __device__ void operator()(int64_t i, uint64_t excl, uint64_t cnt) const
    {
        if (!cnt || (int64_t)excl >= cap) return;
        const int4 *src = reinterpret_cast<const int4 *>(row + i * BWTK_REC_W);
        int4 *dst = reinterpret_cast<int4 *>(out + excl * BWTK_REC_W);
        dst[0] = src[0];
        dst[1] = src[1];
    }
};

static int bits_for(int64_t v)
{
    int b = 1;
    while ((1ll << b) <= v) b++;
    return b;
}

static int64_t cand_capacity(int64_t n) { return n / 2 + 65536; }

// All maximal runs of text[j]==text[j+u], u in [u_lo, u_hi], at least (mc-1)*u long, sorted by
// (u descending, start): *sk / *sv (run start in the low `abits` key bits, run end), *m runs.
// Returns BWTK_EWORKSPACE when more than `ccap` runs exist (*m then holds the count).
static int collect_runs(const uint8_t *d_text, int64_t n, int64_t u_lo, int64_t u_hi, int64_t mc,
                        unsigned long long *key0, unsigned long long *key1, uint32_t *val0, uint32_t *val1,
                        unsigned long long *d_count, int64_t ccap, const rsort::Workspace &rws, cudaStream_t st,
                        const unsigned long long **sk, const uint32_t **sv, int64_t *m,
                        const scan::Workspace *sws = nullptr, int64_t min_run_u1 = 0, const uint32_t *hint = nullptr,
                        uint32_t *hint_list = nullptr, unsigned hint_cap = 0)
{
    *m = 0;
    BWTK_CUDA(bwtk::zero_async(d_count, 16, st));
    CandOut out;
    out.key = key0; out.val = val0; out.count = d_count; out.cap = ccap;
    out.abits = bits_for(n); out.umax = u_hi;
    // units with (mc-1)*u < 8 take the per-position search, the rest the ballot kernel
    int64_t small_hi = 7 / (mc - 1);
    if (small_hi > u_hi) small_hi = u_hi;
    // ordered form (no atomics, no sort) when a scan workspace is given and the text allows 16-byte loads
    const bool ordered = sws != nullptr && small_hi >= u_lo && small_hi <= 7 && (((uintptr_t)d_text) & 15) == 0 &&
                         (mc - 1) * small_hi <= 9 && (min_run_u1 <= 9);   // the 24-bit match mask decides runs of <= 9
    if (small_hi >= u_lo && !ordered) {
        find_runs_small_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(d_text, n, u_lo, small_hi, mc, min_run_u1, out);
        BWTK_LAUNCH_CHECK();
    }
    int64_t big_lo = small_hi + 1 > u_lo ? small_hi + 1 : u_lo;
    // units whose minimum run reaches 31 positions go to the 16-position (2-bit) kernel
    int64_t pk_lo = (30 + (mc - 1)) / (mc - 1);
    if (pk_lo < big_lo) pk_lo = big_lo;
    const int64_t byte_hi = pk_lo - 1 < u_hi ? pk_lo - 1 : u_hi;
    if (big_lo <= byte_hi) {
        int64_t tiles = ceil_div(n, TP);
        int64_t nu = byte_hi - big_lo + 1;
        // enough CTAs to fill the machine even for short contigs
        int64_t ysplit = ceil_div((int64_t)NUM_SMS * 8, tiles);
        if (ysplit > nu) ysplit = nu;
        if (ysplit < 1) ysplit = 1;
        if (ysplit > 65535) ysplit = 65535;
        int64_t u_per_block = ceil_div(nu, ysplit);
        ysplit = ceil_div(nu, u_per_block);
        size_t smem = (size_t)(TP + byte_hi + 16);
        // the opt-in maximum once (thread-safe static initialisation: several contigs are scanned concurrently,
        // and a per-call "raise if larger" would let a small contig's call lower the limit under a large one's)
        static const cudaError_t attr_rc =
            cudaFuncSetAttribute(find_runs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM);
        BWTK_CUDA(attr_rc);
        BWTK_REQUIRE(smem <= (size_t)MAX_DYN_SMEM, "unit length too large for the run finder's shared-memory tile");
        dim3 grid((unsigned)tiles, (unsigned)ysplit);
        find_runs_kernel<<<grid, THREADS, smem, st>>>(d_text, n, big_lo, byte_hi, u_per_block, mc, out);
        BWTK_LAUNCH_CHECK();
    }
    if (pk_lo <= u_hi) {
        int64_t tiles = ceil_div(n, TPK);
        int64_t nu = u_hi - pk_lo + 1;
        int64_t ysplit = ceil_div((int64_t)NUM_SMS * 4, tiles);
        if (ysplit > nu) ysplit = nu;
        if (ysplit < 1) ysplit = 1;
        if (ysplit > 65535) ysplit = 65535;
        int64_t u_per_block = ceil_div(nu, ysplit);
        ysplit = ceil_div(nu, u_per_block);
        const size_t span16 = (size_t)((TPK + u_hi + 2 * GK + 15) & ~15ll);
        size_t smem = span16 + span16 / 4 + 64;
        static const cudaError_t attr16_rc =
            cudaFuncSetAttribute(find_runs16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM);
        BWTK_CUDA(attr16_rc);
        BWTK_REQUIRE(smem <= (size_t)MAX_DYN_SMEM, "unit length too large for the run finder's shared-memory tile");
        dim3 grid((unsigned)tiles, (unsigned)ysplit);
        if (hint != nullptr && hint_list != nullptr && u_hi + 2 * GK <= (int64_t)HL_WORDS * 16 && (((uintptr_t)d_text) & 15) == 0) {
            // only the aligned groups whose 16-mer occurs twice (the caller's hint bitmap) are compared at all
            unsigned *d_hcount = reinterpret_cast<unsigned *>(d_count + 1);     // d_count[1] is free (zeroed above)
            hint_list_kernel<<<(unsigned)ceil_div(ceil_div(n, GK), 256), 256, 0, st>>>(hint, n, hint_list, d_hcount, hint_cap);
            BWTK_LAUNCH_CHECK();
            prof::Scope ps("find_runs16_list_kernel", n, st);
            find_runs16_list_kernel<<<NUM_SMS * 8, HL_WARPS * 32, 0, st>>>(d_text, n, hint_list, d_hcount, hint_cap, pk_lo, u_hi, mc,
                                                                        out);
            BWTK_LAUNCH_CHECK();
        } else {
            prof::Scope ps("find_runs16_kernel", n, st);
            find_runs16_kernel<<<grid, THREADS, smem, st>>>(d_text, n, pk_lo, u_hi, u_per_block, mc, out);
            BWTK_LAUNCH_CHECK();
        }
    }
    unsigned long long h_cand = 0;
    { int rc = read_back(&h_cand, d_count, 8, st); if (rc) return rc; }
    *m = (int64_t)h_cand;
    if ((int64_t)h_cand > ccap) return BWTK_EWORKSPACE;
    int in_first = 1;
    if (h_cand > 1) {
        int rc = rsort::sort_pairs<unsigned long long>(key0, val0, key1, val1, (int64_t)h_cand, 0,
                                                       out.abits + bits_for(u_hi), rws, st, &in_first, nullptr);
        if (rc) return rc;
    }
    *sk = in_first ? key0 : key1;
    *sv = in_first ? val0 : val1;
    if (ordered) {
        // the small units follow the sorted large ones in the list: u = small_hi first (smaller key), u_lo last
        CandOut tail = out;
        tail.key = in_first ? key0 : key1;
        tail.val = in_first ? val0 : val1;
        for (int64_t u = small_hi; u >= u_lo; u--) {
            const int64_t L = u == 1 && min_run_u1 > mc - 1 ? min_run_u1 : (mc - 1) * u;
            int rc = BWTK_OK;
            switch (u) {
            case 1: rc = small_runs_append<1>(d_text, n, L, tail, d_count, *sws, st); break;
            case 2: rc = small_runs_append<2>(d_text, n, L, tail, d_count, *sws, st); break;
            case 3: rc = small_runs_append<3>(d_text, n, L, tail, d_count, *sws, st); break;
            case 4: rc = small_runs_append<4>(d_text, n, L, tail, d_count, *sws, st); break;
            case 5: rc = small_runs_append<5>(d_text, n, L, tail, d_count, *sws, st); break;
            case 6: rc = small_runs_append<6>(d_text, n, L, tail, d_count, *sws, st); break;
            default: rc = small_runs_append<7>(d_text, n, L, tail, d_count, *sws, st); break;
            }
            if (rc) return rc;
        }
        { int rc = read_back(&h_cand, d_count, 8, st); if (rc) return rc; }
        *m = (int64_t)h_cand;
        if ((int64_t)h_cand > ccap) return BWTK_EWORKSPACE;
    }
    return BWTK_OK;
}

// bit i: suffix i shares >= min_len symbols with a neighbour in suffix order
__global__ void __launch_bounds__(256)
    repeat_hint_kernel(const int32_t *__restrict__ sa, const int32_t *__restrict__ lcp, int64_t n, int min_len,
                       uint32_t *__restrict__ bits)
{
    const int64_t j0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (j0 >= n) return;
    int v[4] = {0, 0, 0, 0};
    if (j0 + 4 <= n && (((uintptr_t)lcp) & 15) == 0) {
        const int4 q = __ldg(reinterpret_cast<const int4 *>(lcp + j0));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
        for (int k = 0; k < 4 && j0 + k < n; k++) v[k] = lcp[j0 + k];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int64_t j = j0 + k;
        if (j >= 1 && j < n && v[k] >= min_len) {
            const uint32_t a = (uint32_t)__ldg(sa + j), b = (uint32_t)__ldg(sa + j - 1);
            atomicOr(bits + (a >> 5), 1u << (a & 31));
            atomicOr(bits + (b >> 5), 1u << (b & 31));
        }
    }
}

}  // namespace strict

// ===========================================================================
// Tier 1
// ===========================================================================
namespace tier1 {

struct Params {
    const uint8_t *text;
    const uint8_t *seen;
    int64_t n;
    int m;
    int mc;
    int min_len;
    double min_entropy;
    const double *plogp;  // [10][10]: plogp[L*10 + c] = (c/L)*log2(c/L)
};

__device__ __forceinline__ bool acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

// "Would the reference emit a repeat if its scan visited position i in this
// pass?" (bwt.py:1454-1527, with the pre-pass seen mask).  Returns the array
// end, or -1.
//
// `mis` carries the end of the last run walked by this thread (first j with
// text[j] != text[j+m], or n-m): the thread's next, adjacent positions lie inside that run
// and reuse it instead of walking again -- inside a repeat array every position would
// otherwise walk to the array's end, and the whole CTA waits for those threads.
__device__ int64_t would_emit(const Params &p, int64_t i, int64_t &mis)
{
    const int m = p.m;
    const int64_t n = p.n;
    if (i >= n - m) return -1;
    if (p.seen[i]) return -1;
    const uint8_t *t = p.text;
    if (!acgt(__ldg(t + i))) return -1;   // keeps N blocks (megabases in real assemblies) from being walked
    // run of text[j] == text[j+m] from i; copies = 1 + run/m.  Tested before the ACGT check
    // of the motif: on non-repetitive sequence the run ends after one or two symbols.
    int64_t lim = n - m;
    int64_t need = (int64_t)(p.mc - 1) * m;
    int64_t j = mis;
    if (j <= i) {
        j = i;
        while (j < lim && __ldg(t + j) == __ldg(t + j + m)) {
            j++;
        }
        mis = j;
    }
    int64_t run = j - i;
    if (run < need) return -1;
    for (int q = 1; q < m; q++)
        if (!acgt(__ldg(t + i + q))) return -1;
    int64_t copies = 1 + run / m;
    if (copies < p.mc) return -1;
    int64_t length = copies * m;
    if (length < 10) {
        // entropy of the motif, symbols accumulated in first-seen order
        int cnt[4] = {0, 0, 0, 0};
        int order[4], nd = 0;
        for (int q = 0; q < m; q++) {
            uint8_t c = __ldg(t + i + q);
            int k = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
            if (cnt[k] == 0) order[nd++] = k;
            cnt[k]++;
        }
        double e = 0.0;
        for (int q = 0; q < nd; q++) e -= p.plogp[m * 10 + cnt[order[q]]];
        if (e < p.min_entropy) return -1;
    }
    if (length < p.min_len) return -1;
    return i + length;
}

// Tier 1 candidate functor: end of the array the reference would emit at i, or -1
struct Tier1Cand {
    Params p;
    struct State { int64_t mis = -1; };
    __device__ int64_t operator()(int64_t i, State &st) const { return would_emit(p, i, st.mis); }
};

// ---- candidates from maximal runs ----------------------------------------------
// Position i is a Tier 1 candidate of motif length m iff the run of text[j]==text[j+m] from i
// is at least (mc-1)*m long, i.e. iff i lies in a maximal run [ra, rb) found by
// strict::find_runs_* with rb - i >= (mc-1)*m.  The runs of all motif lengths come from one
// launch over the text; a pass then only expands its own runs (a few percent of the
// positions) instead of testing every position.  The remaining tests are would_emit's.
struct RunCand {
    Params p;
    const unsigned long long *rkey;   // runs of this motif length, sorted by start (low abits)
    const uint32_t *rend;
    int abits;
    __device__ __forceinline__ void bounds(int64_t r, int64_t &ra, int64_t &rb, int64_t &hi) const
    {
        ra = (int64_t)(rkey[r] & ((1ull << abits) - 1ull));
        rb = (int64_t)rend[r];
        hi = rb - (int64_t)(p.mc - 1) * p.m;          // last position with enough run left
        if (hi > p.n - p.m - 1) hi = p.n - p.m - 1;
    }
    // the run region has period m: a non-ACGT symbol among its first m is in every window
    __device__ __forceinline__ bool motif_is_acgt(int64_t ra) const
    {
        for (int q = 0; q < p.m; q++)
            if (!acgt(__ldg(p.text + ra + q))) return false;
        return true;
    }
    // array end if position i of a run ending at rb is a candidate, else -1 (would_emit's tail)
    __device__ __forceinline__ int64_t test(int64_t i, int64_t rb) const
    {
        const int m = p.m;
        if (p.seen[i]) return -1;
        const int64_t copies = 1 + (rb - i) / m;
        const int64_t length = copies * m;
        if (length < 10) {
            // entropy of the motif, symbols accumulated in first-seen order
            int cnt[4] = {0, 0, 0, 0};
            int order[4], nd = 0;
            for (int q = 0; q < m; q++) {
                uint8_t c = __ldg(p.text + i + q);
                int k = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
                if (cnt[k] == 0) order[nd++] = k;
                cnt[k]++;
            }
            double e = 0.0;
            for (int q = 0; q < nd; q++) e -= p.plogp[m * 10 + cnt[order[q]]];
            if (e < p.min_entropy) return -1;
        }
        if (length < p.min_len) return -1;
        return i + length;
    }
};

// Candidates of every run.  Most runs span a handful of positions: one THREAD per run walks them; runs of
// more than RUN_SHORT positions are left to a second launch with one WARP per run (32 positions per step).
constexpr int RUN_SHORT = 48;

__global__ void __launch_bounds__(256)
    run_count_kernel(RunCand rc, int64_t nruns, uint32_t *__restrict__ cnt)
{
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= nruns) return;
    int64_t ra, rb, hi;
    rc.bounds(r, ra, rb, hi);
    if (hi - ra >= RUN_SHORT) return;          // the warp kernel's
    uint32_t total = 0;
    if (hi >= ra && rc.motif_is_acgt(ra))
        for (int64_t i = ra; i <= hi; i++) total += rc.test(i, rb) >= 0 ? 1u : 0u;
    cnt[r] = total;
}

__global__ void __launch_bounds__(256)
    run_count_long_kernel(RunCand rc, int64_t nruns, uint32_t *__restrict__ cnt)
{
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < nruns; r += (int64_t)gridDim.x * 8) {
        int64_t ra, rb, hi;
        rc.bounds(r, ra, rb, hi);
        if (hi - ra < RUN_SHORT) continue;
        uint32_t total = 0;
        if (rc.motif_is_acgt(ra)) {
            for (int64_t i0 = ra; i0 <= hi; i0 += 32) {
                const int64_t i = i0 + lane;
                const bool ok = i <= hi && rc.test(i, rb) >= 0;
                total += (uint32_t)__popc(__ballot_sync(0xffffffffu, ok));
            }
        }
        if (lane == 0) cnt[r] = total;
    }
}

// the candidates in ascending position order at off[r]..
__global__ void __launch_bounds__(256)
    run_emit_kernel(RunCand rc, int64_t nruns, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ off,
                    int32_t *__restrict__ cpos, int32_t *__restrict__ cend, unsigned long long *__restrict__ ckey,
                    uint32_t *__restrict__ cidx, int64_t step)
{
    const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= nruns) return;
    if (cnt[r] == 0) return;
    int64_t ra, rb, hi;
    rc.bounds(r, ra, rb, hi);
    if (hi - ra >= RUN_SHORT) return;
    uint32_t s = off[r];
    for (int64_t i = ra; i <= hi; i++) {
        const int64_t e = rc.test(i, rb);
        if (e >= 0) {
            cpos[s] = (int32_t)i;
            cend[s] = (int32_t)e;
            ckey[s] = ((unsigned long long)(i % step) << rc.abits) | (unsigned long long)i;
            cidx[s] = s;
            s++;
        }
    }
}

__global__ void __launch_bounds__(256)
    run_emit_long_kernel(RunCand rc, int64_t nruns, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ off,
                         int32_t *__restrict__ cpos, int32_t *__restrict__ cend, unsigned long long *__restrict__ ckey,
                         uint32_t *__restrict__ cidx, int64_t step)
{
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < nruns; r += (int64_t)gridDim.x * 8) {
        int64_t ra, rb, hi;
        rc.bounds(r, ra, rb, hi);
        if (hi - ra < RUN_SHORT || cnt[r] == 0) continue;
        uint32_t at = off[r];
        for (int64_t i0 = ra; i0 <= hi; i0 += 32) {
            const int64_t i = i0 + lane;
            const int64_t e = i <= hi ? rc.test(i, rb) : -1;
            const unsigned bal = __ballot_sync(0xffffffffu, e >= 0);
            if (e >= 0) {
                const uint32_t s = at + (uint32_t)__popc(bal & lanemask_lt());
                cpos[s] = (int32_t)i;
                cend[s] = (int32_t)e;
                ckey[s] = ((unsigned long long)(i % step) << rc.abits) | (unsigned long long)i;
                cidx[s] = s;
            }
            at += (uint32_t)__popc(bal);
        }
    }
}

struct CountArr {
    const uint32_t *cnt;
    __device__ uint64_t operator()(int64_t r) const { return cnt[r]; }
};
struct StoreOffset {
    uint32_t *off;
    __device__ void operator()(int64_t r, uint64_t excl, uint64_t) const { off[r] = (uint32_t)excl; }
};

// first run of every motif length in the sorted run list: off[u] = lower bound of
// key >= (umax - u) << abits, u = 0..umax+1 (off[0] = number of runs)
__global__ void run_offsets_kernel(const unsigned long long *__restrict__ key, int64_t m, int abits, int umax,
                                   long long *__restrict__ off)
{
    int u = threadIdx.x;
    if (u > umax + 1) return;
    if (u == 0) { off[0] = m; return; }
    if (u == umax + 1) { off[u] = 0; return; }
    const unsigned long long want = (unsigned long long)(umax - u) << abits;
    int64_t lo = 0, hi = m;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (key[mid] < want) lo = mid + 1; else hi = mid;
    }
    off[u] = lo;
}

// ---- generic greedy replay ---------------------------------------------------
// Cand: __device__ int64_t operator()(int64_t i)  -> array end if the reference's
//       scan would emit when it visits i (a pure function of the pre-pass state), else -1
// State: per-thread scratch the candidate test may keep across the thread's adjacent positions.
// The array end found in the count phase is kept for the emit phase (scan.cuh Ctx).
template <typename Cand> struct CandCtx {
    int64_t end[scan::ITEMS];
    typename Cand::State st;
};
template <typename Cand> struct CountCand {
    using Ctx = CandCtx<Cand>;
    Cand c;
    __device__ uint64_t operator()(int64_t i, int k, Ctx &ctx) const
    {
        ctx.end[k] = c(i, ctx.st);
        return ctx.end[k] >= 0 ? 1ull : 0ull;
    }
};
template <typename Cand> struct EmitCand {
    Cand c;
    int32_t *cpos;
    int32_t *cend;
    int64_t step;
    int abits;
    unsigned long long *ckey;  // (pos % step) << abits | pos
    uint32_t *cidx;
    __device__ void operator()(int64_t i, int k, uint64_t excl, uint64_t cnt, CandCtx<Cand> &ctx) const
    {
        if (!cnt) return;
        cpos[excl] = (int32_t)i;
        cend[excl] = (int32_t)ctx.end[k];
        ckey[excl] = ((unsigned long long)(i % step) << abits) | (unsigned long long)i;
        cidx[excl] = (uint32_t)excl;
    }
};

// first index in sorted `key` with key >= want, or K
__device__ __forceinline__ int64_t lower_bound(const unsigned long long *key, int64_t K, unsigned long long want)
{
    int64_t lo = 0, hi = K;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (key[mid] < want) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// candidate visited first when the scan (re)starts at position s: smallest
// candidate position >= s congruent to s modulo step
__device__ __forceinline__ int32_t next_cand(const unsigned long long *skey, const uint32_t *sidx, int64_t K,
                                             int64_t s, int64_t step, int abits, int64_t limit)
{
    if (s >= limit) return -1;  // the reference's scan loop has ended
    unsigned long long r = (unsigned long long)(s % step);
    int64_t at = lower_bound(skey, K, (r << abits) | (unsigned long long)s);
    if (at >= K) return -1;
    if ((skey[at] >> abits) != r) return -1;
    return (int32_t)sidx[at];
}

__global__ void succ_kernel(const int32_t *__restrict__ cend, const unsigned long long *__restrict__ skey,
                            const uint32_t *__restrict__ sidx, int64_t K, int64_t step, int abits,
                            int64_t limit, int32_t *__restrict__ succ, int32_t *__restrict__ first)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) *first = next_cand(skey, sidx, K, 0, step, abits, limit);
    if (k >= K) return;
    succ[k] = next_cand(skey, sidx, K, cend[k], step, abits, limit);
}

__global__ void double_kernel(const int32_t *__restrict__ jin, int32_t *__restrict__ jout, int64_t K)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    int32_t a = jin[k];
    jout[k] = a < 0 ? -1 : jin[a];
}

// anchors every 2^kappa hops along the path from `first`
__global__ void coarse_walk_kernel(const int32_t *__restrict__ jbig, const int32_t *__restrict__ first,
                                   int32_t *__restrict__ anchor, unsigned *__restrict__ nanchor)
{
    if (blockIdx.x || threadIdx.x) return;
    int32_t v = *first;
    unsigned a = 0;
    while (v >= 0) {
        anchor[a++] = v;
        v = jbig[v];
    }
    *nanchor = a;
}

// second level: from every coarse anchor, 2^kappa jumps of 2^kappa hops each -> anchor2[a << kappa | j] (-1: none)
__global__ void mid_walk_kernel(const int32_t *__restrict__ jmid, const int32_t *__restrict__ anchor,
                                const unsigned *__restrict__ nanchor, int kappa, int32_t *__restrict__ anchor2)
{
    int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= (int64_t)*nanchor) return;
    int32_t v = anchor[a];
    int32_t *dst = anchor2 + (a << kappa);
    for (int64_t j = 0; j < (1ll << kappa); j++) {
        dst[j] = v;
        if (v >= 0) v = jmid[v];
    }
}

// third level: 2^kappa single hops from every second-level anchor, marking the path
__global__ void fine_walk_kernel(const int32_t *__restrict__ succ, const int32_t *__restrict__ anchor2,
                                 const unsigned *__restrict__ nanchor, int kappa, uint8_t *__restrict__ onpath)
{
    int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= ((int64_t)*nanchor << kappa)) return;
    int32_t v = anchor2[a];
    for (int64_t h = 0; h < (1ll << kappa) && v >= 0; h++) {
        onpath[v] = 1;
        v = succ[v];
    }
}

struct CountPath {
    const uint8_t *onpath;
    __device__ uint64_t operator()(int64_t k) const { return onpath[k]; }
};
// Writer: __device__ void operator()(int64_t start, int64_t end, int32_t *row)
template <typename Writer> struct EmitPath {
    const int32_t *cpos;
    const int32_t *cend;
    Writer w;
    int32_t *rec;
    int64_t rec_base;
    int64_t cap;
    __device__ void operator()(int64_t k, uint64_t excl, uint64_t cnt) const
    {
        if (!cnt) return;
        int64_t slot = rec_base + (int64_t)excl;
        int32_t row[BWTK_REC_W];
        w(cpos[k], cend[k], row);   // side effects (seen mask) happen even past the capacity
        if (slot < cap) {
            int32_t *dst = rec + slot * BWTK_REC_W;
#pragma unroll
            for (int q = 0; q < BWTK_REC_W; q++) dst[q] = row[q];
        }
    }
};

struct Tier1Writer {
    int m;
    uint8_t *seen;
    __device__ void operator()(int64_t s, int64_t e, int32_t *row) const
    {
        for (int64_t q = s; q < e; q++) seen[q] = 1;
        row[0] = (int32_t)s; row[1] = (int32_t)e; row[2] = m; row[3] = (int32_t)((e - s) / m);
        row[4] = 0; row[5] = 0; row[6] = 0; row[7] = 0;
    }
};

// scratch shared by the passes of one call
struct PassWs {
    int32_t *cpos, *cend, *succ, *j0, *j1, *j2, *anchor, *anchor2;
    uint32_t *cidx0, *cidx1;
    unsigned long long *ckey0, *ckey1;
    uint8_t *onpath;
    int32_t *d_first;
    unsigned *d_nanchor;
    rsort::Workspace rws;
    scan::Workspace sws;
};

static int64_t pass_ws_bytes(int64_t n)
{
    return 10 * align_up(n * 4, 256) + align_up(n * 4 + 4096, 256) + 2 * align_up(n * 8, 256) + align_up(n, 256) + rsort::workspace_bytes(n) +
           scan::workspace_bytes(n) + 4096;
}

static PassWs carve_pass_ws(Carver &c, int64_t n)
{
    PassWs w;
    w.cpos = c.take<int32_t>(n); w.cend = c.take<int32_t>(n);
    w.cidx0 = c.take<uint32_t>(n); w.cidx1 = c.take<uint32_t>(n);
    w.succ = c.take<int32_t>(n); w.j0 = c.take<int32_t>(n); w.j1 = c.take<int32_t>(n); w.j2 = c.take<int32_t>(n);
    w.anchor = c.take<int32_t>(n); w.anchor2 = c.take<int32_t>(2 * n + 1024);
    w.ckey0 = c.take<unsigned long long>(n); w.ckey1 = c.take<unsigned long long>(n);
    w.onpath = c.take<uint8_t>(n);
    w.rws = rsort::carve(c, n);
    w.sws = scan::carve(c, n);
    w.d_first = c.take<int32_t>(4);
    w.d_nanchor = c.take<unsigned>(4);
    return w;
}

// One greedy pass: positions [0, npos) are scanned with stride `step`, an emission
// at c jumps to end(c); the scan stops at `limit`.  Appends rows at rec_base and
// returns the number of emissions in *emitted.
// Replays the reference's walk over K candidates held in w (cpos/cend ascending, ckey0/cidx0).
// `path_bound`: an upper bound of the number of candidates on the visited path (K if unknown).
template <typename Writer>
static int greedy_replay(int64_t K, int64_t path_bound, int64_t limit, int64_t step, int abits, Writer writer,
                         const PassWs &w, int32_t *d_rec, int64_t rec_base, int64_t cap, int64_t *emitted,
                         cudaStream_t st)
{
    *emitted = 0;
    if (K == 0) return BWTK_OK;
    int rc;
    const unsigned long long *skey = w.ckey0;
    const uint32_t *sidx = w.cidx0;
    if (step > 1) {
        int in_first = 1;
        int sbits = 1;
        while ((1ll << sbits) <= step) sbits++;
        rc = rsort::sort_pairs<unsigned long long>(w.ckey0, w.cidx0, w.ckey1, w.cidx1, K, 0, abits + sbits, w.rws, st,
                                                   &in_first, nullptr);
        if (rc) return rc;
        skey = in_first ? w.ckey0 : w.ckey1;
        sidx = in_first ? w.cidx0 : w.cidx1;
    }
    succ_kernel<<<(unsigned)ceil_div(K, 256), 256, 0, st>>>(w.cend, skey, sidx, K, step, abits, limit, w.succ,
                                                          w.d_first);
    BWTK_LAUNCH_CHECK();
    // Three-level walk along the visited path (at most `plen` candidates): jump tables for 2^kappa and 4^kappa
    // hops, kappa ~ log2(plen)/3.  One thread takes plen/4^kappa coarse jumps, then a thread per coarse anchor
    // 2^kappa middle jumps, then a thread per middle anchor 2^kappa single hops -- ~3 plen^(1/3) dependent loads
    // in a row instead of the 2 sqrt(plen) of a two-level walk (a pass over a chromosome has millions of runs).
    const int64_t plen = path_bound > 0 && path_bound < K ? path_bound : K;
    int kappa = 1;
    while ((1ll << (3 * kappa)) < plen) kappa++;
    const int32_t *cur = w.succ, *jmid = w.succ;
    int32_t *bufs[3] = {w.j0, w.j1, w.j2};
    for (int r = 0; r < 2 * kappa; r++) {
        int32_t *dst = nullptr;
        for (int q = 0; q < 3; q++)
            if (bufs[q] != cur && bufs[q] != jmid) { dst = bufs[q]; break; }
        double_kernel<<<(unsigned)ceil_div(K, 256), 256, 0, st>>>(cur, dst, K);
        BWTK_LAUNCH_CHECK();
        cur = dst;
        if (r == kappa - 1) jmid = cur;
    }
    coarse_walk_kernel<<<1, 32, 0, st>>>(cur, w.d_first, w.anchor, w.d_nanchor);
    BWTK_LAUNCH_CHECK();
    BWTK_CUDA(bwtk::zero_async(w.onpath, (size_t)K, st));
    const int64_t max_anchor = (plen >> (2 * kappa)) + 2;
    mid_walk_kernel<<<(unsigned)ceil_div(max_anchor, 128), 128, 0, st>>>(jmid, w.anchor, w.d_nanchor, kappa, w.anchor2);
    BWTK_LAUNCH_CHECK();
    fine_walk_kernel<<<(unsigned)ceil_div(max_anchor << kappa, 128), 128, 0, st>>>(w.succ, w.anchor2, w.d_nanchor, kappa,
                                                                                 w.onpath);
    BWTK_LAUNCH_CHECK();
    CountPath cp{w.onpath};
    EmitPath<Writer> ep{w.cpos, w.cend, writer, d_rec, rec_base, cap};
    rc = scan::run(K, cp, ep, w.sws, st);
    if (rc) return rc;
    unsigned long long hE = 0;
    rc = read_back(&hE, w.sws.total, 8, st);
    if (rc) return rc;
    *emitted = (int64_t)hE;
    return BWTK_OK;
}

// One greedy pass: positions [0, npos) are scanned with stride `step`, an emission
// at c jumps to end(c); the scan stops at `limit`.  Appends rows at rec_base and
// returns the number of emissions in *emitted.  Candidates: cand(i) for every position.
template <typename Cand, typename Writer>
static int greedy_pass(int64_t npos, int64_t limit, int64_t step, int abits, Cand cand, Writer writer,
                       const PassWs &w, int32_t *d_rec, int64_t rec_base, int64_t cap, int64_t *emitted,
                       cudaStream_t st)
{
    *emitted = 0;
    if (npos <= 0) return BWTK_OK;
    CountCand<Cand> cc{cand};
    EmitCand<Cand> ec{cand, w.cpos, w.cend, step, abits, w.ckey0, w.cidx0};
    int rc = scan::run(npos, cc, ec, w.sws, st);
    if (rc) return rc;
    unsigned long long hK = 0;
    rc = read_back(&hK, w.sws.total, 8, st);
    if (rc) return rc;
    return greedy_replay((int64_t)hK, (int64_t)hK, limit, step, abits, writer, w, d_rec, rec_base, cap, emitted, st);
}

// The same pass with the candidates expanded from `nruns` maximal runs (Tier 1): a warp per run
// counts, the counts are scanned, a warp per run writes.  An emission jumps past its run, so the
// visited path has at most `nruns` candidates, which sizes the two-level walk.
template <typename Writer>
static int greedy_pass_runs(const RunCand &rcand, int64_t nruns, int64_t limit, int64_t step, Writer writer,
                            const PassWs &w, int32_t *d_rec, int64_t rec_base, int64_t cap, int64_t *emitted,
                            cudaStream_t st)
{
    *emitted = 0;
    if (nruns <= 0) return BWTK_OK;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(w.j0), *off = reinterpret_cast<uint32_t *>(w.j1);   // free until the replay
    const unsigned long_grid = (unsigned)(ceil_div(nruns, 8) < NUM_SMS * 16 ? ceil_div(nruns, 8) : NUM_SMS * 16);
    run_count_kernel<<<(unsigned)ceil_div(nruns, 256), 256, 0, st>>>(rcand, nruns, cnt);
    BWTK_LAUNCH_CHECK();
    run_count_long_kernel<<<long_grid, 256, 0, st>>>(rcand, nruns, cnt);
    BWTK_LAUNCH_CHECK();
    CountArr ca{cnt};
    StoreOffset so{off};
    int rc = scan::run(nruns, ca, so, w.sws, st);
    if (rc) return rc;
    run_emit_kernel<<<(unsigned)ceil_div(nruns, 256), 256, 0, st>>>(rcand, nruns, cnt, off, w.cpos, w.cend, w.ckey0,
                                                                   w.cidx0, step);
    BWTK_LAUNCH_CHECK();
    run_emit_long_kernel<<<long_grid, 256, 0, st>>>(rcand, nruns, cnt, off, w.cpos, w.cend, w.ckey0, w.cidx0, step);
    BWTK_LAUNCH_CHECK();
    unsigned long long hK = 0;
    rc = read_back(&hK, w.sws.total, 8, st);
    if (rc) return rc;
    return greedy_replay((int64_t)hK, nruns, limit, step, rcand.abits, writer, w, d_rec, rec_base, cap, emitted, st);
}

// ---- strict adjacency with max_mismatch > 0 (bwt.py:1921-1999) -------------------
struct StrictMMCand {
    const uint8_t *text;
    int64_t n;   // without the trailing '$'
    int64_t u, mm, mc;
    struct State {};
    __device__ int64_t operator()(int64_t i, State &) const
    {
        if (i + u * mc > n) return -1;
        int64_t count = 1;
        for (;;) {
            int64_t a = i + (count - 1) * u, b = i + count * u;
            if (b + u > n) break;
            int64_t hd = 0;
            for (int64_t t = 0; t < u && hd <= mm; t++) hd += (__ldg(text + a + t) != __ldg(text + b + t));
            if (hd <= mm) count++; else break;
        }
        return count >= mc ? i + count * u : -1;
    }
};
struct StrictMMWriter {
    const uint8_t *text;
    int64_t u;
    __device__ void operator()(int64_t s, int64_t e, int32_t *row) const
    {
        int64_t prim = u;
        for (int64_t p = 1; p <= u / 2; p++) {
            if (u % p) continue;
            bool ok = true;
            for (int64_t j = p; j < u; j++)
                if (__ldg(text + s + j) != __ldg(text + s + j - p)) { ok = false; break; }
            if (ok) { prim = p; break; }
        }
        row[0] = (int32_t)s; row[1] = (int32_t)e; row[2] = (int32_t)prim;
        row[3] = (int32_t)(prim < u ? (e - s) / prim : (e - s) / u);
        row[4] = 0; row[5] = 0; row[6] = (int32_t)u; row[7] = 0;
    }
};

}  // namespace tier1

// ===========================================================================
// LCP plateaus
// ===========================================================================
namespace plateau {

__global__ void max_kernel(const int32_t *__restrict__ lcp, int64_t n, int *__restrict__ out)
{
    int v = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int x = lcp[i];
        v = x > v ? x : v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        int t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out, v);
}

// hi 31 bits: plateau starts, lo 31 bits: plateau members.  One scan element = a chunk of 16 LCP values
// (four 16-byte loads; the LCP array of an index build is 16-byte aligned): members are rare, and a scan
// element per LCP value spent ten times the array's read time on per-element bookkeeping.
struct CountMembers {
    const int32_t *lcp;
    int64_t n;
    int thr;
    // bit j of *in: lcp[16c + j] >= thr; *prev: the same for lcp[16c - 1]
    __device__ __forceinline__ void masks(int64_t c, uint32_t *in, bool *prev) const
    {
        const int64_t r0 = c * 16;
        uint32_t m = 0;
        if (r0 + 16 <= n && (((uintptr_t)lcp) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int4 v = __ldg(reinterpret_cast<const int4 *>(lcp + r0) + q);
                m |= (uint32_t)(v.x >= thr) << (4 * q) | (uint32_t)(v.y >= thr) << (4 * q + 1) |
                     (uint32_t)(v.z >= thr) << (4 * q + 2) | (uint32_t)(v.w >= thr) << (4 * q + 3);
            }
        } else {
            for (int j = 0; j < 16 && r0 + j < n; j++) m |= (uint32_t)(lcp[r0 + j] >= thr) << j;
        }
        *in = m;
        *prev = r0 > 0 && lcp[r0 - 1] >= thr;
    }
    __device__ uint64_t operator()(int64_t c) const
    {
        uint32_t in;
        bool prev;
        masks(c, &in, &prev);
        const uint32_t starts = in & ~((in << 1) | (prev ? 1u : 0u));
        return ((uint64_t)__popc(starts) << 31) | (uint64_t)__popc(in);
    }
};
struct EmitMembers {
    CountMembers cm;
    const int32_t *sa;
    int abits;
    unsigned long long *key;  // (segment id << abits) | text position
    uint32_t *val;
    __device__ void operator()(int64_t c, uint64_t excl, uint64_t cnt) const
    {
        if (!(cnt & 0x7fffffffull)) return;
        uint32_t in;
        bool prev;
        cm.masks(c, &in, &prev);
        uint64_t seg = excl >> 31;              // starts before this chunk
        uint64_t slot = excl & 0x7fffffffull;
        bool was = prev;
        for (int j = 0; j < 16; j++) {
            const bool is = (in >> j) & 1u;
            if (is) {
                if (!was) seg++;                // starts up to and including this member
                key[slot] = (seg << abits) | (uint64_t)(uint32_t)sa[c * 16 + j];
                val[slot] = 0;
                slot++;
            }
            was = is;
        }
    }
};

// per sorted member: arithmetic progression with difference `period`, validation
// (>= 80 % of positions equal to the first unit), record row
__global__ void __launch_bounds__(256)
    progress_kernel(const uint8_t *__restrict__ text, int64_t n_text, const unsigned long long *__restrict__ key,
                    int64_t M, int abits, int64_t period, int64_t mc, uint8_t *__restrict__ emit,
                    int32_t *__restrict__ rows)
{
    int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= M) return;
    const unsigned long long amask = (1ull << abits) - 1ull;
    unsigned long long ka = key[a];
    unsigned long long seg = ka >> abits;
    int64_t start = (int64_t)(ka & amask);
    int64_t copies = 1;
    for (int64_t b = a + 1; b < M; b++) {
        unsigned long long kb = key[b];
        if ((kb >> abits) != seg) break;
        if ((int64_t)(kb & amask) != start + copies * period) break;
        copies++;
    }
    bool ok = copies >= mc && (start + period <= n_text);
    if (ok) {
        int64_t rep_end = start + copies * period;
        int64_t mlen = (rep_end <= n_text ? rep_end : n_text) - start;
        ok = mlen >= 2 * period;
        if (ok) {
            int64_t match = 0;
            for (int64_t t = 0; t < mlen; t++)
                match += (__ldg(text + start + t) == __ldg(text + start + (t % period)));
            ok = ((double)match / (double)mlen) >= 0.8;
        }
        if (ok) {
            int32_t *row = rows + a * BWTK_REC_W;
            row[0] = (int32_t)start; row[1] = (int32_t)rep_end; row[2] = (int32_t)period;
            row[3] = (int32_t)copies; row[4] = 0; row[5] = 0; row[6] = 0; row[7] = 0;
        }
    }
    emit[a] = ok ? 1 : 0;
}

}  // namespace plateau
}  // namespace bwtk

using namespace bwtk;

// ---------------------------------------------------------------------------
// strict scan entry point
// ---------------------------------------------------------------------------
extern "C" int64_t bwtk_strict_workspace_bytes(int64_t n, int64_t)
{
    if (n < 1) n = 1;
    int64_t cap = strict::cand_capacity(n);
    int64_t exact = 2 * align_up(cap * 8, 256) + 2 * align_up(cap * 4, 256) + align_up(cap, 256) +
                    align_up(cap * BWTK_REC_W * 4, 256) + rsort::workspace_bytes(cap) + scan::workspace_bytes(cap) +
                    align_up((n / strict::GK + 64) * 4, 256) + 8192;
    int64_t general = tier1::pass_ws_bytes(n) + 8192;  // max_mismatch > 0: greedy replay per unit length
    return exact > general ? exact : general;
}

extern "C" int64_t bwtk_repeat_hint_bytes(int64_t n) { return (n / 32 + 2) * 4; }

extern "C" int32_t bwtk_repeat_hint(const int32_t *d_sa, const int32_t *d_lcp, int64_t n, int32_t min_len,
                                    uint32_t *d_bits, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return BWTK_OK;
    BWTK_REQUIRE(d_sa && d_lcp && d_bits, "null pointer");
    BWTK_REQUIRE(min_len >= 1, "min_len must be >= 1");
    BWTK_CUDA(bwtk::zero_async(d_bits, (size_t)bwtk_repeat_hint_bytes(n), st));
    strict::repeat_hint_kernel<<<(unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, st>>>(d_sa, d_lcp, n, min_len, d_bits);
    BWTK_LAUNCH_CHECK();
    return BWTK_OK;
}

extern "C" int32_t bwtk_strict_scan(const uint8_t *d_text, int64_t n_total, int64_t min_unit_len,
                                    int64_t max_unit_len, int64_t max_mismatch, int64_t min_copies,
                                    int32_t *d_rec, int64_t cap, int64_t *h_count, void *d_ws,
                                    int64_t ws_bytes, void *stream)
{
    return bwtk_strict_scan_hinted(d_text, n_total, min_unit_len, max_unit_len, max_mismatch, min_copies, d_rec, cap, h_count,
                                   nullptr, 0, d_ws, ws_bytes, stream);
}

extern "C" int32_t bwtk_strict_scan_hinted(const uint8_t *d_text, int64_t n_total, int64_t min_unit_len,
                                           int64_t max_unit_len, int64_t max_mismatch, int64_t min_copies,
                                           int32_t *d_rec, int64_t cap, int64_t *h_count, const uint32_t *d_hint_bits,
                                           int32_t hint_len, void *d_ws, int64_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    BWTK_REQUIRE(h_count, "null count");
    *h_count = 0;
    BWTK_REQUIRE(min_copies >= 2, "min_copies must be >= 2");
    BWTK_REQUIRE(min_unit_len >= 1, "min_unit_len must be >= 1");
    BWTK_REQUIRE(max_mismatch >= 0, "max_mismatch must be >= 0");
    if (n_total <= 0) return BWTK_OK;
    BWTK_REQUIRE(d_text && d_ws && (d_rec || cap == 0), "null pointer");
    BWTK_REQUIRE(n_total < (1ll << 30), "n must be < 2^30");
    int64_t n = n_total;
    {   // exclude a trailing '$' (bwt.py:1915-1916)
        uint8_t last = 0;
        BWTK_CUDA(cudaMemcpyAsync(&last, d_text + n - 1, 1, cudaMemcpyDeviceToHost, st));
        BWTK_CUDA(cudaStreamSynchronize(st));
        if (last == 36) n--;
    }
    int64_t umax = n / min_copies < max_unit_len ? n / min_copies : max_unit_len;
    if (umax < min_unit_len || n <= 0) return BWTK_OK;
    BWTK_REQUIRE(umax <= 60000, "max_unit_len above 60000 is not supported on device");
    if (ws_bytes < bwtk_strict_workspace_bytes(n_total, max_unit_len)) {
        set_error("strict workspace: need %lld bytes", (long long)bwtk_strict_workspace_bytes(n_total, max_unit_len));
        return BWTK_EWORKSPACE;
    }
    if (max_mismatch > 0) {
        // Hamming-tolerant adjacency: candidates "count(i) >= min_copies" per unit
        // length, greedy order replayed by pointer jumping (same engine as Tier 1).
        Carver cg(d_ws, ws_bytes);
        tier1::PassWs pw = tier1::carve_pass_ws(cg, n_total);
        if (!cg.ok()) { set_error("strict workspace carve overflow"); return BWTK_EWORKSPACE; }
        BWTK_CUDA(bwtk::zero_async(pw.rws.err, sizeof(int), st));
        BWTK_CUDA(bwtk::zero_async(pw.sws.err, sizeof(int), st));
        const int abits_g = strict::bits_for(n);
        int64_t total = 0;
        for (int64_t u = umax; u >= min_unit_len; u--) {
            int64_t npos = n - u * min_copies + 1;  // positions the while-loop can visit
            if (npos <= 0) continue;
            tier1::StrictMMCand cand{d_text, n, u, max_mismatch, min_copies};
            tier1::StrictMMWriter wr{d_text, u};
            int64_t emitted = 0;
            int rc = tier1::greedy_pass(npos, npos, 1, abits_g, cand, wr, pw, d_rec, total, cap, &emitted, st);
            if (rc) return rc;
            total += emitted;
        }
        int h_err[2] = {0, 0};
        BWTK_CUDA(cudaMemcpyAsync(&h_err[0], pw.rws.err, 4, cudaMemcpyDeviceToHost, st));
        BWTK_CUDA(cudaMemcpyAsync(&h_err[1], pw.sws.err, 4, cudaMemcpyDeviceToHost, st));
        BWTK_CUDA(cudaStreamSynchronize(st));
        if (h_err[0] || h_err[1]) { set_error("look-back spin limit hit in strict scan"); return BWTK_EINTERNAL; }
        *h_count = total;
        if (total > cap) {
            set_error("strict scan: %lld records exceed capacity %lld", (long long)total, (long long)cap);
            return BWTK_EOVERFLOW;
        }
        return BWTK_OK;
    }
    const int64_t ccap = strict::cand_capacity(n_total);
    Carver c(d_ws, ws_bytes);
    unsigned long long *key0 = c.take<unsigned long long>(ccap);
    unsigned long long *key1 = c.take<unsigned long long>(ccap);
    uint32_t *val0 = c.take<uint32_t>(ccap);
    uint32_t *val1 = c.take<uint32_t>(ccap);
    uint8_t *emit = c.take<uint8_t>(ccap);
    int32_t *rows = c.take<int32_t>(ccap * BWTK_REC_W);
    rsort::Workspace rws = rsort::carve(c, ccap);
    scan::Workspace sws = scan::carve(c, ccap);
    unsigned long long *d_count = c.take<unsigned long long>(2);
    const int64_t hint_cap = n_total / strict::GK + 64;
    uint32_t *hint_list = c.take<uint32_t>(hint_cap);
    if (!c.ok()) { set_error("strict workspace carve overflow"); return BWTK_EWORKSPACE; }
    // a hint of h symbols filters groups of 16 matching positions iff h <= 16 (a repeated 16-mer repeats its prefixes)
    const uint32_t *hint = (d_hint_bits != nullptr && hint_len >= 1 && hint_len <= strict::GK) ? d_hint_bits : nullptr;
    BWTK_CUDA(bwtk::zero_async(rws.err, sizeof(int), st));
    BWTK_CUDA(bwtk::zero_async(sws.err, sizeof(int), st));
    const unsigned long long *sk = nullptr;
    const uint32_t *sv = nullptr;
    int64_t m = 0;
    int rc = strict::collect_runs(d_text, n, min_unit_len, umax, min_copies, key0, key1, val0, val1, d_count, ccap, rws,
                                  st, &sk, &sv, &m, &sws, 0, hint, hint_list, (unsigned)hint_cap);
    if (rc == BWTK_EWORKSPACE) {
        set_error("strict scan: %lld candidate runs exceed the workspace capacity %lld", (long long)m, (long long)ccap);
        return rc;
    }
    if (rc) return rc;
    if (m == 0) return BWTK_OK;
    const int abits_runs = strict::bits_for(n);
    strict::Resolved res{emit, rows};
    strict::resolve_walk_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(sk, sv, m, abits_runs, umax, min_copies, res);
    BWTK_LAUNCH_CHECK();
    strict::resolve_rows_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(d_text, sk, m, abits_runs, umax, res);
    BWTK_LAUNCH_CHECK();
    strict::CountEmit ce{emit};
    strict::WriteRows wr{rows, d_rec, cap};
    rc = scan::run(m, ce, wr, sws, st);
    if (rc) return rc;
    unsigned long long h_total = 0;
    int h_err[2] = {0, 0};
    BWTK_CUDA(cudaMemcpyAsync(&h_total, sws.total, 8, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaMemcpyAsync(&h_err[0], rws.err, 4, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaMemcpyAsync(&h_err[1], sws.err, 4, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaStreamSynchronize(st));
    if (h_err[0] || h_err[1]) { set_error("look-back spin limit hit in strict scan"); return BWTK_EINTERNAL; }
    *h_count = (int64_t)h_total;
    if ((int64_t)h_total > cap) {
        set_error("strict scan: %llu records exceed capacity %lld", h_total, (long long)cap);
        return BWTK_EOVERFLOW;
    }
    return BWTK_OK;
}

// ---------------------------------------------------------------------------
// Tier 1 entry point
// ---------------------------------------------------------------------------
extern "C" int64_t bwtk_tier1_workspace_bytes(int64_t n)
{
    if (n < 1) n = 1;
    const int64_t ccap = strict::cand_capacity(n);             // maximal runs of the motif lengths 1..9
    return align_up(n, 256) + tier1::pass_ws_bytes(n) + 8192 +  // seen mask + pass scratch + entropy table
           2 * align_up(ccap * 8, 256) + 2 * align_up(ccap * 4, 256) + 1024;
}

extern "C" int32_t bwtk_tier1_scan(const uint8_t *d_text, int64_t n, int32_t max_motif_len,
                                   int32_t min_copies, int32_t min_array_len, double min_entropy,
                                   int32_t *d_rec, int64_t cap, int64_t *h_count, uint8_t *d_seen_out,
                                   void *d_ws, int64_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    BWTK_REQUIRE(h_count, "null count");
    *h_count = 0;
    if (n <= 0) return BWTK_OK;
    BWTK_REQUIRE(d_text && d_ws && (d_rec || cap == 0), "null pointer");
    BWTK_REQUIRE(n < (1ll << 30), "n must be < 2^30");
    BWTK_REQUIRE(min_copies >= 1, "min_copies must be >= 1");
    if (ws_bytes < bwtk_tier1_workspace_bytes(n)) {
        set_error("tier1 workspace: need %lld bytes", (long long)bwtk_tier1_workspace_bytes(n));
        return BWTK_EWORKSPACE;
    }
    Carver c(d_ws, ws_bytes);
    uint8_t *seen = c.take<uint8_t>(n);
    tier1::PassWs pw = tier1::carve_pass_ws(c, n);
    double *d_plogp = c.take<double>(100);
    const int64_t ccap = strict::cand_capacity(n);
    unsigned long long *rkey0 = c.take<unsigned long long>(ccap);
    unsigned long long *rkey1 = c.take<unsigned long long>(ccap);
    uint32_t *rval0 = c.take<uint32_t>(ccap);
    uint32_t *rval1 = c.take<uint32_t>(ccap);
    unsigned long long *d_runcount = c.take<unsigned long long>(2);
    long long *d_runoff = c.take<long long>(16);
    if (!c.ok()) { set_error("tier1 workspace carve overflow"); return BWTK_EWORKSPACE; }
    BWTK_CUDA(bwtk::zero_async(seen, (size_t)n, st));
    BWTK_CUDA(bwtk::zero_async(pw.rws.err, sizeof(int), st));
    BWTK_CUDA(bwtk::zero_async(pw.sws.err, sizeof(int), st));
    // (c/L)*log2(c/L) in IEEE double, the same expression the host evaluates
    double plogp[100];
    for (int L = 0; L < 10; L++)
        for (int k = 0; k < 10; k++) {
            double v = 0.0;
            if (L > 0 && k > 0 && k <= L) { double p = (double)k / (double)L; v = p * log2(p); }
            plogp[L * 10 + k] = v;
        }
    BWTK_CUDA(cudaMemcpyAsync(d_plogp, plogp, sizeof(plogp), cudaMemcpyHostToDevice, st));
    BWTK_CUDA(cudaStreamSynchronize(st));

    // adaptive sampling of the reference (bwt.py:1441-1448)
    int64_t step = n > 10000000 ? 50 : (n > 5000000 ? 20 : 1);
    const int abits = strict::bits_for(n);
    int64_t total = 0;
    int mmax = max_motif_len < 9 ? max_motif_len : 9;
    // maximal runs of text[j]==text[j+m] for every motif length at once (min_copies >= 2: a
    // candidate then sits inside a run; with min_copies == 1 every position is one)
    const unsigned long long *rk = nullptr;
    const uint32_t *rv = nullptr;
    long long runoff[16] = {0};
    bool by_runs = false;
    if (min_copies >= 2 && mmax >= 1) {
        int64_t nruns = 0;
        // a one-symbol motif has entropy 0: with a positive entropy floor a homopolymer array is only ever
        // emitted from 10 copies on (bwt.py:1499-1508), i.e. from runs of 9 -- which spares the list the
        // ~5 % of all positions that start a run of 2..8
        const int64_t min_run_u1 = min_entropy > 0.0 ? 9 : 0;
        int rc = strict::collect_runs(d_text, n, 1, mmax, min_copies, rkey0, rkey1, rval0, rval1, d_runcount, ccap,
                                      pw.rws, st, &rk, &rv, &nruns, &pw.sws, min_run_u1);
        if (rc && rc != BWTK_EWORKSPACE) return rc;
        if (rc == BWTK_OK) {
            by_runs = true;
            if (nruns > 0) {
                tier1::run_offsets_kernel<<<1, 32, 0, st>>>(rk, nruns, abits, mmax, d_runoff);
                BWTK_LAUNCH_CHECK();
                rc = read_back(runoff, d_runoff, sizeof(long long) * (size_t)(mmax + 2), st);
                if (rc) return rc;
            }
        }   // more runs than the buffer holds (pathological text): per-position candidates below
    }
    for (int m = mmax; m >= 1; m--) {
        if (n - m <= 0) continue;
        tier1::Params p{d_text, seen, n, m, min_copies, min_array_len, min_entropy, d_plogp};
        tier1::Tier1Writer wr{m, seen};
        int64_t emitted = 0;
        int rc;
        if (by_runs) {
            // runs of length m occupy [off[m], off[m-1]) of the list sorted by (m descending, start)
            const int64_t lo = runoff[m], hi = runoff[m - 1];
            tier1::RunCand rcand{p, rk + lo, rv + lo, abits};
            rc = tier1::greedy_pass_runs(rcand, hi - lo, n - m, step, wr, pw, d_rec, total, cap, &emitted, st);
        } else {
            tier1::Tier1Cand cand{p};
            rc = tier1::greedy_pass(n - m, n - m, step, abits, cand, wr, pw, d_rec, total, cap, &emitted, st);
        }
        if (rc) return rc;
        total += emitted;
    }
    int h_err[2] = {0, 0};
    BWTK_CUDA(cudaMemcpyAsync(&h_err[0], pw.rws.err, 4, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaMemcpyAsync(&h_err[1], pw.sws.err, 4, cudaMemcpyDeviceToHost, st));
    if (d_seen_out) BWTK_CUDA(cudaMemcpyAsync(d_seen_out, seen, (size_t)n, cudaMemcpyDeviceToDevice, st));
    BWTK_CUDA(cudaStreamSynchronize(st));
    if (h_err[0] || h_err[1]) { set_error("look-back spin limit hit in tier1 scan"); return BWTK_EINTERNAL; }
    *h_count = total;
    if (total > cap) {
        set_error("tier1 scan: %lld records exceed capacity %lld", (long long)total, (long long)cap);
        return BWTK_EOVERFLOW;
    }
    return BWTK_OK;
}

// ---------------------------------------------------------------------------
// LCP plateau entry point
// ---------------------------------------------------------------------------
extern "C" int64_t bwtk_plateau_workspace_bytes(int64_t n)
{
    if (n < 1) n = 1;
    return 2 * align_up(n * 8, 256) + 2 * align_up(n * 4, 256) + align_up(n, 256) +
           align_up(n * BWTK_REC_W * 4, 256) + rsort::workspace_bytes(n) + scan::workspace_bytes(n) + 8192;
}

extern "C" int32_t bwtk_lcp_plateaus(const uint8_t *d_text, int64_t n_text, const int32_t *d_sa,
                                     const int32_t *d_lcp, int64_t n, int64_t min_period, int64_t max_period,
                                     int64_t min_copies, int32_t *d_rec, int64_t cap, int64_t *h_count,
                                     int64_t *h_threshold, void *d_ws, int64_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    BWTK_REQUIRE(h_count && h_threshold, "null output");
    *h_count = 0;
    *h_threshold = -1;
    if (n <= 0) return BWTK_OK;
    BWTK_REQUIRE(d_text && d_sa && d_lcp && d_ws && (d_rec || cap == 0), "null pointer");
    BWTK_REQUIRE(n < (1ll << 30) && n_text < (1ll << 30), "n must be < 2^30");
    if (ws_bytes < bwtk_plateau_workspace_bytes(n)) {
        set_error("plateau workspace: need %lld bytes", (long long)bwtk_plateau_workspace_bytes(n));
        return BWTK_EWORKSPACE;
    }
    Carver c(d_ws, ws_bytes);
    unsigned long long *key0 = c.take<unsigned long long>(n);
    unsigned long long *key1 = c.take<unsigned long long>(n);
    uint32_t *val0 = c.take<uint32_t>(n);
    uint32_t *val1 = c.take<uint32_t>(n);
    uint8_t *emit = c.take<uint8_t>(n);
    int32_t *rows = c.take<int32_t>(n * BWTK_REC_W);
    rsort::Workspace rws = rsort::carve(c, n);
    scan::Workspace sws = scan::carve(c, n);
    int *d_max = c.take<int>(4);
    if (!c.ok()) { set_error("plateau workspace carve overflow"); return BWTK_EWORKSPACE; }
    BWTK_CUDA(bwtk::zero_async(d_max, sizeof(int), st));
    BWTK_CUDA(bwtk::zero_async(rws.err, sizeof(int), st));
    BWTK_CUDA(bwtk::zero_async(sws.err, sizeof(int), st));
    int grid = (int)(ceil_div(n, 256 * 8) < NUM_SMS * 8 ? ceil_div(n, 256 * 8) : NUM_SMS * 8);
    plateau::max_kernel<<<grid, 256, 0, st>>>(d_lcp, n, d_max);
    BWTK_LAUNCH_CHECK();
    int h_max = 0;
    BWTK_CUDA(cudaMemcpyAsync(&h_max, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaStreamSynchronize(st));
    // threshold (bwt.py:2124-2129)
    if (h_max < min_period) return BWTK_OK;
    int64_t thr = max_period < h_max ? max_period : h_max;
    if (thr > 20) thr = 20;
    if (thr < min_period) thr = min_period;
    *h_threshold = thr;
    if (thr <= 0) {
        set_error("plateau threshold %lld is not positive", (long long)thr);
        return BWTK_EINVAL;
    }
    const int abits = strict::bits_for(n_text > n ? n_text : n);
    plateau::CountMembers cm{d_lcp, n, (int)thr};
    plateau::EmitMembers em{cm, d_sa, abits, key0, val0};
    int rc = scan::run(ceil_div(n, 16), cm, em, sws, st);
    if (rc) return rc;
    unsigned long long h_tot = 0;
    BWTK_CUDA(cudaMemcpyAsync(&h_tot, sws.total, 8, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaStreamSynchronize(st));
    int64_t M = (int64_t)(h_tot & 0x7fffffffull);
    int64_t segs = (int64_t)(h_tot >> 31);
    if (M == 0) return BWTK_OK;
    int in_first = 1;
    rc = rsort::sort_pairs<unsigned long long>(key0, val0, key1, val1, M, 0, abits + strict::bits_for(segs), rws,
                                               st, &in_first, nullptr);
    if (rc) return rc;
    const unsigned long long *sk = in_first ? key0 : key1;
    plateau::progress_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, st>>>(d_text, n_text, sk, M, abits, thr,
                                                                        min_copies, emit, rows);
    BWTK_LAUNCH_CHECK();
    strict::CountEmit ce{emit};
    strict::WriteRows wr{rows, d_rec, cap};
    rc = scan::run(M, ce, wr, sws, st);
    if (rc) return rc;
    unsigned long long h_total = 0;
    int h_err[2] = {0, 0};
    BWTK_CUDA(cudaMemcpyAsync(&h_total, sws.total, 8, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaMemcpyAsync(&h_err[0], rws.err, 4, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaMemcpyAsync(&h_err[1], sws.err, 4, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaStreamSynchronize(st));
    if (h_err[0] || h_err[1]) { set_error("look-back spin limit hit in plateau scan"); return BWTK_EINTERNAL; }
    *h_count = (int64_t)h_total;
    if ((int64_t)h_total > cap) {
        set_error("plateau scan: %llu records exceed capacity %lld", h_total, (long long)cap);
        return BWTK_EOVERFLOW;
    }
    return BWTK_OK;
}
