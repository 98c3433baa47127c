// extend.cu -- mismatch-tolerant extension, majority-vote consensus and the
// period scan that drives them:
//   * MotifUtils.build_consensus_motif_array      (reference bwt.py:1207-1256)
//   * Tier2LCPFinder._extend_with_mismatches       (reference bwt.py:2392-2498)
//   * Tier2LCPFinder._extend_tandem_fm             (reference bwt.py:2697-2805)
//   * Tier2LCPFinder._find_repeats_simple          (reference bwt.py:2177-2390)
//
// The reference re-derives the consensus from scratch after every tentative
// copy; here a per-column symbol tally is kept and updated incrementally
// (total mismatches = cells - sum of column maxima; consensus = first maximum =
// smallest byte on ties), which is the same decision at O(copies*period).
// The period scan runs one CTA per period pass: 128 speculative visits per
// iteration, the first survivor evaluated in full by a whole warp.  All passes run
// at once with the full 100 000-visit budget; the reference's GLOBAL budget is
// applied afterwards from the per-pass visit counts (rows past a pass's share are
// dropped), and a pass stops early as soon as its predecessors have finished and it
// knows what they left it.
#include "common.cuh"

namespace bwtk {
namespace ext {

constexpr int MAXP = 1024;    // longest period handled by the batch kernels
constexpr int SLOTS = 8;      // distinct byte values tallied per column

// Per-column tally of byte values (up to SLOTS distinct values; real DNA needs 5).
struct Tally {
    uint8_t *sym;     // [period][SLOTS]
    uint16_t *cnt;    // [period][SLOTS]
    uint8_t *nsym;    // [period]
    int period;
    bool overflow;

    __device__ void reset(int p)
    {
        period = p;
        overflow = false;
        for (int q = 0; q < p; q++) nsym[q] = 0;
    }
    __device__ void add(int col, uint8_t b, int delta)
    {
        uint8_t *s = sym + col * SLOTS;
        uint16_t *c = cnt + col * SLOTS;
        int k = nsym[col];
        for (int j = 0; j < k; j++)
            if (s[j] == b) { c[j] = (uint16_t)(c[j] + delta); return; }
        if (k >= SLOTS || delta < 0) { overflow = true; return; }
        s[k] = b; c[k] = (uint16_t)delta; nsym[col] = (uint8_t)(k + 1);
    }
    // majority symbol of a column: highest count, smallest byte on ties
    __device__ uint8_t best(int col, int *best_cnt) const
    {
        const uint8_t *s = sym + col * SLOTS;
        const uint16_t *c = cnt + col * SLOTS;
        int k = nsym[col], bc = -1;
        uint8_t bs = 0;
        for (int j = 0; j < k; j++) {
            int v = c[j];
            if (v > bc || (v == bc && s[j] < bs)) { bc = v; bs = s[j]; }
        }
        *best_cnt = bc < 0 ? 0 : bc;
        return bs;
    }
    __device__ void add_copy(const uint8_t *text, int64_t at, int delta)
    {
        for (int q = 0; q < period; q++) add(q, __ldg(text + at + q), delta);
    }
    // total Hamming distance of `copies` tallied copies to the majority consensus
    __device__ int64_t total_mm(int64_t copies) const
    {
        int64_t cells = copies * period, keep = 0;
        for (int q = 0; q < period; q++) { int bc; best(q, &bc); keep += bc; }
        return cells - keep;
    }
};

__device__ __forceinline__ int64_t mm_budget(int64_t motif_len, int64_t copies)
{
    // _get_max_mismatches_for_array (bwt.py:2003-2025), float64 like numpy
    if (motif_len == 1) return 0;
    double total = (double)(motif_len * copies);
    int64_t v = (int64_t)ceil((motif_len <= 6 ? 0.05 : 0.08) * total);
    return v > 1 ? v : 1;
}

__device__ __forceinline__ bool is_transversion(uint8_t b1, uint8_t b2)
{
    // count_transversions_array (bwt.py:780-800)
    if (b1 == b2) return false;
    uint8_t c1 = (b1 >= 65 && b1 <= 84) ? b1 : 'N';
    uint8_t c2 = (b2 >= 65 && b2 <= 84) ? b2 : 'N';
    if (c1 == c2) return false;
    if ((c1 == 'A' && c2 == 'G') || (c1 == 'G' && c2 == 'A')) return false;
    if ((c1 == 'C' && c2 == 'T') || (c1 == 'T' && c2 == 'C')) return false;
    return true;
}

// any tallied symbol that is a transversion away from its column's consensus?
__device__ bool has_transversion(const Tally &t)
{
    for (int q = 0; q < t.period; q++) {
        int bc;
        uint8_t cons = t.best(q, &bc);
        const uint8_t *s = t.sym + q * SLOTS;
        const uint16_t *c = t.cnt + q * SLOTS;
        for (int j = 0; j < t.nsym[q]; j++)
            if (c[j] > 0 && is_transversion(s[j], cons)) return true;
    }
    return false;
}

struct ExtOut { int64_t array_start, array_end, copies, full_start, full_end; };

// _extend_with_mismatches (bwt.py:2392-2498).  `cons` receives the consensus of
// the accepted copies (period bytes).
__device__ ExtOut extend_with_mismatches(const uint8_t *s, int64_t start_pos, int period, int64_t n,
                                         bool allow_mm, Tally &t, uint8_t *cons)
{
    t.reset(period);
    int64_t start = start_pos, end = start_pos + period, copies = 1;
    t.add_copy(s, start_pos, 1);
    while (end + period <= n) {
        t.add_copy(s, end, 1);
        int64_t bud = allow_mm ? mm_budget(period, copies + 1) : 0;
        if (t.total_mm(copies + 1) <= bud) { copies++; end += period; }
        else { t.add_copy(s, end, -1); break; }
    }
    while (start - period >= 0) {
        t.add_copy(s, start - period, 1);
        int64_t bud = allow_mm ? mm_budget(period, copies + 1) : 0;
        if (t.total_mm(copies + 1) <= bud) { copies++; start -= period; }
        else { t.add_copy(s, start - period, -1); break; }
    }
    for (int q = 0; q < period; q++) { int bc; cons[q] = t.best(q, &bc); }
    ExtOut o;
    o.copies = copies; o.full_start = start; o.full_end = end;
    int64_t pr = 0;
    while (pr < period && end + pr < n && __ldg(s + end + pr) == cons[pr % period]) pr++;
    int64_t pl = 0;
    while (pl < period && start - pl - 1 >= 0 &&
           __ldg(s + start - pl - 1) == cons[period - 1 - (pl % period)]) pl++;
    o.array_start = start - pl;
    o.array_end = end + pr;
    return o;
}

__device__ bool homopolymer(const uint8_t *s, int64_t at, int len)
{
    uint8_t c = __ldg(s + at);
    for (int i = 1; i < len; i++)
        if (__ldg(s + at + i) != c) return false;
    return true;
}

// _extend_tandem_fm (bwt.py:2697-2805): n = text_arr.size
__device__ ExtOut extend_tandem_fm(const uint8_t *s, int64_t n, int64_t seed, int mlen, Tally &t)
{
    t.reset(mlen);
    int64_t start = seed, end = seed + mlen, copies = 1;
    t.add_copy(s, seed, 1);
    while (end + mlen <= n) {
        if (mlen > 1 && homopolymer(s, end, mlen)) break;
        t.add_copy(s, end, 1);
        if (t.total_mm(copies + 1) <= mm_budget(mlen, copies + 1) && !has_transversion(t)) { copies++; end += mlen; }
        else { t.add_copy(s, end, -1); break; }
    }
    while (start - mlen >= 0) {
        if (mlen > 1 && homopolymer(s, start - mlen, mlen)) break;
        t.add_copy(s, start - mlen, 1);
        if (t.total_mm(copies + 1) <= mm_budget(mlen, copies + 1) && !has_transversion(t)) { copies++; start -= mlen; }
        else { t.add_copy(s, start - mlen, -1); break; }
    }
    ExtOut o;
    o.array_start = start; o.array_end = end; o.copies = copies; o.full_start = start; o.full_end = end;
    return o;
}

// build_consensus_motif_array (bwt.py:1207-1256): copies that would run past
// text_size are dropped; returns copies used.
__device__ int64_t consensus(const uint8_t *text, int64_t text_size, int64_t start, int period, int64_t copies,
                             Tally &t, uint8_t *cons, int64_t *total_mm, int64_t *max_mm)
{
    *total_mm = 0; *max_mm = 0;
    if (copies <= 0 || period <= 0) return 0;
    int64_t used = 0;
    while (used < copies && start + (used + 1) * period <= text_size) used++;
    if (!used) return 0;
    t.reset(period);
    for (int64_t c = 0; c < used; c++) t.add_copy(text, start + c * period, 1);
    for (int q = 0; q < period; q++) { int bc; cons[q] = t.best(q, &bc); }
    for (int64_t c = 0; c < used; c++) {
        int64_t mm = 0;
        for (int q = 0; q < period; q++) mm += (__ldg(text + start + c * period + q) != cons[q]);
        *total_mm += mm;
        if (mm > *max_mm) *max_mm = mm;
    }
    return used;
}

// scratch carved per thread: tally (period*SLOTS*(1+2) + period) + 2*period consensus bytes
__host__ __device__ inline int64_t scratch_per_thread(int maxp)
{
    return (((int64_t)maxp * (SLOTS * 3 + 1 + 2) + 64) + 15) / 16 * 16;  // keeps the uint16 tally aligned
}

__device__ Tally make_tally(uint8_t *scratch, int maxp)
{
    Tally t;
    t.cnt = reinterpret_cast<uint16_t *>(scratch);              // 2-byte aligned region first
    t.sym = scratch + (int64_t)maxp * SLOTS * 2;
    t.nsym = t.sym + (int64_t)maxp * SLOTS;
    t.period = 0;
    t.overflow = false;
    return t;
}

__global__ void extend_batch_kernel(const uint8_t *__restrict__ text, int64_t n, const int32_t *__restrict__ seed,
                                    const int32_t *__restrict__ period, const int32_t *__restrict__ flags,
                                    int64_t m, int mode, int maxp, uint8_t *scratch, int32_t *__restrict__ out,
                                    int *err)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint8_t *my = scratch + i * scratch_per_thread(maxp);
    Tally t = make_tally(my, maxp);
    uint8_t *cons = t.nsym + maxp;
    int p = period[i];
    int32_t *row = out + i * 8;
    for (int q = 0; q < 8; q++) row[q] = 0;
    if (p <= 0 || p > maxp || seed[i] < 0 || (int64_t)seed[i] + p > n) { row[7] = -1; return; }
    ExtOut o = mode == 0 ? extend_with_mismatches(text, seed[i], p, n, (flags[i] & 1) != 0, t, cons)
                         : extend_tandem_fm(text, n, seed[i], p, t);
    if (t.overflow) *err = 1;
    if (mode == 0) {
        row[0] = (int32_t)o.array_start; row[1] = (int32_t)o.array_end; row[2] = (int32_t)o.copies;
        row[3] = (int32_t)o.full_start; row[4] = (int32_t)o.full_end;
    } else {
        row[0] = (int32_t)o.array_start; row[1] = (int32_t)o.array_end; row[2] = (int32_t)o.copies;
    }
}

__global__ void consensus_batch_kernel(const uint8_t *__restrict__ text, int64_t text_size,
                                       const int32_t *__restrict__ start, const int32_t *__restrict__ period,
                                       const int32_t *__restrict__ copies, const int64_t *__restrict__ cons_off,
                                       int64_t m, int maxp, uint8_t *scratch, uint8_t *__restrict__ cons_out,
                                       int32_t *__restrict__ mm_out, int *err)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint8_t *my = scratch + i * scratch_per_thread(maxp);
    Tally t = make_tally(my, maxp);
    uint8_t *cons = t.nsym + maxp;
    int p = period[i];
    int32_t *row = mm_out + i * 4;
    row[0] = row[1] = row[2] = row[3] = 0;
    if (p <= 0 || p > maxp || start[i] < 0) return;
    int64_t tmm, mmm;
    int64_t used = consensus(text, text_size, start[i], p, copies[i], t, cons, &tmm, &mmm);
    if (t.overflow) *err = 1;
    row[0] = (int32_t)tmm; row[1] = (int32_t)mmm; row[2] = (int32_t)used;
    if (used)
        for (int q = 0; q < p; q++) cons_out[cons_off[i] + q] = cons[q];
}

// ---------------------------------------------------------------- period scan
struct ScanCfg {
    const uint8_t *s;
    int64_t n;        // text length without the trailing '$'
    int64_t n_total;  // text_arr.size
    int64_t min_p, max_p, per_step, pos_step;
    int allow_mm;
    int64_t min_copies, min_array_len;
    double min_entropy;
    const uint8_t *mask;
    const double *plogp;
    int64_t dim;
};

__device__ double entropy_of(const uint8_t *s, int64_t at, int len, const double *plogp, int64_t dim)
{
    // MotifUtils.calculate_entropy (bwt.py:729-745): symbols in first-seen order
    uint8_t sym[32];
    int cnt[32];
    int nd = 0;
    for (int i = 0; i < len; i++) {
        uint8_t c = __ldg(s + at + i);
        int j = 0;
        for (; j < nd; j++)
            if (sym[j] == c) { cnt[j]++; break; }
        if (j == nd && nd < 32) { sym[nd] = c; cnt[nd] = 1; nd++; }
    }
    double e = 0.0;
    for (int j = 0; j < nd; j++) e -= plogp[(int64_t)cnt[j] * dim + len];
    return e;
}

__device__ int smallest_period(const uint8_t *s, int64_t at, int len)
{
    for (int p = 1; p <= len / 2; p++) {
        if (len % p) continue;
        bool ok = true;
        for (int j = p; j < len; j++)
            if (__ldg(s + at + j) != __ldg(s + at + j - p)) { ok = false; break; }
        if (ok) return p;
    }
    return len;
}

__device__ int smallest_period_buf(const uint8_t *b, int len)
{
    for (int p = 1; p <= len / 2; p++) {
        if (len % p) continue;
        bool ok = true;
        for (int j = p; j < len; j++)
            if (b[j] != b[j - p]) { ok = false; break; }
        if (ok) return p;
    }
    return len;
}

constexpr int64_t MAX_ITER = 100000;

constexpr int PASS_THREADS = 128;

// Tests of one visit that need no extension state; true = the reference would
// just step on (bwt.py:2259-2275 plus the two-copy pre-check below).
__device__ bool visit_rejected(const ScanCfg &c, int p, int64_t i)
{
    const uint8_t *s = c.s;
    if (c.mask && c.mask[i]) return true;
    if (c.min_copies > 2) {
        // With two copies the vote ties on every differing column, so the first right
        // (left) extension is accepted iff the Hamming distance to the neighbouring copy
        // is within the budget.  If neither neighbour qualifies the array stays at one
        // copy and can never reach min_copies (>= 3).
        const int64_t bud = (c.allow_mm && p <= 64) ? mm_budget(p, 2) : 0;
        int64_t hr = 0, hl = bud + 1;
        for (int q = 0; q < p && hr <= bud; q++) hr += (__ldg(s + i + q) != __ldg(s + i + p + q));
        if (hr > bud && i - p >= 0) {
            hl = 0;
            for (int q = 0; q < p && hl <= bud; q++) hl += (__ldg(s + i + q) != __ldg(s + i - p + q));
        }
        if (hr > bud && hl > bud) return true;
    }
    for (int q = 0; q < p; q++) {
        uint8_t ch = __ldg(s + i + q);
        if (ch == 36 || ch == 78) return true;
    }
    return entropy_of(s, i, p, c.plogp, c.dim) < c.min_entropy;
}

// Full evaluation of one visit that passed the cheap tests (bwt.py:2277-2386): extension,
// primitive-period re-extension, consensus.  A pure function of the position.  Returns
// whether a record is emitted; *next_i is where the scan continues.  One-thread form.
__device__ bool visit_full(const ScanCfg &c, int p, int64_t i, Tally &t, uint8_t *cons, int32_t *row,
                           int64_t *next_i)
{
    const uint8_t *s = c.s;
    const int64_t n = c.n;
    *next_i = i + c.pos_step;
    ExtOut o = extend_with_mismatches(s, i, p, n, c.allow_mm && p <= 64, t, cons);
    int64_t a_len = o.array_end - o.array_start;
    if (a_len < c.min_array_len) return false;
    int64_t part = a_len - o.copies * p;
    if (part < 0) part = 0;
    int64_t eff = o.copies + (((double)part / (double)p) >= 0.75 ? 1 : 0);
    if (!(o.copies >= c.min_copies || eff >= c.min_copies)) return false;
    int prim = smallest_period(s, o.full_start, p);
    int p_eff = prim < p ? prim : p;
    o = extend_with_mismatches(s, o.full_start, p_eff, n, c.allow_mm && p_eff <= 64, t, cons);
    a_len = o.array_end - o.array_start;
    part = a_len - o.copies * p_eff;
    if (part < 0) part = 0;
    eff = o.copies + (((double)part / (double)p_eff) >= 0.75 ? 1 : 0);
    if (o.copies < c.min_copies && eff < c.min_copies) return false;
    int64_t tmm, mmm;
    int64_t a_start = o.array_start, a_end = o.array_end, copies_full = o.copies, cons_start = o.full_start;
    int64_t used = consensus(s, c.n_total, cons_start, p_eff, copies_full, t, cons, &tmm, &mmm);
    if (!used) return false;
    int prim2 = smallest_period_buf(cons, p_eff);
    if (prim2 < p_eff) {
        p_eff = prim2;
        copies_full = (a_end - a_start) / p_eff;
        if (copies_full < 1) copies_full = 1;
        a_end = a_start + copies_full * p_eff;
        cons_start = a_start;
        used = consensus(s, c.n_total, a_start, p_eff, copies_full, t, cons, &tmm, &mmm);
        if (!used) return false;
    }
    row[0] = (int32_t)a_start; row[1] = (int32_t)a_end; row[2] = p_eff;
    row[3] = (int32_t)copies_full; row[4] = (int32_t)tmm; row[5] = (int32_t)mmm;
    row[6] = (int32_t)cons_start; row[7] = (int32_t)used;
    *next_i = a_end;   // jump past the array (bwt.py:2386)
    return true;
}

// ---- warp-cooperative evaluation (periods <= 32: every large contig, bwt.py:2192-2200) ----
// Lane q owns column q of the tally, in registers; one copy is added per step by all lanes
// at once and the mismatch total is a warp reduction, so an extension step costs ~100 cycles
// instead of the ~4000 of the one-thread form above.  Results are identical to Tally's.
struct WTally {
    uint8_t sym[SLOTS];
    uint16_t cnt[SLOTS];
    int nsym;
    bool overflow;
    __device__ __forceinline__ void reset() { nsym = 0; overflow = false; }
    __device__ __forceinline__ void add(uint8_t b, int delta)
    {
        bool done = false;
#pragma unroll
        for (int j = 0; j < SLOTS; j++)
            if (!done && j < nsym && sym[j] == b) { cnt[j] = (uint16_t)(cnt[j] + delta); done = true; }
        if (done) return;
        if (nsym >= SLOTS || delta < 0) { overflow = true; return; }
#pragma unroll
        for (int j = 0; j < SLOTS; j++)
            if (j == nsym) { sym[j] = b; cnt[j] = (uint16_t)delta; }
        nsym++;
    }
    __device__ __forceinline__ uint8_t best(int *best_cnt) const
    {
        int bc = -1;
        uint8_t bs = 0;
#pragma unroll
        for (int j = 0; j < SLOTS; j++) {
            if (j < nsym) {
                int v = cnt[j];
                if (v > bc || (v == bc && sym[j] < bs)) { bc = v; bs = sym[j]; }
            }
        }
        *best_cnt = bc < 0 ? 0 : bc;
        return bs;
    }
};

__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void w_add_copy(WTally &t, const uint8_t *s, int64_t at, int p, int delta)
{
    const int lane = threadIdx.x & 31;
    if (lane < p) t.add(__ldg(s + at + lane), delta);
}

__device__ __forceinline__ int64_t w_total_mm(const WTally &t, int p, int64_t copies)
{
    const int lane = threadIdx.x & 31;
    int bc = 0;
    if (lane < p) t.best(&bc);
    return copies * p - (int64_t)warp_sum(bc);
}

// _extend_with_mismatches by a whole warp; `cons` = this lane's consensus byte (lane < period)
__device__ ExtOut w_extend_with_mismatches(const uint8_t *s, int64_t start_pos, int period, int64_t n, bool allow_mm,
                                           WTally &t, uint8_t &cons)
{
    const int lane = threadIdx.x & 31;
    t.reset();
    int64_t start = start_pos, end = start_pos + period, copies = 1;
    w_add_copy(t, s, start_pos, period, 1);
    while (end + period <= n) {
        w_add_copy(t, s, end, period, 1);
        int64_t bud = allow_mm ? mm_budget(period, copies + 1) : 0;
        if (w_total_mm(t, period, copies + 1) <= bud) { copies++; end += period; }
        else { w_add_copy(t, s, end, period, -1); break; }
    }
    while (start - period >= 0) {
        w_add_copy(t, s, start - period, period, 1);
        int64_t bud = allow_mm ? mm_budget(period, copies + 1) : 0;
        if (w_total_mm(t, period, copies + 1) <= bud) { copies++; start -= period; }
        else { w_add_copy(t, s, start - period, period, -1); break; }
    }
    int bc;
    cons = lane < period ? t.best(&bc) : (uint8_t)0;
    ExtOut o;
    o.copies = copies; o.full_start = start; o.full_end = end;
    // partial copies: leading run of matches against the consensus, right then left
    bool mr = lane < period && end + lane < n && __ldg(s + end + lane) == cons;
    unsigned br = __ballot_sync(0xffffffffu, mr);
    int pr = (~br) ? (__ffs(~br) - 1) : 32;
    int src = period - 1 - lane;
    uint8_t crev = __shfl_sync(0xffffffffu, cons, src < 0 ? 0 : src);
    bool ml = lane < period && start - lane - 1 >= 0 && __ldg(s + start - lane - 1) == crev;
    unsigned bl = __ballot_sync(0xffffffffu, ml);
    int pl = (~bl) ? (__ffs(~bl) - 1) : 32;
    o.array_start = start - pl;
    o.array_end = end + pr;
    return o;
}

// smallest period of the `len` (<= 32) bytes held one per lane
__device__ int w_smallest_period_lanes(uint8_t c, int len)
{
    const int lane = threadIdx.x & 31;
    for (int p = 1; p <= len / 2; p++) {
        if (len % p) continue;
        uint8_t prev = __shfl_sync(0xffffffffu, c, lane >= p ? lane - p : 0);
        unsigned bad = __ballot_sync(0xffffffffu, lane >= p && lane < len && c != prev);
        if (!bad) return p;
    }
    return len;
}

// build_consensus_motif_array by a whole warp
__device__ int64_t w_consensus(const uint8_t *text, int64_t text_size, int64_t start, int period, int64_t copies,
                               WTally &t, uint8_t &cons, int64_t *total_mm, int64_t *max_mm)
{
    const int lane = threadIdx.x & 31;
    *total_mm = 0; *max_mm = 0;
    if (copies <= 0 || period <= 0) return 0;
    int64_t used = 0;
    while (used < copies && start + (used + 1) * period <= text_size) used++;
    if (!used) return 0;
    t.reset();
    for (int64_t c = 0; c < used; c++) w_add_copy(t, text, start + c * period, period, 1);
    int bc;
    cons = lane < period ? t.best(&bc) : (uint8_t)0;
    for (int64_t c = 0; c < used; c++) {
        bool ne = lane < period && __ldg(text + start + c * period + lane) != cons;
        int64_t mm = __popc(__ballot_sync(0xffffffffu, ne));
        *total_mm += mm;
        if (mm > *max_mm) *max_mm = mm;
    }
    return used;
}

// visit_full by a whole warp (p <= 32); all outputs are warp-uniform
__device__ bool w_visit_full(const ScanCfg &c, int p, int64_t i, int32_t *row, int64_t *next_i, int *err)
{
    const int lane = threadIdx.x & 31;
    const uint8_t *s = c.s;
    const int64_t n = c.n;
    WTally t;
    uint8_t cons = 0;
    bool emitted = false;
    *next_i = i + c.pos_step;
    do {
        ExtOut o = w_extend_with_mismatches(s, i, p, n, c.allow_mm && p <= 64, t, cons);
        int64_t a_len = o.array_end - o.array_start;
        if (a_len < c.min_array_len) break;
        int64_t part = a_len - o.copies * p;
        if (part < 0) part = 0;
        int64_t eff = o.copies + (((double)part / (double)p) >= 0.75 ? 1 : 0);
        if (!(o.copies >= c.min_copies || eff >= c.min_copies)) break;
        uint8_t first = lane < p ? __ldg(s + o.full_start + lane) : (uint8_t)0;
        int prim = w_smallest_period_lanes(first, p);
        int p_eff = prim < p ? prim : p;
        o = w_extend_with_mismatches(s, o.full_start, p_eff, n, c.allow_mm && p_eff <= 64, t, cons);
        a_len = o.array_end - o.array_start;
        part = a_len - o.copies * p_eff;
        if (part < 0) part = 0;
        eff = o.copies + (((double)part / (double)p_eff) >= 0.75 ? 1 : 0);
        if (o.copies < c.min_copies && eff < c.min_copies) break;
        int64_t tmm, mmm;
        int64_t a_start = o.array_start, a_end = o.array_end, copies_full = o.copies, cons_start = o.full_start;
        int64_t used = w_consensus(s, c.n_total, cons_start, p_eff, copies_full, t, cons, &tmm, &mmm);
        if (!used) break;
        int prim2 = w_smallest_period_lanes(cons, p_eff);
        if (prim2 < p_eff) {
            p_eff = prim2;
            copies_full = (a_end - a_start) / p_eff;
            if (copies_full < 1) copies_full = 1;
            a_end = a_start + copies_full * p_eff;
            cons_start = a_start;
            used = w_consensus(s, c.n_total, a_start, p_eff, copies_full, t, cons, &tmm, &mmm);
            if (!used) break;
        }
        row[0] = (int32_t)a_start; row[1] = (int32_t)a_end; row[2] = p_eff;
        row[3] = (int32_t)copies_full; row[4] = (int32_t)tmm; row[5] = (int32_t)mmm;
        row[6] = (int32_t)cons_start; row[7] = (int32_t)used;
        emitted = true;
        *next_i = a_end;   // jump past the array (bwt.py:2386)
    } while (false);
    if (__any_sync(0xffffffffu, t.overflow)) *err = 1;
    return emitted;
}

// One CTA per period pass.  The reference's scan visits i, i+step, ... and only
// jumps after an emission, and every rejecting test is free of side effects, so
// the CTA tests PASS_THREADS consecutive visits at once; the first visit that is
// not rejected is evaluated in full by warp 0 (extension, consensus, record; one
// thread for periods above 32) and sets the next position.  budget[pass] = visits
// this pass may spend (MAX_ITER in the counting run).  Rows are appended to `tmp`
// with aux = (pass, sequence number within the pass, visit index).
__global__ void __launch_bounds__(PASS_THREADS)
    period_pass_kernel(ScanCfg c, int64_t npass, const int64_t *__restrict__ budget,
                       int64_t *__restrict__ visits, int64_t *__restrict__ emits,
                       int32_t *__restrict__ tmp, int32_t *__restrict__ tmp_aux, int64_t tmp_cap,
                       unsigned long long *tmp_count, uint8_t *scratch, int maxp, int *err,
                       volatile long long *done)
{
    const int64_t pass = blockIdx.x;
    if (pass >= npass) return;
    __shared__ int64_t s_i, s_it, s_seq;
    // speculative run (budget == nullptr): done[q] = visits of pass q + 1 once it has finished.
    // A pass adds up its predecessors as they finish and, once all have, knows what the global
    // budget leaves it and stops there instead of walking the whole contig.
    __shared__ int64_t s_allowed, s_prev_sum;
    __shared__ int s_known, s_learned;
    __shared__ int s_first[PASS_THREADS / 32];
    __shared__ int s_stop;
    constexpr int MEMO = 4;
    __shared__ int64_t s_memo_x[MEMO], s_memo_ni[MEMO], s_memo_it[MEMO];
    __shared__ int s_memo_emit[MEMO], s_memo_next, s_last_full;
    __shared__ int32_t s_memo_row[MEMO][BWTK_REC_W];
    __shared__ int64_t s_ff, s_ff_seq0, s_ff_it0, s_ff_cyc;
    __shared__ int s_ff_hit;
    __shared__ int64_t s_ev_i, s_ev_it;
    __shared__ int s_ev_in;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = (int)(c.min_p + pass * c.per_step);
    uint8_t *my = scratch + pass * scratch_per_thread(maxp);
    Tally t = make_tally(my, maxp);
    uint8_t *cons = t.nsym + maxp;
    const int64_t n = c.n;
    if (tid == 0) {
        s_allowed = budget ? budget[pass] : MAX_ITER;
        s_prev_sum = 0; s_known = 0; s_learned = 0;
        s_i = 0; s_it = 0; s_seq = 0; s_memo_next = 0; s_last_full = -1; s_ff = 0;
        for (int q = 0; q < MEMO; q++) { s_memo_x[q] = -1; s_memo_it[q] = -1; }
    }
    __syncthreads();
    while (p > 0) {
        if (done && tid == 0 && !s_learned) {
            while (s_known < (int)pass) {
                const long long d = done[s_known];
                if (d == 0) break;
                s_prev_sum += d - 1;
                s_known++;
            }
            if (s_known == (int)pass) {
                int64_t left = MAX_ITER - s_prev_sum;
                if (left < 0) left = 0;
                if (left < s_allowed) s_allowed = left;
                s_learned = 1;
            }
        }
        __syncthreads();
        const int64_t allowed = s_allowed;
        const int64_t i0 = s_i, it0 = s_it;
        if (i0 + 2 * (int64_t)p > n || it0 >= allowed) break;
        // visit number it0 + tid at position i0 + tid*step
        const int64_t x = i0 + (int64_t)tid * c.pos_step;
        const bool in_range = (x + 2 * (int64_t)p <= n) && (it0 + tid < allowed);
        const bool stop = !in_range || !visit_rejected(c, p, x);
        unsigned bal = __ballot_sync(0xffffffffu, stop);
        if (lane == 0) s_first[warp] = bal ? (warp * 32 + __ffs(bal) - 1) : PASS_THREADS;
        __syncthreads();
        if (tid == 0) {
            int f = PASS_THREADS;
            for (int w = 0; w < PASS_THREADS / 32; w++) f = s_first[w] < f ? s_first[w] : f;
            s_stop = f;
        }
        __syncthreads();
        const int f = s_stop;
        if (f == PASS_THREADS) {   // all of them step on
            if (tid == 0) { s_i = i0 + (int64_t)PASS_THREADS * c.pos_step; s_it = it0 + PASS_THREADS; }
            __syncthreads();
            continue;
        }
        // visits 0..f-1 were rejected; visit f either ends the scan or is evaluated in full
        if (tid == f) { s_ev_i = x; s_ev_it = it0 + f; s_ev_in = in_range ? 1 : 0; }
        __syncthreads();
        if (warp == 0) {
            int64_t it = s_ev_it;
            const int64_t i = s_ev_i;
            if (lane == 0) s_ff = 0;
            if (!s_ev_in) {
                if (lane == 0) { s_i = i; s_it = it; }   // loop condition fails on re-entry (end of text or budget)
            } else {
                it++;
                // The outcome of a full evaluation is a pure function of the position, and the
                // reference can loop (array_end <= i after the primitive-period re-extension,
                // bwt.py:2305-2386) until the iteration cap: remember the last few outcomes.
                int hit = -1;
                for (int q = 0; q < MEMO; q++)
                    if (s_memo_x[q] == i) hit = q;
                if (hit < 0) {
                    int64_t next_i = i + c.pos_step;
                    int32_t row[BWTK_REC_W];
#pragma unroll
                    for (int q = 0; q < BWTK_REC_W; q++) row[q] = 0;
                    bool emitted = false;
                    if (p <= 32) {
                        emitted = w_visit_full(c, p, i, row, &next_i, err);
                    } else if (lane == 0) {
                        emitted = visit_full(c, p, i, t, cons, row, &next_i);
                        if (t.overflow) *err = 1;
                    }
                    hit = s_memo_next % MEMO;
                    __syncwarp();
                    if (lane == 0) {
                        s_memo_next++;
                        s_memo_x[hit] = i;
                        s_memo_ni[hit] = next_i;
                        s_memo_emit[hit] = emitted ? 1 : 0;
                        s_memo_it[hit] = -1;
                        for (int q = 0; q < BWTK_REC_W; q++) s_memo_row[hit][q] = emitted ? row[q] : 0;
                    }
                }
                if (lane == 0) {
                    int64_t seq = s_seq;
                    const bool emitted = s_memo_emit[hit] != 0;
                    if (emitted) {
                        if (tmp) {
                            unsigned long long slot = atomicAdd(tmp_count, 1ull);
                            if ((int64_t)slot < tmp_cap) {
                                int32_t *row = tmp + slot * BWTK_REC_W;
                                for (int q = 0; q < BWTK_REC_W; q++) row[q] = s_memo_row[hit][q];
                                int32_t *ax = tmp_aux + slot * 4;
                                ax[0] = (int32_t)pass; ax[1] = (int32_t)seq; ax[2] = (int32_t)it; ax[3] = 0;
                            }
                        }
                        seq++;
                    }
                    // Simple cycle: the same position was the previous full evaluation too, so the
                    // visits in between repeat verbatim; skip whole cycles up to the budget.
                    if (s_last_full == hit && s_memo_it[hit] >= 0) {
                        int64_t cyc = it - s_memo_it[hit];
                        int64_t reps = cyc > 0 ? (allowed - it) / cyc : 0;
                        if (reps > 0) {
                            s_ff = emitted ? reps : 0;
                            s_ff_hit = hit;
                            s_ff_seq0 = seq;
                            s_ff_it0 = it;
                            s_ff_cyc = cyc;
                            it += reps * cyc;
                            if (emitted) seq += reps;
                        }
                    }
                    s_memo_it[hit] = it;
                    s_last_full = hit;
                    s_seq = seq;
                    s_i = s_memo_ni[hit];
                    s_it = it;
                }
            }
        }
        __syncthreads();
        if (s_ff > 0 && tmp) {
            // duplicate rows of the skipped cycles (the host dedups them like the reference's `seen` set)
            const int64_t reps = s_ff;
            __shared__ unsigned long long s_ff_base;
            if (tid == 0) s_ff_base = atomicAdd(tmp_count, (unsigned long long)reps);
            __syncthreads();
            for (int64_t r = tid; r < reps; r += PASS_THREADS) {
                int64_t slot = (int64_t)s_ff_base + r;
                if (slot < tmp_cap) {
                    int32_t *row = tmp + slot * BWTK_REC_W;
                    for (int q = 0; q < BWTK_REC_W; q++) row[q] = s_memo_row[s_ff_hit][q];
                    int32_t *ax = tmp_aux + slot * 4;
                    ax[0] = (int32_t)pass; ax[1] = (int32_t)(s_ff_seq0 + r);
                    ax[2] = (int32_t)(s_ff_it0 + (r + 1) * s_ff_cyc); ax[3] = 0;
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        // a pass cut short by the learned budget would have gone on: report more than it was left,
        // which is all budget_kernel needs to know (it clamps, and notes that the cap was hit)
        int64_t v = s_it;
        if (done && s_learned && p > 0 && s_it >= s_allowed && s_i + 2 * (int64_t)p <= n) v = s_allowed + 1;
        if (visits) visits[pass] = v;
        if (emits) emits[pass] = s_seq;
        if (done) {
            __threadfence();
            done[pass] = v + 1;
        }
    }
}

// Applies the global iteration budget: pass q may spend max(0, MAX_ITER - visits
// of earlier passes) iterations (the reference returns at the first iteration
// whose running count exceeds MAX_ITER).
__global__ void budget_kernel(const int64_t *__restrict__ visits, int64_t npass, int64_t *__restrict__ budget,
                              int64_t *__restrict__ total_iter)
{
    if (blockIdx.x || threadIdx.x) return;
    int64_t used = 0;
    bool hit = false;
    for (int64_t q = 0; q < npass; q++) {
        int64_t left = MAX_ITER - used;
        if (left < 0) left = 0;
        int64_t v = visits[q];
        if (v > left) { v = left; hit = true; }
        budget[q] = hit && v == left ? left : v;
        used += v;
        if (hit) { for (int64_t r = q + 1; r < npass; r++) budget[r] = 0; break; }
    }
    // the reference increments once more before it notices the overflow
    *total_iter = hit ? MAX_ITER + 1 : used;
}

__global__ void offsets_kernel(const int64_t *__restrict__ emits, int64_t npass, int64_t *__restrict__ offs)
{
    if (blockIdx.x || threadIdx.x) return;
    int64_t acc = 0;
    for (int64_t q = 0; q < npass; q++) { offs[q] = acc; acc += emits[q]; }
    offs[npass] = acc;
}

// rows of every pass that fall inside its budget (aux = pass, sequence number, visit count)
__global__ void count_within_budget_kernel(const int32_t *__restrict__ aux, int64_t m,
                                           const int64_t *__restrict__ budget, unsigned long long *emits)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int pass = aux[i * 4];
    if ((int64_t)aux[i * 4 + 2] <= budget[pass]) atomicAdd(&emits[pass], 1ull);
}

__global__ void place_rows_kernel(const int32_t *__restrict__ tmp, const int32_t *__restrict__ aux, int64_t m,
                                  const int64_t *__restrict__ offs, const int64_t *__restrict__ budget,
                                  int32_t *__restrict__ out, int64_t cap)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (budget && (int64_t)aux[i * 4 + 2] > budget[aux[i * 4]]) return;   // emitted past the pass's budget
    int64_t at = offs[aux[i * 4]] + aux[i * 4 + 1];
    if (at >= cap) return;
    for (int q = 0; q < BWTK_REC_W; q++) out[at * BWTK_REC_W + q] = tmp[i * BWTK_REC_W + q];
}

}  // namespace ext
}  // namespace bwtk

using namespace bwtk;

// The scratch of these entry points comes from the device's default stream-ordered pool.  With
// the default release threshold (0) the pool hands everything back to the driver at every
// synchronisation, so each call paid a fresh allocation of tens of MB (measured: 15 ms calls
// with 100-500 ms outliers).  Keep freed blocks cached instead.
static cudaError_t pool_alloc(void **p, size_t bytes, cudaStream_t st)
{
    static thread_local int tuned_device = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (tuned_device != dev) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        tuned_device = dev;
    }
    return cudaMallocAsync(p, bytes, st);
}

static int max_of_periods(const int32_t *d_period, int64_t m, cudaStream_t st, int *out)
{
    // small batches: read the periods back to size the scratch
    int32_t *h = (int32_t *)malloc((size_t)m * 4);
    if (!h) { set_error("out of host memory"); return BWTK_EINTERNAL; }
    cudaError_t e = cudaMemcpyAsync(h, d_period, (size_t)m * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { free(h); set_error("period read-back: %s", cudaGetErrorString(e)); return BWTK_ECUDA; }
    int mx = 1;
    for (int64_t i = 0; i < m; i++) if (h[i] > mx) mx = h[i];
    free(h);
    *out = mx;
    return BWTK_OK;
}

extern "C" int32_t bwtk_extend_batch(const uint8_t *d_text, int64_t n, const int32_t *d_seed,
                                     const int32_t *d_period, const int32_t *d_flags, int64_t m, int32_t mode,
                                     int32_t *d_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (m == 0) return BWTK_OK;
    BWTK_REQUIRE(d_text && d_seed && d_period && d_flags && d_out, "null pointer");
    BWTK_REQUIRE(mode == 0 || mode == 1, "mode must be 0 or 1");
    int maxp = 1;
    int rc = max_of_periods(d_period, m, st, &maxp);
    if (rc) return rc;
    BWTK_REQUIRE(maxp <= 65535, "period too long");
    uint8_t *scratch = nullptr;
    int *d_err = nullptr;
    int64_t bytes = m * ext::scratch_per_thread(maxp) + 16;
    BWTK_CUDA(pool_alloc((void **)&scratch, (size_t)bytes, st));
    d_err = (int *)(scratch + bytes - 16);
    BWTK_CUDA(bwtk::zero_async(d_err, 4, st));
    ext::extend_batch_kernel<<<(unsigned)ceil_div(m, 64), 64, 0, st>>>(d_text, n, d_seed, d_period, d_flags, m,
                                                                      mode, maxp, scratch, d_out, d_err);
    BWTK_LAUNCH_CHECK();
    int h_err = 0;
    BWTK_CUDA(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaStreamSynchronize(st));
    cudaFreeAsync(scratch, st);
    if (h_err) { set_error("extend_batch: more than %d distinct symbols in one column", ext::SLOTS); return BWTK_EINVAL; }
    return BWTK_OK;
}

extern "C" int32_t bwtk_consensus_batch(const uint8_t *d_text, int64_t text_size, const int32_t *d_start,
                                        const int32_t *d_period, const int32_t *d_copies,
                                        const int64_t *d_cons_off, int64_t m, uint8_t *d_cons, int32_t *d_mm,
                                        void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (m == 0) return BWTK_OK;
    BWTK_REQUIRE(d_text && d_start && d_period && d_copies && d_cons_off && d_cons && d_mm, "null pointer");
    int maxp = 1;
    int rc = max_of_periods(d_period, m, st, &maxp);
    if (rc) return rc;
    BWTK_REQUIRE(maxp <= 65535, "period too long");
    uint8_t *scratch = nullptr;
    int64_t bytes = m * ext::scratch_per_thread(maxp) + 16;
    BWTK_CUDA(pool_alloc((void **)&scratch, (size_t)bytes, st));
    int *d_err = (int *)(scratch + bytes - 16);
    BWTK_CUDA(bwtk::zero_async(d_err, 4, st));
    ext::consensus_batch_kernel<<<(unsigned)ceil_div(m, 64), 64, 0, st>>>(
        d_text, text_size, d_start, d_period, d_copies, d_cons_off, m, maxp, scratch, d_cons, d_mm, d_err);
    BWTK_LAUNCH_CHECK();
    int h_err = 0;
    BWTK_CUDA(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaStreamSynchronize(st));
    cudaFreeAsync(scratch, st);
    if (h_err) { set_error("consensus_batch: more than %d distinct symbols in one column", ext::SLOTS); return BWTK_EINVAL; }
    return BWTK_OK;
}

extern "C" int32_t bwtk_period_scan(const uint8_t *d_text, int64_t n_total, int64_t min_period,
                                    int64_t max_period, int32_t allow_mismatches, int64_t min_copies,
                                    int64_t min_array_len, double min_entropy, const uint8_t *d_tier1_mask,
                                    const double *d_plogp, int64_t plogp_dim, int32_t *d_rec, int64_t cap,
                                    int64_t *h_count, int64_t *h_iterations, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    BWTK_REQUIRE(h_count && h_iterations, "null output");
    *h_count = 0;
    *h_iterations = 0;
    if (n_total <= 0) return BWTK_OK;
    BWTK_REQUIRE(d_text && d_plogp && (d_rec || cap == 0), "null pointer");
    int64_t n = n_total;
    {
        uint8_t last = 0;
        BWTK_CUDA(cudaMemcpyAsync(&last, d_text + n - 1, 1, cudaMemcpyDeviceToHost, st));
        BWTK_CUDA(cudaStreamSynchronize(st));
        if (last == 36) n--;
    }
    // clamps and steps exactly as bwt.py:2192-2230
    int64_t half = n / 2 > 1 ? n / 2 : 1;
    int64_t max_p = max_period < half ? max_period : half;
    int64_t lim = n > 100000 ? 30 : n > 10000 ? 50 : n > 1000 ? 100 : 200;
    if (max_p > lim) max_p = lim;
    int64_t min_p = min_period < max_p ? min_period : max_p;
    int64_t pos_step, per_step;
    if (n > 10000000) { pos_step = 500; per_step = 20; }
    else if (n > 5000000) { pos_step = 200; per_step = 10; }
    else if (n > 1000000) { pos_step = 100; per_step = 5; }
    else if (n > 100000) { pos_step = 50; per_step = 2; }
    else if (n > 10000) { pos_step = 20; per_step = 1; }
    else { pos_step = 10; per_step = 1; }
    if (max_p < min_p || max_p < 1) return BWTK_OK;
    if (min_p < 1) {
        // range(min_p, ...) with non-positive periods: those passes do nothing useful
        // (i + 2p <= n holds forever for p <= 0 in the reference); refuse instead of hanging.
        set_error("min_period must be >= 1");
        return BWTK_EINVAL;
    }
    BWTK_REQUIRE(plogp_dim > max_p, "entropy table smaller than the longest period");
    int64_t npass = (max_p - min_p) / per_step + 1;
    int maxp = (int)max_p;
    ext::ScanCfg cfg;
    cfg.s = d_text; cfg.n = n; cfg.n_total = n_total; cfg.min_p = min_p; cfg.max_p = max_p;
    cfg.per_step = per_step; cfg.pos_step = pos_step; cfg.allow_mm = allow_mismatches;
    cfg.min_copies = min_copies; cfg.min_array_len = min_array_len; cfg.min_entropy = min_entropy;
    cfg.mask = d_tier1_mask; cfg.plogp = d_plogp; cfg.dim = plogp_dim;

    int64_t tmp_cap = cap > 65536 ? cap : 65536;
    int64_t sc_bytes = npass * ext::scratch_per_thread(maxp);
    int64_t bytes = align_up(sc_bytes, 256) + align_up(tmp_cap * BWTK_REC_W * 4, 256) +
                    align_up(tmp_cap * 16, 256) + 5 * align_up((npass + 2) * 8, 256) + 1024;
    uint8_t *buf = nullptr;
    BWTK_CUDA(pool_alloc((void **)&buf, (size_t)bytes, st));
    Carver c(buf, bytes);
    uint8_t *scratch = c.take<uint8_t>(sc_bytes);
    int32_t *tmp = c.take<int32_t>(tmp_cap * BWTK_REC_W);
    int32_t *aux = c.take<int32_t>(tmp_cap * 4);
    int64_t *visits = c.take<int64_t>(npass + 2);
    int64_t *emits = c.take<int64_t>(npass + 2);
    int64_t *budget = c.take<int64_t>(npass + 2);
    int64_t *offs = c.take<int64_t>(npass + 2);
    unsigned long long *tmp_count = c.take<unsigned long long>(2);
    long long *d_done = c.take<long long>(npass + 2);
    int64_t *d_iter = c.take<int64_t>(2);
    int *d_err = c.take<int>(4);
    BWTK_CUDA(bwtk::zero_async(tmp_count, 16, st));
    BWTK_CUDA(bwtk::zero_async(d_err, 4, st));
    unsigned grid = (unsigned)npass;
    // One speculative run: every pass walks with the full budget and appends its rows with the
    // visit count at which each was emitted.  The reference's global 100 000-visit budget is then
    // applied afterwards: pass q may spend what the earlier passes left, and its rows past that
    // point are dropped (they form a suffix of the pass).  Only if the speculative rows overflow
    // the buffer (cycles fast-forwarded under a budget they will not get) is the walk repeated
    // with the real budgets.
    BWTK_CUDA(bwtk::zero_async(d_done, (size_t)(npass + 2) * 8, st));
    ext::period_pass_kernel<<<grid, ext::PASS_THREADS, 0, st>>>(cfg, npass, nullptr, visits, emits, tmp, aux, tmp_cap,
                                                 tmp_count, scratch, maxp, d_err, d_done);
    BWTK_LAUNCH_CHECK();
    ext::budget_kernel<<<1, 32, 0, st>>>(visits, npass, budget, d_iter);
    BWTK_LAUNCH_CHECK();
    unsigned long long h_m = 0;
    int rc = read_back(&h_m, tmp_count, 8, st);
    if (rc) { cudaFreeAsync(buf, st); return rc; }
    const int64_t *row_budget = budget;
    if ((int64_t)h_m > tmp_cap) {
        BWTK_CUDA(bwtk::zero_async(tmp_count, 16, st));
        ext::period_pass_kernel<<<grid, ext::PASS_THREADS, 0, st>>>(cfg, npass, budget, nullptr, emits, tmp, aux, tmp_cap,
                                                     tmp_count, scratch, maxp, d_err, nullptr);
        BWTK_LAUNCH_CHECK();
        rc = read_back(&h_m, tmp_count, 8, st);
        if (rc) { cudaFreeAsync(buf, st); return rc; }
        row_budget = nullptr;   // every row of this run is inside its budget
    } else if (h_m > 0) {
        BWTK_CUDA(bwtk::zero_async(emits, (size_t)(npass + 2) * 8, st));
        ext::count_within_budget_kernel<<<(unsigned)ceil_div((int64_t)h_m, 256), 256, 0, st>>>(
            aux, (int64_t)h_m, budget, reinterpret_cast<unsigned long long *>(emits));
        BWTK_LAUNCH_CHECK();
    } else {
        BWTK_CUDA(bwtk::zero_async(emits, (size_t)(npass + 2) * 8, st));
    }
    ext::offsets_kernel<<<1, 32, 0, st>>>(emits, npass, offs);
    BWTK_LAUNCH_CHECK();
    int h_err = 0;
    int64_t h_it = 0, h_total = 0;
    rc = read_back(&h_err, d_err, 4, st);
    if (!rc) rc = read_back(&h_it, d_iter, 8, st);
    if (!rc) rc = read_back(&h_total, offs + npass, 8, st);
    if (rc) { cudaFreeAsync(buf, st); return rc; }
    if (h_err) {
        set_error("period scan: more than %d distinct symbols in one column", ext::SLOTS);
        rc = BWTK_EINVAL;
    } else if ((int64_t)h_m > tmp_cap) {
        *h_count = (int64_t)h_m;
        set_error("period scan: %llu records exceed capacity %lld", h_m, (long long)tmp_cap);
        rc = BWTK_EOVERFLOW;
    } else {
        *h_count = h_total;
        *h_iterations = h_it;
        if (h_total > cap) {
            set_error("period scan: %lld records exceed capacity %lld", (long long)h_total, (long long)cap);
            rc = BWTK_EOVERFLOW;
        } else if (h_total > 0) {
            ext::place_rows_kernel<<<(unsigned)ceil_div((int64_t)h_m, 256), 256, 0, st>>>(tmp, aux, (int64_t)h_m,
                                                                                         offs, row_budget, d_rec, cap);
            bwtk::count_launch();
            if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
                set_error("place_rows failed");
                rc = BWTK_ECUDA;
            }
        }
    }
    cudaFreeAsync(buf, st);
    return rc;
}
