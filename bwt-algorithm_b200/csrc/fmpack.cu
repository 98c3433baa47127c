// fmpack.cu -- the search side of the FM index as 64-byte rank blocks, and rank / backward
// search / the 1..kmax motif sweep over it (reference bwt.py:335-389: rank = checkpoint +
// remainder scan, backward_search = two ranks per pattern character).
//
// Block b (64 B, one 128-B line never straddled, two 32-B sectors) covers BWT positions
// [192 b, 192 b + 192):
//     bytes  0..15   4 x u32: #A, #C, #G, #T in bwt[0 : 192 b]   (30 bits each; bit 31 of the
//                    first word = "this block holds exception symbols")
//     bytes 16..63   192 symbols, 2 bits each, symbol j of a word at bits [2j, 2j+1]
// so a rank is ONE random access: the counts and the first 64 symbols share sector 0, the other
// 128 symbols are sector 1 (read only when the position is past symbol 64).  Symbols other than
// A/C/G/T ('$', N, IUPAC, lower case) are stored as code 0 and listed twice: by position (to
// correct #A inside a flagged block) and grouped by byte value (their own rank is a binary
// search).  A chr1-sized BWT becomes 83 MB and stays resident in the 126 MB L2, where the
// byte-per-symbol BWT + Occ rows (280 MB) do not.
#include "common.cuh"
#include "scan.cuh"

namespace bwtk {
namespace fmp {

constexpr int BLK = 192;

struct Packed {
    const uint4 *blocks;
    int32_t n;
    const uint32_t *exc_pos;      // all exception positions, ascending
    const uint32_t *exc_by_code;  // the same positions grouped by byte value, ascending inside a group
    int32_t n_exc;
    const int64_t *code_off;      // [257] group offsets into exc_by_code
    const int64_t *C;             // [256]
    const int64_t *tot;           // [256]
    int32_t C4[4], tot4[4];       // C / totals of A, C, G, T (kernel parameter space)
    const int32_t *ftab_sp, *ftab_ep;  // intervals of every ACGT k-mer (optional)
    int32_t ftab_k;
};

__device__ __forceinline__ int code2_of(int c)
{
    return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1;
}

// bit 2j set where symbol j of w equals the symbol replicated in pat
__device__ __forceinline__ uint32_t eq_syms(uint32_t w, uint32_t pat)
{
    uint32_t x = w ^ pat;
    return ~(x | (x >> 1)) & 0x55555555u;
}

// keeps the first cnt (0..16 after clamping) symbols of a word's match mask
__device__ __forceinline__ uint32_t first_syms(int cnt)
{
    return cnt >= 16 ? 0x55555555u : cnt <= 0 ? 0u : (0x55555555u & ((1u << (2 * cnt)) - 1u));
}

// matches of pat among the first cnt (<= 64) symbols of v
__device__ __forceinline__ int count_vec(uint4 v, uint32_t pat, int cnt)
{
    return __popc(eq_syms(v.x, pat) & first_syms(cnt)) + __popc(eq_syms(v.y, pat) & first_syms(cnt - 16)) +
           __popc(eq_syms(v.z, pat) & first_syms(cnt - 32)) + __popc(eq_syms(v.w, pat) & first_syms(cnt - 48));
}

// per-symbol counts (A, C, G, T) among the first cnt symbols of v, added to r[4]
__device__ __forceinline__ void count_vec_all(uint4 v, int cnt, int r[4])
{
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t m = first_syms(cnt - 16 * k);
        const uint32_t lo = w[k] & 0x55555555u, hi = (w[k] >> 1) & 0x55555555u;
        const int t = __popc(lo & hi & m), g = __popc(hi & ~lo & m), c = __popc(lo & ~hi & m);
        r[3] += t; r[2] += g; r[1] += c;
        r[0] += __popc(m) - t - g - c;
    }
}

__device__ __forceinline__ int lower_bound_u32(const uint32_t *a, int lo, int hi, uint32_t key)
{
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// exceptions of any byte value in [from, to)
__device__ __noinline__ int exceptions_in(const Packed &ix, uint32_t from, uint32_t to)
{
    int a = lower_bound_u32(ix.exc_pos, 0, ix.n_exc, from);
    return lower_bound_u32(ix.exc_pos, a, ix.n_exc, to) - a;
}

// #symbol c2 (0..3) in bwt[0:pos], 0 <= pos <= n
__device__ __forceinline__ int rank4(const Packed &ix, int c2, int pos)
{
    const uint32_t blk = (uint32_t)pos / BLK;
    const int off = pos - (int)blk * BLK;
    const uint4 *b = ix.blocks + (size_t)blk * 4;
    const uint4 h = __ldg(b);
    const uint4 s0 = __ldg(b + 1);
    const uint32_t pat = 0x55555555u * (uint32_t)c2;
    const uint32_t base = c2 == 0 ? (h.x & 0x3fffffffu) : c2 == 1 ? h.y : c2 == 2 ? h.z : h.w;
    int r = count_vec(s0, pat, off);
    if (off > 64) {
        const uint4 s1 = __ldg(b + 2), s2 = __ldg(b + 3);
        r += count_vec(s1, pat, off - 64) + count_vec(s2, pat, off - 128);
    }
    if (c2 == 0 && (h.x >> 31) && off > 0) r -= exceptions_in(ix, blk * BLK, (uint32_t)pos);
    return (int)base + r;
}

// ranks of all four symbols at pos
__device__ __forceinline__ void rank4_all(const Packed &ix, int pos, int r[4])
{
    const uint32_t blk = (uint32_t)pos / BLK;
    const int off = pos - (int)blk * BLK;
    const uint4 *b = ix.blocks + (size_t)blk * 4;
    const uint4 h = __ldg(b);
    const uint4 s0 = __ldg(b + 1);
    r[0] = (int)(h.x & 0x3fffffffu); r[1] = (int)h.y; r[2] = (int)h.z; r[3] = (int)h.w;
    count_vec_all(s0, off, r);
    if (off > 64) {
        const uint4 s1 = __ldg(b + 2), s2 = __ldg(b + 3);
        count_vec_all(s1, off - 64, r);
        count_vec_all(s2, off - 128, r);
    }
    if ((h.x >> 31) && off > 0) r[0] -= exceptions_in(ix, blk * BLK, (uint32_t)pos);
}

// rank of any byte value (bwt.py:335-357: pos <= 0 -> 0, pos > n -> n, unknown byte -> 0)
__device__ __forceinline__ int rank_any(const Packed &ix, int c, int64_t pos64)
{
    if (pos64 <= 0) return 0;
    const int pos = pos64 > ix.n ? ix.n : (int)pos64;
    const int c2 = code2_of(c);
    if (c2 >= 0) return rank4(ix, c2, pos);
    const int a = (int)__ldg(ix.code_off + c), e = (int)__ldg(ix.code_off + c + 1);
    return lower_bound_u32(ix.exc_by_code, a, e, (uint32_t)pos) - a;
}

// ---- build -------------------------------------------------------------------------
// one thread per block: packs 192 BWT bytes, leaves the block's own symbol counts in the header
__global__ void __launch_bounds__(128)
    pack_blocks_kernel(const uint8_t *__restrict__ bwt, int32_t n, int32_t nblk, uint4 *__restrict__ blocks,
                       uint32_t *__restrict__ nexc)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    const int64_t base = (int64_t)b * BLK;
    uint32_t w[12];
    uint32_t cnt[4] = {0, 0, 0, 0};
    uint32_t exc = 0;
    const bool vec = base + BLK <= n && ((uintptr_t)(bwt + base) & 15) == 0;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        uint32_t bytes4[4];
        if (vec) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(bwt + base) + k);
            bytes4[0] = v.x; bytes4[1] = v.y; bytes4[2] = v.z; bytes4[3] = v.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t x = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int64_t i = base + k * 16 + q * 4 + j;
                    x |= (uint32_t)(i < n ? bwt[i] : (uint8_t)'A') << (8 * j);   // padding: code 0, not counted
                }
                bytes4[q] = x;
            }
        }
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int ch = (bytes4[j >> 2] >> (8 * (j & 3))) & 0xff;
            const bool inside = base + k * 16 + j < n;
            const int c2 = code2_of(ch);
            if (inside) {
                if (c2 >= 0) cnt[c2]++; else exc++;
            }
            word |= (uint32_t)(c2 > 0 ? c2 : 0) << (2 * j);
        }
        w[k] = word;
    }
    uint4 *o = blocks + (size_t)b * 4;
    o[0] = make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]);
    o[1] = make_uint4(w[0], w[1], w[2], w[3]);
    o[2] = make_uint4(w[4], w[5], w[6], w[7]);
    o[3] = make_uint4(w[8], w[9], w[10], w[11]);
    nexc[b] = exc;
}

// scans of the per-block counts: two 31-bit counters per pass
struct CountPair {
    const uint4 *blocks;
    int second;      // 0: (A, C)   1: (G, T)
    __device__ unsigned long long operator()(int64_t b) const
    {
        const uint4 h = blocks[(size_t)b * 4];
        const unsigned long long lo = second ? h.z : h.x, hi = second ? h.w : h.y;
        return lo | (hi << 31);
    }
};
struct WritePair {
    uint32_t *out;   // [nblk][2]
    __device__ void operator()(int64_t b, unsigned long long excl, unsigned long long) const
    {
        out[2 * b] = (uint32_t)(excl & 0x7fffffffull);
        out[2 * b + 1] = (uint32_t)(excl >> 31);
    }
};
__global__ void write_headers_kernel(uint4 *__restrict__ blocks, const uint32_t *__restrict__ ac,
                                     const uint32_t *__restrict__ gt, const uint32_t *__restrict__ nexc, int32_t nblk)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblk) return;
    blocks[(size_t)b * 4] = make_uint4(ac[2 * b] | (nexc[b] ? 0x80000000u : 0u), ac[2 * b + 1], gt[2 * b], gt[2 * b + 1]);
}

// exceptions: positions in ascending order + their byte values
struct CountExc {
    const uint32_t *nexc;
    __device__ unsigned long long operator()(int64_t b) const { return nexc[b]; }
};
struct EmitExc {
    const uint8_t *bwt;
    int32_t n;
    uint32_t *pos;
    uint8_t *code;
    int64_t cap;
    __device__ void operator()(int64_t b, unsigned long long excl, unsigned long long cnt) const
    {
        if (cnt == 0) return;
        int64_t o = (int64_t)excl;
        const int64_t base = b * BLK;
        for (int j = 0; j < BLK && base + j < n; j++) {
            const int ch = bwt[base + j];
            if (code2_of(ch) < 0) {
                if (o < cap) { pos[o] = (uint32_t)(base + j); code[o] = (uint8_t)ch; }
                o++;
            }
        }
    }
};
struct CountCode {
    const uint8_t *code;
    int want;
    __device__ unsigned long long operator()(int64_t i) const { return code[i] == want; }
};
struct EmitCode {
    const uint32_t *pos;
    uint32_t *out;
    __device__ void operator()(int64_t i, unsigned long long excl, unsigned long long cnt) const
    {
        if (cnt) out[excl] = pos[i];
    }
};

// ---- queries -------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    search_kernel(Packed ix, const uint8_t *__restrict__ pats, int64_t stride, const int32_t *__restrict__ lens,
                  int64_t nq, int32_t *__restrict__ sp_out, int32_t *__restrict__ ep_out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int m = lens[q];
    const uint8_t *p = pats + q * stride;
    int sp = 0, ep = ix.n - 1;
    int i = m - 1;
    bool seeded = false;
    if (ix.ftab_k > 0 && m >= ix.ftab_k) {
        // start from the interval of the pattern's last k characters when they are all ACGT
        uint32_t v = 0;
        bool ok = true;
        for (int j = 0; j < ix.ftab_k; j++) {
            const int c2 = code2_of(__ldg(p + m - ix.ftab_k + j));
            ok = ok && c2 >= 0;
            v = (v << 2) | (uint32_t)(c2 & 3);
        }
        if (ok) {
            sp = __ldg(ix.ftab_sp + v);
            ep = __ldg(ix.ftab_ep + v);
            i = m - ix.ftab_k - 1;
            if (sp < 0) i = -1;
            seeded = true;
        }
    }
    if (!seeded && m > 0) {
        // the last character's interval is its C-array bucket (bwt.py:374-377)
        const int c = __ldg(p + m - 1);
        const long long t = __ldg(ix.tot + c);
        if (t == 0) { sp = ep = -1; i = -1; }
        else { sp = (int)__ldg(ix.C + c); ep = sp + (int)t - 1; i = m - 2; }
    }
    for (; i >= 0; i--) {
        const int c = __ldg(p + i);
        const int c2 = code2_of(c);
        if (c2 >= 0) {
            if (ix.tot4[c2] == 0) { sp = ep = -1; break; }
            const int a = rank4(ix, c2, sp), b = rank4(ix, c2, ep + 1);
            sp = ix.C4[c2] + a;
            ep = ix.C4[c2] + b - 1;
        } else {
            if (__ldg(ix.tot + c) == 0) { sp = ep = -1; break; }
            const int cc = (int)__ldg(ix.C + c);
            const int a = rank_any(ix, c, sp), b = rank_any(ix, c, (int64_t)ep + 1);
            sp = cc + a;
            ep = cc + b - 1;
        }
        if (sp > ep) { sp = ep = -1; break; }
    }
    sp_out[q] = sp;
    ep_out[q] = ep;
}

// Warp-cooperative form: the 32 lanes own 32 queries, but every rank is answered by a GROUP OF FOUR
// lanes that load one 16-byte quarter of the 64-byte block each, so a warp-wide load instruction
// touches 8 lines (8 complete blocks) instead of 32 and a rank costs one L1 wavefront instead of
// 3.3 (header, first sector's symbols, and for two positions in three the second sector): the
// thread-per-query kernel is bound by exactly that (ncu: L1tex wavefronts; B300_MICROARCH "LDG").
// The group serves its four members in turn: one shuffle broadcasts (position, symbol), the quarter
// counts are summed with two xor-shuffles.  Lanes whose quarter lies past the position, and groups
// whose member is finished or is on a non-ACGT character, issue no load.
__global__ void __launch_bounds__(256)
    search_coop_kernel(Packed ix, const uint8_t *__restrict__ pats, int64_t stride, const int32_t *__restrict__ lens,
                       int64_t nq, int32_t *__restrict__ sp_out, int32_t *__restrict__ ep_out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, g = lane & 3u;
    const bool have = q < nq;
    const int m = have ? lens[q] : 0;
    const uint8_t *p = pats + (have ? q : 0) * stride;
    int sp = 0, ep = ix.n - 1;
    int i = m - 1;
    bool seeded = false;
    if (have && ix.ftab_k > 0 && m >= ix.ftab_k) {
        uint32_t v = 0;
        bool ok = true;
        for (int j = 0; j < ix.ftab_k; j++) {
            const int c2 = code2_of(__ldg(p + m - ix.ftab_k + j));
            ok = ok && c2 >= 0;
            v = (v << 2) | (uint32_t)(c2 & 3);
        }
        if (ok) {
            sp = __ldg(ix.ftab_sp + v);
            ep = __ldg(ix.ftab_ep + v);
            i = m - ix.ftab_k - 1;
            if (sp < 0) i = -1;
            seeded = true;
        }
    }
    if (have && !seeded && m > 0) {
        const int c = __ldg(p + m - 1);
        const long long t = __ldg(ix.tot + c);
        if (t == 0) { sp = ep = -1; i = -1; }
        else { sp = (int)__ldg(ix.C + c); ep = sp + (int)t - 1; i = m - 2; }
    }
    while (__any_sync(0xffffffffu, i >= 0)) {
        // this lane's request for the step: two positions and a 2-bit symbol, or nothing
        uint32_t msg0 = 0xffffffffu, msg1 = 0xffffffffu;
        int c2 = -1;
        if (i >= 0) {
            const int c = __ldg(p + i);
            c2 = code2_of(c);
            if (c2 >= 0) {
                if (ix.tot4[c2] == 0) { sp = ep = -1; i = -1; c2 = -1; }
                else {
                    msg0 = (uint32_t)sp | ((uint32_t)c2 << 30);
                    msg1 = (uint32_t)(ep + 1) | ((uint32_t)c2 << 30);
                }
            } else {
                // '$', N, IUPAC ...: binary search in the exception lists, by the lane itself
                if (__ldg(ix.tot + c) == 0) { sp = ep = -1; i = -1; }
                else {
                    const int cc = (int)__ldg(ix.C + c);
                    const int a = rank_any(ix, c, sp), b = rank_any(ix, c, (int64_t)ep + 1);
                    sp = cc + a;
                    ep = cc + b - 1;
                    if (sp > ep) { sp = ep = -1; i = -1; } else i--;
                }
            }
        }
        // the group serves its four members in turn; the two ranks of a member (sp and ep + 1) are loaded
        // together (two independent requests in flight per lane; issuing all eight at once costs 74
        // registers and a third of the occupancy, and measured slower: 2.65 vs 3.33 G queries/s)
        int r0 = 0, r1 = 0;
        unsigned flagged = 0;  // requests whose block holds exception symbols (seen by the g == 0 lanes)
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const unsigned owner = (lane & ~3u) | (unsigned)o;
            const uint32_t ma = __shfl_sync(0xffffffffu, msg0, owner);
            const uint32_t mb = __shfl_sync(0xffffffffu, msg1, owner);
            uint4 va = make_uint4(0u, 0u, 0u, 0u), vb = va;
            int ca = -1, cb = -1;          // symbols of this lane's quarter that lie before the position
            if (ma != 0xffffffffu) {       // (both requests of a member are present or absent together)
                const uint32_t pa = ma & 0x3fffffffu, pb = mb & 0x3fffffffu;
                const uint32_t ba = pa / BLK, bb = pb / BLK;
                const int oa = (int)(pa - ba * BLK), ob = (int)(pb - bb * BLK);
                const int mine_a = g == 0 ? 0 : oa - 64 * ((int)g - 1);
                const int mine_b = g == 0 ? 0 : ob - 64 * ((int)g - 1);
                if (g == 0 || mine_a > 0) { va = __ldg(ix.blocks + (size_t)ba * 4 + g); ca = mine_a; }
                if (g == 0 || mine_b > 0) { vb = __ldg(ix.blocks + (size_t)bb * 4 + g); cb = mine_b; }
            }
            const uint32_t s2 = ma >> 30;
            int pa_cnt = 0, pb_cnt = 0;
            bool fa = false, fb = false;
            if (g == 0) {
                if (ca >= 0) {
                    pa_cnt = (int)(s2 == 0 ? (va.x & 0x3fffffffu) : s2 == 1 ? va.y : s2 == 2 ? va.z : va.w);
                    pb_cnt = (int)(s2 == 0 ? (vb.x & 0x3fffffffu) : s2 == 1 ? vb.y : s2 == 2 ? vb.z : vb.w);
                    fa = s2 == 0 && (va.x >> 31);
                    fb = s2 == 0 && (vb.x >> 31);
                }
            } else {
                if (ca >= 0) pa_cnt = count_vec(va, 0x55555555u * s2, ca);
                if (cb >= 0) pb_cnt = count_vec(vb, 0x55555555u * s2, cb);
            }
            pa_cnt += __shfl_xor_sync(0xffffffffu, pa_cnt, 1);
            pb_cnt += __shfl_xor_sync(0xffffffffu, pb_cnt, 1);
            pa_cnt += __shfl_xor_sync(0xffffffffu, pa_cnt, 2);
            pb_cnt += __shfl_xor_sync(0xffffffffu, pb_cnt, 2);
            const unsigned fba = __ballot_sync(0xffffffffu, fa), fbb = __ballot_sync(0xffffffffu, fb);
            if (lane == owner) {
                r0 = pa_cnt;
                r1 = pb_cnt;
                if ((fba >> (lane & ~3u)) & 1u) flagged |= 1u;
                if ((fbb >> (lane & ~3u)) & 1u) flagged |= 2u;
            }
        }
        if (flagged) {
            // '$' / N / IUPAC stored as code 0 inside the block: take them out of the A count (rare)
            if (flagged & 1u) {
                const uint32_t pos = msg0 & 0x3fffffffu;
                if (pos % BLK) r0 -= exceptions_in(ix, pos - pos % BLK, pos);
            }
            if (flagged & 2u) {
                const uint32_t pos = msg1 & 0x3fffffffu;
                if (pos % BLK) r1 -= exceptions_in(ix, pos - pos % BLK, pos);
            }
        }
        if (c2 >= 0) {
            sp = ix.C4[c2] + r0;
            ep = ix.C4[c2] + r1 - 1;
            if (sp > ep) { sp = ep = -1; i = -1; } else i--;
        }
    }
    if (have) {
        sp_out[q] = sp;
        ep_out[q] = ep;
    }
}

__global__ void __launch_bounds__(256)
    rank_kernel(Packed ix, const int32_t *__restrict__ codes, const int64_t *__restrict__ pos, int64_t nq,
                int64_t *__restrict__ out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int c = codes[q];
    out[q] = (c < 0 || c > 255) ? 0 : rank_any(ix, c, pos[q]);
}

// level k >= 2: thread per parent motif X of length k-1; the two blocks at sp(X) and ep(X)+1 give the
// intervals of all four children cX at once
__global__ void __launch_bounds__(256)
    sweep_level_kernel(Packed ix, int k, int64_t parents, int64_t parent_off, int64_t child_off,
                       int32_t *__restrict__ sp_arr, int32_t *__restrict__ ep_arr)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= parents) return;
    if (k == 1) {
        sp_arr[p] = ix.tot4[p] ? ix.C4[p] : -1;
        ep_arr[p] = ix.tot4[p] ? ix.C4[p] + ix.tot4[p] - 1 : -1;
        return;
    }
    const int psp = sp_arr[parent_off + p], pep = ep_arr[parent_off + p];
    int ra[4] = {0, 0, 0, 0}, rb[4] = {0, 0, 0, 0};
    if (psp >= 0) {
        rank4_all(ix, psp, ra);
        rank4_all(ix, pep + 1, rb);
    }
#pragma unroll
    for (int ci = 0; ci < 4; ci++) {
        int sp = -1, ep = -1;
        if (psp >= 0 && ix.tot4[ci] != 0) {
            sp = ix.C4[ci] + ra[ci];
            ep = ix.C4[ci] + rb[ci] - 1;
            if (sp > ep) { sp = -1; ep = -1; }
        }
        const int64_t child = child_off + (int64_t)ci * parents + p;
        sp_arr[child] = sp;
        ep_arr[child] = ep;
    }
}

}  // namespace fmp
}  // namespace bwtk

using namespace bwtk;

extern "C" int64_t bwtk_fm_pack_bytes(int64_t n)
{
    if (n < 0) n = 0;
    return (n / fmp::BLK + 1) * 64;
}

extern "C" int64_t bwtk_fm_pack_workspace_bytes(int64_t n, int64_t n_exc)
{
    if (n < 0) n = 0;
    if (n_exc < 0) n_exc = 0;
    const int64_t nblk = n / fmp::BLK + 1;
    return 3 * align_up(nblk * 8, 256) + align_up(n_exc, 256) + scan::workspace_bytes(nblk) +
           scan::workspace_bytes(n_exc + 1) + 4096;
}

extern "C" int32_t bwtk_fm_pack(const uint8_t *d_bwt, int64_t n, const int64_t *h_totals, void *d_blocks,
                                uint32_t *d_exc_pos, uint32_t *d_exc_by_code, int64_t exc_cap, int64_t *d_code_off,
                                int64_t *h_exc_count, void *d_ws, int64_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    BWTK_REQUIRE(d_bwt && h_totals && d_blocks && d_code_off && h_exc_count && d_ws, "null pointer");
    BWTK_REQUIRE(n >= 1 && n < (1ll << 30), "n must be in [1, 2^30)");
    int64_t n_exc = n;
    for (int c : {'A', 'C', 'G', 'T'}) n_exc -= h_totals[c];
    BWTK_REQUIRE(n_exc >= 0, "byte totals do not match n");
    *h_exc_count = n_exc;
    if (n_exc > exc_cap) {
        set_error("fm_pack: %lld exception symbols exceed capacity %lld", (long long)n_exc, (long long)exc_cap);
        return BWTK_EOVERFLOW;
    }
    BWTK_REQUIRE(n_exc == 0 || (d_exc_pos && d_exc_by_code), "null exception arrays");
    if (ws_bytes < bwtk_fm_pack_workspace_bytes(n, n_exc)) {
        set_error("fm_pack workspace: need %lld bytes", (long long)bwtk_fm_pack_workspace_bytes(n, n_exc));
        return BWTK_EWORKSPACE;
    }
    const int32_t nblk = (int32_t)(n / fmp::BLK + 1);
    Carver c(d_ws, ws_bytes);
    uint32_t *ac = c.take<uint32_t>(2 * (int64_t)nblk);
    uint32_t *gt = c.take<uint32_t>(2 * (int64_t)nblk);
    uint32_t *nexc = c.take<uint32_t>(nblk);
    uint8_t *exc_code = c.take<uint8_t>(n_exc + 1);
    scan::Workspace sws = scan::carve(c, nblk);
    scan::Workspace ews = scan::carve(c, n_exc + 1);
    if (!c.ok()) { set_error("fm_pack workspace carve overflow"); return BWTK_EWORKSPACE; }
    BWTK_CUDA(bwtk::zero_async(sws.err, sizeof(int), st));
    BWTK_CUDA(bwtk::zero_async(ews.err, sizeof(int), st));
    uint4 *blocks = (uint4 *)d_blocks;
    {
        prof::Scope ps("fm_pack_blocks", n + (int64_t)nblk * 64, st);
        fmp::pack_blocks_kernel<<<(unsigned)ceil_div(nblk, 128), 128, 0, st>>>(d_bwt, (int32_t)n, nblk, blocks, nexc);
        BWTK_LAUNCH_CHECK();
    }
    int rc = scan::run(nblk, fmp::CountPair{blocks, 0}, fmp::WritePair{ac}, sws, st);
    if (rc) return rc;
    rc = scan::run(nblk, fmp::CountPair{blocks, 1}, fmp::WritePair{gt}, sws, st);
    if (rc) return rc;
    if (n_exc > 0) {
        rc = scan::run(nblk, fmp::CountExc{nexc}, fmp::EmitExc{d_bwt, (int32_t)n, d_exc_pos, exc_code, n_exc}, sws, st);
        if (rc) return rc;
    }
    fmp::write_headers_kernel<<<(unsigned)ceil_div(nblk, 256), 256, 0, st>>>(blocks, ac, gt, nexc, nblk);
    BWTK_LAUNCH_CHECK();
    // group the exception positions by byte value (ascending inside a group: stable compaction)
    int64_t off[257];
    int64_t run = 0;
    for (int b = 0; b < 256; b++) {
        off[b] = run;
        if (b != 'A' && b != 'C' && b != 'G' && b != 'T') run += h_totals[b];
    }
    off[256] = run;
    BWTK_CUDA(cudaMemcpyAsync(d_code_off, off, sizeof(off), cudaMemcpyHostToDevice, st));
    for (int b = 0; b < 256 && n_exc > 0; b++) {
        if (b == 'A' || b == 'C' || b == 'G' || b == 'T' || h_totals[b] == 0) continue;
        rc = scan::run(n_exc, fmp::CountCode{exc_code, b}, fmp::EmitCode{d_exc_pos, d_exc_by_code + off[b]}, ews, st);
        if (rc) return rc;
    }
    int h_err[2] = {0, 0};
    BWTK_CUDA(cudaMemcpyAsync(&h_err[0], sws.err, 4, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaMemcpyAsync(&h_err[1], ews.err, 4, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaStreamSynchronize(st));   // also: `off` must outlive its upload
    if (h_err[0] || h_err[1]) { set_error("look-back spin limit hit in fm_pack"); return BWTK_EINTERNAL; }
    return BWTK_OK;
}

static int make_packed(const bwtk_fm_index *fx, fmp::Packed *ix)
{
    BWTK_REQUIRE(fx && fx->d_blocks && fx->d_code_off && fx->d_C && fx->d_tot, "null pointer in bwtk_fm_index");
    BWTK_REQUIRE(fx->n >= 1 && fx->n < (1ll << 30), "n must be in [1, 2^30)");
    BWTK_REQUIRE(fx->n_exc == 0 || (fx->d_exc_pos && fx->d_exc_by_code), "null exception arrays");
    BWTK_REQUIRE(fx->ftab_k == 0 || (fx->ftab_k >= 1 && fx->ftab_k <= 12 && fx->d_ftab_sp && fx->d_ftab_ep),
                 "bad k-mer table");
    ix->blocks = (const uint4 *)fx->d_blocks;
    ix->n = (int32_t)fx->n;
    ix->exc_pos = fx->d_exc_pos;
    ix->exc_by_code = fx->d_exc_by_code;
    ix->n_exc = (int32_t)fx->n_exc;
    ix->code_off = fx->d_code_off;
    ix->C = fx->d_C;
    ix->tot = fx->d_tot;
    for (int k = 0; k < 4; k++) {
        ix->C4[k] = (int32_t)fx->acgt_C[k];
        ix->tot4[k] = (int32_t)fx->acgt_tot[k];
    }
    ix->ftab_sp = fx->d_ftab_sp;
    ix->ftab_ep = fx->d_ftab_ep;
    ix->ftab_k = fx->ftab_k;
    return BWTK_OK;
}

// Keeps the rank blocks resident in L2 while queries stream through (cudaAccessPolicyWindow);
// best effort: devices / drivers that refuse simply run without it.
static void l2_window(cudaStream_t st, const void *base, size_t bytes, bool on)
{
    static thread_local int persist_max = -1;
    if (persist_max < 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, dev);
        persist_max = v;
        if (v > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)v);
        cudaGetLastError();
    }
    if (persist_max <= 0) return;
    int dev = 0, win_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&win_max, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    if (on) {
        size_t win = bytes < (size_t)win_max ? bytes : (size_t)win_max;
        attr.accessPolicyWindow.base_ptr = const_cast<void *>(base);
        attr.accessPolicyWindow.num_bytes = win;
        double ratio = (double)persist_max / (double)(win ? win : 1);
        attr.accessPolicyWindow.hitRatio = (float)(ratio > 1.0 ? 1.0 : ratio);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        attr.accessPolicyWindow.num_bytes = 0;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
}

extern "C" int32_t bwtk_fm_search_batch(const bwtk_fm_index *fx, const uint8_t *d_pats, int64_t stride,
                                        const int32_t *d_lens, int64_t nq, int32_t *d_sp, int32_t *d_ep,
                                        int32_t flags, void *stream)
{
    if (nq == 0) return BWTK_OK;
    cudaStream_t st = (cudaStream_t)stream;
    fmp::Packed ix;
    int rc = make_packed(fx, &ix);
    if (rc) return rc;
    BWTK_REQUIRE(d_lens && d_sp && d_ep, "null pointer");
    const bool win = (flags & BWTK_FM_L2_PERSIST) != 0;
    if (win) l2_window(st, fx->d_blocks, (size_t)bwtk_fm_pack_bytes(fx->n), true);
    {
        prof::Scope ps("fm_search_kernel", nq * (stride + 12), st);
        if (flags & BWTK_FM_THREAD_PER_QUERY)
            fmp::search_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(ix, d_pats, stride, d_lens, nq, d_sp, d_ep);
        else
            fmp::search_coop_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, st>>>(ix, d_pats, stride, d_lens, nq, d_sp,
                                                                                d_ep);
        count_launch();
    }
    cudaError_t e = cudaGetLastError();
    if (win) l2_window(st, nullptr, 0, false);
    if (e != cudaSuccess) { set_error("fm_search_batch: kernel launch -> %s", cudaGetErrorString(e)); return BWTK_ECUDA; }
    return BWTK_OK;
}

extern "C" int32_t bwtk_fm_rank_batch(const bwtk_fm_index *fx, const int32_t *d_codes, const int64_t *d_pos,
                                      int64_t nq, int64_t *d_out, void *stream)
{
    if (nq == 0) return BWTK_OK;
    fmp::Packed ix;
    int rc = make_packed(fx, &ix);
    if (rc) return rc;
    BWTK_REQUIRE(d_codes && d_pos && d_out, "null pointer");
    fmp::rank_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, (cudaStream_t)stream>>>(ix, d_codes, d_pos, nq, d_out);
    BWTK_LAUNCH_CHECK();
    return BWTK_OK;
}

extern "C" int32_t bwtk_fm_motif_sweep(const bwtk_fm_index *fx, int32_t kmax, int32_t *d_sp, int32_t *d_ep,
                                       int32_t flags, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    BWTK_REQUIRE(kmax >= 1 && kmax <= 12, "kmax must be in 1..12");
    BWTK_REQUIRE(d_sp && d_ep, "null pointer");
    fmp::Packed ix;
    int rc = make_packed(fx, &ix);
    if (rc) return rc;
    ix.ftab_k = 0;
    const bool win = (flags & BWTK_FM_L2_PERSIST) != 0;
    if (win) l2_window(st, fx->d_blocks, (size_t)bwtk_fm_pack_bytes(fx->n), true);
    int64_t parents = 1, parent_off = 0, child_off = 0;
    cudaError_t e = cudaSuccess;
    for (int k = 1; k <= kmax && e == cudaSuccess; k++) {
        // level k holds 4^k motifs starting at child_off = (4^k - 4)/3
        const int64_t threads = k == 1 ? 4 : parents;
        fmp::sweep_level_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, st>>>(ix, k, threads, parent_off, child_off,
                                                                                 d_sp, d_ep);
        count_launch();
        e = cudaGetLastError();
        parent_off = child_off;
        parents = k == 1 ? 4 : parents * 4;
        child_off += parents;
    }
    if (win) l2_window(st, nullptr, 0, false);
    if (e != cudaSuccess) { set_error("fm_motif_sweep: kernel launch -> %s", cudaGetErrorString(e)); return BWTK_ECUDA; }
    return BWTK_OK;
}
