// search.cu -- FM-index rank / backward search (reference bwt.py:335-389),
// batched over patterns, plus the trie sweep over every ACGT motif of length
// 1..kmax (one LF step per motif: interval(cX) is derived from interval(X)).
#include "common.cuh"

namespace bwtk {
namespace fm {

struct Index {
    const uint8_t *bwt;
    const int32_t *occ;  // [nrows][ncp]
    int64_t ncp;
    int occ_rate;
    const int64_t *C;    // [256]
    const int64_t *tot;  // [256]
    const int32_t *row;  // [256]
    int64_t n;
};

// exact count of bytes equal to `code` among the 4 bytes of w
__device__ __forceinline__ int eq_bytes(uint32_t w, uint32_t pat4)
{
    uint32_t x = w ^ pat4;
    uint32_t y = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    y = ~(y | x | 0x7f7f7f7fu);  // 0x80 in every byte of x that is zero
    return __popc(y);
}

// #code in bwt[0:pos]  (bwt.py:335-357; remainder scan bwt.py:46-53)
__device__ __forceinline__ int64_t rank_of(const Index &ix, int code, int64_t pos)
{
    if (pos <= 0) return 0;
    if (pos > ix.n) pos = ix.n;
    int r = __ldg(ix.row + code);
    if (r < 0) return 0;
    int64_t idx = pos / ix.occ_rate;
    int64_t p = idx * ix.occ_rate;
    int64_t base = __ldg(ix.occ + (int64_t)r * ix.ncp + idx);
    const uint8_t *b = ix.bwt;
    uint32_t pat4 = (uint32_t)code * 0x01010101u;
    // byte-wise up to 4-byte alignment, then words, then tail
    while (p < pos && (((uintptr_t)(b + p)) & 3)) { base += (b[p] == code); p++; }
    while (p + 4 <= pos) {
        base += eq_bytes(__ldg(reinterpret_cast<const uint32_t *>(b + p)), pat4);
        p += 4;
    }
    while (p < pos) { base += (__ldg(b + p) == code); p++; }
    return base;
}

__global__ void __launch_bounds__(256)
    bsearch_kernel(Index ix, const uint8_t *__restrict__ pats, int64_t stride,
                   const int32_t *__restrict__ lens, int64_t nq, int32_t *__restrict__ sp_out,
                   int32_t *__restrict__ ep_out)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int m = lens[q];
    const uint8_t *p = pats + q * stride;
    int64_t sp, ep;
    if (m == 0) {
        sp = 0; ep = ix.n - 1;
    } else {
        int c = p[m - 1];
        if (__ldg(ix.tot + c) == 0) { sp = -1; ep = -1; }
        else {
            sp = __ldg(ix.C + c);
            ep = sp + __ldg(ix.tot + c) - 1;
            for (int i = m - 2; i >= 0; i--) {
                c = p[i];
                if (__ldg(ix.tot + c) == 0) { sp = -1; ep = -1; break; }
                int64_t cc = __ldg(ix.C + c);
                sp = cc + rank_of(ix, c, sp);
                ep = cc + rank_of(ix, c, ep + 1) - 1;
                if (sp > ep) { sp = -1; ep = -1; break; }
            }
        }
    }
    sp_out[q] = (int32_t)sp;
    ep_out[q] = (int32_t)ep;
}

__global__ void __launch_bounds__(256)
    rank_kernel(Index ix, const int32_t *__restrict__ codes, const int64_t *__restrict__ pos, int64_t nq,
                int64_t *__restrict__ out)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int c = codes[q];
    out[q] = (c < 0 || c > 255) ? 0 : rank_of(ix, c, pos[q]);
}

// level k (k >= 2): thread per parent motif X of length k-1, emits cX for c in ACGT
__global__ void __launch_bounds__(256)
    sweep_level_kernel(Index ix, int k, int64_t parents, int64_t parent_off, int64_t child_off,
                       int32_t *__restrict__ sp_arr, int32_t *__restrict__ ep_arr)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= parents) return;
    const int codes[4] = {'A', 'C', 'G', 'T'};
    if (k == 1) {
        int c = codes[p];
        int64_t t = __ldg(ix.tot + c);
        sp_arr[p] = t ? (int32_t)__ldg(ix.C + c) : -1;
        ep_arr[p] = t ? (int32_t)(__ldg(ix.C + c) + t - 1) : -1;
        return;
    }
    int64_t psp = sp_arr[parent_off + p], pep = ep_arr[parent_off + p];
#pragma unroll
    for (int ci = 0; ci < 4; ci++) {
        int c = codes[ci];
        int64_t sp = -1, ep = -1;
        if (psp >= 0 && __ldg(ix.tot + c) != 0) {
            int64_t cc = __ldg(ix.C + c);
            sp = cc + rank_of(ix, c, psp);
            ep = cc + rank_of(ix, c, pep + 1) - 1;
            if (sp > ep) { sp = -1; ep = -1; }
        }
        int64_t child = child_off + (int64_t)ci * parents + p;
        sp_arr[child] = (int32_t)sp;
        ep_arr[child] = (int32_t)ep;
    }
}

}  // namespace fm
}  // namespace bwtk

using namespace bwtk;

static fm::Index make_index(const uint8_t *d_bwt, const int32_t *d_occ, int64_t ncp, int32_t occ_rate,
                            const int64_t *d_C, const int64_t *d_tot, const int32_t *d_row, int64_t n)
{
    fm::Index ix;
    ix.bwt = d_bwt; ix.occ = d_occ; ix.ncp = ncp; ix.occ_rate = occ_rate;
    ix.C = d_C; ix.tot = d_tot; ix.row = d_row; ix.n = n;
    return ix;
}

extern "C" int32_t bwtk_bsearch_batch(const uint8_t *d_bwt, const int32_t *d_occ, int64_t ncp,
                                      int32_t occ_rate, const int64_t *d_C, const int64_t *d_tot,
                                      const int32_t *d_row_of_code, int64_t n, const uint8_t *d_pats,
                                      int64_t stride, const int32_t *d_lens, int64_t nq, int32_t *d_sp,
                                      int32_t *d_ep, void *stream)
{
    if (nq == 0) return BWTK_OK;
    BWTK_REQUIRE(d_bwt && d_occ && d_C && d_tot && d_row_of_code && d_lens && d_sp && d_ep, "null pointer");
    BWTK_REQUIRE(occ_rate >= 1 && n >= 1, "bad occ_rate / n");
    fm::Index ix = make_index(d_bwt, d_occ, ncp, occ_rate, d_C, d_tot, d_row_of_code, n);
    fm::bsearch_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, (cudaStream_t)stream>>>(
        ix, d_pats, stride, d_lens, nq, d_sp, d_ep);
    BWTK_LAUNCH_CHECK();
    return BWTK_OK;
}

extern "C" int32_t bwtk_rank_batch(const uint8_t *d_bwt, const int32_t *d_occ, int64_t ncp,
                                   int32_t occ_rate, const int32_t *d_row_of_code, int64_t n,
                                   const int32_t *d_codes, const int64_t *d_pos, int64_t nq,
                                   int64_t *d_out, void *stream)
{
    if (nq == 0) return BWTK_OK;
    BWTK_REQUIRE(d_bwt && d_occ && d_row_of_code && d_codes && d_pos && d_out, "null pointer");
    fm::Index ix = make_index(d_bwt, d_occ, ncp, occ_rate, nullptr, nullptr, d_row_of_code, n);
    fm::rank_kernel<<<(unsigned)ceil_div(nq, 256), 256, 0, (cudaStream_t)stream>>>(ix, d_codes, d_pos, nq,
                                                                                 d_out);
    BWTK_LAUNCH_CHECK();
    return BWTK_OK;
}

extern "C" int32_t bwtk_bsearch_motif_sweep(const uint8_t *d_bwt, const int32_t *d_occ, int64_t ncp,
                                            int32_t occ_rate, const int64_t *d_C, const int64_t *d_tot,
                                            const int32_t *d_row_of_code, int64_t n, int32_t kmax,
                                            int32_t *d_sp, int32_t *d_ep, void *stream)
{
    BWTK_REQUIRE(kmax >= 1 && kmax <= 12, "kmax must be in 1..12");
    BWTK_REQUIRE(d_bwt && d_occ && d_C && d_tot && d_row_of_code && d_sp && d_ep, "null pointer");
    fm::Index ix = make_index(d_bwt, d_occ, ncp, occ_rate, d_C, d_tot, d_row_of_code, n);
    int64_t parents = 1, parent_off = 0, child_off = 0;
    for (int k = 1; k <= kmax; k++) {
        // level k holds 4^k motifs starting at child_off = (4^k - 4)/3
        int64_t threads = k == 1 ? 4 : parents;
        fm::sweep_level_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, (cudaStream_t)stream>>>(
            ix, k, threads, parent_off, child_off, d_sp, d_ep);
        BWTK_LAUNCH_CHECK();
        parent_off = child_off;
        parents = k == 1 ? 4 : parents * 4;
        child_off += parents;
    }
    return BWTK_OK;
}
