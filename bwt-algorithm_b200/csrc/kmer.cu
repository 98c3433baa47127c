// kmer.cu -- 8-mer position index (replaces BWTCore._build_kmer_hash,
// reference bwt.py:138-171): bucket offsets per 16-bit code + window start
// positions, ascending inside a bucket.  Semantics kept from the reference:
//   * A,N -> 0, C -> 1, G -> 2, T -> 3 (either case); every other byte is
//     skipped WITHOUT resetting the rolling window, so a window is "the last
//     <= 8 valid symbols" and its recorded position is (index of its last
//     symbol) - 7;
//   * position 0 is recorded only when the first 8 bytes are all valid;
//   * nothing is recorded for n < 8.
// Pipeline: validity scan + compaction of 2-bit codes -> 16-bit window codes ->
// stable LSD radix sort (2 passes) of (code, position) -> bucket offsets.
#include "radix_sort.cuh"
#include "scan.cuh"

namespace bwtk {
namespace kmer {

__device__ __forceinline__ int base_bits(uint8_t c)
{
    switch (c) {
    case 'A': case 'a': case 'N': case 'n': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

struct CountValid {
    const uint8_t *text;
    __device__ uint64_t operator()(int64_t i) const { return base_bits(__ldg(text + i)) >= 0 ? 1ull : 0ull; }
};

struct EmitValid {
    const uint8_t *text;
    uint8_t *cc;      // compacted 2-bit codes of the valid symbols
    int32_t *pos;     // recorded window positions, in recording order
    uint32_t *vend;   // number of valid symbols up to and including the window's last symbol
    __device__ void operator()(int64_t i, uint64_t excl64, uint64_t cnt) const
    {
        if (!cnt) return;
        uint32_t excl = (uint32_t)excl64;
        cc[excl] = (uint8_t)base_bits(__ldg(text + i));
        if (i < 7) return;
        // V7 = valid symbols among text[0..7]
        uint32_t v7 = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) v7 += base_bits(__ldg(text + j)) >= 0;
        uint32_t first_ok = v7 == 8 ? 1u : 0u;
        if (i == 7) {
            if (first_ok) { pos[0] = 0; vend[0] = 8; }
            return;
        }
        uint32_t r = excl - v7 + first_ok;  // valid symbols in [8, i) + the first window
        pos[r] = (int32_t)(i - 7);
        vend[r] = excl + 1;
    }
};

__global__ void window_codes_kernel(const uint8_t *__restrict__ cc, const uint32_t *__restrict__ vend,
                                    int64_t R, uint32_t *__restrict__ key)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    int64_t v = vend[r];
    uint32_t w = 0;
#pragma unroll
    for (int j = 8; j >= 1; j--) {
        int64_t q = v - j;
        w = (w << 2) | (q >= 0 ? (uint32_t)cc[q] : 0u);
    }
    key[r] = w & 0xffffu;
}

// bucket_off[c] = first r with key[r] >= c (keys sorted); bucket_off[65536] = R
__global__ void bucket_offsets_kernel(const uint32_t *__restrict__ key, int64_t R, int32_t *__restrict__ off)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > R) return;
    uint32_t lo = r == 0 ? 0u : key[r - 1] + 1u;
    uint32_t hi = r == R ? 65536u : key[r];
    for (uint32_t c = lo; c <= hi; c++) off[c] = (int32_t)r;
}

}  // namespace kmer
}  // namespace bwtk

using namespace bwtk;

extern "C" int64_t bwtk_kmer8_workspace_bytes(int64_t n)
{
    if (n < 1) n = 1;
    return align_up(n, 256) + 4 * align_up(n * 4, 256) + scan::workspace_bytes(n) +
           rsort::workspace_bytes(n) + 8192;
}

extern "C" int32_t bwtk_kmer8_index(const uint8_t *d_text, int64_t n, int32_t *d_bucket_off,
                                    int32_t *d_pos, int64_t *h_count, void *d_ws, int64_t ws_bytes,
                                    void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    BWTK_REQUIRE(d_bucket_off && h_count, "null pointer");
    *h_count = 0;
    BWTK_CUDA(bwtk::zero_async(d_bucket_off, 65537 * sizeof(int32_t), st));
    if (n < 8) { BWTK_CUDA(cudaStreamSynchronize(st)); return BWTK_OK; }
    BWTK_REQUIRE(d_text && d_pos && d_ws, "null pointer");
    BWTK_REQUIRE(n < (1ll << 30), "n must be < 2^30");
    if (ws_bytes < bwtk_kmer8_workspace_bytes(n)) {
        set_error("kmer8 workspace: need %lld bytes", (long long)bwtk_kmer8_workspace_bytes(n));
        return BWTK_EWORKSPACE;
    }
    Carver c(d_ws, ws_bytes);
    uint8_t *cc = c.take<uint8_t>(n);
    uint32_t *vend = c.take<uint32_t>(n);
    uint32_t *key0 = c.take<uint32_t>(n);
    uint32_t *key1 = c.take<uint32_t>(n);
    uint32_t *val1 = c.take<uint32_t>(n);
    scan::Workspace sws = scan::carve(c, n);
    rsort::Workspace rws = rsort::carve(c, n);
    if (!c.ok()) { set_error("kmer8 workspace carve overflow"); return BWTK_EWORKSPACE; }
    BWTK_CUDA(bwtk::zero_async(sws.err, sizeof(int), st));
    BWTK_CUDA(bwtk::zero_async(rws.err, sizeof(int), st));

    kmer::CountValid cv{d_text};
    kmer::EmitValid ev{d_text, cc, d_pos, vend};
    int rc = scan::run(n, cv, ev, sws, st);
    if (rc) return rc;
    // recorded windows = valid symbols at i >= 8, plus the first window
    unsigned long long h_valid = 0;
    uint8_t first8[8];
    BWTK_CUDA(cudaMemcpyAsync(&h_valid, sws.total, 8, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaMemcpyAsync(first8, d_text, 8, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaStreamSynchronize(st));
    int v7 = 0;
    for (int j = 0; j < 8; j++) {
        uint8_t ch = first8[j];
        v7 += (ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T' || ch == 'N' || ch == 'a' || ch == 'c' ||
               ch == 'g' || ch == 't' || ch == 'n');
    }
    int64_t R = (int64_t)h_valid - v7 + (v7 == 8 ? 1 : 0);
    *h_count = R;
    if (R > 0) {
        kmer::window_codes_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, st>>>(cc, vend, R, key0);
        BWTK_LAUNCH_CHECK();
        int in_first = 1;
        rc = rsort::sort_pairs<uint32_t>(key0, (uint32_t *)d_pos, key1, val1, R, 0, 16, rws, st, &in_first,
                                         nullptr);
        if (rc) return rc;
        const uint32_t *sk = in_first ? key0 : key1;
        if (!in_first)
            BWTK_CUDA(cudaMemcpyAsync(d_pos, val1, (size_t)R * 4, cudaMemcpyDeviceToDevice, st));
        kmer::bucket_offsets_kernel<<<(unsigned)ceil_div(R + 1, 256), 256, 0, st>>>(sk, R, d_bucket_off);
        BWTK_LAUNCH_CHECK();
    }
    int h_err[2] = {0, 0};
    BWTK_CUDA(cudaMemcpyAsync(&h_err[0], sws.err, sizeof(int), cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaMemcpyAsync(&h_err[1], rws.err, sizeof(int), cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaStreamSynchronize(st));
    if (h_err[0] || h_err[1]) { set_error("look-back spin limit hit in kmer8 index"); return BWTK_EINTERNAL; }
    return BWTK_OK;
}
