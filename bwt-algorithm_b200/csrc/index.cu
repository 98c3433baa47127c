// index.cu -- byte histogram (C array), bit-packing, BWT + Occ checkpoints, LCP,
// and the fused index build that shares one histogram / one packed text between
// the suffix array, the BWT and the LCP array.
#include "common.cuh"

#include <stdarg.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace bwtk {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int k) { g_launches += k; }

// ---- small read-backs that bypass the copy engines --------------------------------
// Control values (active counts, flags, the byte histogram) are read by the host several
// times per build.  As cudaMemcpyAsync they queue on the device-to-host copy engine behind
// whatever bulk download another stream has in flight (420 MB per chr21-sized contig in the
// streaming pipeline) and stall the build for milliseconds.  Instead a one-warp kernel
// stores them into mapped pinned host memory; the host reads it after the stream sync.
namespace {
constexpr size_t READ_BACK_MAX = 4096;
__global__ void read_back_kernel(const unsigned char *__restrict__ src, unsigned char *dst, int bytes)
{
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
}
struct Staging {
    unsigned char *host = nullptr;
    ~Staging() { if (host) cudaFreeHost(host); }
};
}  // namespace

namespace {
__global__ void zero_kernel(unsigned char *p, size_t head, size_t vecs, size_t tail)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < head) p[t] = 0;
    uint4 *v = reinterpret_cast<uint4 *>(p + head);
    for (size_t i = t; i < vecs; i += stride) v[i] = make_uint4(0u, 0u, 0u, 0u);
    if (t < tail) p[head + vecs * 16 + t] = 0;
}
}  // namespace

cudaError_t zero_async(void *p, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return cudaSuccess;
    size_t head = (16 - ((uintptr_t)p & 15)) & 15;
    if (head > bytes) head = bytes;
    const size_t vecs = (bytes - head) / 16, tail = bytes - head - vecs * 16;
    size_t want = (vecs + 255) / 256;
    if (want < 1) want = 1;
    if (want > (size_t)NUM_SMS * 8) want = (size_t)NUM_SMS * 8;
    zero_kernel<<<(unsigned)want, 256, 0, st>>>((unsigned char *)p, head, vecs, tail);
    count_launch();
    return cudaGetLastError();
}

int read_back(void *dst, const void *d_src, size_t bytes, cudaStream_t st)
{
    static thread_local Staging stg;
    if (bytes == 0) return BWTK_OK;
    if (bytes > READ_BACK_MAX) {
        BWTK_CUDA(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
        BWTK_CUDA(cudaStreamSynchronize(st));
        return BWTK_OK;
    }
    if (!stg.host)
        BWTK_CUDA(cudaHostAlloc((void **)&stg.host, READ_BACK_MAX, cudaHostAllocPortable | cudaHostAllocMapped));
    unsigned char *d_alias = nullptr;
    BWTK_CUDA(cudaHostGetDevicePointer((void **)&d_alias, stg.host, 0));
    read_back_kernel<<<1, 128, 0, st>>>((const unsigned char *)d_src, d_alias, (int)bytes);
    BWTK_LAUNCH_CHECK();
    BWTK_CUDA(cudaStreamSynchronize(st));
    memcpy(dst, stg.host, bytes);
    return BWTK_OK;
}

namespace prof {
struct Rec {
    const char *name;
    int64_t bytes;
    cudaEvent_t e0, e1;
};
static bool g_on = false;
static std::vector<Rec> g_recs;
static std::mutex g_mu;
bool enabled() { return g_on; }
void begin(const char *name, int64_t algo_bytes, cudaStream_t st)
{
    Rec r;
    r.name = name;
    r.bytes = algo_bytes;
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
    std::lock_guard<std::mutex> g(g_mu);
    g_recs.push_back(r);
}
void end(cudaStream_t st)
{
    std::lock_guard<std::mutex> g(g_mu);
    if (!g_recs.empty()) cudaEventRecord(g_recs.back().e1, st);
}
}  // namespace prof

int64_t sa_core_workspace_bytes(int64_t n);
int sa_build_core(const uint32_t *packed, int64_t n, int bits, bool fast, int32_t *d_sa, int32_t *d_isa_out,
                  void *d_ws, int64_t ws_bytes, int64_t *h_stats, cudaStream_t st, const uint32_t **d_skey0_out,
                  void **d_free_out, int64_t *free_bytes_out);

// ------------------------------------------------------------------ histogram
// out[0..255] = byte counts, out[256] = the last byte of the text
__global__ void __launch_bounds__(256) byte_hist_kernel(const uint8_t *__restrict__ text, int64_t n,
                                                        unsigned long long *__restrict__ out)
{
    __shared__ unsigned int s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    // aligned 16-byte body, byte-wise head and tail
    uintptr_t addr = (uintptr_t)text;
    int64_t head = (int64_t)((16 - (addr & 15)) & 15);
    if (head > n) head = n;
    int64_t body = (n - head) / 16;
    const uint4 *v = reinterpret_cast<const uint4 *>(text + head);
    int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    uint32_t last = 256, run = 0;
    auto add = [&](uint32_t b) {
        if (b == last) {
            run++;
        } else {
            if (run) atomicAdd(&s_h[last], run);
            last = b;
            run = 1;
        }
    };
    for (int64_t i = gtid; i < body; i += gstride) {
        uint4 q = __ldg(v + i);
        uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            add(w[j] & 255u); add((w[j] >> 8) & 255u); add((w[j] >> 16) & 255u); add(w[j] >> 24);
        }
    }
    if (gtid < head) add(text[gtid]);
    int64_t tail0 = head + body * 16;
    if (tail0 + gtid < n && gtid < 16) add(text[tail0 + gtid]);
    if (run) atomicAdd(&s_h[last], run);
    __syncthreads();
    unsigned int c = s_h[threadIdx.x];
    if (c) atomicAdd(&out[threadIdx.x], (unsigned long long)c);
    if (gtid == 0 && n > 0) out[256] = text[n - 1];
}

// d_scratch: 257 x u64.  Synchronises.  h_last (may be null) receives the last byte.
int byte_histogram(const uint8_t *d_text, int64_t n, int64_t *h_totals, int *h_last, unsigned long long *d_scratch,
                   cudaStream_t st)
{
    BWTK_CUDA(bwtk::zero_async(d_scratch, 257 * sizeof(unsigned long long), st));
    if (n > 0) {
        int64_t want = ceil_div(n, 256 * 64);
        int grid = (int)(want < 1 ? 1 : (want > NUM_SMS * 8 ? NUM_SMS * 8 : want));
        {
            prof::Scope ps("byte_hist_kernel", n, st);
            byte_hist_kernel<<<grid, 256, 0, st>>>(d_text, n, d_scratch);
        }
        BWTK_LAUNCH_CHECK();
    }
    int64_t host[257];
    {
        int rc = read_back(host, d_scratch, 257 * sizeof(int64_t), st);
        if (rc) return rc;
    }
    memcpy(h_totals, host, 256 * sizeof(int64_t));
    if (h_last) *h_last = (int)host[256];
    return BWTK_OK;
}

// ------------------------------------------------------------------ packing
int64_t packed_words(int64_t n, int bits) { return ceil_div(n * bits, 32) + 4; }

struct Lut {
    uint8_t v[256];
};

// one thread per output word; bits in {1,2,4,8}; the code table travels as a kernel parameter
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ text, int64_t n, Lut lut, int bits,
                                                   uint32_t *__restrict__ packed, int64_t nwords)
{
    __shared__ uint8_t s_lut[256];
    __shared__ __align__(16) uint8_t s_in[256 * 32];
    s_lut[threadIdx.x] = lut.v[threadIdx.x];
    const int spw = 32 / bits;  // symbols per word
    const int64_t w0 = (int64_t)blockIdx.x * 256;
    const int64_t b0 = w0 * spw;
    const int tile_bytes = 256 * spw;
    if ((((uintptr_t)(text + b0)) & 15) == 0 && b0 + tile_bytes <= n) {
        const uint4 *src = reinterpret_cast<const uint4 *>(text + b0);
        for (int i = threadIdx.x; i < tile_bytes / 16; i += 256) reinterpret_cast<uint4 *>(s_in)[i] = __ldg(src + i);
    } else {
        for (int i = threadIdx.x; i < tile_bytes; i += 256) {
            int64_t g = b0 + i;
            s_in[i] = g < n ? __ldg(text + g) : 0;
        }
    }
    __syncthreads();
    int64_t w = w0 + threadIdx.x;
    if (w >= nwords) return;
    uint32_t acc = 0;
    const uint8_t *src = s_in + threadIdx.x * spw;
    int64_t sym0 = w * spw;
    for (int j = 0; j < spw; j++) {
        uint32_t code = (sym0 + j < n) ? s_lut[src[j]] : 0u;
        acc = (acc << bits) | code;
    }
    packed[w] = acc;
}

// Chooses the packing of a text: returns bits, fills lut; *fast = ACGT$ layout
// ('$', the unique last symbol, shares code 0 with 'A').
static int choose_packing(int64_t n, const int64_t *totals, int last_byte, Lut *lut, bool *fast)
{
    memset(lut->v, 0, 256);
    int sigma = 0;
    for (int b = 0; b < 256; b++) sigma += totals[b] > 0;
    int64_t acgt = totals['A'] + totals['C'] + totals['G'] + totals['T'];
    *fast = (n > 0 && totals['$'] == 1 && acgt == n - 1 && last_byte == '$');
    if (*fast) {
        lut->v['A'] = 0; lut->v['C'] = 1; lut->v['G'] = 2; lut->v['T'] = 3; lut->v['$'] = 0;
        return 2;
    }
    int d = 0;
    for (int b = 0; b < 256; b++)
        if (totals[b] > 0) lut->v[b] = (uint8_t)d++;
    return sigma <= 2 ? 1 : sigma <= 4 ? 2 : sigma <= 16 ? 4 : 8;
}

// Histogram (one host sync) + packing of the text.  d_packed needs packed_words(n, 8)
// words, d_hist_scratch 257 x u64.
int prepare_text(const uint8_t *d_text, int64_t n, uint32_t *d_packed, unsigned long long *d_hist_scratch,
                 int64_t *h_totals, int *bits_out, bool *fast_out, cudaStream_t st)
{
    int last = 0;
    int rc = byte_histogram(d_text, n, h_totals, &last, d_hist_scratch, st);
    if (rc) return rc;
    Lut lut;
    int bits = choose_packing(n, h_totals, last, &lut, fast_out);
    *bits_out = bits;
    int64_t nwords = packed_words(n, bits);
    {
        prof::Scope ps("pack_kernel", n + nwords * 4, st);
        pack_kernel<<<(unsigned)ceil_div(nwords, 256), 256, 0, st>>>(d_text, n, lut, bits, d_packed, nwords);
    }
    BWTK_LAUNCH_CHECK();
    return BWTK_OK;
}

// ------------------------------------------------------------------ BWT + Occ
struct RowMap {
    int16_t row[256];  // row of each byte value in the Occ matrix, -1 if absent
};

// One warp per checkpoint block: gathers the block's BWT bytes, stores them, and
// writes the block's per-row symbol counts into occ[row][blk+1] (prefix-summed
// later, in place).  PACKED (ACGT$ texts): the symbol comes from the 2-bit text -- a quarter of
// the bytes, so that the random gathers of a chromosome-sized contig (62 MB packed against
// 249 MB of bytes) stay inside the 126 MB L2; the '$' that precedes suffix 0 is the one special case.
template <bool PACKED>
__global__ void __launch_bounds__(256)
    bwt_block_kernel(const uint8_t *__restrict__ text, const uint32_t *__restrict__ packed,
                     const int32_t *__restrict__ sa, int64_t n, int occ_rate,
                     RowMap rows, int nrows, uint8_t *__restrict__ bwt, int32_t *__restrict__ occ, int64_t ncp,
                     int64_t nblk)
{
    __shared__ int s_row[256];
    s_row[threadIdx.x] = rows.row[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int64_t warp_g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t blk = warp_g; blk < nblk; blk += nwarps) {
        int64_t j0 = blk * occ_rate;
        int64_t j1 = j0 + occ_rate < n ? j0 + occ_rate : n;
        if (PACKED) {
            // five symbols: counts by ballot (every lane ends up with the block's totals), one store per row
            int c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
            for (int64_t jb = j0; jb < j1; jb += 32) {
                const int64_t j = jb + lane;
                uint32_t code = 5u;                                   // past the end of the block
                if (j < j1) {
                    const int64_t p = (int64_t)__ldg(sa + j) - 1;
                    code = p >= 0 ? (__ldg(packed + (p >> 4)) >> (30 - 2 * (int)(p & 15))) & 3u : 4u;
                    bwt[j] = (uint8_t)((0x2454474341ull >> (8 * code)) & 0xffu);   // "ACGT$"
                }
                c0 += __popc(__ballot_sync(0xffffffffu, code == 0u));
                c1 += __popc(__ballot_sync(0xffffffffu, code == 1u));
                c2 += __popc(__ballot_sync(0xffffffffu, code == 2u));
                c3 += __popc(__ballot_sync(0xffffffffu, code == 3u));
                c4 += __popc(__ballot_sync(0xffffffffu, code == 4u));
            }
            if (lane < 5) {
                const int v = lane == 0 ? c0 : lane == 1 ? c1 : lane == 2 ? c2 : lane == 3 ? c3 : c4;
                const int r = s_row[(int)((0x2454474341ull >> (8 * lane)) & 0xffu)];
                if (r >= 0) occ[(int64_t)r * ncp + blk + 1] = v;
            }
            continue;
        }
        for (int rg = 0; rg < nrows; rg += 8) {
            int cnt[8];
#pragma unroll
            for (int r = 0; r < 8; r++) cnt[r] = 0;
            for (int64_t j = j0 + lane; j < j1; j += 32) {
                uint8_t c;
                if (rg == 0) {
                    int64_t p = (int64_t)__ldg(sa + j) - 1;
                    if (p < 0) p += n;
                    c = __ldg(text + p);
                    bwt[j] = c;
                } else {
                    c = bwt[j];
                }
                int r = s_row[c] - rg;
#pragma unroll
                for (int q = 0; q < 8; q++) cnt[q] += (r == q);
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                int v = cnt[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == q && rg + q < nrows) occ[(int64_t)(rg + q) * ncp + blk + 1] = v;
            }
            __syncwarp();
        }
    }
}

// In-place inclusive prefix sum of occ[row][1..nblk]; occ[row][0] = 0.  Each row is
// cut into OCC_SEGS segments: segment sums first, then every segment scans itself
// starting from the sum of the segments before it.
constexpr int OCC_SEGS = 32;

__global__ void __launch_bounds__(256) occ_partial_kernel(const int32_t *__restrict__ occ, int64_t ncp, int64_t nblk,
                                                          int64_t seg_len, int *__restrict__ partial)
{
    const int seg = blockIdx.x, r = blockIdx.y;
    const int32_t *row = occ + (int64_t)r * ncp + 1;
    int64_t lo = (int64_t)seg * seg_len, hi = lo + seg_len < nblk ? lo + seg_len : nblk;
    int sum = 0;
    for (int64_t i = lo + threadIdx.x; i < hi; i += 256) sum += row[i];
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __shared__ int s_w[8];
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; w++) t += s_w[w];
        partial[r * OCC_SEGS + seg] = t;
    }
}

__global__ void __launch_bounds__(1024) occ_scan_kernel(int32_t *__restrict__ occ, int64_t ncp, int64_t nblk,
                                                        int64_t seg_len, const int *__restrict__ partial)
{
    __shared__ int s_w[32];
    __shared__ int s_carry;
    const int seg = blockIdx.x, r = blockIdx.y;
    int32_t *row = occ + (int64_t)r * ncp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        int c = 0;
        for (int q = 0; q < seg; q++) c += partial[r * OCC_SEGS + q];
        s_carry = c;
        if (seg == 0) row[0] = 0;
    }
    __syncthreads();
    int64_t lo = (int64_t)seg * seg_len, hi = lo + seg_len < nblk ? lo + seg_len : nblk;
    constexpr int IT = 4;
    for (int64_t base = lo; base < hi; base += 1024 * IT) {
        int v[IT], sum = 0;
#pragma unroll
        for (int k = 0; k < IT; k++) {
            int64_t i = base + (int64_t)tid * IT + k;
            v[k] = i < hi ? row[1 + i] : 0;
            sum += v[k];
        }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = s_w[lane];
            int winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            s_w[lane] = winc - w;  // exclusive warp offsets
        }
        __syncthreads();
        int carry = s_carry;
        int run = carry + s_w[warp] + inc - sum;
#pragma unroll
        for (int k = 0; k < IT; k++) {
            int64_t i = base + (int64_t)tid * IT + k;
            run += v[k];
            if (i < hi) row[1 + i] = run;
        }
        __syncthreads();
        if (tid == 1023) s_carry = run;
        __syncthreads();
    }
}

// d_partial: nrows * OCC_SEGS ints of scratch
static int launch_bwt_occ(const uint8_t *d_text, const uint32_t *d_packed2, const int32_t *d_sa, int64_t n, int occ_rate,
                          const RowMap &rows, int nrows, uint8_t *d_bwt, int32_t *d_occ, int *d_partial, cudaStream_t st)
{
    int64_t nblk = ceil_div(n, occ_rate);
    int64_t ncp = n / occ_rate + 1 + (n % occ_rate != 0);
    int64_t grid = ceil_div(nblk, 8);
    if (grid > NUM_SMS * 16) grid = NUM_SMS * 16;
    {
        prof::Scope ps("bwt_block_kernel", n * 5 + (int64_t)nrows * ncp * 4, st);
        if (d_packed2)   // 2-bit ACGT$ layout of the text (position n-1, the '$', is packed as code 0 and never read)
            bwt_block_kernel<true><<<(unsigned)grid, 256, 0, st>>>(d_text, d_packed2, d_sa, n, occ_rate, rows, nrows, d_bwt,
                                                                   d_occ, ncp, nblk);
        else
            bwt_block_kernel<false><<<(unsigned)grid, 256, 0, st>>>(d_text, nullptr, d_sa, n, occ_rate, rows, nrows, d_bwt,
                                                                    d_occ, ncp, nblk);
    }
    BWTK_LAUNCH_CHECK();
    {
        prof::Scope ps("occ_scan_kernel", (int64_t)nrows * ncp * 12, st);
        int64_t seg_len = ceil_div(nblk, OCC_SEGS);
        dim3 grid(OCC_SEGS, nrows);
        occ_partial_kernel<<<grid, 256, 0, st>>>(d_occ, ncp, nblk, seg_len, d_partial);
        count_launch();
        occ_scan_kernel<<<grid, 1024, 0, st>>>(d_occ, ncp, nblk, seg_len, d_partial);
    }
    BWTK_LAUNCH_CHECK();
    return BWTK_OK;
}

// ------------------------------------------------------------------ LCP (lcp.cu)
int64_t lcp_scratch_bytes(int64_t n);
int launch_lcp(const uint32_t *packed, const int32_t *d_sa, const uint32_t *d_skey0, int64_t n, int bits, bool fast,
               int32_t *d_lcp, void *d_scratch, int64_t scratch_bytes, cudaStream_t st);

}  // namespace bwtk

using namespace bwtk;

extern "C" int32_t bwtk_version(void) { return 102; }

static __global__ void __launch_bounds__(256)
    upload_kernel(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst, size_t head, size_t vecs,
                  size_t tail)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < head) dst[t] = src[t];
    const uint4 *sv = reinterpret_cast<const uint4 *>(src + head);
    uint4 *dv = reinterpret_cast<uint4 *>(dst + head);
    for (size_t i = t; i < vecs; i += stride) dv[i] = sv[i];
    if (t < tail) dst[head + vecs * 16 + t] = src[head + vecs * 16 + t];
}

// ---- device memory for hosts that bring no CUDA allocator of their own (the CLI's start-up path: a fresh
// `python bwt.py` answers a small FASTA before `import torch` would have finished) ------------------------------
extern "C" int32_t bwtk_device_count(int32_t *count)
{
    BWTK_REQUIRE(count, "null pointer");
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();          // no driver / no device is an answer (0), not a sticky error
        n = 0;
    }
    *count = n;
    return BWTK_OK;
}

extern "C" int32_t bwtk_dev_alloc(void **d_ptr, int64_t bytes)
{
    BWTK_REQUIRE(d_ptr && bytes >= 0, "bad arguments");
    *d_ptr = nullptr;
    BWTK_CUDA(cudaMalloc(d_ptr, (size_t)(bytes > 0 ? bytes : 1)));
    return BWTK_OK;
}

extern "C" int32_t bwtk_dev_free(void *d_ptr)
{
    if (d_ptr) BWTK_CUDA(cudaFree(d_ptr));
    return BWTK_OK;
}

extern "C" int32_t bwtk_copy_to_device(void *d_dst, const void *h_src, int64_t bytes, void *stream)
{
    if (bytes <= 0) return BWTK_OK;
    BWTK_REQUIRE(d_dst && h_src, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    BWTK_CUDA(cudaMemcpyAsync(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice, st));
    BWTK_CUDA(cudaStreamSynchronize(st));
    return BWTK_OK;
}

extern "C" int32_t bwtk_copy_to_host(void *h_dst, const void *d_src, int64_t bytes, void *stream)
{
    if (bytes <= 0) return BWTK_OK;
    BWTK_REQUIRE(h_dst && d_src, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    BWTK_CUDA(cudaMemcpyAsync(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaStreamSynchronize(st));
    return BWTK_OK;
}

extern "C" int32_t bwtk_download(const void *d_src, void *h_pinned, int64_t bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (bytes <= 0) return BWTK_OK;
    BWTK_REQUIRE(d_src && h_pinned, "null pointer");
    cudaPointerAttributes at;
    BWTK_CUDA(cudaPointerGetAttributes(&at, h_pinned));
    BWTK_REQUIRE(at.type == cudaMemoryTypeHost && at.devicePointer, "destination must be pinned host memory");
    unsigned char *dst = (unsigned char *)at.devicePointer;
    BWTK_REQUIRE(((((uintptr_t)dst) | ((uintptr_t)d_src)) & 15) == 0, "pointers must be 16-byte aligned");
    const size_t vecs = (size_t)bytes / 16, tail = (size_t)bytes - vecs * 16;
    // few CTAs: PCIe is saturated long before the SMs are, and the next contig's kernels need them
    upload_kernel<<<64, 256, 0, st>>>((const unsigned char *)d_src, dst, 0, vecs, tail);
    BWTK_LAUNCH_CHECK();
    return BWTK_OK;
}

extern "C" int32_t bwtk_upload_text(const uint8_t *h_pinned, uint8_t *d_dst, int64_t n, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0) return BWTK_OK;
    BWTK_REQUIRE(h_pinned && d_dst, "null pointer");
    cudaPointerAttributes at;
    BWTK_CUDA(cudaPointerGetAttributes(&at, h_pinned));
    BWTK_REQUIRE(at.type == cudaMemoryTypeHost && at.devicePointer, "source must be pinned host memory");
    const unsigned char *src = (const unsigned char *)at.devicePointer;
    size_t head = (16 - ((uintptr_t)src & 15)) & 15;
    if (head > (size_t)n || (((uintptr_t)d_dst + head) & 15)) head = (size_t)n;   // misaligned pair: byte copy
    size_t vecs = ((size_t)n - head) / 16, tail = (size_t)n - head - vecs * 16;
    if (head == (size_t)n) {
        // byte path through the vector loop is not possible; use the head/tail lanes in chunks
        BWTK_CUDA(cudaMemcpyAsync(d_dst, h_pinned, (size_t)n, cudaMemcpyHostToDevice, st));
        return BWTK_OK;
    }
    upload_kernel<<<NUM_SMS * 8, 256, 0, st>>>(src, d_dst, head, vecs, tail);
    BWTK_LAUNCH_CHECK();
    return BWTK_OK;
}

extern "C" int32_t bwtk_last_error(char *buf, int32_t buflen)
{
    if (!buf || buflen <= 0) return BWTK_EINVAL;
    strncpy(buf, g_err, (size_t)buflen - 1);
    buf[buflen - 1] = 0;
    return BWTK_OK;
}

extern "C" int64_t bwtk_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int32_t bwtk_profile_enable(int32_t on)
{
    prof::g_on = on != 0;
    return BWTK_OK;
}

// Synchronises the device, aggregates the recorded launches by name and writes
// "name\tlaunches\ttotal_ms\talgorithmic_bytes\n" lines; clears the records.
extern "C" int32_t bwtk_profile_report(char *buf, int32_t buflen)
{
    if (!buf || buflen <= 0) return BWTK_EINVAL;
    cudaDeviceSynchronize();
    struct Agg { int64_t n = 0; double ms = 0; int64_t bytes = 0; };
    std::map<std::string, Agg> agg;
    std::vector<std::string> order;
    {
        std::lock_guard<std::mutex> g(prof::g_mu);
        for (auto &r : prof::g_recs) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, r.e0, r.e1);
            cudaEventDestroy(r.e0);
            cudaEventDestroy(r.e1);
            if (!agg.count(r.name)) order.push_back(r.name);
            Agg &a = agg[r.name];
            a.n++; a.ms += ms; a.bytes += r.bytes;
        }
        prof::g_recs.clear();
    }
    std::string out;
    char line[256];
    for (auto &nm : order) {
        Agg &a = agg[nm];
        snprintf(line, sizeof(line), "%s\t%lld\t%.6f\t%lld\n", nm.c_str(), (long long)a.n, a.ms, (long long)a.bytes);
        out += line;
    }
    strncpy(buf, out.c_str(), (size_t)buflen - 1);
    buf[buflen - 1] = 0;
    return BWTK_OK;
}

extern "C" int32_t bwtk_byte_histogram(const uint8_t *d_text, int64_t n, int64_t *h_totals, void *stream)
{
    BWTK_REQUIRE(h_totals && n >= 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *scratch = nullptr;
    BWTK_CUDA(cudaMallocAsync((void **)&scratch, 257 * sizeof(unsigned long long), st));
    int rc = byte_histogram(d_text, n, h_totals, nullptr, scratch, st);
    cudaFreeAsync(scratch, st);
    return rc;
}

static int make_rowmap(const int32_t *h_row_of_code, RowMap *rm)
{
    for (int b = 0; b < 256; b++) {
        int r = h_row_of_code[b];
        if (r < -1 || r > 255) return -1;
        rm->row[b] = (int16_t)r;
    }
    return 0;
}

extern "C" int64_t bwtk_bwt_occ_workspace_bytes(int64_t, int32_t, int32_t nrows)
{
    return (int64_t)(nrows > 0 ? nrows : 1) * OCC_SEGS * 4 + 256;
}

extern "C" int32_t bwtk_bwt_occ(const uint8_t *d_text, const int32_t *d_sa, int64_t n, int32_t occ_rate,
                                const int32_t *h_row_of_code, int32_t nrows, uint8_t *d_bwt,
                                int32_t *d_occ, void *d_ws, int64_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return BWTK_OK;
    BWTK_REQUIRE(d_text && d_sa && d_bwt && d_occ && h_row_of_code && d_ws, "null pointer");
    BWTK_REQUIRE(occ_rate >= 1 && nrows >= 1 && nrows <= 256, "bad occ_rate/nrows");
    BWTK_REQUIRE(ws_bytes >= bwtk_bwt_occ_workspace_bytes(n, occ_rate, nrows), "workspace too small");
    RowMap rm;
    BWTK_REQUIRE(make_rowmap(h_row_of_code, &rm) == 0, "row_of_code entries must be in [-1, 255]");
    return launch_bwt_occ(d_text, nullptr, d_sa, n, occ_rate, rm, nrows, d_bwt, d_occ, (int *)d_ws, st);
}

extern "C" int64_t bwtk_lcp_workspace_bytes(int64_t n)
{
    if (n < 1) n = 1;
    return align_up(packed_words(n, 8) * 4, 256) + 8192 + lcp_scratch_bytes(n);
}

extern "C" int32_t bwtk_lcp_build(const uint8_t *d_text, const int32_t *d_sa, int64_t n, int32_t *d_lcp,
                                  void *d_ws, int64_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return BWTK_OK;
    BWTK_REQUIRE(d_text && d_sa && d_lcp && d_ws, "null pointer");
    if (ws_bytes < bwtk_lcp_workspace_bytes(n)) {
        set_error("lcp workspace: need %lld bytes", (long long)bwtk_lcp_workspace_bytes(n));
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
    return launch_lcp(packed, d_sa, nullptr, n, bits, fast, d_lcp, (char *)d_ws + c.off, ws_bytes - c.off, st);
}

// ---- fused index build: a3 + a4 + a5 + a6 + a10 with one histogram, one packed
// text and no host round trips other than the alphabet read-back and the
// per-batch suffix-array round counts -------------------------------------------
extern "C" int64_t bwtk_index_workspace_bytes(int64_t n)
{
    if (n < 1) n = 1;
    return sa_core_workspace_bytes(n) + align_up(packed_words(n, 8) * 4, 256) + 256 * OCC_SEGS * 4 + 8192;
}

extern "C" int32_t bwtk_index_build(const uint8_t *d_text, int64_t n, int32_t occ_rate, int32_t *d_sa,
                                    int32_t *d_isa, uint8_t *d_bwt, int32_t *d_occ, int32_t occ_rows_cap,
                                    int32_t *d_lcp, int64_t *h_totals, int32_t *h_row_of_code, int64_t *h_stats,
                                    void *d_ws, int64_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (h_stats) memset(h_stats, 0, 8 * sizeof(int64_t));
    BWTK_REQUIRE(h_totals && h_row_of_code, "null host output");
    memset(h_totals, 0, 256 * sizeof(int64_t));
    for (int b = 0; b < 256; b++) h_row_of_code[b] = -1;
    BWTK_REQUIRE(n >= 0 && n < (1ll << 30), "n must be in [0, 2^30)");
    if (n == 0) return BWTK_OK;
    BWTK_REQUIRE(d_text && d_sa && d_ws, "null pointer");
    BWTK_REQUIRE(occ_rate >= 1, "bad occ_rate");
    if (ws_bytes < bwtk_index_workspace_bytes(n)) {
        set_error("index workspace: need %lld bytes, got %lld", (long long)bwtk_index_workspace_bytes(n),
                  (long long)ws_bytes);
        return BWTK_EWORKSPACE;
    }
    Carver c(d_ws, ws_bytes);
    uint32_t *packed = c.take<uint32_t>(packed_words(n, 8));
    unsigned long long *d_hist = c.take<unsigned long long>(260);
    int *d_partial = c.take<int>(256 * OCC_SEGS);
    int bits;
    bool fast;
    int rc = prepare_text(d_text, n, packed, d_hist, h_totals, &bits, &fast, st);
    if (rc) return rc;
    int nrows = 0;
    RowMap rm;
    for (int b = 0; b < 256; b++) {
        rm.row[b] = -1;
        if (h_totals[b] > 0) { h_row_of_code[b] = nrows; rm.row[b] = (int16_t)nrows; nrows++; }
    }
    if (d_bwt && d_occ && nrows > occ_rows_cap) {
        if (h_stats) h_stats[7] = nrows;
        set_error("Occ matrix has room for %d rows, the text has %d distinct bytes", occ_rows_cap, nrows);
        return BWTK_EOVERFLOW;
    }
    c.off = align_up(c.off, 256);
    const uint32_t *d_skey0 = nullptr;
    void *d_free = nullptr;          // the part of the suffix-sort workspace that is dead once the SA stands
    int64_t free_bytes = 0;
    rc = sa_build_core(packed, n, bits, fast, d_sa, d_isa, (char *)d_ws + c.off, ws_bytes - c.off, h_stats, st, &d_skey0,
                       &d_free, &free_bytes);
    if (rc) return rc;
    if (h_stats) h_stats[7] = nrows;
    if (d_bwt && d_occ) {
        rc = launch_bwt_occ(d_text, fast ? packed : nullptr, d_sa, n, occ_rate, rm, nrows, d_bwt, d_occ, d_partial, st);
        if (rc) return rc;
    }
    if (d_lcp) {
        rc = launch_lcp(packed, d_sa, d_skey0, n, bits, fast, d_lcp, d_free, free_bytes, st);
        if (rc) return rc;
    }
    return BWTK_OK;
}
