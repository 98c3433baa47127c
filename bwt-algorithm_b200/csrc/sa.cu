// sa.cu -- suffix-array construction (replaces BWTCore._build_suffix_array,
// reference bwt.py:212-264) by prefix doubling over a bit-packed text:
//
//   round 0 : 32-bit key = the first S symbols of every suffix (S = 32/bits,
//             16 bases for the ACGT$ fast path), LSD radix sort of (key, index)
//             pairs fed in DECREASING index order, so that the stable sort puts
//             a suffix that runs off the end of the text before longer ones
//             with the same zero-padded key ("shorter suffix first", the
//             reference's key2 = -1 rule);
//   regroup : head flags -> group-head position (max-scan) -> rank[] scatter,
//             SA write, compaction of suffixes whose group is not yet a
//             singleton (one single-pass decoupled look-back scan);
//   round r : for the active suffixes only, key = (group head, rank[i+h]+1),
//             LSD radix sort, regroup; h doubles.  Stops when nothing is active.
//
// The result is the unique suffix array, so it is bit-identical to the
// reference's regardless of the refinement path taken.
#include "radix_sort.cuh"

namespace bwtk {

int pack_text(const uint8_t *d_text, int64_t n, const uint8_t *h_lut, int bits, uint32_t *d_packed,
              uint8_t *d_lut_scratch, cudaStream_t st);
int64_t packed_words(int64_t n, int bits);
int byte_histogram(const uint8_t *d_text, int64_t n, int64_t *h_totals, unsigned long long *d_scratch,
                   cudaStream_t st);
int choose_packing(const uint8_t *d_text, int64_t n, const int64_t *totals, uint8_t *lut, bool *fast,
                   cudaStream_t st);

namespace sa {

constexpr int RG_THREADS = 256;
constexpr int RG_ITEMS = 8;
constexpr int RG_TILE = RG_THREADS * RG_ITEMS;

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

__global__ void init_keys_kernel(const uint32_t *__restrict__ packed, int64_t n, int bits, int S,
                                 uint32_t *__restrict__ key, uint32_t *__restrict__ val)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int64_t i = n - 1 - t;
    uint32_t w = window32(packed, i * bits);
    int used = S * bits;
    key[t] = used == 32 ? w : (w >> (32 - used));
    val[t] = (uint32_t)i;
}

__global__ void build_keys_kernel(const uint32_t *__restrict__ suf, const int32_t *__restrict__ grp,
                                  const int32_t *__restrict__ rank, int64_t m, int64_t n, int64_t h,
                                  int kbits, uint64_t *__restrict__ key)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    int64_t s = (int64_t)suf[t] + h;
    uint64_t k2 = s < n ? (uint64_t)(__ldg(rank + s) + 1) : 0ull;
    key[t] = ((uint64_t)(uint32_t)grp[t] << kbits) | k2;
}

// One pass over the sorted (key, suffix) list: see file header.
template <typename KeyT, bool FIRST>
__global__ void __launch_bounds__(RG_THREADS)
    regroup_kernel(const KeyT *__restrict__ skey, const uint32_t *__restrict__ ssuf,
                   const int32_t *__restrict__ pos, int64_t m, int64_t short_from,
                   int32_t *__restrict__ sa, int32_t *__restrict__ rank,
                   int32_t *__restrict__ npos, uint32_t *__restrict__ nsuf, int32_t *__restrict__ ngrp,
                   unsigned long long *status, unsigned *tile_counter, unsigned *out_count, int *err)
{
    __shared__ unsigned s_tile;
    __shared__ unsigned long long s_warp[RG_THREADS / 32];
    __shared__ unsigned long long s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t t0 = tile * RG_TILE + (int64_t)tid * RG_ITEMS;

    // keys/suffixes t0-1 .. t0+ITEMS (heads need the predecessor, activity the successor)
    KeyT k[RG_ITEMS + 2];
    uint32_t sf[RG_ITEMS + 2];
#pragma unroll
    for (int j = 0; j < RG_ITEMS + 2; j++) {
        int64_t t = t0 - 1 + j;
        bool ok = t >= 0 && t < m;
        k[j] = ok ? skey[t] : (KeyT)0;
        sf[j] = ok ? ssuf[t] : 0u;
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
    unsigned long long agg = rg_pack(lmax, lsum);
    unsigned long long inc = agg;
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

    if (tid == RG_THREADS - 1) {
        unsigned long long tile_agg = rg_combine(wpre, inc);
        volatile unsigned long long *st = status;
        unsigned long long excl = 0;
        if (tile == 0) {
            st[0] = RG_INCL | tile_agg;
        } else {
            st[tile] = RG_AGG | tile_agg;
            int64_t t = tile - 1;
            while (true) {
                unsigned long long s = st[t];
                int spins = 0;
                while ((s & RG_FLAGS) == 0ull) {
                    if (++spins > rsort::SPIN_LIMIT) {
                        *err = 2;
                        s = RG_INCL;
                        break;
                    }
                    __nanosleep(32);
                    s = st[t];
                }
                excl = rg_combine(s & ~RG_FLAGS, excl);
                if (s & RG_INCL) break;
                t--;
            }
            st[tile] = RG_INCL | rg_combine(excl, tile_agg);
        }
        s_prefix = excl;
        if ((tile + 1) * (int64_t)RG_TILE >= m) *out_count = rg_sum(rg_combine(excl, tile_agg));
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
            sa[p[j]] = (int32_t)s;
            rank[s] = hp;
            if (!(head[j] && head[j + 1])) {
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
}  // namespace bwtk

using namespace bwtk;

extern "C" int64_t bwtk_sa_workspace_bytes(int64_t n)
{
    if (n < 1) n = 1;
    int64_t b = 0;
    b += align_up(n * 4, 256);                        // rank
    b += 2 * align_up(n * 8, 256);                    // key buffers A, B
    b += 2 * align_up(n * 4, 256);                    // value buffers
    b += 2 * align_up(n * 4, 256);                    // position buffers
    b += align_up(n * 4, 256);                        // group heads
    b += align_up(packed_words(n, 8) * 4, 256);       // packed text (worst case 8 bits)
    b += align_up(ceil_div(n, sa::RG_TILE) * 8 + 256, 256);  // regroup status
    b += rsort::workspace_bytes(n);
    b += 4096;                                        // counters, lut, histogram scratch
    return b;
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
    int32_t *rank = c.take<int32_t>(n);
    uint64_t *keyA = c.take<uint64_t>(n);
    uint64_t *keyB = c.take<uint64_t>(n);
    uint32_t *val0 = c.take<uint32_t>(n);
    uint32_t *val1 = c.take<uint32_t>(n);
    int32_t *pos0 = c.take<int32_t>(n);
    int32_t *pos1 = c.take<int32_t>(n);
    int32_t *grp = c.take<int32_t>(n);
    uint32_t *packed = c.take<uint32_t>(packed_words(n, 8));
    int64_t rg_tiles_max = ceil_div(n, sa::RG_TILE);
    unsigned long long *rg_status = c.take<unsigned long long>(rg_tiles_max + 8);
    rsort::Workspace rws = rsort::carve(c, n);
    unsigned *counters = c.take<unsigned>(16);  // [0] regroup tile counter, [1] active count
    uint8_t *d_lut = c.take<uint8_t>(256);
    unsigned long long *d_hist = c.take<unsigned long long>(256);
    if (!c.ok()) {
        set_error("sa workspace carve overflow");
        return BWTK_EWORKSPACE;
    }
    if (d_isa_out) rank = d_isa_out;

    // alphabet -> bits per symbol
    int64_t totals[256];
    int rc = byte_histogram(d_text, n, totals, d_hist, st);
    if (rc) return rc;
    uint8_t lut[256];
    bool fast;
    // ACGT$ layout: '$' shares code 0 with 'A'; it is the unique last symbol, so
    // every suffix whose first S symbols reach past it is made a singleton in
    // round 0, and the one ending exactly on it is ordered by the past-the-end
    // key (0) in round 1.
    int bits = choose_packing(d_text, n, totals, lut, &fast, st);
    if (bits < 0) { set_error("choose_packing failed"); return BWTK_ECUDA; }
    const int S = 32 / bits;
    rc = pack_text(d_text, n, lut, bits, packed, d_lut, st);
    if (rc) return rc;

    BWTK_CUDA(cudaMemsetAsync(rws.err, 0, sizeof(int), st));
    int64_t passes = 0, sum_active = 0, rounds = 0;
    int in_first = 1;
    uint32_t *key32a = (uint32_t *)keyA, *key32b = (uint32_t *)keyA + n;
    if (n == 1) {
        BWTK_CUDA(cudaMemsetAsync(d_sa, 0, 4, st));
        BWTK_CUDA(cudaMemsetAsync(rank, 0, 4, st));
        BWTK_CUDA(cudaStreamSynchronize(st));
        return BWTK_OK;
    }
    {
        int thr = 256;
        { prof::Scope ps("init_keys_kernel", n * 8 + n * bits / 8, st);
        sa::init_keys_kernel<<<(unsigned)ceil_div(n, thr), thr, 0, st>>>(packed, n, bits, S, key32a, val0); }
        BWTK_LAUNCH_CHECK();
        rc = rsort::sort_pairs<uint32_t>(key32a, val0, key32b, val1, n, 0, S * bits, rws, st, &in_first,
                                         &passes);
        if (rc) return rc;
    }
    uint32_t *skey32 = in_first ? key32a : key32b;
    uint32_t *sval = in_first ? val0 : val1;
    int32_t *pos_in = pos0, *pos_out = pos1;
    uint32_t *suf_other = in_first ? val1 : val0;  // free value buffer receives the active suffixes

    unsigned h_count = 0;
    {
        int64_t tiles = ceil_div(n, sa::RG_TILE);
        BWTK_CUDA(cudaMemsetAsync(rg_status, 0, (size_t)tiles * 8, st));
        BWTK_CUDA(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned), st));
        { prof::Scope ps("regroup_first", n * 16, st);
        sa::regroup_kernel<uint32_t, true><<<(unsigned)tiles, sa::RG_THREADS, 0, st>>>(
            skey32, sval, nullptr, n, n - S + 1, d_sa, rank, pos_out, suf_other, grp, rg_status,
            counters, counters + 1, rws.err); }
        BWTK_LAUNCH_CHECK();
        BWTK_CUDA(cudaMemcpyAsync(&h_count, counters + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        BWTK_CUDA(cudaStreamSynchronize(st));
    }
    int64_t active0 = h_count;
    rounds = 1;
    // after round 0 the active list lives in (pos_out, suf_other, grp)
    uint32_t *suf_in = suf_other;
    uint32_t *suf_free = sval;
    { int32_t *t = pos_in; pos_in = pos_out; pos_out = t; }
    const int kbits = sa::bits_for(n);
    const int gbits = sa::bits_for(n - 1);
    int64_t h = S;
    while (h_count > 0) {
        int64_t m = h_count;
        sum_active += m;
        rounds++;
        { prof::Scope ps("build_keys_kernel", m * 20, st);
        sa::build_keys_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, st>>>(suf_in, grp, rank, m, n, h,
                                                                         kbits, keyA); }
        BWTK_LAUNCH_CHECK();
        rc = rsort::sort_pairs<uint64_t>(keyA, suf_in, keyB, suf_free, m, 0, kbits + gbits, rws, st,
                                         &in_first, &passes);
        if (rc) return rc;
        uint64_t *sk = in_first ? keyA : keyB;
        uint32_t *ss = in_first ? suf_in : suf_free;
        uint32_t *sn = in_first ? suf_free : suf_in;  // the other value buffer takes the next list
        int64_t tiles = ceil_div(m, sa::RG_TILE);
        BWTK_CUDA(cudaMemsetAsync(rg_status, 0, (size_t)tiles * 8, st));
        BWTK_CUDA(cudaMemsetAsync(counters, 0, 2 * sizeof(unsigned), st));
        { prof::Scope ps("regroup_round", m * 36, st);
        sa::regroup_kernel<uint64_t, false><<<(unsigned)tiles, sa::RG_THREADS, 0, st>>>(
            sk, ss, pos_in, m, 0, d_sa, rank, pos_out, sn, grp, rg_status, counters, counters + 1,
            rws.err); }
        BWTK_LAUNCH_CHECK();
        BWTK_CUDA(cudaMemcpyAsync(&h_count, counters + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        BWTK_CUDA(cudaStreamSynchronize(st));
        suf_in = sn;
        suf_free = ss;
        { int32_t *t = pos_in; pos_in = pos_out; pos_out = t; }
        h <<= 1;
        if (h > 2 * n && h_count > 0) {
            set_error("suffix array refinement did not converge");
            return BWTK_EINTERNAL;
        }
    }
    int h_err = 0;
    BWTK_CUDA(cudaMemcpyAsync(&h_err, rws.err, sizeof(int), cudaMemcpyDeviceToHost, st));
    BWTK_CUDA(cudaStreamSynchronize(st));
    if (h_err) {
        set_error("look-back spin limit hit (code %d)", h_err);
        return BWTK_EINTERNAL;
    }
    if (h_stats) {
        h_stats[0] = rounds; h_stats[1] = bits; h_stats[2] = S; h_stats[3] = active0;
        h_stats[4] = sum_active; h_stats[5] = passes; h_stats[6] = fast ? 1 : 0;
    }
    return BWTK_OK;
}
