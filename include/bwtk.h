/*
 * bwtk.h -- C ABI of libbwtk.so: the sm_100a kernels under the `bwt` module.
 *
 * The reference (wyim-pgl/bwt-algorithm, bwt.py) is pure Python and has no
 * FFI; the drop-in boundary is its Python module surface.
 * This header is the new seam UNDER those Python names: every entry point
 * below names the reference method (file:line) whose arithmetic it replaces.
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would
 * add to call them.
 *
 * Conventions
 *  - every pointer named d_* is a DEVICE pointer owned by the caller (the
 *    Python side allocates with torch); the library never returns memory it
 *    allocated and keeps no global state except a thread-local error string.
 *  - h_* pointers are HOST pointers (small outputs the call synchronises for).
 *  - `stream` is a cudaStream_t passed as void* (0 = default stream).
 *  - return value: 0 = OK, negative = BWTK_E*; on error outputs are undefined
 *    and bwtk_last_error() describes the failure.
 *  - record capacity overflow returns BWTK_EOVERFLOW with the required count
 *    in the count slot; results are deterministic, so the caller re-allocates
 *    and calls again.
 *  - n < 2^30 for every text (int32 suffix arrays as in the reference).
 */
#ifndef BWTK_H
#define BWTK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BWTK_OK 0
#define BWTK_EINVAL (-1)
#define BWTK_ECUDA (-2)
#define BWTK_EWORKSPACE (-3)
#define BWTK_EOVERFLOW (-4)
#define BWTK_EINTERNAL (-5)

#define BWTK_REC_W 8 /* int32 per record row: start,end,period,copies,total_mm,max_mm,aux0,aux1 */

int32_t bwtk_version(void);
/* copies the calling thread's last error message (NUL terminated) */
int32_t bwtk_last_error(char *buf, int32_t buflen);
/* number of kernels this library has launched in the calling process */
int64_t bwtk_launch_count(void);

/* Optional per-kernel timing for bench.py: when enabled, CUDA events are
 * recorded on the launching stream around the main kernels.
 * bwtk_profile_report synchronises the device, writes one
 * "name\tlaunches\ttotal_ms\talgorithmic_bytes\n" line per kernel into buf and
 * clears the records.  (No counterpart in the reference.) */
int32_t bwtk_profile_enable(int32_t on);
int32_t bwtk_profile_report(char *buf, int32_t buflen);

/* ---- text ingest: BWTCore.__init__'s text_arr (bwt.py:122) ---------------
 * Copies n bytes from PINNED host memory (cudaHostAlloc / torch pin_memory,
 * device-accessible under unified addressing) to d_dst with a kernel instead of
 * a copy engine, so that the upload of one contig is never queued behind the
 * bulk downloads of another stream (streaming.IndexPipeline).  Asynchronous. */
int32_t bwtk_upload_text(const uint8_t *h_pinned, uint8_t *d_dst, int64_t n, void *stream);
/* The other direction (suffix_array / bwt_arr / LCP as host arrays, bwt.py:127-128, 55-72):
 * `bytes` from device memory into PINNED host memory by a kernel that stores straight
 * over PCIe, for the same reason.  Both pointers 16-byte aligned.  Asynchronous. */
int32_t bwtk_download(const void *d_src, void *h_pinned, int64_t bytes, void *stream);

/* ---- device memory for hosts without a CUDA allocator of their own ---------------------------------
 * The entry points of this header take raw device pointers from any allocator (torch, cudaMalloc, a pool).
 * A host that has none -- the drop-in CLI answering `python bwt.py small.fa` (bwt.py:4201-4370) before an
 * `import torch` would have finished, or a C caller -- gets buffers here.  bwtk_device_count: visible CUDA
 * devices (0 without a driver; never an error).  bwtk_dev_alloc / bwtk_dev_free: cudaMalloc / cudaFree on the
 * current device (256-byte aligned).  bwtk_copy_to_device / bwtk_copy_to_host: pageable or pinned host memory,
 * on `stream`, synchronous (they return when the bytes have arrived). */
int32_t bwtk_device_count(int32_t *count);
int32_t bwtk_dev_alloc(void **d_ptr, int64_t bytes);
int32_t bwtk_dev_free(void *d_ptr);
int32_t bwtk_copy_to_device(void *d_dst, const void *h_src, int64_t bytes, void *stream);
int32_t bwtk_copy_to_host(void *h_dst, const void *d_src, int64_t bytes, void *stream);

/* ---- FASTA ingest: TandemRepeatFinder.load_reference (bwt.py:3713-3756), host code -----
 * Two passes over the file's bytes with the reference's line rules (strip, '>' headers, first
 * token = name, upper-case, lines before the first header dropped; line ends \n, \r\n, \r).
 * bwtk_fasta_index: h_rec[count][5] = name_off, name_len, body_off, body_end, seq_len;
 * BWTK_EOVERFLOW with the record count when cap is too small; BWTK_EINVAL for a header without
 * a name; *h_flags bit 0 = non-ASCII bytes present (caller must use the reference's own loop).
 * bwtk_fasta_sequence: symbols [skip, skip+take) of one record, upper-cased, into dst (e.g.
 * pinned host memory), '$' appended when sentinel != 0; returns bytes written or < 0. */
int32_t bwtk_fasta_index(const uint8_t *buf, int64_t len, int64_t *h_rec, int64_t cap, int64_t *h_count,
                         int32_t *h_flags);
int64_t bwtk_fasta_sequence(const uint8_t *buf, int64_t body_off, int64_t body_end, int64_t skip,
                            int64_t take, int32_t sentinel, uint8_t *dst);

/* ---- post-processing on rows: TandemRepeatFinder._suppress_nested_short_calls (bwt.py:3402-3497),
 * host code.  One entry per call of ONE contig (start, end, len(motif), mismatch_rate > 0);
 * keep[i] = 1 when the call survives: not overlapped by a KEPT call with a strictly longer motif
 * for >= 80 % of its length (homopolymers), >= 10 % / 30 % (motif >= 10x / 5x longer) or
 * >= overlap_threshold otherwise.  Same survivors as the reference's quadratic loop. */
int32_t bwtk_suppress_nested(const int32_t *start, const int32_t *end, const int32_t *motif_len,
                             const uint8_t *imperfect, int64_t n, double overlap_threshold, uint8_t *keep);

/* ---- post-processing: MotifUtils.align_repeat_region (bwt.py:997-1102) with its per-copy aligner
 * MotifUtils._align_unit_to_window (bwt.py:828-995), host code, on the bytes of one contig.  The walk
 * starts at `start` with the template `motif` (k bytes), takes runs of exact copies in one step, aligns every
 * other copy by the banded unit-cost DP (ties substitution > deletion > insertion; end column = first
 * minimum in [k - max_indel, k + max_indel]; rejected above `tol` substitutions or `max_indel` inserted /
 * deleted bases) and re-derives the consensus (per column the first maximum in first-seen order) after it.
 * tol and max_indel are the values the reference derives from mismatch_fraction / its default rule.
 * out = int64[8]: copies, consumed length, sum of per-copy errors, max errors per copy, inserted bases,
 * deleted bases, number of variation notes, bytes of notes.  consensus: k bytes.  notes: the reference's
 * `variations` strings ("copy:pos:X>Y", "copy:pos:ins(S)", "copy:pos:del(n)") joined by '\n'.
 * Returns 1 (summary written), 0 (the reference returns None), BWTK_EOVERFLOW (notes_cap too small, out[7]
 * holds the size), BWTK_EWORKSPACE (a column saw more than 8 distinct symbols: use the generic walk). */
int32_t bwtk_align_repeat_region(const uint8_t *seq, int64_t total, int64_t start, int64_t end,
                                 const uint8_t *motif, int32_t k, int32_t tol, int32_t max_indel,
                                 int32_t min_copies, uint8_t *consensus, int64_t *out, uint8_t *notes,
                                 int64_t notes_cap);

/* ---- a5: BWTCore._build_char_counts (bwt.py:276-286) ------------------
 * byte histogram of the text; h_totals[256] (host) receives the counts.
 * The exclusive prefix sum over present bytes (the FM "C" array) is a 256-entry
 * host loop done by the caller.  Synchronises the stream. */
int32_t bwtk_byte_histogram(const uint8_t *d_text, int64_t n, int64_t *h_totals, void *stream);

/* ---- a3: BWTCore._build_suffix_array (bwt.py:212-264) -----------------
 * suffix array under byte order, shorter suffix first (key2 = -1 rule).
 * d_isa_out may be NULL; otherwise receives the inverse suffix array.
 * h_stats (may be NULL) receives int64[8]: rounds, bits/symbol, symbols/key,
 * active after round 0, sum of active over rounds, radix passes, path flags, 0.
 * Path flags: bit 0 = ACGT$ 2-bit layout, bit 1 = round 0 by the MSD bucket sort
 * (2-bit texts from 2^24 symbols on -- below that the four LSD passes are faster; environment BWTK_MSD=0 keeps the LSD sort,
 * BWTK_MSD_MIN_N=<n> moves the threshold), bit 2 = first regroup pass fused into it
 * (BWTK_MSD_FUSE=0 switches the fusion off).  The suffix array does not depend on the path.
 * Synchronises the stream (one 4-byte read-back per doubling round). */
int64_t bwtk_sa_workspace_bytes(int64_t n);
int32_t bwtk_sa_build(const uint8_t *d_text, int64_t n, int32_t *d_sa, int32_t *d_isa_out,
                      void *d_ws, int64_t ws_bytes, int64_t *h_stats, void *stream);

/* ---- a1: BWTCore.__init__ index build (bwt.py:106-136), fused ------------
 * One call = byte histogram (a5) + suffix array (a3, + inverse) + BWT and Occ
 * checkpoints (a4, a6) + LCP (a10), sharing one histogram and one bit-packed
 * text.  d_isa, d_bwt/d_occ and d_lcp may be NULL to skip that output.
 * h_totals[256] / h_row_of_code[256] (host) receive the byte counts and the Occ
 * row of every byte value (-1 if absent; rows are in byte order).  d_occ must
 * hold occ_rows_cap rows of ncp = n/occ_rate + 1 + (n%occ_rate != 0) int32; if
 * the text has more distinct bytes the call returns BWTK_EOVERFLOW with the
 * row count in h_stats[7].  h_stats as in bwtk_sa_build (+ [7] = Occ rows).
 * Host syncs: the alphabet read-back and one per batch of doubling rounds; the
 * BWT/Occ/LCP kernels are left in flight on the stream.  d_sa and d_ws must be
 * 16-byte aligned (every cudaMalloc'ed pointer is); BWTK_EINVAL otherwise. */
int64_t bwtk_index_workspace_bytes(int64_t n);
int32_t bwtk_index_build(const uint8_t *d_text, int64_t n, int32_t occ_rate, int32_t *d_sa, int32_t *d_isa,
                         uint8_t *d_bwt, int32_t *d_occ, int32_t occ_rows_cap, int32_t *d_lcp,
                         int64_t *h_totals, int32_t *h_row_of_code, int64_t *h_stats, void *d_ws,
                         int64_t ws_bytes, void *stream);

/* ---- a4+a6: _build_bwt_array + _build_occurrence_checkpoints ----------
 * (bwt.py:266-274, 288-326).  d_bwt[i] = text[(sa[i]-1) mod n].
 * h_row_of_code[256]: row index in d_occ for each byte value, -1 if the byte
 * does not occur.  d_occ is [nrows][ncp] int32 row-major with
 * ncp = n/occ_rate + 1 + (n%occ_rate != 0): cp[0]=0, cp[m+1]=#code in
 * bwt[0:(m+1)*occ_rate], plus the total when the last block is partial.
 * d_ws needs bwtk_bwt_occ_workspace_bytes(n, occ_rate, nrows) bytes. */
int64_t bwtk_bwt_occ_workspace_bytes(int64_t n, int32_t occ_rate, int32_t nrows);
int32_t bwtk_bwt_occ(const uint8_t *d_text, const int32_t *d_sa, int64_t n, int32_t occ_rate,
                     const int32_t *h_row_of_code, int32_t nrows, uint8_t *d_bwt, int32_t *d_occ,
                     void *d_ws, int64_t ws_bytes, void *stream);

/* ---- a10: _kasai_lcp_uint8 (bwt.py:55-72) ------------------------------
 * lcp[0]=0, lcp[r]=LCP(text[sa[r-1]:], text[sa[r]:]). */
int64_t bwtk_lcp_workspace_bytes(int64_t n);
int32_t bwtk_lcp_build(const uint8_t *d_text, const int32_t *d_sa, int64_t n, int32_t *d_lcp,
                       void *d_ws, int64_t ws_bytes, void *stream);

/* ---- a2: _build_kmer_hash(k=8) (bwt.py:138-171) -------------------------
 * d_bucket_off[65537]: start of each 16-bit code's bucket in d_pos;
 * d_pos: window start positions, ascending inside a bucket.  *h_count receives
 * the number of recorded windows (0 when n < 8).  Synchronises. */
int64_t bwtk_kmer8_workspace_bytes(int64_t n);
int32_t bwtk_kmer8_index(const uint8_t *d_text, int64_t n, int32_t *d_bucket_off, int32_t *d_pos,
                         int64_t *h_count, void *d_ws, int64_t ws_bytes, void *stream);

/* ---- a8/a9: rank + backward_search (bwt.py:335-389) ---------------------
 * Batched: pattern q is d_pats[q*stride : q*stride+d_lens[q]] (bytes).
 * d_C[256]/d_tot[256]: C array / totals (int64), d_row_of_code[256] int32.
 * Outputs inclusive (sp,ep), (-1,-1) if absent, (0,n-1) for an empty pattern. */
int32_t bwtk_bsearch_batch(const uint8_t *d_bwt, const int32_t *d_occ, int64_t ncp, int32_t occ_rate,
                           const int64_t *d_C, const int64_t *d_tot, const int32_t *d_row_of_code,
                           int64_t n, const uint8_t *d_pats, int64_t stride, const int32_t *d_lens,
                           int64_t nq, int32_t *d_sp, int32_t *d_ep, void *stream);
/* Trie sweep over every ACGT motif of length 1..kmax (kmax <= 12).  Motif m of
 * length k with base-4 value v (first character most significant) is stored
 * at index (4^k - 4)/3 + v.  One LF step per motif. */
int32_t bwtk_bsearch_motif_sweep(const uint8_t *d_bwt, const int32_t *d_occ, int64_t ncp,
                                 int32_t occ_rate, const int64_t *d_C, const int64_t *d_tot,
                                 const int32_t *d_row_of_code, int64_t n, int32_t kmax,
                                 int32_t *d_sp, int32_t *d_ep, void *stream);
/* rank(code,pos) probes (bwt.py:335-357) */
int32_t bwtk_rank_batch(const uint8_t *d_bwt, const int32_t *d_occ, int64_t ncp, int32_t occ_rate,
                        const int32_t *d_row_of_code, int64_t n, const int32_t *d_codes,
                        const int64_t *d_pos, int64_t nq, int64_t *d_out, void *stream);

/* ---- a8/a9 on a packed index: rank / backward_search (bwt.py:335-389) ---------
 * The search side of the FM index as 64-byte rank blocks: block b covers BWT
 * positions [192 b, 192 b + 192) and holds #A, #C, #G, #T before the block
 * (4 x u32) followed by the 192 symbols at 2 bits each, so one rank is one
 * random access (counts + first 64 symbols in the first 32-B sector).  Bytes
 * other than A/C/G/T are "exceptions": stored as code 0 and listed by position
 * (d_exc_pos, ascending) and grouped by byte value (d_exc_by_code with group
 * offsets d_code_off[257]).  Answers are identical to bwtk_bsearch_batch /
 * bwtk_rank_batch / bwtk_bsearch_motif_sweep over the byte BWT + Occ rows.
 *
 * bwtk_fm_pack builds the arrays from the BWT bytes.  h_totals[256]: byte counts
 * (as returned by bwtk_index_build); the number of exceptions is
 * n - totals[A,C,G,T] and is written to *h_exc_count (BWTK_EOVERFLOW if it
 * exceeds exc_cap).  d_blocks: bwtk_fm_pack_bytes(n) bytes, 16-byte aligned.
 * Synchronises the stream. */
typedef struct bwtk_fm_index {
    const void *d_blocks;          /* bwtk_fm_pack_bytes(n) bytes */
    int64_t n;                     /* BWT length */
    const uint32_t *d_exc_pos;     /* [n_exc] */
    const uint32_t *d_exc_by_code; /* [n_exc] */
    int64_t n_exc;
    const int64_t *d_code_off;     /* [257] */
    const int64_t *d_C;            /* [256] C array */
    const int64_t *d_tot;          /* [256] byte totals */
    int64_t acgt_C[4];             /* host copies of C / totals of A, C, G, T */
    int64_t acgt_tot[4];
    const int32_t *d_ftab_sp;      /* optional: intervals of all 4^ftab_k ACGT k-mers, index = base-4 */
    const int32_t *d_ftab_ep;      /*   value, first character most significant (a level of the motif sweep) */
    int32_t ftab_k;                /* 0 = no table */
} bwtk_fm_index;
#define BWTK_FM_L2_PERSIST 1       /* flags: keep the rank blocks resident in L2 (access policy window) */
#define BWTK_FM_THREAD_PER_QUERY 2 /* flags: one thread per query (default: four lanes share every block load) */
int64_t bwtk_fm_pack_bytes(int64_t n);
int64_t bwtk_fm_pack_workspace_bytes(int64_t n, int64_t n_exc);
int32_t bwtk_fm_pack(const uint8_t *d_bwt, int64_t n, const int64_t *h_totals, void *d_blocks,
                     uint32_t *d_exc_pos, uint32_t *d_exc_by_code, int64_t exc_cap, int64_t *d_code_off,
                     int64_t *h_exc_count, void *d_ws, int64_t ws_bytes, void *stream);
/* as bwtk_bsearch_batch; with a k-mer table, patterns of length >= ftab_k whose last ftab_k
 * characters are ACGT start from that k-mer's interval (same answers, fewer LF steps) */
int32_t bwtk_fm_search_batch(const bwtk_fm_index *ix, const uint8_t *d_pats, int64_t stride,
                             const int32_t *d_lens, int64_t nq, int32_t *d_sp, int32_t *d_ep,
                             int32_t flags, void *stream);
/* as bwtk_rank_batch */
int32_t bwtk_fm_rank_batch(const bwtk_fm_index *ix, const int32_t *d_codes, const int64_t *d_pos,
                           int64_t nq, int64_t *d_out, void *stream);
/* as bwtk_bsearch_motif_sweep: both blocks of a parent motif serve its four children */
int32_t bwtk_fm_motif_sweep(const bwtk_fm_index *ix, int32_t kmax, int32_t *d_sp, int32_t *d_ep,
                            int32_t flags, void *stream);

/* ---- a12: Tier1STRFinder._find_simple_tandems_kmer (bwt.py:1426-1531) ---
 * rows: start,end,motif_len,copies,0,0,0,0 in the reference's emission order.
 * d_seen_out (may be NULL): n bytes, the final seen mask. */
int64_t bwtk_tier1_workspace_bytes(int64_t n);
int32_t bwtk_tier1_scan(const uint8_t *d_text, int64_t n, int32_t max_motif_len, int32_t min_copies,
                        int32_t min_array_len, double min_entropy, int32_t *d_rec, int64_t cap,
                        int64_t *h_count, uint8_t *d_seen_out, void *d_ws, int64_t ws_bytes,
                        void *stream);

/* ---- a13: Tier2LCPFinder.find_long_unit_repeats_strict (bwt.py:1891-2001)
 * rows: start,end,primitive_period,copies,0,0,unit_len,0, ordered by unit
 * length descending then start ascending (the reference's order). */
int64_t bwtk_strict_workspace_bytes(int64_t n, int64_t max_unit_len);
int32_t bwtk_strict_scan(const uint8_t *d_text, int64_t n, int64_t min_unit_len, int64_t max_unit_len,
                         int64_t max_mismatch, int64_t min_copies, int32_t *d_rec, int64_t cap,
                         int64_t *h_count, void *d_ws, int64_t ws_bytes, void *stream);
/* The same scan with a hint from the index of the same text: bit i of d_hint_bits set iff suffix i shares
 * >= hint_len symbols with a neighbour in suffix order (bwtk_repeat_hint over SA + LCP; bwtk_repeat_hint_bytes(n)
 * bytes).  With hint_len <= 16 the unit lengths >= 16 -- 98 % of the scan's compares -- only look at the aligned
 * 16-position groups whose 16-mer occurs more than once (6 % of them on planted random sequence).  Same rows.
 * d_hint_bits == NULL: plain bwtk_strict_scan.  Text and hint must be 16-byte aligned for the fast path. */
int64_t bwtk_repeat_hint_bytes(int64_t n);
int32_t bwtk_repeat_hint(const int32_t *d_sa, const int32_t *d_lcp, int64_t n, int32_t min_len,
                         uint32_t *d_bits, void *stream);
int32_t bwtk_strict_scan_hinted(const uint8_t *d_text, int64_t n, int64_t min_unit_len, int64_t max_unit_len,
                                int64_t max_mismatch, int64_t min_copies, int32_t *d_rec, int64_t cap,
                                int64_t *h_count, const uint32_t *d_hint_bits, int32_t hint_len, void *d_ws,
                                int64_t ws_bytes, void *stream);

/* ---- a15: _detect_lcp_plateaus (bwt.py:2118-2145, 2500-2560) ------------
 * rows: start,end,period(=threshold),copies,0,0,0,0; *h_threshold = -1 when
 * no scan happens (max LCP < min_period). */
int64_t bwtk_plateau_workspace_bytes(int64_t n);
int32_t bwtk_lcp_plateaus(const uint8_t *d_text, int64_t n_text, const int32_t *d_sa,
                          const int32_t *d_lcp, int64_t n, int64_t min_period, int64_t max_period,
                          int64_t min_copies, int32_t *d_rec, int64_t cap, int64_t *h_count,
                          int64_t *h_threshold, void *d_ws, int64_t ws_bytes, void *stream);

/* ---- a16/a14/a17: extension + consensus batches --------------------------
 * mode 0: _extend_with_mismatches (bwt.py:2392-2498), out row int32[8]:
 *         array_start,array_end,copies,full_start,full_end,0,0,0;
 *         d_flags[i] bit0 = allow_mismatches.
 * mode 1: _extend_tandem_fm (bwt.py:2697-2805), out row: start,end,copies,... */
int32_t bwtk_extend_batch(const uint8_t *d_text, int64_t n, const int32_t *d_seed,
                          const int32_t *d_period, const int32_t *d_flags, int64_t m, int32_t mode,
                          int32_t *d_out, void *stream);
/* MotifUtils.build_consensus_motif_array (bwt.py:1207-1256): consensus bytes
 * for item i at d_cons[d_cons_off[i] : +period]; d_mm row int32[4]:
 * total_mm, max_mm_per_copy, copies_used, 0. */
int32_t bwtk_consensus_batch(const uint8_t *d_text, int64_t text_size, const int32_t *d_start,
                             const int32_t *d_period, const int32_t *d_copies,
                             const int64_t *d_cons_off, int64_t m, uint8_t *d_cons, int32_t *d_mm,
                             void *stream);

/* ---- a16: _find_repeats_simple (bwt.py:2177-2390), clock frozen ----------
 * rows: array_start,array_end,p_eff,copies_full,total_mm,max_mm,cons_start,
 * copies_used, before the host-side (start,end,canonical) dedup.
 * d_plogp: (max_p+1)^2 doubles, plogp[c*(max_p+1)+L] = (c/L)*log2(c/L) as the
 * host computes it (keeps the entropy gate bit-identical to numpy's). */
int32_t bwtk_period_scan(const uint8_t *d_text, int64_t n, int64_t min_period, int64_t max_period,
                         int32_t allow_mismatches, int64_t min_copies, int64_t min_array_len,
                         double min_entropy, const uint8_t *d_tier1_mask, const double *d_plogp,
                         int64_t plogp_dim, int32_t *d_rec, int64_t cap, int64_t *h_count,
                         int64_t *h_iterations, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BWTK_H */
