// rowchain.cu -- row-level form of TandemRepeatFinder._suppress_nested_short_calls (reference
// bwt.py:3402-3497), host code.  The reference walks the calls of a contig sorted by (has
// mismatches, motif length descending) and drops a call when a KEPT call with a strictly longer motif
// overlaps it by more than an adaptive fraction of its own length; every kept call is then tested
// against all later ones (quadratic).  A contig yields millions of integer rows, so the same predicate
// runs here over the rows themselves: kept spans are filed under the 256-bp buckets they touch and a
// row is only tested against the spans of its own buckets.  Calls of one motif length never affect
// each other (only strictly longer motifs suppress), so the result does not depend on their order.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "../../include/bwtk.h"

namespace {
struct Span {
    int32_t start, end, k;
};
}  // namespace

// start/end/motif_len/imperfect: one entry per call of ONE contig (imperfect[i] != 0: mismatch_rate > 0).
// keep[i] = 1 when the call survives.  overlap_threshold is the reference's default cut (0.5).
extern "C" int32_t bwtk_suppress_nested(const int32_t *start, const int32_t *end, const int32_t *motif_len,
                                        const uint8_t *imperfect, int64_t n, double overlap_threshold, uint8_t *keep)
{
    if (n == 0) return BWTK_OK;
    if (!start || !end || !motif_len || !keep || n < 0) return BWTK_EINVAL;
    std::vector<int64_t> order((size_t)n);
    std::iota(order.begin(), order.end(), (int64_t)0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        const int ia = imperfect && imperfect[a] ? 1 : 0, ib = imperfect && imperfect[b] ? 1 : 0;
        if (ia != ib) return ia < ib;
        return motif_len[a] > motif_len[b];
    });
    int32_t max_end = 0;
    for (int64_t i = 0; i < n; i++) max_end = std::max(max_end, std::max(end[i], start[i] + 1));
    const int64_t nb = ((int64_t)max_end >> 8) + 2;
    // bucket lists as singly linked chains in arrival order (head/tail per bucket, next per entry)
    std::vector<int32_t> head((size_t)nb, -1), tail((size_t)nb, -1);
    std::vector<Span> spans;
    std::vector<int32_t> next, span_of;
    spans.reserve((size_t)n / 4 + 16);
    for (int64_t oi = 0; oi < n; oi++) {
        const int64_t i = order[(size_t)oi];
        const int32_t rs = start[i], re = end[i], k = motif_len[i];
        const double span = (double)(re - rs);
        const int64_t first = (int64_t)std::max(rs, 0) >> 8;
        const int64_t last = (int64_t)std::max(std::max(re - 1, rs), 0) >> 8;
        bool nested = false;
        for (int64_t b = first; b <= last && !nested; b++) {
            for (int32_t e = head[(size_t)b]; e >= 0; e = next[(size_t)e]) {
                const Span &o = spans[(size_t)span_of[(size_t)e]];
                if (o.k <= k) continue;
                const int32_t ov = std::min(re, o.end) - std::max(rs, o.start);
                if (ov <= 0) continue;
                if (k == 1 && o.k > 1 && (double)ov / span >= 0.8) { nested = true; break; }
                const double ratio = (double)o.k / (double)k;
                const double cut = ratio >= 10 ? 0.1 : ratio >= 5 ? 0.3 : overlap_threshold;
                if ((double)ov / span >= cut) { nested = true; break; }
            }
        }
        keep[i] = nested ? 0 : 1;
        if (!nested) {
            const int32_t si = (int32_t)spans.size();
            spans.push_back(Span{rs, re, k});
            for (int64_t b = first; b <= last; b++) {
                const int32_t e = (int32_t)next.size();
                next.push_back(-1);
                span_of.push_back(si);
                if (tail[(size_t)b] >= 0) next[(size_t)tail[(size_t)b]] = e; else head[(size_t)b] = e;
                tail[(size_t)b] = e;
            }
        }
    }
    return BWTK_OK;
}

// ---- MotifUtils.align_repeat_region on the bytes of one contig (reference bwt.py:828-1102), host code ------
// The merge / refine stages of the chain re-derive a call by walking it copy by copy: every copy is aligned
// against the running consensus by a banded unit-cost DP (ties: substitution > deletion > insertion, end
// column = first minimum in [k - max_indel, k + max_indel]), the per-column tallies are updated, and the
// consensus (first maximum in first-seen order per column) becomes the next template.  A chr21-sized contig
// asks for ~75 k such walks; this is the same walk over the contig's bytes, answer for answer.
namespace {
struct Tally {                 // one consensus column: symbols in first-seen order
    static const int CAP = 8;
    uint8_t sym[CAP];
    int64_t cnt[CAP];
    int n;
};
inline bool tally_add(Tally &t, uint8_t s, int64_t by)
{
    for (int i = 0; i < t.n; i++)
        if (t.sym[i] == s) { t.cnt[i] += by; return true; }
    if (t.n == Tally::CAP) return false;
    t.sym[t.n] = s; t.cnt[t.n] = by; t.n++;
    return true;
}
inline void consensus_of(const std::vector<Tally> &tal, std::vector<uint8_t> &tmpl)
{
    for (size_t c = 0; c < tal.size(); c++) {
        const Tally &t = tal[c];
        if (!t.n) continue;                       // empty column: the template's own base stays
        int best = 0;
        for (int i = 1; i < t.n; i++)
            if (t.cnt[i] > t.cnt[best]) best = i;
        tmpl[c] = t.sym[best];
    }
}
struct Notes {
    uint8_t *buf;
    int64_t cap, len;
    bool first;
    void put(const char *s, size_t n)
    {
        if (buf && len + (int64_t)n <= cap) memcpy(buf + len, s, n);
        len += (int64_t)n;
    }
    void begin()
    {
        if (!first) put("\n", 1);
        first = false;
    }
};
struct AlignScratch {
    std::vector<int32_t> cost;
    std::vector<uint8_t> move;
    std::vector<int32_t> col_r, col_q;            // traceback columns: motif index / window index, -1 = gap
};
}  // namespace

// out (int64[8]): copies, consumed length, sum of per-copy errors, max errors per copy, inserted bases,
// deleted bases, number of variation notes, bytes of notes.  consensus: k bytes.  notes: the `variations`
// strings joined by '\n' (NULL / too small: out[7] holds the size needed and BWTK_EOVERFLOW is returned).
// Returns 1 with a summary, 0 for the reference's None, BWTK_EWORKSPACE when a consensus column sees more
// than 8 distinct symbols (the caller then takes the Python walk), BWTK_EINVAL on bad arguments.
extern "C" int32_t bwtk_align_repeat_region(const uint8_t *seq, int64_t total, int64_t start, int64_t end,
                                            const uint8_t *motif, int32_t k, int32_t tol, int32_t max_indel,
                                            int32_t min_copies, uint8_t *consensus, int64_t *out, uint8_t *notes,
                                            int64_t notes_cap)
{
    if (!seq || !motif || !consensus || !out || k < 0 || total < 0 || max_indel < 0) return BWTK_EINVAL;
    for (int i = 0; i < 8; i++) out[i] = 0;
    if (k == 0 || total == 0) return 0;
    if (start < 0) start = 0;
    end = end > start ? end : total;
    if (end > total) end = total;
    if (tol < 0) tol = 0;
    static thread_local AlignScratch sc;
    std::vector<Tally> tal((size_t)k);
    for (auto &t : tal) t.n = 0;
    std::vector<uint8_t> tmpl(motif, motif + k);
    Notes nt{notes, notes ? notes_cap : 0, 0, true};
    int64_t n_notes = 0, units = 0, sum_err = 0, max_err = 0, n_ins = 0, n_del = 0;
    int64_t pos = start;
    int64_t stop = std::max(end, start + (int64_t)k * min_copies) + std::max((int64_t)k * 3, (int64_t)max_indel * 4);
    if (stop > total) stop = total;
    char tmp[64];
    while (pos < stop) {
        if (pos + k <= total && memcmp(seq + pos, tmpl.data(), (size_t)k) == 0) {
            int64_t run = 1, nxt = pos + k;
            while (nxt < stop && nxt + k <= total && memcmp(seq + nxt, tmpl.data(), (size_t)k) == 0) { run++; nxt += k; }
            units += run;
            for (int c = 0; c < k; c++)
                if (!tally_add(tal[(size_t)c], tmpl[(size_t)c], run)) return BWTK_EWORKSPACE;
            pos = nxt;
            continue;
        }
        const int64_t wend = std::min(total, pos + k + max_indel);
        const int32_t n = (int32_t)(wend - pos), m = k;
        if (n < k - max_indel) break;
        const uint8_t *win = seq + pos;
        // ---- _align_unit_to_window (bwt.py:828-995)
        if (n == 0) break;
        const int32_t lo = std::max(0, m - max_indel), hi = std::min(n, m + max_indel);
        if (lo > hi) break;
        const int32_t big = m + n + 10, band = max_indel + 2, W = n + 1;
        if (sc.cost.size() < (size_t)(m + 1) * W) { sc.cost.resize((size_t)(m + 1) * W); sc.move.resize((size_t)(m + 1) * W); }
        int32_t *cost = sc.cost.data();
        uint8_t *move = sc.move.data();
        // cells outside the band keep (big, 0); only those next to the band are ever read
        for (int32_t i = 1; i <= m; i++) {
            const int32_t a = std::max(0, i - band - 1), b = std::min(n, i + band + 1);
            for (int32_t j = a; j <= b; j++) { cost[(size_t)i * W + j] = big; move[(size_t)i * W + j] = 0; }
            cost[(size_t)i * W] = i; move[(size_t)i * W] = 2;
        }
        for (int32_t j = 0; j <= n; j++) { cost[j] = j; move[j] = 3; }
        move[0] = 0;
        for (int32_t i = 1; i <= m; i++) {
            int32_t *ci = cost + (size_t)i * W;
            const int32_t *cp = cost + (size_t)(i - 1) * W;
            uint8_t *mi = move + (size_t)i * W;
            const uint8_t a = tmpl[(size_t)i - 1];
            const int32_t j1 = std::min(n, i + band);
            for (int32_t j = std::max(1, i - band); j <= j1; j++) {
                int32_t best = cp[j - 1] + (a != win[j - 1]);
                uint8_t mv = 1;
                int32_t d = cp[j] + 1;
                if (d < best) { best = d; mv = 2; }
                d = ci[j - 1] + 1;
                if (d < best) { best = d; mv = 3; }
                ci[j] = best; mi[j] = mv;
            }
        }
        int32_t end_j = -1, end_cost = big;
        for (int32_t j = lo; j <= hi; j++)
            if (cost[(size_t)m * W + j] < end_cost) { end_cost = cost[(size_t)m * W + j]; end_j = j; }
        if (end_j <= 0 || end_cost >= big) break;
        sc.col_r.clear(); sc.col_q.clear();
        {
            int32_t i = m, j = end_j;
            while (i > 0 || j > 0) {
                const uint8_t mv = move[(size_t)i * W + j];
                if (mv == 1) { sc.col_r.push_back(i - 1); sc.col_q.push_back(j - 1); i--; j--; }
                else if (mv == 2) { sc.col_r.push_back(i - 1); sc.col_q.push_back(-1); i--; }
                else if (mv == 3) { sc.col_r.push_back(-1); sc.col_q.push_back(j - 1); j--; }
                else break;
            }
        }
        // pass 1: the totals decide whether the copy is accepted at all
        int32_t subs = 0, ins_total = 0, del_total = 0;
        for (size_t c = sc.col_r.size(); c-- > 0;) {
            const int32_t r = sc.col_r[c], q = sc.col_q[c];
            if (r < 0) ins_total++;
            else if (q < 0) del_total++;
            else if (tmpl[(size_t)r] != win[q]) subs++;
        }
        if (subs > tol || ins_total > max_indel || del_total > max_indel) break;
        if (end_j == 0) break;
        // pass 2: variation notes in the reference's order, tallies of the aligned bases
        const int64_t copy_no = units + 1;
        {
            int32_t ref_pos = 0, ins_at = 0, ins_from = -1, ins_len = 0, del_run = 0, del_at = 0;
            auto flush_ins = [&]() {
                if (!ins_len) return;
                nt.begin();
                nt.put(tmp, (size_t)snprintf(tmp, sizeof tmp, "%lld:%d:ins(", (long long)copy_no, ins_at));
                nt.put(reinterpret_cast<const char *>(win + ins_from), (size_t)ins_len);
                nt.put(")", 1);
                n_notes++;
                ins_len = 0; ins_from = -1; ins_at = 0;
            };
            auto flush_del = [&]() {
                if (!del_run) return;
                nt.begin();
                nt.put(tmp, (size_t)snprintf(tmp, sizeof tmp, "%lld:%d:del(%d)", (long long)copy_no, del_at, del_run));
                n_notes++;
                del_run = 0;
            };
            for (size_t c = sc.col_r.size(); c-- > 0;) {
                const int32_t r = sc.col_r[c], q = sc.col_q[c];
                if (r < 0) {
                    if (!ins_len) { ins_at = ref_pos; ins_from = q; }
                    ins_len++;
                    continue;
                }
                flush_ins();
                ref_pos++;
                if (q < 0) {
                    if (!del_run) del_at = ref_pos;
                    del_run++;
                    continue;
                }
                flush_del();
                if (tmpl[(size_t)r] != win[q]) {
                    nt.begin();
                    nt.put(tmp, (size_t)snprintf(tmp, sizeof tmp, "%lld:%d:%c>%c", (long long)copy_no, ref_pos,
                                                 (char)tmpl[(size_t)r], (char)win[q]));
                    n_notes++;
                }
                if (ref_pos - 1 < k && !tally_add(tal[(size_t)ref_pos - 1], win[q], 1)) return BWTK_EWORKSPACE;
            }
            flush_ins();
            flush_del();
        }
        const int64_t err = (int64_t)subs + ins_total + del_total;
        units++;
        sum_err += err;
        if (err > max_err) max_err = err;
        n_ins += ins_total;
        n_del += del_total;
        pos += end_j;
        consensus_of(tal, tmpl);
    }
    if (units < min_copies || pos - start <= 0) return 0;
    consensus_of(tal, tmpl);
    memcpy(consensus, tmpl.data(), (size_t)k);
    out[0] = units; out[1] = pos - start; out[2] = sum_err; out[3] = max_err; out[4] = n_ins; out[5] = n_del;
    out[6] = n_notes; out[7] = nt.len;
    if (nt.len > nt.cap) return BWTK_EOVERFLOW;
    return 1;
}
