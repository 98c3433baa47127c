// rowchain.cu -- row-level form of TandemRepeatFinder._suppress_nested_short_calls (reference
// bwt.py:3402-3497), host code.  The reference walks the calls of a contig sorted by (has
// mismatches, motif length descending) and drops a call when a KEPT call with a strictly longer motif
// overlaps it by more than an adaptive fraction of its own length; every kept call is then tested
// against all later ones (quadratic).  A contig yields millions of integer rows, so the same predicate
// runs here over the rows themselves: kept spans are filed under the 256-bp buckets they touch and a
// row is only tested against the spans of its own buckets.  Calls of one motif length never affect
// each other (only strictly longer motifs suppress), so the result does not depend on their order.
#include <stdint.h>

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
