// fasta.cu -- FASTA ingest (host code): TandemRepeatFinder.load_reference, reference bwt.py:3713-3756.
//
// The reference reads the file line by line in text mode: every line is stripped of leading and
// trailing whitespace; a stripped line starting with '>' opens a record whose name is the first
// whitespace-separated token after '>'; every other non-empty stripped line is upper-cased and
// appended to the current record (interior whitespace stays); lines before the first header are
// dropped.  Here the same rules run over the file's bytes in two passes: bwtk_fasta_index finds
// the records and their sequence lengths, bwtk_fasta_sequence writes one record's sequence --
// upper-cased, optionally flank-trimmed and '$'-terminated -- straight into the caller's buffer
// (pinned host memory, so the upload needs no further copy).  Line ends are '\n', '\r\n' and a
// lone '\r' (Python's universal newlines).  Only ASCII files take this path: str.upper() and
// str.strip() have non-ASCII cases (the index call reports them and the caller falls back to the
// reference's own loop).
#include <stdint.h>
#include <string.h>

#include "../../include/bwtk.h"

namespace {

// str.strip() / str.split() whitespace, ASCII part: \t \n \v \f \r, FS GS RS US, space
inline bool is_space(uint8_t c) { return (c >= 9 && c <= 13) || (c >= 28 && c <= 32); }

struct Line {
    int64_t a, b;      // stripped content [a, b)
    int64_t next;      // start of the next line
};

inline Line next_line(const uint8_t *buf, int64_t pos, int64_t len)
{
    int64_t e = pos;
    while (e < len && buf[e] != '\n' && buf[e] != '\r') e++;
    Line ln;
    ln.next = e;
    if (e < len) ln.next = (buf[e] == '\r' && e + 1 < len && buf[e + 1] == '\n') ? e + 2 : e + 1;
    int64_t a = pos, b = e;
    while (a < b && is_space(buf[a])) a++;
    while (b > a && is_space(buf[b - 1])) b--;
    ln.a = a;
    ln.b = b;
    return ln;
}

}  // namespace

// rec rows: name_off, name_len, body_off (first byte after the header line), body_end (start of the
// next header line or len), seq_len (symbols the record's lines contribute).  Returns BWTK_EINVAL for a
// header without a name (the reference raises IndexError there) and BWTK_EOVERFLOW (count = records
// needed) when cap is too small; *h_flags bit 0 = the file holds non-ASCII bytes.
extern "C" int32_t bwtk_fasta_index(const uint8_t *buf, int64_t len, int64_t *h_rec, int64_t cap, int64_t *h_count,
                                    int32_t *h_flags)
{
    if (!buf || !h_count || !h_flags || (cap > 0 && !h_rec)) return BWTK_EINVAL;
    int64_t count = 0, pos = 0, cur = -1;
    int32_t flags = 0;
    for (int64_t i = 0; i < len; i++)
        if (buf[i] & 0x80) { flags |= 1; break; }
    while (pos < len) {
        const Line ln = next_line(buf, pos, len);
        if (ln.b > ln.a && buf[ln.a] == '>') {
            int64_t na = ln.a + 1;
            while (na < ln.b && is_space(buf[na])) na++;
            int64_t nb = na;
            while (nb < ln.b && !is_space(buf[nb])) nb++;
            if (nb == na) return BWTK_EINVAL;
            if (cur >= 0 && cur < cap) h_rec[cur * 5 + 3] = pos;
            cur = count++;
            if (cur < cap) {
                h_rec[cur * 5 + 0] = na;
                h_rec[cur * 5 + 1] = nb - na;
                h_rec[cur * 5 + 2] = ln.next;
                h_rec[cur * 5 + 3] = len;
                h_rec[cur * 5 + 4] = 0;
            }
        } else if (cur >= 0 && cur < cap) {
            h_rec[cur * 5 + 4] += ln.b - ln.a;
        }
        pos = ln.next;
    }
    *h_count = count;
    *h_flags = flags;
    return count > cap ? BWTK_EOVERFLOW : BWTK_OK;
}

// Writes symbols [skip, skip + take) of the record whose lines lie in buf[body_off, body_end), upper-cased,
// to dst, followed by '$' when `sentinel` is non-zero.  Returns the number of bytes written or a negative code.
extern "C" int64_t bwtk_fasta_sequence(const uint8_t *buf, int64_t body_off, int64_t body_end, int64_t skip,
                                       int64_t take, int32_t sentinel, uint8_t *dst)
{
    if (!buf || !dst || body_off < 0 || body_end < body_off || skip < 0 || take < 0) return BWTK_EINVAL;
    int64_t pos = body_off, seen = 0, out = 0;
    while (pos < body_end && out < take) {
        const Line ln = next_line(buf, pos, body_end);
        int64_t a = ln.a, b = ln.b;
        const int64_t l = b - a;
        if (l > 0) {
            if (seen + l > skip) {
                if (seen < skip) a += skip - seen;
                int64_t m = b - a;
                if (m > take - out) m = take - out;
                const uint8_t *src = buf + a;
                uint8_t *d = dst + out;
                for (int64_t i = 0; i < m; i++) {
                    const uint8_t c = src[i];
                    d[i] = (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c;
                }
                out += m;
            }
            seen += l;
        }
        pos = ln.next;
    }
    if (out != take) return BWTK_EINVAL;
    if (sentinel) dst[out++] = '$';
    return out;
}
