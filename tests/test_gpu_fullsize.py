"""Full-size runs (BASELINE.json configs[2] and configs[3]): a chr21-sized and a
chr1-sized synthetic contig, checked through size-independent properties --
the CPU oracle needs minutes for a 46.7 Mb suffix array, so only its cheap
scans are compared row for row here:

  * SA is a permutation and ISA its inverse;
  * LCP[i] is exactly the common prefix of suffixes SA[i-1], SA[i] (rolling
    hash of both prefixes, mod 2^64) and the character after it orders them,
    which proves the suffix order;
  * BWT[i] == text[SA[i]-1]; every Occ checkpoint == a prefix count of the BWT;
  * the 8-mer table is the stable bucket sort of the valid windows;
  * the FM interval of every ACGT motif of length 1..10 has the size of that
    motif's occurrence count, and its ends carry the motif;
  * Tier 1 and period-scan rows == oracle rows on the whole contig; strict
    adjacency rows verify against the text and agree with the scan of a prefix.

torch is used here only as the checker.
"""
import numpy as np
import pytest

from tests.util import gen_contig

pytestmark = pytest.mark.gpu

CHR21 = 46_709_983
CHR1 = 248_956_422


@pytest.fixture(scope="module")
def torch():
    import torch as t

    import bwt_algorithm_b200  # noqa: F401

    return t


def _powers(torch, n, base, dev):
    """base**j (mod 2^64) for j in [0, n] by square-and-multiply over the bits of j."""
    j = torch.arange(n + 1, dtype=torch.int64, device=dev)
    p = torch.ones(n + 1, dtype=torch.int64, device=dev)
    b = torch.tensor(base, dtype=torch.int64, device=dev)
    k = 0
    while (1 << k) <= n:
        sel = ((j >> k) & 1).bool()
        p = torch.where(sel, p * b, p)
        b = b * b
        k += 1
    return p


def _check_suffix_order_and_lcp(torch, text_t, sa, lcp):
    n = text_t.numel()
    dev = text_t.device
    B = 0x9E3779B97F4A7C15 - (1 << 64)          # odd multiplier, as a signed int64
    P = _powers(torch, n, B, dev)
    H = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    H[1:] = torch.cumsum((text_t.long() + 1) * P[:n], 0)
    a, b, l = sa[:-1].long(), sa[1:].long(), lcp[1:].long()
    assert int(lcp[0]) == 0
    assert bool((l >= 0).all()) and bool((a + l <= n).all()) and bool((b + l <= n).all())
    seg_a = (H[a + l] - H[a]) * P[b]
    seg_b = (H[b + l] - H[b]) * P[a]
    assert bool((seg_a == seg_b).all()), "LCP overstates a common prefix"
    del seg_a, seg_b, H, P
    pad = torch.cat([text_t.to(torch.int16), torch.full((1,), -1, dtype=torch.int16, device=dev)])
    ca, cb = pad[a + l], pad[b + l]
    assert bool((ca < cb).all()), "suffix order / LCP maximality violated"


def _check_index_properties(torch, ix, text_t):
    n = ix.n
    dev = text_t.device
    sa = ix.sa
    assert bool((torch.bincount(sa.long(), minlength=n) == 1).all()), "SA is not a permutation"
    ar = torch.arange(n, dtype=torch.int32, device=dev)
    assert torch.equal(ix.isa[sa.long()], ar), "ISA is not the inverse of SA"
    _check_suffix_order_and_lcp(torch, text_t, sa, ix.lcp)
    prev = sa.long() - 1
    prev[prev < 0] = n - 1
    assert torch.equal(ix.bwt, text_t[prev]), "BWT != text[SA-1]"
    hist = torch.bincount(text_t.long(), minlength=256).cpu().numpy()
    assert np.array_equal(hist, ix.totals)
    at = torch.clamp(torch.arange(ix.ncp, dtype=torch.int64, device=dev) * ix.occ_rate, max=n)
    for code in ix.codes:
        pref = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        pref[1:] = torch.cumsum((ix.bwt == code).to(torch.int32), 0, dtype=torch.int32)
        assert torch.equal(ix.occ[ix.row_of_code[code], : ix.ncp], pref[at]), f"Occ row of byte {code}"


def _kmer_codes(torch, text_t, k):
    """(codes, valid) of every window of k bases (first base most significant)."""
    n = text_t.numel()
    lut = torch.full((256,), -1, dtype=torch.int64, device=text_t.device)
    for v, ch in enumerate(b"ACGT"):
        lut[ch] = v
    sym = lut[text_t.long()]
    m = n - k + 1
    code = torch.zeros(m, dtype=torch.int64, device=text_t.device)
    valid = torch.ones(m, dtype=torch.bool, device=text_t.device)
    for j in range(k):
        s = sym[j:j + m]
        valid &= s >= 0
        code = code * 4 + torch.clamp(s, min=0)
    return code, valid


def _check_kmer_table(torch, ix, text_t):
    code, valid = _kmer_codes(torch, text_t, 8)
    pos = torch.nonzero(valid).flatten()
    c = code[pos]
    order = torch.sort(c, stable=True).indices
    assert ix.kmer_count == pos.numel()
    assert torch.equal(ix.kmer_pos.long(), pos[order])
    off = torch.zeros(65537, dtype=torch.int64, device=text_t.device)
    off[1:] = torch.cumsum(torch.bincount(c, minlength=65536), 0)
    assert torch.equal(ix.kmer_off.long(), off)


def _check_motif_sweep(torch, ix, text_t, kmax=10):
    sp, ep = ix.motif_sweep(kmax)
    sa = ix.sa.long()
    for k in range(1, kmax + 1):
        code, valid = _kmer_codes(torch, text_t, k)
        want = torch.bincount(code[valid], minlength=4 ** k)
        base = (4 ** k - 4) // 3
        s, e = sp[base:base + 4 ** k].long(), ep[base:base + 4 ** k].long()
        got = torch.where(s >= 0, e - s + 1, torch.zeros_like(s))
        assert torch.equal(got, want), f"interval sizes differ for k={k}"
        present = torch.nonzero(s >= 0).flatten()
        for end in (s, e):                                  # both ends of the interval carry the motif
            at = sa[end[present]]
            assert bool((at + k <= text_t.numel()).all())
            assert torch.equal(code[at], present) and bool(valid[at].all())


@pytest.fixture(scope="module")
def chr21(torch):
    from bwt_algorithm_b200.device_index import DeviceIndex

    s = gen_contig(CHR21, 2024)
    text = np.concatenate([s, np.frombuffer(b"$", np.uint8)])
    ix = DeviceIndex(text, build_isa=True, build_lcp=True)
    return text, ix


def test_chr21_index_properties(torch, chr21):
    text, ix = chr21
    assert ix.n == CHR21 + 1 and (ix.sa_stats[6] & 1) == 1
    _check_index_properties(torch, ix, ix.text)
    torch.cuda.empty_cache()


def test_chr21_kmer_table_and_motif_sweep(torch, chr21):
    text, ix = chr21
    _check_kmer_table(torch, ix, ix.text)
    _check_motif_sweep(torch, ix, ix.text)
    torch.cuda.empty_cache()


def test_chr21_tier1_rows_equal_oracle(torch, chr21, oracle):
    from bwt_algorithm_b200 import detect

    text, ix = chr21
    tb = text.tobytes()
    want = oracle.tier1_scan(tb)
    got = detect.tier1_rows(tb)
    assert len(want) > 1000 and np.array_equal(got, want)


def test_chr21_period_scan_equal_oracle(torch, chr21, oracle):
    from bwt_algorithm_b200 import detect

    text, ix = chr21
    tb = text.tobytes()
    want, wit = oracle.period_scan(tb)
    got, git = detect.period_scan_rows(tb)
    assert git == wit and np.array_equal(got, want)


def test_chr21_plateaus_equal_oracle_on_device_arrays(torch, chr21, oracle):
    """The SA/LCP were proven above; the oracle's plateau pass runs on them (its own SA takes minutes)."""
    from bwt_algorithm_b200 import detect

    text, ix = chr21
    tb = text.tobytes()
    sa, lcp = ix.sa.cpu().numpy(), ix.lcp.cpu().numpy()
    for min_period in (1, 10):
        want, wthr = oracle.lcp_plateaus(tb, sa, lcp, min_period=min_period)
        got, gthr = detect.plateau_rows(tb, sa, lcp, min_period=min_period)
        assert gthr == wthr and np.array_equal(got, want)


def test_chr21_strict_rows_verify_against_text(torch, chr21):
    """Worker parameters (bwt.py:3310-3327): unit 1..1000, exact adjacency, >= 3 copies."""
    from bwt_algorithm_b200 import detect

    text, ix = chr21
    tb = text.tobytes()
    rows = detect.strict_rows(tb, 1, 1000, 0, 3)
    again = detect.strict_rows(tb, 1, 1000, 0, 3)
    assert np.array_equal(rows, again), "scan is not deterministic"
    assert len(rows) > 100_000
    start, end, unit, copies = (rows[:, i].astype(np.int64) for i in range(4))
    assert np.all(start >= 0) and np.all(end <= CHR21) and np.all(end - start >= 3 * unit)
    assert np.all(copies >= 3) and np.all(rows[:, 4] == 0) and np.all(rows[:, 5] == 0)
    # scan order (bwt.py:1921-1999): unit lengths from long to short, left to right within one
    scanned = rows[:, 6].astype(np.int64)
    assert np.all(scanned % unit == 0) and np.all(scanned <= 1000)
    order_key = (1000 - scanned) * (1 << 32) + start
    assert np.all(np.diff(order_key) > 0), "rows are not in scan order"
    t = ix.text
    dev = t.device
    d_start = torch.from_numpy(start).to(dev)
    d_end = torch.from_numpy(end).to(dev)
    d_unit = torch.from_numpy(unit).to(dev)
    for u in np.unique(unit):
        u = int(u)
        sel = torch.nonzero(d_unit == u).flatten()
        bad = torch.zeros(t.numel() - u + 1, dtype=torch.int32, device=dev)
        bad[1:] = torch.cumsum((t[:-u] != t[u:]).to(torch.int32), 0, dtype=torch.int32)
        s, e = d_start[sel], d_end[sel]
        assert bool((bad[e - u] - bad[s] == 0).all()), f"a unit-{u} row is not an exact tandem"
    # the scan is a left-to-right greedy: a prefix scans to the same rows away from its cut
    cut = 400_000
    pre = detect.strict_rows(tb[:cut] + b"$", 1, 1000, 0, 3)
    lim = cut - 4000
    assert np.array_equal(pre[pre[:, 1] < lim], rows[rows[:, 1] < lim])


def _oracle_strict_window(args):
    from oracle import oracle as orc

    window, = args
    return orc.strict_scan(window, 1, 1000, 0, 3)


def test_chr21_strict_rows_equal_oracle_on_32_windows(torch, chr21, oracle):
    """K = 32 windows of 20 kb cut at seeded offsets from the chr21-sized contig, each scanned as a
    standalone contig by the oracle (all host cores) and by the CUDA path -- and the whole-contig rows that lie
    strictly inside a window, away from its cuts, must be the window's rows (the scan is a left-to-right greedy
    per unit length, so a row depends only on the text from its unit's previous emission onwards)."""
    import multiprocessing as mp
    import os

    from bwt_algorithm_b200 import detect

    text, ix = chr21
    W, K = 20_000, 32
    offs = np.sort(np.random.default_rng(2021).integers(0, CHR21 - W, K))
    windows = [text[o:o + W].tobytes() + b"$" for o in offs]
    with mp.get_context("fork").Pool(min(K, os.cpu_count() or 1)) as pool:
        want = pool.map(_oracle_strict_window, [(w,) for w in windows])
    whole = detect.strict_rows(ix.text, 1, 1000, 0, 3)
    n_rows = 0
    for o, w, exp in zip(offs, windows, want):
        got = detect.strict_rows(w, 1, 1000, 0, 3)
        assert np.array_equal(got, exp), f"window at {o}: strict rows differ from the oracle"
        n_rows += len(exp)
        # rows of the whole-contig scan well inside the window == the window's rows there (shifted)
        lo, hi = o + 3000, o + W - 3000
        inside = whole[(whole[:, 0] >= lo) & (whole[:, 1] <= hi)].copy()
        inside[:, 0] -= o
        inside[:, 1] -= o
        sub = exp[(exp[:, 0] >= 3000) & (exp[:, 1] <= W - 3000)]
        assert np.array_equal(inside, sub), f"window at {o}: whole-contig rows differ inside the window"
    assert n_rows > 20_000


def test_chr1_sized_contig_index_and_search(torch):
    """configs[3]: all 1-10 bp motifs on a chr1-sized contig."""
    from bwt_algorithm_b200.device_index import DeviceIndex

    s = gen_contig(CHR1, 77)
    text = np.concatenate([s, np.frombuffer(b"$", np.uint8)])
    del s
    ix = DeviceIndex(text, build_isa=True, build_lcp=True, build_kmer=False)
    assert ix.n == CHR1 + 1
    _check_index_properties(torch, ix, ix.text)
    torch.cuda.empty_cache()
    _check_motif_sweep(torch, ix, ix.text)
