"""GPU parity of the detector kernels
against the CPU oracle: identical integer record rows, in the same order."""
import numpy as np
import pytest

from tests.util import gen_contig, text_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def detect():
    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import detect as d

    return d


def _scan_texts():
    rng = np.random.default_rng(5)
    out = [(nm, tx) for nm, tx in text_cases() if len(tx) >= 2 and nm not in ("binaryish",)]
    out.append(("planted_6000", gen_contig(6000, 8).tobytes() + b"$"))
    s = gen_contig(9000, 12)
    s[3000:3400] = ord("N")
    s[5000:5090] = np.frombuffer(b"ACGTTGCATG" * 9, np.uint8)
    out.append(("planted_9000_N", s.tobytes() + b"$"))
    out.append(("low_complexity", bytes(b"AC"[x] for x in rng.integers(0, 2, 4000)) + b"$"))
    out.append(("homopolymers", b"".join(bytes([b"ACGT"[i % 4]]) * int(l) for i, l in enumerate(rng.integers(1, 40, 300))) + b"$"))
    return out


@pytest.mark.parametrize("name,text", _scan_texts(), ids=[c[0] for c in _scan_texts()])
def test_tier1_rows(detect, oracle, name, text):
    for t in (text, text.rstrip(b"$")):
        if not t:
            continue
        want = oracle.tier1_scan(t)
        got = detect.tier1_rows(t)
        assert np.array_equal(got, want), f"{name}: tier1 rows differ ({len(got)} vs {len(want)})"
    want = oracle.tier1_scan(text, max_motif_length=12, min_copies=2, min_array_length=4, min_entropy=0.5)
    got = detect.tier1_rows(text, max_motif_len=12, min_copies=2, min_array_len=4, min_entropy=0.5)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name,text", _scan_texts(), ids=[c[0] for c in _scan_texts()])
def test_strict_rows_worker_call(detect, oracle, name, text):
    n = len(text.rstrip(b"$"))
    eff = max(120, min(n // 3, 1000))
    want = oracle.strict_scan(text, 1, eff, 0, 3)
    got = detect.strict_rows(text, 1, eff, 0, 3)
    assert got.shape == want.shape, f"{name}: {got.shape} vs {want.shape}"
    assert np.array_equal(got, want), f"{name}: strict rows differ"


@pytest.mark.parametrize("params", [(20, 120, 0, 3), (2, 40, 0, 4), (1, 9, 0, 2), (5, 300, 0, 5)])
def test_strict_rows_other_params(detect, oracle, params):
    text = gen_contig(12000, 21).tobytes() + b"CAGGT" * 70 + b"ACGTTGCAAT" * 40 + b"$"
    want = oracle.strict_scan(text, *params)
    got = detect.strict_rows(text, *params)
    assert np.array_equal(got, want)


def test_strict_rows_on_one_long_chain_of_runs(detect, oracle):
    """Two motif variants that differ in one base, alternating every few copies: every switch costs the unit-3 match
    run one position, so all candidate runs of that unit are linked (each starts within a unit of the end of the run
    before) and every entry depends on the emission before it (the reference's `i = end` jump).  The replay walks
    the chain once from its first run: rows equal the oracle's, and 6 Mb of the same text -- one chain of 10^6 runs,
    on which a walk back from every run would be quadratic -- finishes in a moment."""
    import time

    def chain_text(n, reps):
        blk = b"AAC" * reps + b"AAG" * reps
        return (blk * (n // len(blk) + 1))[:n] + b"$"

    for reps, params in ((2, (1, 12, 0, 2)), (3, (1, 9, 0, 3)), (2, (2, 7, 0, 2)), (4, (1, 30, 0, 3))):
        text = chain_text(30_000, reps)
        want = oracle.strict_scan(text, *params)
        got = detect.strict_rows(text, *params)
        assert np.array_equal(got, want), (reps, params)
        assert len(want) > 1000
    big = chain_text(6_000_000, 2)
    detect.strict_rows(big, 2, 12, 0, 2)
    t0 = time.perf_counter()
    rows = detect.strict_rows(big, 2, 12, 0, 2)
    assert time.perf_counter() - t0 < 2.0 and len(rows) > 900_000


def test_strict_rows_150k(detect, oracle):
    text = gen_contig(150_000, 42).tobytes() + b"$"
    want = oracle.strict_scan(text, 1, 1000, 0, 3)
    got = detect.strict_rows(text, 1, 1000, 0, 3)
    assert np.array_equal(got, want)
    assert len(got) > 5000


@pytest.mark.parametrize("name,text", _scan_texts(), ids=[c[0] for c in _scan_texts()])
def test_plateau_rows(detect, oracle, name, text):
    sa = oracle.suffix_array(text)
    lcp = oracle.kasai_lcp(text, sa)
    for min_period in (1, 10):
        want, wthr = oracle.lcp_plateaus(text, sa, lcp, min_period=min_period)
        got, gthr = detect.plateau_rows(text, sa, lcp, min_period=min_period)
        assert gthr == wthr
        assert np.array_equal(got, want), f"{name}: plateau rows differ (min_period={min_period})"


def test_tier1_sampling_regimes(detect, oracle):
    # n just above 5 M (step 20) exercises the adaptive position_step replay
    s = gen_contig(5_000_400, 33)
    want = oracle.tier1_scan(s)
    got = detect.tier1_rows(s)
    assert np.array_equal(got, want)
    assert len(got) > 100


def _ext_texts():
    s = gen_contig(8000, 17, sub_rate=0.06)
    s[2500:2600] = ord("N")
    return [("planted_8000", s.tobytes() + b"$"),
            ("imperfect_runs", (b"ACGTAC" * 20 + b"ACGAAC" + b"ACGTAC" * 9 + b"TTTTTTTTTT" + b"CAG" * 30 + b"CAA" + b"CAG" * 8) * 3 + b"$"),
            ("AC_period2", b"AC" * 700 + b"A$")]


@pytest.mark.parametrize("name,text", _ext_texts(), ids=[c[0] for c in _ext_texts()])
def test_extend_and_consensus_batches(detect, oracle, name, text):
    rng = np.random.default_rng(len(text))
    n_total = len(text)
    n = n_total - 1
    seeds, periods, flags = [], [], []
    for _ in range(600):
        p = int(rng.integers(1, 40))
        if n < 3 * p:
            continue
        seeds.append(int(rng.integers(0, n - 2 * p)))
        periods.append(p)
        flags.append(int(rng.integers(0, 2)))
    got = detect.extend_batch(text, n, seeds, periods, flags, mode=0)
    for i, (s0, p, f) in enumerate(zip(seeds, periods, flags)):
        want = oracle.extend_with_mismatches(text, s0, p, n, bool(f))
        assert tuple(int(x) for x in got[i, :5]) == want, (name, s0, p, f)
    got = detect.extend_batch(text, n_total, seeds, periods, None, mode=1)
    for i, (s0, p) in enumerate(zip(seeds, periods)):
        want = oracle.extend_tandem_fm(text, s0, p)
        assert tuple(int(x) for x in got[i, :3]) == want, (name, s0, p)
    copies = [int(x) for x in rng.integers(1, 12, len(seeds))]
    cons, mm = detect.consensus_batch(text, seeds, periods, copies)
    for i, (s0, p, c) in enumerate(zip(seeds, periods, copies)):
        wc, wrate, wmax = oracle.consensus(text, s0, p, c)
        assert np.array_equal(cons[i][: wc.size], wc) and (wc.size == 0 or cons[i].size == wc.size)
        used = int(mm[i, 2])
        rate = (mm[i, 0] / (used * p)) if used else 0.0
        assert rate == wrate and int(mm[i, 1]) == wmax


@pytest.mark.parametrize("n,seed", [(600, 3), (1200, 4), (3000, 1), (9000, 12), (20_000, 5), (120_000, 9)])
def test_period_scan_rows(detect, oracle, n, seed):
    s = gen_contig(n, seed, sub_rate=0.04)
    if n >= 9000:
        s[n // 2: n // 2 + 300] = ord("N")
    text = s.tobytes() + b"$"
    want, wit = oracle.period_scan(text)
    got, git = detect.period_scan_rows(text)
    assert git == wit
    assert np.array_equal(got, want), f"n={n}: {len(got)} vs {len(want)} rows"
    # with a Tier 1 mask and without mismatches
    t1 = oracle.tier1_scan(text)
    mask = np.zeros(len(text) - 1, np.uint8)
    for r in t1:
        mask[r[0]:min(r[1], mask.size)] = 1
    want, wit = oracle.period_scan(text, allow_mismatches=False, tier1_mask=mask, min_period=2, max_period=60)
    got, git = detect.period_scan_rows(text, allow_mismatches=False, tier1_mask=mask, min_period=2, max_period=60)
    assert git == wit and np.array_equal(got, want)


@pytest.mark.parametrize("params", [(20, 120, 2, 3), (2, 40, 1, 4), (3, 60, 3, 3)])
def test_strict_rows_with_mismatches(detect, oracle, params):
    # Hamming-tolerant adjacency (the API default max_mismatch=2), greedy-replay engine
    s = gen_contig(9000, 31, sub_rate=0.05)
    text = s.tobytes() + b"ACGTTGCAATGGCCATTGCA" * 12 + b"ACGTTGCAATGGCCATTGCT" * 3 + b"$"
    want = oracle.strict_scan(text, *params)
    got = detect.strict_rows(text, *params)
    assert np.array_equal(got, want), f"{len(got)} vs {len(want)}"
    assert len(want) > 0


def test_pinned_row_pool_reuses_any_large_enough_block(detect):
    """Detector rows of >= 8 MB land in pinned blocks that return to the pool when the array and its
    views are gone; a smaller download then reuses the larger block instead of allocating its class."""
    import gc

    import torch

    pool = detect._PinnedPool()
    dev = torch.device("cuda", torch.cuda.current_device())
    big = torch.arange(3_000_000 * 8, dtype=torch.int32, device=dev).reshape(-1, 8)
    small = big[:400_000] * 3
    a = pool.lend(torch, big)
    assert a.shape == (3_000_000, 8) and a.dtype == np.int32 and np.array_equal(a, big.cpu().numpy())
    base_a = a.__array_interface__["data"][0]
    b = pool.lend(torch, small)                     # `a` is alive: a second block
    assert b.__array_interface__["data"][0] != base_a and np.array_equal(b, small.cpu().numpy())
    view = a[10:20, :4]
    del a
    gc.collect()
    assert len(pool._free) == 0                     # a view keeps the block on loan
    assert np.array_equal(view, big[10:20, :4].cpu().numpy())
    del view
    gc.collect()
    assert len(pool._free) == 1
    c = pool.lend(torch, small)                     # served by the big block
    assert c.__array_interface__["data"][0] == base_a and np.array_equal(c, small.cpu().numpy())
    assert np.array_equal(b, small.cpu().numpy())   # the other loan is untouched
    del b, c
    gc.collect()
    assert len(pool._free) == 2
    # the public path goes through the module's pool
    rows = detect._rows_to_host(torch, big)
    assert np.array_equal(rows, big.cpu().numpy())


# ---- n-dependent regimes (bwt.py:1441-1447 Tier 1 position_step; 2193-2228 period scan max_p / step clamps) ----
def _boundary_sizes():
    out = []
    for thr in (10_000, 100_000, 1_000_000, 5_000_000, 10_000_000):
        out += [thr - 1, thr, thr + 1, thr + 2]      # the text's n with and without '$' lands on both sides
    return out


@pytest.mark.parametrize("n", _boundary_sizes())
def test_tier1_and_period_scan_at_regime_boundaries(detect, oracle, n):
    """Just below / at / above every n threshold of the reference, with the sentinel (n + 1 symbols) and without."""
    body = gen_contig(n, 1000 + n % 977)
    for text in (body.tobytes() + b"$", body.tobytes()):
        want = oracle.tier1_scan(text)
        got, seen = detect.tier1_rows(text, want_seen=True)
        assert np.array_equal(got, want), f"n={len(text)}: tier 1 rows differ"
        want_p, want_it = oracle.period_scan(text)
        got_p, got_it = detect.period_scan_rows(text)
        assert got_it == want_it and np.array_equal(got_p, want_p), f"n={len(text)}: period scan differs"
        mask = np.zeros(len(text), np.uint8)
        for a, b in want[:, :2].tolist():
            mask[a:b] = 1
        m = mask[: len(text.rstrip(b"$"))]
        # the kernel's seen mask is the union of the Tier 1 calls: what GenomeScanner hands to the period scan
        assert np.array_equal(seen[: m.size] != 0, m != 0), f"n={len(text)}: seen mask != union of the Tier 1 rows"
        want_m, it_m = oracle.period_scan(text, tier1_mask=m)
        got_m, git_m = detect.period_scan_rows(text, tier1_mask=m)
        assert git_m == it_m and np.array_equal(got_m, want_m), f"n={len(text)}: masked period scan differs"


def test_strict_and_tier1_on_an_unaligned_text_view(oracle):
    """The ordered small-unit run finder needs 16-byte aligned text; a view that starts at an odd address
    takes the per-position kernel + sort instead.  Both must give the oracle's rows."""
    import torch

    from bwt_algorithm_b200 import detect

    text = gen_contig(120_000, 11).tobytes() + b"$"
    buf = torch.zeros(len(text) + 64, dtype=torch.uint8, device="cuda")
    for off in (0, 1, 7, 16):
        view = buf[off:off + len(text)]
        view.copy_(torch.from_numpy(np.frombuffer(text, np.uint8).copy()))
        assert view.data_ptr() % 16 == (buf.data_ptr() + off) % 16
        rows = detect.strict_rows(view, 1, 1000, 0, 3)
        assert np.array_equal(rows, oracle.strict_scan(text, 1, 1000, 0, 3)), f"strict rows differ at offset {off}"
        t1 = detect.tier1_rows(view)
        assert np.array_equal(t1, oracle.tier1_scan(text)), f"tier 1 rows differ at offset {off}"


def test_tier1_homopolymer_floor_only_with_positive_entropy(oracle):
    """min_entropy = 0 keeps short homopolymer arrays (the run list then holds every run of >= 2)."""
    from bwt_algorithm_b200 import detect

    rng = np.random.default_rng(5)
    s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 60_000)].copy()
    for k in range(40):
        s[1000 * k + 10:1000 * k + 10 + 3 + k % 12] = ord("ACGT"[k % 4])
    text = s.tobytes() + b"$"
    for ent in (0.0, 0.5, 1.0):
        got = detect.tier1_rows(text, 9, 3, 6, ent)
        assert np.array_equal(got, oracle.tier1_scan(text, 9, 3, 6, ent)), f"min_entropy={ent}"


def test_strict_scan_with_the_index_hint_gives_the_same_rows(oracle):
    """bwtk_strict_scan_hinted: unit lengths >= 16 only look at aligned groups whose 16-mer occurs twice (bitmap from
    SA + LCP).  Same rows as the plain scan and the oracle: planted contig, long repeats (every group active), an N
    block (bytes that alias in the 2-bit codes), period-2 text, a hint of fewer symbols, and a hint that is too long
    to be a filter (ignored)."""
    import torch

    from bwt_algorithm_b200 import detect
    from bwt_algorithm_b200.device_index import DeviceIndex

    rng = np.random.default_rng(3)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    s1 = gen_contig(150_000, 42)
    s2 = gen_contig(60_000, 8)
    s2[10_000:14_000] = ord("N")
    s2[30_000:36_000] = np.tile(acgt[rng.integers(0, 4, 300)], 20)          # unit 300 x 20
    s2[40_000:41_000] = np.tile(acgt[rng.integers(0, 4, 17)], 59)[:1000]      # unit 17
    s3 = np.tile(np.frombuffer(b"AC", np.uint8), 6000)
    s4 = acgt[rng.integers(0, 4, 40_000)].copy()
    s4[5_000:5_000 + 3 * 997] = np.tile(acgt[rng.integers(0, 4, 997)], 3)     # unit 997, three copies
    for name, s in (("planted_150k", s1), ("N_and_long_units", s2), ("period2", s3), ("unit_997", s4)):
        text = s.tobytes() + b"$"
        ix = DeviceIndex(text, build_lcp=True)
        n = len(text)
        want = oracle.strict_scan(text, 1, 1000, 0, 3)
        plain = detect.strict_rows(ix.text, 1, 1000, 0, 3)
        assert np.array_equal(plain, want), name
        for hl in (16, 12, 20):
            hint = detect.repeat_hint(ix.sa, ix.lcp, n, hl)
            got = detect.strict_rows(ix.text, 1, 1000, 0, 3, hint=hint, hint_len=hl)
            assert np.array_equal(got, want), f"{name}: hinted rows differ (hint_len={hl})"
        # the bitmap itself: bit i <=> the 16-mer at i occurs again
        hint = detect.repeat_hint(ix.sa, ix.lcp, n, 16).cpu().numpy().view(np.uint32)
        bits = ((hint[np.arange(n) >> 5] >> (np.arange(n) & 31)) & 1).astype(bool)
        lcp, sa = ix.lcp.cpu().numpy(), ix.sa.cpu().numpy()
        ref = np.zeros(n, bool)
        j = np.flatnonzero(lcp[1:] >= 16) + 1
        ref[sa[j]] = True
        ref[sa[j - 1]] = True
        assert np.array_equal(bits, ref), name
    del torch
