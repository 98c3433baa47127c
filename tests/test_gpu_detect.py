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


@pytest.mark.parametrize("name,text", _scan_texts())
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


@pytest.mark.parametrize("name,text", _scan_texts())
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


def test_strict_rows_150k(detect, oracle):
    text = gen_contig(150_000, 42).tobytes() + b"$"
    want = oracle.strict_scan(text, 1, 1000, 0, 3)
    got = detect.strict_rows(text, 1, 1000, 0, 3)
    assert np.array_equal(got, want)
    assert len(got) > 5000


@pytest.mark.parametrize("name,text", _scan_texts())
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
