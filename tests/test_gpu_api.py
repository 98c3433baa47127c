"""The drop-in module surface on the GPU against the REFERENCE's own outputs.

tests/golden/*.npz hold what the unmodified reference returned for each API
call (oracle/gen_golden.py); here the same calls go through
``bwt_algorithm_b200.bwt`` -> ctypes -> libbwtk.so and must return the same
arrays and the same records (floats within 1e-9, far inside the 1e-6 bar).
"""
import glob
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLDEN]


def _load(path):
    z = np.load(path)
    return z, json.loads(bytes(z["meta"]).decode())


def _cmp_records(got, want, where):
    assert len(got) == len(want), f"{where}: {len(got)} records vs {len(want)} in the reference"
    for i, (g, w) in enumerate(zip(got, want)):
        for key, wv in w.items():
            gv = getattr(g, key)
            tag = f"{where}[{i}].{key}"
            if isinstance(wv, float):
                assert abs(float(gv) - wv) <= 1e-9, f"{tag}: {gv} vs {wv}"
            elif isinstance(wv, dict):
                for k2 in wv:
                    assert abs(float(gv[k2]) - float(wv[k2])) <= 1e-9, f"{tag}[{k2}]"
            elif isinstance(wv, int) and not isinstance(wv, bool):
                assert int(gv) == wv, f"{tag}: {gv} vs {wv}"
            else:
                assert gv == wv, f"{tag}: {gv!r} vs {wv!r}"


@pytest.fixture(scope="module")
def bwt():
    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import bwt as mod

    return mod


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_bwtcore_matches_reference(bwt, path):
    z, meta = _load(path)
    text = bytes(z["text"]).decode()
    core = bwt.BWTCore(text)
    assert core.n == len(text)
    if "sa" in z.files:
        assert core.suffix_array.dtype == np.int32 and np.array_equal(core.suffix_array, z["sa"])
        assert core.bwt_arr.dtype == np.uint8 and np.array_equal(core.bwt_arr, z["bwt"])
        occ = core.occ_checkpoints
        assert sorted(occ) == meta["occ_codes"]
        for r, code in enumerate(meta["occ_codes"]):
            assert occ[code].dtype == np.int32 and np.array_equal(occ[code], z["occ"][r])
        lcp = bwt.Tier2LCPFinder(core)._compute_lcp_array()
        assert lcp.dtype == np.int32 and np.array_equal(lcp, z["lcp"])
        assert np.array_equal(bwt._kasai_lcp_uint8(core.text_arr, core.suffix_array), z["lcp"])
        kc, kp = [], []
        for w in sorted(core.kmer_hash.keys()):
            kc += [w] * len(core.kmer_hash[w])
            kp += core.kmer_hash[w]
        assert kc == z["kmer_codes"].tolist() and kp == z["kmer_pos"].tolist()
    assert core.alphabet == meta["alphabet"]
    assert core.char_counts == meta["char_counts"] and core.char_totals == meta["char_totals"]
    assert all(int(v) == int(core.suffix_array[i]) for i, v in core.sampled_sa.items())
    assert sorted(core.sampled_sa) == list(range(0, core.n, 32))
    pats = meta["patterns"]
    sp, ep = core.backward_search_many([p.encode("utf-8") for p in pats])
    assert [[int(a), int(b)] for a, b in zip(sp, ep)] == meta["search"]
    for i in range(0, len(pats), 9):
        assert list(core.backward_search(pats[i])) == meta["search"][i]
        assert core.count_occurrences(pats[i]) == meta["count"][i]
    for idx, want in zip(meta["locate_idx"], meta["locate"]):
        assert core.locate_positions(pats[idx]) == want
    for ch, pos, want in meta["rank_probe"][:16]:
        assert core.rank(ch, pos) == want
    for q, want in zip(meta["kmer_queries"], meta["kmer_lookup"]):
        assert list(core.get_kmer_positions(q)) == want, f"get_kmer_positions({q!r})"
    if 0 < core.n <= 600:
        for i in range(0, core.n, 7):
            assert int(core._get_suffix_position(i)) == int(core.suffix_array[i])
    core.clear()
    assert core.suffix_array.size == 0 and core.text == "" and core.occ_checkpoints == {}


LEVEL1 = [p for p in GOLDEN if _load(p)[1]["level"] >= 1]


@pytest.mark.parametrize("path", LEVEL1, ids=[os.path.basename(p)[:-4] for p in LEVEL1])
def test_detectors_match_reference(bwt, path):
    z, meta = _load(path)
    text = bytes(z["text"]).decode()
    seq = text[:-1]
    name = meta["name"]
    core = bwt.BWTCore(text)
    s_arr = np.frombuffer(seq.encode(), np.uint8)
    _cmp_records(bwt.Tier1STRFinder(s_arr).find_strs(name), meta["tier1_nosent"], "tier1_nosent")
    _cmp_records(bwt.Tier1STRFinder(core.text_arr).find_strs(name), meta["tier1_sent"], "tier1_sent")
    _cmp_records(bwt.Tier1STRFinder(s_arr, max_motif_length=12).find_strs(name), meta["tier1_m12"], "tier1_m12")
    t2 = bwt.Tier2LCPFinder(core, min_period=1)
    if "strict_worker" in meta:
        eff = max(120, min(len(seq) // 3, 1000))
        _cmp_records(t2.find_long_unit_repeats_strict(name, min_unit_len=1, max_unit_len=eff, max_mismatch=0,
                                                      min_copies=3), meta["strict_worker"], "strict_worker")
        _cmp_records(t2.find_long_unit_repeats_strict(name), meta["strict_default"], "strict_default")
        _cmp_records(t2.find_long_unit_repeats_strict(name, min_unit_len=2, max_unit_len=40, max_mismatch=1,
                                                      min_copies=4), meta["strict_mm1"], "strict_mm1")
    lcp = t2._compute_lcp_array()
    got = t2._detect_lcp_plateaus(lcp, name)
    _cmp_records(got, meta["plateau_p1"], "plateau_p1")
    assert all(isinstance(r.start, np.int32) for r in got)
    _cmp_records(bwt.Tier2LCPFinder(core, min_period=10)._detect_lcp_plateaus(lcp, name), meta["plateau_p10"],
                 "plateau_p10")
    ns = len(seq)
    for s0, p, allow, *want in meta["extend_mm"][:12]:
        assert list(t2._extend_with_mismatches(core.text_arr, s0, p, ns, allow_mismatches=allow)) == want
    for s0, p, *want in meta["extend_fm"][:12]:
        assert list(t2._extend_tandem_fm(core.text_arr, s0, seq[s0:s0 + p], p, 0)) == want
    for s0, p, nc, cons, rate, mx in meta["consensus"][:12]:
        c, r, m = bwt.MotifUtils.build_consensus_motif_array(core.text_arr, s0, p, nc)
        assert c.tobytes().decode("latin1") == cons and r == rate and m == mx
    if meta["level"] >= 2:
        _cmp_records(t2.find_long_repeats(name, set()), meta["period_scan"], "period_scan")
        seen = {(r["start"], r["end"]) for r in meta["tier1_sent"]}
        _cmp_records(t2.find_long_repeats(name, seen), meta["period_scan_seen"], "period_scan_seen")
    if meta["level"] >= 3:
        _cmp_records(t2.find_short_imperfect_repeats(name, set()), meta["short_imperfect"], "short_imperfect")


def test_short_imperfect_matches_oracle_on_planted_contig(bwt, oracle):
    from tests.util import gen_contig

    s = gen_contig(4000, 23, sub_rate=0.02)
    s[1000:1063] = np.frombuffer(b"AAGGCTC" * 9, np.uint8)
    s[2500:2581] = np.frombuffer(b"GATTACAGG" * 4 + b"GATTGCAGG" + b"GATTACAGG" * 4, np.uint8)
    text = s.tobytes() + b"$"
    core = bwt.BWTCore(text.decode())
    from bwt_algorithm_b200 import seeded

    got = seeded.short_imperfect_arrays(bwt.Tier2LCPFinder(core, min_period=1), set())
    want = oracle.short_imperfect_arrays(text)
    assert [(a[0], a[1], a[2], a[3], a[4].tobytes(), a[5], a[6], a[7]) for a in got] == \
           [(a[0], a[1], a[2], a[3], a[4].tobytes(), a[5], a[6], a[7]) for a in want]
    assert len(want) >= 1


def test_worker_and_helpers(bwt, oracle):
    from tests.util import gen_contig

    seq = gen_contig(6000, 8).tobytes().decode()
    cfg = {"sa_sample_rate": 32, "enable_tier1": True, "enable_tier2": True, "allow_mismatches": True,
           "max_motif_length": 9, "min_period": 10, "max_period": 1000, "min_copies": 3, "min_entropy": 1.0,
           "show_progress": False, "max_unit_len": 120}
    got = bwt._process_chromosome_worker(("c", seq, cfg))
    want = oracle.strict_scan(seq + "$", 1, 1000, 0, 3)
    assert [(r.start, r.end, len(r.motif), int(r.copies)) for r in got] == [tuple(x) for x in want[:, :4].tolist()]
    cfg["enable_tier2"] = False
    assert bwt._process_chromosome_worker(("c", seq, cfg)) == []
    arr = np.frombuffer(seq.encode(), np.uint8)
    assert bwt._count_equal_range(arr, 10, 300, ord("A")) == int(np.count_nonzero(arr[10:300] == ord("A")))
    assert bwt.HAVE_NUMBA is False
