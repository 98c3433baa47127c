"""Pins the CPU oracle (oracle/oracle.c, oracle/oracle.py) and the host-side
row -> record conversion of the product package to the REFERENCE ITSELF.

tests/golden/*.npz were produced by oracle/gen_golden.py, which imports the
unmodified reference (bwt.py) in the build container with the wall clock
frozen, on the reference's own FASTA fixtures, on seeded planted contigs
 and on adversarial texts.  Everything here runs on CPU.
"""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
FLOAT_TOL = 1e-9


def _load(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _cmp_records(got, want, where):
    assert len(got) == len(want), f"{where}: {len(got)} records vs {len(want)} in the reference"
    for i, (g, w) in enumerate(zip(got, want)):
        for key, wv in w.items():
            gv = getattr(g, key)
            tag = f"{where}[{i}].{key}"
            if isinstance(wv, float):
                assert abs(float(gv) - wv) <= FLOAT_TOL, f"{tag}: {gv} vs {wv}"
            elif isinstance(wv, dict):
                assert set(gv) == set(wv), tag
                for k2 in wv:
                    assert abs(float(gv[k2]) - float(wv[k2])) <= FLOAT_TOL, f"{tag}[{k2}]"
            elif isinstance(wv, int) and not isinstance(wv, bool):
                assert int(gv) == wv, f"{tag}: {gv} vs {wv}"
            else:
                assert gv == wv, f"{tag}: {gv!r} vs {wv!r}"


@pytest.fixture(scope="module")
def pkg():
    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import finders, motifs, seeded

    return finders, motifs, seeded


def test_golden_files_present():
    assert len(GOLDEN) >= 35


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_index_arrays_match_reference(oracle, path):
    z, meta = _load(path)
    text = bytes(z["text"])
    ix = oracle.OracleIndex(text)
    lcp = ix.lcp()
    codes = meta["occ_codes"]
    occ = np.stack([ix.occ[c] for c in codes]) if codes else np.zeros((0, 0), np.int32)
    for key, arr in (("sa", ix.sa), ("bwt", ix.bwt), ("lcp", lcp), ("occ", occ.astype(np.int32))):
        assert _sha(arr) == meta[key + "_sha256"], f"{key} differs from the reference"
        if key in z.files:
            assert np.array_equal(arr, z[key])
    assert sorted(ix.occ.keys()) == codes
    for ch, v in meta["char_totals"].items():
        assert int(ix.totals[ord(ch)]) == v
    for ch, v in meta["char_counts"].items():
        assert int(ix.counts[ord(ch)]) == v
    assert meta["sampled_sa_ok"]
    # backward search / count / locate (bwt.py:359-410)
    pats = [p.encode("utf-8") for p in meta["patterns"]]
    stride = max(1, max(len(p) for p in pats))
    mat = np.zeros((len(pats), stride), np.uint8)
    lens = np.array([len(p) for p in pats], np.int32)
    for i, p in enumerate(pats):
        mat[i, : len(p)] = np.frombuffer(p, np.uint8)
    sp, ep = ix.backward_search_batch(mat, lens)
    want = np.array(meta["search"], np.int64).reshape(-1, 2)
    assert np.array_equal(sp, want[:, 0]) and np.array_equal(ep, want[:, 1])
    counts = np.where(sp == -1, 0, ep - sp + 1)
    assert counts.tolist() == meta["count"]
    for idx, wantpos in zip(meta["locate_idx"], meta["locate"]):
        got = [] if sp[idx] == -1 else sorted(int(x) for x in ix.sa[sp[idx]:ep[idx] + 1])
        assert got == wantpos
    for ch, pos, want_rank in meta["rank_probe"]:
        p = min(max(pos, 0), ix.n)
        assert int(np.count_nonzero(ix.bwt[:p] == ord(ch))) == want_rank
    # 8-mer table (bwt.py:138-193)
    table = oracle.kmer8_table(text)
    kc, kp = [], []
    for w in sorted(table):
        kc += [w] * len(table[w])
        kp += table[w]
    assert _sha(np.array(kc + kp, np.int64)) == meta["kmer_sha256"]
    bits = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 0}
    for q, wantpos in zip(meta["kmer_queries"], meta["kmer_lookup"]):
        if len(q) > 8 or not table:
            s1, e1 = ix.backward_search(q)
            got = [] if s1 == -1 else sorted(int(x) for x in ix.sa[s1:e1 + 1])
        else:
            w, ok = 0, True
            for b in q.upper():
                if b not in bits:
                    ok = False
                    break
                w = (w << 2) | bits[b]
            got = table.get(w, []) if ok else []
        assert got == wantpos, f"k-mer lookup {q!r}"


LEVEL1 = [p for p in GOLDEN if _load(p)[1]["level"] >= 1]


@pytest.mark.parametrize("path", LEVEL1, ids=[os.path.basename(p)[:-4] for p in LEVEL1])
def test_detectors_match_reference(oracle, pkg, path):
    finders, motifs, _ = pkg
    z, meta = _load(path)
    text = bytes(z["text"])
    seq = text[:-1]
    name = meta["name"]
    t_arr = np.frombuffer(text, np.uint8)
    s_arr = np.frombuffer(seq, np.uint8)
    # Tier 1 (bwt.py:1426-1531)
    _cmp_records(finders.tier1_records(s_arr, oracle.tier1_scan(seq), name), meta["tier1_nosent"], "tier1_nosent")
    _cmp_records(finders.tier1_records(t_arr, oracle.tier1_scan(text), name), meta["tier1_sent"], "tier1_sent")
    _cmp_records(finders.tier1_records(s_arr, oracle.tier1_scan(seq, max_motif_length=12), name), meta["tier1_m12"],
                 "tier1_m12")
    # strict adjacency (bwt.py:1891-2001)
    if "strict_worker" in meta:
        eff = max(120, min(len(seq) // 3, 1000))
        _cmp_records(finders.strict_records(t_arr, oracle.strict_scan(text, 1, eff, 0, 3), name, 0),
                     meta["strict_worker"], "strict_worker")
        _cmp_records(finders.strict_records(t_arr, oracle.strict_scan(text, 20, 120, 2, 3), name, 2),
                     meta["strict_default"], "strict_default")
        _cmp_records(finders.strict_records(t_arr, oracle.strict_scan(text, 2, 40, 1, 4), name, 1),
                     meta["strict_mm1"], "strict_mm1")
    # LCP plateaus (bwt.py:2118-2145, 2500-2560)
    sa = oracle.suffix_array(text)
    lcp = oracle.kasai_lcp(text, sa)
    for key, mp in (("plateau_p1", 1), ("plateau_p10", 10)):
        rows, _thr = oracle.lcp_plateaus(text, sa, lcp, min_period=mp)
        _cmp_records(finders.plateau_records(t_arr, rows, name), meta[key], key)
    # extension / consensus probes (bwt.py:2392-2498, 2697-2805, 1207-1256)
    ns = len(seq)
    for s0, p, allow, *want in meta["extend_mm"]:
        assert list(oracle.extend_with_mismatches(text, s0, p, ns, allow)) == want
    for s0, p, *want in meta["extend_fm"]:
        assert list(oracle.extend_tandem_fm(text, s0, p)) == want
    for s0, p, nc, cons, rate, mx in meta["consensus"]:
        c, r, m = oracle.consensus(text, s0, p, nc)
        assert c.tobytes().decode("latin1") == cons and r == rate and m == mx


LEVEL2 = [p for p in GOLDEN if _load(p)[1]["level"] >= 2]


@pytest.mark.parametrize("path", LEVEL2, ids=[os.path.basename(p)[:-4] for p in LEVEL2])
def test_period_scan_matches_reference(oracle, pkg, path):
    finders, _, _ = pkg
    z, meta = _load(path)
    text = bytes(z["text"])
    t_arr = np.frombuffer(text, np.uint8)
    name = meta["name"]

    def records(mask):
        rows, _ = oracle.period_scan(text, tier1_mask=mask)
        cons = [oracle.consensus(text, int(r[6]), int(r[2]), int(r[3]))[0] for r in rows]
        return finders.period_records(t_arr, rows, cons, name)

    _cmp_records(records(None), meta["period_scan"], "period_scan")
    mask = np.zeros(len(text) - 1, np.uint8)
    for r in meta["tier1_sent"]:
        mask[r["start"]:min(r["end"], mask.size)] = 1
    _cmp_records(records(mask), meta["period_scan_seen"], "period_scan_seen")


LEVEL3 = [p for p in GOLDEN if _load(p)[1]["level"] >= 3]


@pytest.mark.parametrize("path", LEVEL3, ids=[os.path.basename(p)[:-4] for p in LEVEL3])
def test_short_imperfect_matches_reference(oracle, pkg, path):
    _, _, seeded = pkg
    z, meta = _load(path)
    text = bytes(z["text"])
    arrays = oracle.short_imperfect_arrays(text)
    recs = seeded.records_from_arrays(np.frombuffer(text, np.uint8), arrays, meta["name"])
    _cmp_records(recs, meta["short_imperfect"], "short_imperfect")


def test_motif_helpers_match_reference_vectors(pkg):
    """The reference's own unit checks (test_imperfect_repeats.py:188-253)."""
    _, motifs, _ = pkg
    M = motifs.MotifUtils
    assert M.hamming_distance("ATCG", "ATCG") == 0 and M.hamming_distance("ATCG", "ATGG") == 1
    assert M.hamming_distance("AAAA", "TTTT") == 4 and M.hamming_distance("ATCG", "GCTA") == 4
    assert M.calculate_entropy("AAAA") == 0.0 and M.calculate_entropy("ATAT") == 1.0
    assert abs(M.calculate_entropy("ATCG") - 2.0) < 1e-12
    assert M.build_consensus_motif(["ATCG", "ATCG", "ATGG", "ATCG"])[0] == "ATCG"
    assert M.reverse_complement("ATCG") == "CGAT"
    assert M.get_canonical_motif("TCGA") == "ATCG"
    assert [len(list(M.enumerate_motifs(k))) for k in range(1, 7)] == [4, 6, 20, 60, 204, 670]
