#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED Python reference.

Run in the build container only (it imports /root/reference/bwt.py, which does
not exist on the GPU box):

    python oracle/gen_golden.py [--only NAME] [--skip-slow]

The wall clock seen by the reference is frozen by
rebinding the module attribute ``bwt.time`` to a shim, so the 30 s cap of
``_find_repeats_simple`` (bwt.py:2238-2257) never fires and outputs are
deterministic; the 100 000-iteration cap still applies.

Every case stores the input text and the reference's outputs for the hot-path
functions of SURVEY.md §8(a).  Large arrays are stored as sha256 digests.
"""
from __future__ import annotations

import argparse
import hashlib
import importlib.util
import json
import os
import sys
import time as _real_time
import types

import numpy as np

REF_DIR = os.environ.get("BWT_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


def load_reference():
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_bwt", os.path.join(REF_DIR, "bwt.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_bwt"] = mod
    spec.loader.exec_module(mod)
    shim = types.SimpleNamespace(**{k: getattr(_real_time, k) for k in dir(_real_time) if not k.startswith("_")})
    shim.time = lambda: 0.0
    mod.time = shim
    return mod


def gen_contig(n, seed, sub_rate=0.03):
    """SURVEY.md Appendix B."""
    rng = np.random.default_rng(seed)
    A = np.frombuffer(b"ACGT", dtype=np.uint8)
    s = A[rng.integers(0, 4, n)]
    pos = 500
    while pos < n - 400:
        k = int(rng.integers(1, 10))
        m = A[rng.integers(0, 4, k)]
        c = int(rng.integers(4, 25))
        arr = np.tile(m, c)
        mut = rng.random(arr.size) < sub_rate
        arr[mut] = A[rng.integers(0, 4, int(mut.sum()))]
        s[pos:pos + arr.size] = arr
        pos += arr.size + int(rng.integers(300, 1500))
    return s


def read_fasta(path):
    out, name, seq = [], None, []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith(">"):
                if name is not None:
                    out.append((name, "".join(seq)))
                name, seq = line[1:].split()[0], []
            elif line:
                seq.append(line.upper())
    if name is not None:
        out.append((name, "".join(seq)))
    return out


def rec_to_dict(r):
    return {
        "start": int(r.start), "end": int(r.end), "motif": r.motif, "copies": float(r.copies),
        "length": int(r.length), "tier": int(r.tier), "confidence": float(r.confidence),
        "consensus_motif": r.consensus_motif, "mismatch_rate": float(r.mismatch_rate),
        "max_mismatches_per_copy": int(r.max_mismatches_per_copy),
        "n_copies_evaluated": int(r.n_copies_evaluated), "strand": r.strand,
        "percent_matches": float(r.percent_matches), "percent_indels": float(r.percent_indels),
        "score": int(r.score), "composition": r.composition, "entropy": float(r.entropy),
        "actual_sequence": r.actual_sequence, "variations": r.variations,
    }


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def motif_patterns(rng, text_no_sentinel):
    pats = [""]
    for k in range(1, 5):
        for v in range(4 ** k):
            pats.append("".join("ACGT"[(v >> (2 * (k - 1 - j))) & 3] for j in range(k)))
    pats += ["N", "AN", "NA", "$", "A$", "ANNA", "X", "AXA", "NNNN", "acgt"]
    for L in (5, 7, 9, 10, 12):
        for _ in range(40):
            pats.append("".join("ACGT"[x] for x in rng.integers(0, 4, L)))
    n = len(text_no_sentinel)
    for _ in range(60):
        if n < 2:
            break
        L = int(rng.integers(1, min(12, n) + 1))
        p = int(rng.integers(0, n - L + 1))
        pats.append(text_no_sentinel[p:p + L])
    return pats


def run_case(ref, name, seq, level):
    """level 0: index only (hashed); 1: + cheap detectors; 2: + slow detectors."""
    t0 = _real_time.time()
    text = seq + "$"
    core = ref.BWTCore(text)
    out = {"name": name, "level": level}
    arrs = {"text": np.frombuffer(text.encode("utf-8"), np.uint8)}
    n = core.n
    codes = sorted(core.occ_checkpoints.keys())
    occ = np.stack([core.occ_checkpoints[c] for c in codes]) if codes else np.zeros((0, 0), np.int32)
    t2 = ref.Tier2LCPFinder(core, min_period=1)
    lcp = t2._compute_lcp_array()
    big = n > 40000
    for key, a in (("sa", core.suffix_array.astype(np.int32)), ("bwt", core.bwt_arr),
                   ("lcp", lcp.astype(np.int32)), ("occ", occ.astype(np.int32))):
        out[key + "_sha256"] = sha(a)
        if not big:
            arrs[key] = a
    out["occ_codes"] = codes
    out["alphabet"] = core.alphabet
    out["char_counts"] = {k: int(v) for k, v in core.char_counts.items()}
    out["char_totals"] = {k: int(v) for k, v in core.char_totals.items()}
    out["sampled_sa_ok"] = all(int(core.sampled_sa[i]) == int(core.suffix_array[i]) for i in range(0, n, 32))

    rng = np.random.default_rng(12345)
    pats = motif_patterns(rng, seq)
    out["patterns"] = pats
    out["search"] = [list(map(int, core.backward_search(p))) for p in pats]
    out["count"] = [int(core.count_occurrences(p)) for p in pats]
    loc_idx = list(range(0, len(pats), 7))
    out["locate_idx"] = loc_idx
    out["locate"] = [[int(x) for x in core.locate_positions(pats[i])] for i in loc_idx]
    out["rank_probe"] = []
    for _ in range(64):
        ch = "ACGTN$"[int(rng.integers(0, 6))]
        pos = int(rng.integers(-2, n + 3))
        out["rank_probe"].append([ch, pos, int(core.rank(ch, pos))])

    # k-mer table: full table as sorted (code,pos) pairs + lookups
    kc, kp = [], []
    for w in sorted(core.kmer_hash):
        for p in core.kmer_hash[w]:
            kc.append(w)
            kp.append(p)
    out["kmer_sha256"] = sha(np.array(kc + kp, np.int64))
    if not big:
        arrs["kmer_codes"] = np.array(kc, np.int32)
        arrs["kmer_pos"] = np.array(kp, np.int32)
    kq = ["A", "AT", "N", "ACG", "TTTT", "AAAAAAAA", "ACGTACGT", "X", "AXG", "at", "ACGTACGTA", "CAG", "GT"]
    for _ in range(20):
        L = int(rng.integers(1, 10))
        kq.append("".join("ACGTN"[x] for x in rng.integers(0, 5, L)))
    out["kmer_queries"] = kq
    out["kmer_lookup"] = [[int(x) for x in core.get_kmer_positions(q)] for q in kq]

    if level >= 1:
        text_arr = np.frombuffer(seq.encode(), np.uint8)
        out["tier1_nosent"] = [rec_to_dict(r) for r in ref.Tier1STRFinder(text_arr).find_strs(name)]
        out["tier1_sent"] = [rec_to_dict(r) for r in ref.Tier1STRFinder(core.text_arr).find_strs(name)]
        out["tier1_m12"] = [rec_to_dict(r) for r in ref.Tier1STRFinder(text_arr, max_motif_length=12).find_strs(name)]
        # strict scan as the worker calls it (bwt.py:3089-3106) and with defaults
        eff = max(120, min(len(seq) // 3, 1000))
        if len(seq) <= 6000:
            out["strict_worker"] = [rec_to_dict(r) for r in t2.find_long_unit_repeats_strict(
                name, min_unit_len=1, max_unit_len=eff, max_mismatch=0, min_copies=3)]
            out["strict_default"] = [rec_to_dict(r) for r in t2.find_long_unit_repeats_strict(name)]
            out["strict_mm1"] = [rec_to_dict(r) for r in t2.find_long_unit_repeats_strict(
                name, min_unit_len=2, max_unit_len=40, max_mismatch=1, min_copies=4)]
        out["plateau_p1"] = [rec_to_dict(r) for r in t2._detect_lcp_plateaus(lcp, name)]
        t2b = ref.Tier2LCPFinder(core, min_period=10)
        out["plateau_p10"] = [rec_to_dict(r) for r in t2b._detect_lcp_plateaus(lcp, name)]
        # direct extension / consensus probes
        ext, extfm, cons = [], [], []
        ns = len(seq)
        for _ in range(40):
            p = int(rng.integers(1, 12))
            if ns < 3 * p:
                continue
            s0 = int(rng.integers(0, ns - 2 * p))
            allow = bool(rng.integers(0, 2))
            ext.append([s0, p, allow] + [int(x) for x in t2._extend_with_mismatches(core.text_arr, s0, p, ns, allow_mismatches=allow)])
            motif = seq[s0:s0 + p]
            extfm.append([s0, p] + [int(x) for x in t2._extend_tandem_fm(core.text_arr, s0, motif, p, 0)])
            nc = int(rng.integers(1, 8))
            c, mm, mx = ref.MotifUtils.build_consensus_motif_array(core.text_arr, s0, p, nc)
            cons.append([s0, p, nc, c.tobytes().decode("latin1"), float(mm), int(mx)])
        out["extend_mm"] = ext
        out["extend_fm"] = extfm
        out["consensus"] = cons
    if level >= 2:
        out["period_scan"] = [rec_to_dict(r) for r in t2.find_long_repeats(name, set())]
        t1seen = {(r["start"], r["end"]) for r in out["tier1_sent"]}
        out["period_scan_seen"] = [rec_to_dict(r) for r in t2.find_long_repeats(name, t1seen)]
    if level >= 3:
        out["short_imperfect"] = [rec_to_dict(r) for r in t2.find_short_imperfect_repeats(name, set())]
    out["gen_seconds"] = round(_real_time.time() - t0, 2)
    np.savez_compressed(os.path.join(OUT_DIR, name + ".npz"),
                        meta=np.frombuffer(json.dumps(out).encode(), np.uint8), **arrs)
    print(f"{name}: n={n} level={level} {out['gen_seconds']}s", flush=True)


def cases():
    """(name, sequence, level)"""
    cs = []
    for fa, lvl in (("test.fa", 2), ("test2.fa", 2), ("test_synthetic.fasta", 3),
                    ("test_long_motif.fa", 2), ("test1.fa", 3)):
        for i, (nm, seq) in enumerate(read_fasta(os.path.join(REF_DIR, fa))):
            cs.append((f"ref_{fa.split('.')[0]}_{i}", seq, lvl))
    b2s = lambda a: a.tobytes().decode()
    cs.append(("syn_600_s3", b2s(gen_contig(600, 3)), 3))
    cs.append(("syn_1200_s4", b2s(gen_contig(1200, 4)), 2))
    cs.append(("syn_3000_s1", b2s(gen_contig(3000, 1)), 2))
    s = gen_contig(5000, 2)
    s[2000:2300] = ord("N")
    cs.append(("syn_5000_s2_N", b2s(s), 1))
    cs.append(("syn_20000_s5", b2s(gen_contig(20000, 5)), 1))
    s = gen_contig(30000, 6)
    s[10000:14000] = ord("N")
    cs.append(("syn_30000_s6_N", b2s(s), 1))
    cs.append(("syn_150000_s42", b2s(gen_contig(150000, 42)), 0))
    cs.append(("adv_allA_300", "A" * 300, 2))
    cs.append(("adv_AC_511", "AC" * 255 + "A", 2))          # n incl. '$' = 512 (multiple of 128)
    cs.append(("adv_n7", "ACGTACG"[:6], 1))                 # n incl. '$' = 7 (< 8: no k-mer table)
    cs.append(("adv_n8", "ACGTACG", 1))                     # n incl. '$' = 8
    cs.append(("adv_n127", b2s(gen_contig(127, 9)), 1))     # n = 128
    cs.append(("adv_Nruns", "ACGT" * 10 + "N" * 50 + "ACGTTGCA" * 8 + "N" * 9 + "GATTACA" * 6, 2))
    cs.append(("adv_iupac", "ACGTRYACGTACGTKMACGTACGTACGTSWACGTNNACGTACGTACGTACGT" * 3, 1))
    cs.append(("adv_period3", "CAG" * 120 + "T" + "CAG" * 60, 2))
    # seeded seed-and-extend (a14) cases that actually yield records
    rng = np.random.default_rng(77)
    rnd = lambda n: "".join("ACGT"[x] for x in rng.integers(0, 4, n))
    cs.append(("a14_k9_transition", rnd(60) + "GATTACAGG" * 4 + "GATTGCAGG" + "GATTACAGG" * 5 + rnd(50), 3))
    cs.append(("a14_k8_two_loci", rnd(40) + "AACCGGTT" * 3 + "AACCAGTT" + "AACCGGTT" * 4 + rnd(30)
               + "ACGGTCAT" * 7 + rnd(35), 3))
    cs.append(("a14_k7_aliased", rnd(30) + "AAGGCTC" * 9 + rnd(25) + "AAGGCTC" * 3 + "AAGACTC" + "AAGGCTC" * 4 + rnd(20), 3))
    cs.append(("a14_mixed", rnd(50) + "AAAAAAACACACACACACACAC" + rnd(20) + "AAAAAAACAAAAAAAC" + "CAGCAGCAGCAACAGCAGCAGCAG"
               + rnd(30) + "TTGACCGTA" * 6 + "TTGACTGTA" + "TTGACCGTA" * 3 + rnd(40), 3))
    # Found by probing the reference (texts with planted arrays; the ones kept are those whose reference output is
    # non-empty): nine more seeded seed-and-extend (a14) cases with one short_imperfect record each, and three
    # LCP-plateau (a15) cases with period-20 arrays (21, 64 and 30 plateau records).
    cs.append(("a14x_k7_AAGGCTC_transition", "TTGTCAATAATCTACCGCAAATAAGAGGAAAGGCTCAAGGCTCAAGGCTCAAGGCTCAAGGCTCAAGGCTCAAGGTTCAAGGCTCAAGGCTCCCATAAGTCAAATGTGATCCCGGTTTGT", 3))
    cs.append(("a14x_k7_AGATTCA_transition", "TAAATCAGTAAGTGTGCGGCTTGACCGCCGTAATTTTCGCAGATTCAAGATTCAAGATTCAAGATTCAAGACTCAAGATTCAAGATTCAAGATTCACACATAGTTTCTTCGACTTGTGAAACGCGGAGGCGC", 3))
    cs.append(("a14x_k7_ATTGCCA_pure", "ATTATAGCCCTTTGGCACCGTGCGTTCTACCGACCGATGTTTACCATTGCCAATTGCCAATTGCCAATTGCCAATTGCCAATTGCCAATTGCCAATTGCCAATTGCCAATTGCCATTTGCGGCGGTAGATGGAGTAGCAACCATAATTGTGTAGCATAACGCAT", 3))
    cs.append(("a14x_k8_AACCGGTA_transition", "TTAAATGTTAGGAAGCAGTGATAATCCCGACCCAACCGGTAAACCGGTAGACCGGTAAACCGGTAAACCGGTAAACCGGTAAACCGGTAAACCGGTAGCTATTGTTTGGCCAAGTGGCCAGGCCTGCGA", 3))
    cs.append(("a14x_k8_ACGGTCAT_pure", "GTTGGTATGTTACCTGCACTAGGCCGGATGCTATCCGGCCTAACGGTCATACGGTCATACGGTCATACGGTCATACGGTCATACGGTCATACGGTCATACGGTCATGCCGAGTTCAACCCTTAGACCCCTCACCAAGTTGATG", 3))
    cs.append(("a14x_k8_AGGATCCA_transition", "TAATAGTGTATGAATATGGCGTCGCTCCTCTCTGTATAGGATCCAAGGATCCAAGGATCCAAGGATCCAAGGATCCAAGGATCCAAGGGTCCAAGGATCCAAGGATCCACCCTCAACCGAAGCACGGCTGTCGCCAAGGGCGGAACTGCGTAA", 3))
    cs.append(("a14x_k8_AAGCTTGC_transition", "TCATCACAGATAAGCTCCGGTCGAAGACCGTGAAGCTTGCAAGCTTGCAAGCTTGCAAGCTTGCAAGCTTGCAAACTTGCAAGCTTGCAAGCTTGCAAGCTTGCAAGCTTGCCTATGGGGAGGCCGAGCCCTATCCGGGAGGCTTGCCACTCTCG", 3))
    cs.append(("a14x_k9_GATTACAGG_pure", "GGGCTTACCCGCTATATACAAACCTGTATGATAAACCAACTTTGCGATTACAGGGATTACAGGGATTACAGGGATTACAGGGATTACAGGGATTACAGGGATTACAGGGATTACAGGGATTACAGGCGGTAAAAGGGCGGTACGGTCGAGCAACTTGAAGAGCCGAGGTGGTG", 3))
    cs.append(("a14x_k9_AACGGTCAT_pure", "GATATTTGCCAATCGCCCACCAATACATGTTTTGTTGCACATCAGAAACGGTCATAACGGTCATAACGGTCATAACGGTCATAACGGTCATAACGGTCATTGAATTTGGAAAGTGCCGGGCGTGACTCCACCATACTAGATAGTGTG", 3))
    cs.append(("plat_p20_x5", "AGGGCCATATAAAGTTTGCGCAGCAAGGTCCAAGCTTGCAGACGAATACCTGATTTACACTGTGTGGTAGTTTGCATGCTTGTGTGGTAGTTTGCATGCTTGTGTGGTAGTTTGCATGCTTGTGTGGTAGTTTGCATGCTTGTGTGGTAGTTTGCATGCTACCGGAACGGGTTGCTAAGGGTTGTCGCTTAGTCCTAGTGGCATACCGAACTTCGGTCAA", 1))
    cs.append(("plat_p20_two", "GACGGGCCGATTAGCAGGCGCGAAAACACGACGAAGCCACGTGCCTACGCCTCCCGTGAGGTGCCTACGCCTCCCGTGAGGTGCCTACGCCTCCCGTGAGGTGCCTACGCCTCCCGTGAGGTTCTTGTACAGGTCCCTTGAGGATTCCTCATTACTCACGCAGCTATCAGAGGCGTCAACCCGGGCCCCTTAGTAGACAAACCGGACCGATAGTAGACAAACCGGACCGATAGTAGACAAACCGGACCGATAGTAGACAAACCGGACCGATAGTAGACAAACCGGACCGATAGTAGACAAACCGGACCGATAGTAGACAAACCGGACCGAGACTTTGCTTATGCTAGGCTCGCTACCACT", 1))
    cs.append(("plat_p20_mut", "CTAGCAAGTTGGAGGTGCATTGATGGTAATCCGGCATGATCTACGGTCCCACCAGGCGTATGCCTCTTAGACCAGGCGTATGCCTCTTAGACCAGGCGTATGCCTCTTAGACCAGGCGTAAGCCTCTTAGACCAGGCGTATGCCTCTTAGACCAGGCGTATGCCTCTTAGACCAGGCGTATGCCTCTTAGACCAGGCGTATGCCTCTTAGTCTATGAAAACGAGACATCGGAAAATGATTTCCCATGCCGAACGTACGAG", 1))
    cs.append(("adv_empty", "", 0))
    cs.append(("adv_one", "G", 0))
    return cs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--max-level", type=int, default=3)
    args = ap.parse_args()
    os.makedirs(OUT_DIR, exist_ok=True)
    ref = load_reference()
    for name, seq, lvl in cases():
        if args.only and args.only not in name:
            continue
        run_case(ref, name, seq, min(lvl, args.max_level))


if __name__ == "__main__":
    main()
