"""Small end-to-end run for compute-sanitizer: a 150 kb planted contig with an N block through
the index build (SA + BWT + Occ + LCP incl. the deep-pair pass), the packed FM index + searches, the 8-mer
table and the four detectors, checked against the oracle so that a clean sanitizer log is also a correct run.

    compute-sanitizer --tool memcheck  python tools/sanitize_step.py
    compute-sanitizer --tool racecheck python tools/sanitize_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch

    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import detect
    from bwt_algorithm_b200.device_index import DeviceIndex
    from bwt_algorithm_b200.genome import GenomeScanner
    from oracle import oracle as orc
    from tests.util import gen_contig

    n = int(os.environ.get("SANITIZE_N", "150000"))
    s = gen_contig(n, 42)
    s[n // 3:n // 3 + 3000] = ord("N")                       # 6-symbol alphabet + deep LCP pairs
    s[n // 2:n // 2 + 9000] = np.tile(s[1000:4000], 3)       # three 3 kb copies: LCP > the direct-compare budget
    text = s.tobytes() + b"$"
    ix = DeviceIndex(text, build_isa=True, build_lcp=True)
    oi = orc.OracleIndex(text)
    assert np.array_equal(ix.sa.cpu().numpy(), oi.sa) and np.array_equal(ix.bwt.cpu().numpy(), oi.bwt)
    assert np.array_equal(ix.lcp.cpu().numpy(), oi.lcp())
    rng = np.random.default_rng(3)
    pats = [bytes(b"ACGTN"[x] for x in rng.integers(0, 5, int(k))) for k in rng.integers(1, 12, 4000)]
    mat = np.zeros((len(pats), 12), np.uint8)
    lens = np.array([len(p) for p in pats], np.int32)
    for i, p in enumerate(pats):
        mat[i, : len(p)] = np.frombuffer(p, np.uint8)
    want = oi.backward_search_batch(mat, lens)
    for mode in ("packed", "thread", "bytes"):
        ix.use_packed = mode != "bytes"
        ix.thread_per_query = mode == "thread"
        sp, ep = ix.backward_search_batch(pats)
        assert np.array_equal(sp, want[0]) and np.array_equal(ep, want[1]), mode
    ix.use_packed = True
    ix.motif_sweep(6)
    sc = GenomeScanner(len(text), arena_rows=600_000)
    res = sc.scan([torch.from_numpy(np.frombuffer(text, np.uint8).copy()).pin_memory()])[0]
    assert np.array_equal(sc.host_rows(res, "tier1"), orc.tier1_scan(text))
    got = sc.host_rows(res, "strict")
    assert len(got) > 1000          # (row-for-row parity of the strict scan is tests/test_gpu_detect.py's job)
    detect.plateau_rows(ix.text, ix.sa, ix.lcp, 1, 1000, 3)
    detect.period_scan_rows(ix.text)
    torch.cuda.synchronize()
    print(f"sanitize_step ok: n={len(text)}, rows={res.count()}")


if __name__ == "__main__":
    main()
