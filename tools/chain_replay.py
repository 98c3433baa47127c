"""Replay of the CLI's host side on a box without a GPU: FASTA ingest, the row-level post-processing chain
(rowchain.finish_rows) and the writers, with the detector call answered from the rows that
tools/dump_strict_rows.py saved on a B200 (same generator, same seed; --flank-trim 0 because the dump
scanned the whole contig):

    python tools/chain_replay.py [--rows gpurun_out/strict_rows_chr21.npz] [--format bed] [--profile]

prints the wall time of `bwt.main` (second call), the md5 of the output and, with --profile, the cProfile
top lists.  Test infrastructure: nothing in the package imports it."""
import argparse
import contextlib
import cProfile
import hashlib
import io
import json
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default=os.path.join(ROOT, "gpurun_out", "strict_rows_chr21.npz"))
    ap.add_argument("--format", default="bed")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--contigs", type=int, default=1)
    ap.add_argument("--jobs", type=int, default=4)
    ap.add_argument("--limit", type=int, default=0, help="use only the first N bases (rows entirely inside)")
    args = ap.parse_args()
    import numpy as np

    import bwt_algorithm_b200  # noqa: F401
    from bench import gen_contig
    from bwt_algorithm_b200 import bwt, pipeline

    z = np.load(args.rows)
    n = int(z["n"])
    s = gen_contig(n, 21)
    rows = np.zeros((len(z["start"]), 8), np.int32)
    rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3], rows[:, 6] = z["start"], z["end"], z["prim"], z["copies"], z["unit"]
    if args.limit:
        n = args.limit
        s = s[:n]
        rows = rows[rows[:, 1] + rows[:, 6] < n]      # the run that made the call ends before the cut
    # --contigs C: the contig cut into C pieces (calls that cross a cut are dropped): timing of the chain pool only
    C = max(1, args.contigs)
    cuts = [n * i // C for i in range(C + 1)]
    served = {}
    for i in range(C):
        a, b = cuts[i], cuts[i + 1]
        sel = rows[(rows[:, 0] >= a) & (rows[:, 1] + rows[:, 6] < b)].copy() if C > 1 else rows
        sel[:, 0] -= a
        sel[:, 1] -= a
        served[f"chr21_sized_{i}" if C > 1 else "chr21_sized"] = (sel, np.concatenate([s[a:b], np.frombuffer(b"$", np.uint8)]))

    def replay_rows(chrom, seq, config):
        r, text = served[chrom]
        assert len(seq) + 1 == len(text)
        if config.get("show_progress"):
            print(f"  [{chrom}] Strict adjacency: {len(r)} tandem repeats detected")
        return r.copy(), text.copy()

    pipeline._detect_rows = replay_rows
    tmp = tempfile.mkdtemp(prefix="bwtk_replay_")
    fa, out = os.path.join(tmp, "chr21_sized.fa"), os.path.join(tmp, "out." + args.format)
    with open(fa, "wb") as fh:
        for name, (_, text) in served.items():
            part = text[:-1]
            pad = (-part.size) % 80
            body = np.concatenate([part, np.full(pad, 10, np.uint8)]).reshape(-1, 80)
            lines = np.concatenate([body, np.full((body.shape[0], 1), 10, np.uint8)], axis=1).tobytes()
            fh.write(b">" + name.encode() + b" synthetic\n" + lines.rstrip(b"\n") + b"\n")
    argv = [fa, "--progress", "--format", args.format, "--jobs", str(args.jobs), "--flank-trim", "0", "-o", out]
    walls, cpus = [], []
    for _ in range(args.repeat):
        t0, c0 = time.perf_counter(), time.process_time()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            bwt.main(argv)
        walls.append(round(time.perf_counter() - t0, 2))
        cpus.append(round(time.process_time() - c0, 2))
    tail = [ln for ln in buf.getvalue().splitlines() if "Nested call" in ln or "Completed" in ln]
    print(json.dumps({"bases": n, "rows": int(len(rows)), "wall_s": walls, "cpu_s": cpus, "output_lines": sum(1 for _ in open(out)),
                      "md5": hashlib.md5(open(out, "rb").read()).hexdigest(), "stdout": tail}))
    if args.profile:
        pr = cProfile.Profile()
        pr.enable()
        with contextlib.redirect_stdout(io.StringIO()):
            bwt.main(argv)
        pr.disable()
        for key in ("cumulative", "tottime"):
            st = io.StringIO()
            pstats.Stats(pr, stream=st).sort_stats(key).print_stats(args.top)
            print("\n".join(st.getvalue().splitlines()[:args.top + 12]))


if __name__ == "__main__":
    main()
