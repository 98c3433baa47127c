"""Strict-scan rows of the chr21-sized synthetic contig (the CLI's detector call, pipeline._detect_rows) written as
a compressed .npz, so that the host post-processing chain can be replayed and profiled on a box without a GPU
(tools/chain_replay.py):

    python tools/dump_strict_rows.py [--n 46709983] [--out gpurun_out/strict_rows_chr21.npz]

Also prints the wall time of the scan (third call) as one JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=46_709_983)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "strict_rows_chr21.npz"))
    ap.add_argument("--compare", default="", help="an earlier dump: only report whether the rows are the same")
    args = ap.parse_args()
    import numpy as np
    import torch

    import bwt_algorithm_b200  # noqa: F401
    from bench import gen_contig
    from bwt_algorithm_b200 import detect

    s = gen_contig(args.n, 21)
    text = np.concatenate([s, np.frombuffer(b"$", np.uint8)])
    d = torch.from_numpy(text).cuda()
    unit_cap = max(1000, min(args.n // 3, 1000))
    ms = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows = detect.strict_rows(d, 1, unit_cap, 0, 3)
        ms.append(round((time.perf_counter() - t0) * 1e3, 2))
    assert not rows[:, [4, 5, 7]].any()
    if args.compare:
        z = np.load(args.compare)
        same = (int(z["n"]) == args.n and len(z["start"]) == len(rows)
                and all(np.array_equal(z[k], rows[:, c]) for k, c in (("start", 0), ("end", 1), ("prim", 2), ("copies", 3), ("unit", 6))))
        print(json.dumps({"rows": int(len(rows)), "strict_rows_ms": ms, "same_as": args.compare, "same": bool(same)}))
        sys.exit(0 if same else 1)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    np.savez_compressed(args.out, start=rows[:, 0], end=rows[:, 1], prim=rows[:, 2].astype(np.int16),
                        copies=rows[:, 3], unit=rows[:, 6].astype(np.int16), n=np.int64(args.n))
    print(json.dumps({"rows": int(len(rows)), "strict_rows_ms": ms, "bytes": os.path.getsize(args.out)}))


if __name__ == "__main__":
    main()
