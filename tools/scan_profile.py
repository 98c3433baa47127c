"""Times the Tier 1 / Tier 2 scans on one chr21-sized contig (warm, CUDA events +
wall clock) -- run it plainly for numbers, under `ncu --metrics gpu__time_duration.sum`
for the launch list committed under profiles/."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=46_709_983)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch

    import bwt_algorithm_b200  # noqa: F401
    from bench import gen_contig
    from bwt_algorithm_b200 import detect
    from bwt_algorithm_b200.device_index import DeviceIndex

    s = gen_contig(args.n, 21)
    text = np.concatenate([s, np.frombuffer(b"$", np.uint8)])
    d_text = torch.from_numpy(text).cuda()
    ix = DeviceIndex(d_text, text_is_device=True, build_lcp=True, build_kmer=False)
    scans = {
        "tier1": lambda: detect.tier1_rows(d_text),
        "strict_u1_1000": lambda: detect.strict_rows(d_text, 1, 1000, 0, 3),
        "lcp_plateaus": lambda: detect.plateau_rows(d_text, ix.sa, ix.lcp, 1, 1000, 3)[0],
        "period_scan": lambda: detect.period_scan_rows(d_text)[0],
    }
    for name, fn in scans.items():
        if args.only and name not in args.only.split(","):
            continue
        ts = []
        for _ in range(args.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rows = fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        print(name, "records", len(rows), "ms", " ".join(f"{t:.2f}" for t in ts), flush=True)


if __name__ == "__main__":
    main()
