"""Wall-clock profile of the CLI-live worker (index build + strict adjacency scan + record
conversion) on one synthetic contig: where does a user's time go?"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=46_709_983)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--cli", action="store_true", help="run the whole CLI (FASTA in, strfinder table out)")
    args = ap.parse_args()
    import torch

    import bwt_algorithm_b200  # noqa: F401
    from bench import gen_contig
    from bwt_algorithm_b200 import pipeline

    seq = gen_contig(args.n, 21).tobytes().decode("ascii")
    if args.cli:
        fa = "/tmp/pipeline_profile.fa"
        with open(fa, "w") as f:
            f.write(">chrS\n")
            for a in range(0, len(seq), 80):
                f.write(seq[a:a + 80] + "\n")
        argv = [fa, "-o", "/tmp/pipeline_profile.tab", "--jobs", "0"]
        t0 = time.perf_counter()
        if args.profile:
            pr = cProfile.Profile()
            pr.enable()
        pipeline.main(argv)
        if args.profile:
            pr.disable()
        dt = time.perf_counter() - t0
        print(f"cli: {dt:.2f} s ({args.n / dt / 1e6:.2f} Mbases/s)")
        if args.profile:
            pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
        return
    config = {"sa_sample_rate": 32, "enable_tier1": False, "enable_tier2": True, "enable_tier3": False,
              "max_motif_length": 9, "min_period": 10, "max_period": 1000, "min_copies": 3, "min_entropy": 1.0,
              "allow_mismatches": True, "show_progress": True, "max_unit_len": 120}
    pipeline._process_chromosome_worker(("warm", seq[:200_000], config))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if args.profile:
        pr = cProfile.Profile()
        pr.enable()
    recs = pipeline._process_chromosome_worker(("chrS", seq, config))
    if args.profile:
        pr.disable()
    dt = time.perf_counter() - t0
    print(f"worker: {len(recs)} records in {dt:.2f} s ({args.n / dt / 1e6:.2f} Mbases/s)")
    if args.profile:
        pstats.Stats(pr).sort_stats("cumulative").print_stats(18)


if __name__ == "__main__":
    main()
