"""The command profiled under ncu for profiles/r2_*: a warm and a measured index build of a chr21-sized
contig (BASELINE configs[2]), the four detectors on it, and the FM search kernels on its index.  Only the
phase named by --profile lies between cudaProfilerStart/Stop, so with `--profile-from-start off` the ncu
launch list holds exactly that phase:

    python tools/ncu_step.py --profile index                   # plain run (must exit 0 first)
    ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/r2_ncu_index.csv \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct \
        python tools/ncu_step.py --profile index
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=46_709_983)
    ap.add_argument("--queries-log2", type=int, default=22)
    ap.add_argument("--profile", default="index", choices=["index", "scans", "fm", "genome"])
    args = ap.parse_args()
    import torch

    import bwt_algorithm_b200  # noqa: F401
    from bench import IndexStep, device_contig
    from bwt_algorithm_b200 import _lib, detect
    from bwt_algorithm_b200.device_index import FMReplica

    L = _lib.lib()
    dev = torch.device("cuda", 0)
    d_text = device_contig(torch, args.n, 21, dev)
    n = int(d_text.numel())
    step = IndexStep(torch, L, _lib, n, dev)
    prof = torch.cuda.profiler

    def region(name, fn, warm=True):
        if warm:
            fn()
        torch.cuda.synchronize()
        if args.profile == name:
            prof.start()
        out = fn()
        torch.cuda.synchronize()
        if args.profile == name:
            prof.stop()
        return out

    region("index", lambda: step.run(d_text))
    if args.profile == "scans":
        def scans():
            detect.tier1_rows(d_text)
            detect.strict_rows(d_text, 1, 1000, 0, 3)
            detect.plateau_rows(d_text, step.sa, step.lcp, 1, 1000, 3)
            detect.period_scan_rows(d_text)
        region("scans", scans)
    if args.profile == "fm":
        rep = FMReplica(step.bwt, step.occ, step.totals, step.row, n, 128, device=dev)
        rep.ensure_packed()
        nq = 1 << args.queries_log2
        g = torch.Generator(device=dev)
        g.manual_seed(7)
        lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
        pats = lut[torch.randint(0, 4, (nq, 10), generator=g, device=dev)]
        lens = torch.full((nq,), 10, dtype=torch.int32, device=dev)

        def fm():
            rep.motif_sweep(10)
            rep.thread_per_query = False
            rep.search_device(pats, 10, lens, nq)
            rep.thread_per_query = True
            rep.search_device(pats, 10, lens, nq)
            rep.thread_per_query = False
            rep.use_packed = False
            rep.search_device(pats, 10, lens, nq)
            rep.use_packed = True
        region("fm", fm)
    print(f"ncu_step ok: profile={args.profile} n={n}", flush=True)


if __name__ == "__main__":
    main()
