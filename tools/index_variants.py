"""Index build of one contig under the round-0 variants (LSD radix sort / MSD bucket sort unfused /
fused regroup), for one or more builds of the library: step time (CUDA events, L2 flushed), per-kernel
profile, and a check that every variant produces the same SA, BWT and LCP.

    python tools/index_variants.py [--n 46709983] [--libs default gpurun_variants/libbwtk_x.so ...]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MODES = {
    "lsd": {"BWTK_MSD": "0"},
    "msd_unfused": {"BWTK_MSD": "1", "BWTK_MSD_FUSE": "0"},
    "msd_fused": {"BWTK_MSD": "1", "BWTK_MSD_FUSE": "1"},
}


def child(args):
    import torch

    import bwt_algorithm_b200  # noqa: F401
    from bench import IndexStep, device_contig, profile_report
    from bwt_algorithm_b200 import _lib

    L = _lib.lib()
    dev = torch.device("cuda", 0)
    d_text = device_contig(torch, args.n, args.seed, dev)
    if args.skew:
        # low-complexity stretches: poly-A and (CA)n blocks make oversize buckets
        k = args.n // 50
        d_text[1000:1000 + k] = ord("A")
        ca = torch.tensor(list(b"CA"), dtype=torch.uint8, device=dev).repeat(k // 2)
        d_text[3 * k:3 * k + ca.numel()] = ca
    n = int(d_text.numel())
    step = IndexStep(torch, L, _lib, n, dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ref = None
    out = {"lib": os.environ.get("BWTK_LIB", "default"), "n": n}
    for mode in args.modes.split(","):
        for k in ("BWTK_MSD", "BWTK_MSD_FUSE"):
            os.environ.pop(k, None)
        os.environ.update(MODES[mode])
        for _ in range(3):
            step.run(d_text)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step.run(d_text)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        L.bwtk_profile_enable(1)
        step.run(d_text)
        prof = profile_report(L)
        L.bwtk_profile_enable(0)
        cur = (step.sa.clone(), step.bwt.clone(), step.lcp.clone())
        same = None
        if ref is None:
            ref = cur
        else:
            same = all(bool(torch.equal(x, y)) for x, y in zip(ref, cur))
        out[mode] = {"ms": round(statistics.median(ts), 4), "min_ms": round(min(ts), 4), "same_as_first": same,
                     "path_bits": int(step.stats[6]), "rounds": int(step.stats[0]), "active0": int(step.stats[3]),
                     "kernels": {r["kernel"]: [r["launches"], round(r["ms"], 4)] for r in prof}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=46_709_983)
    ap.add_argument("--seed", type=int, default=21)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--skew", action="store_true")
    ap.add_argument("--modes", default="lsd,msd_unfused,msd_fused")
    ap.add_argument("--libs", nargs="*", default=["default"])
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args)
    for lib in args.libs:
        env = dict(os.environ)
        if lib != "default":
            env["BWTK_LIB"] = os.path.abspath(lib)
        cmd = [sys.executable, os.path.abspath(__file__), "--child", "--n", str(args.n), "--seed", str(args.seed),
               "--reps", str(args.reps), "--modes", args.modes] + (["--skew"] if args.skew else [])
        res = subprocess.run(cmd, env=env, capture_output=True, text=True)
        line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else ""
        try:
            d = json.loads(line)
        except Exception:
            print(lib, "FAILED", res.stdout[-1500:], res.stderr[-3000:])
            continue
        print(f"== {d['lib']}  n={d['n']}")
        for mode in args.modes.split(","):
            m = d[mode]
            top = sorted(m["kernels"].items(), key=lambda kv: -kv[1][1])
            print(f"  {mode:12s} {m['ms']:8.4f} ms (min {m['min_ms']:.4f}) same={m['same_as_first']} path={m['path_bits']} "
                  f"rounds={m['rounds']} active0={m['active0']}")
            print("     " + "  ".join(f"{k}:{v[1]:.3f}x{v[0]}" for k, v in top[:14]))


if __name__ == "__main__":
    main()
