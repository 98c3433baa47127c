"""BASELINE.json configs[4]: a synthetic 3.1 Gb, 24-contig human-scale genome (hg38 contig lengths),
indexed (SA + BWT + Occ + LCP) and Tier 1 + strict-adjacency scanned, contigs dealt to the GPUs by
LPT (sharding.lpt_partition).  Rows stay NumPy (the C-ABI result); no Python records are built.

    python tools/genome_bench.py                       # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29533 tools/genome_bench.py      # contigs sharded over 8 GPUs

Contigs are generated on the device (random ACGT + planted imperfect tandem arrays every ~900 bp,
the density of SURVEY Appendix B's generator, vectorised)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
        133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
        58617616, 64444167, 46709983, 50818468, 156040895, 57227415]


def device_contig(torch, n, seed, dev):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    text = lut[torch.randint(0, 4, (n,), generator=g, device=dev)]
    K = max(1, n // 970)
    gaps = torch.randint(300, 1500, (K,), generator=g, device=dev)
    k = torch.randint(1, 10, (K,), generator=g, device=dev)
    c = torch.randint(4, 25, (K,), generator=g, device=dev)
    lens = k * c
    starts = 500 + torch.cumsum(gaps + lens, 0) - lens
    keep = starts + lens < n - 400
    starts, k, lens = starts[keep], k[keep], lens[keep]
    motifs = torch.randint(0, 4, (starts.numel(), 9), generator=g, device=dev)
    arr = torch.repeat_interleave(torch.arange(starts.numel(), device=dev), lens)
    first = torch.cumsum(lens, 0) - lens
    j = torch.arange(arr.numel(), device=dev) - first[arr]
    val = motifs[arr, j % k[arr]]
    mut = torch.rand(arr.numel(), generator=g, device=dev) < 0.03
    val = torch.where(mut, torch.randint(0, 4, (arr.numel(),), generator=g, device=dev), val)
    text[starts[arr] + j] = lut[val]
    return torch.cat([text, torch.tensor([36], dtype=torch.uint8, device=dev)])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every contig (smoke runs)")
    ap.add_argument("--no-scans", action="store_true")
    ap.add_argument("--cold", action="store_true", help="skip the untimed allocator warm-up on the largest contig")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import bwt_algorithm_b200  # noqa: F401
    from bench import IndexStep
    from bwt_algorithm_b200 import _lib, detect, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lengths = [max(1000, int(x * args.scale)) for x in HG38]
    bins = sharding.lpt_partition(lengths, world)
    mine = bins[rank]
    L = _lib.lib()
    # warm-up on a small contig (module load, first-call attributes)
    w = device_contig(torch, 2_000_000, 1, dev)
    IndexStep(torch, L, _lib, int(w.numel()), dev).run(w)
    detect.tier1_rows(w)
    detect.strict_rows(w, 1, 1000, 0, 3)
    torch.cuda.synchronize()
    texts = {i: device_contig(torch, lengths[i], 100 + i, dev) for i in mine}
    torch.cuda.synchronize()
    if not args.cold and mine:
        # allocator warm-up (device and pinned host blocks of the largest sizes), untimed
        big = texts[max(mine, key=lambda i: lengths[i])]
        IndexStep(torch, L, _lib, int(big.numel()), dev).run(big)
        if not args.no_scans:
            detect.tier1_rows(big)
            detect.strict_rows(big, 1, 1000, 0, 3)
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    per = []
    t_index = t_t1 = t_strict = 0.0
    rows_t1 = rows_strict = 0
    for i in mine:
        d_text = texts[i]
        n = int(d_text.numel())
        a = time.perf_counter()
        step = IndexStep(torch, L, _lib, n, dev)
        step.run(d_text)
        torch.cuda.synchronize()
        b = time.perf_counter()
        r1 = rs = None
        if not args.no_scans:
            r1 = detect.tier1_rows(d_text)
            torch.cuda.synchronize()
        c = time.perf_counter()
        if not args.no_scans:
            rs = detect.strict_rows(d_text, 1, 1000, 0, 3)
            torch.cuda.synchronize()
        d = time.perf_counter()
        t_index += b - a
        t_t1 += c - b
        t_strict += d - c
        rows_t1 += 0 if r1 is None else len(r1)
        rows_strict += 0 if rs is None else len(rs)
        per.append({"contig": i, "bases": n - 1, "index_ms": round((b - a) * 1e3, 2), "tier1_ms": round((c - b) * 1e3, 2),
                    "strict_ms": round((d - c) * 1e3, 2)})
        del step
    torch.cuda.synchronize()
    mine_s = time.perf_counter() - t0
    stats = torch.tensor([mine_s, t_index, t_t1, t_strict, float(rows_t1), float(rows_strict),
                          float(sum(lengths[i] for i in mine))], dtype=torch.float64, device=dev)
    if world > 1:
        allst = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allst, stats)
    else:
        allst = [stats]
    if rank == 0:
        wall = max(float(s[0]) for s in allst)
        total = sum(lengths)
        out = {"workload": "24 synthetic contigs with hg38 lengths (BASELINE configs[4])", "bases": total, "n_gpus": world,
               "scale": args.scale, "wall_s_max_over_ranks": round(wall, 3),
               "gbases_per_s": round(total / wall / 1e9, 3),
               "index_s_sum": round(sum(float(s[1]) for s in allst), 3),
               "tier1_s_sum": round(sum(float(s[2]) for s in allst), 3),
               "strict_s_sum": round(sum(float(s[3]) for s in allst), 3),
               "tier1_rows": int(sum(float(s[4]) for s in allst)), "strict_rows": int(sum(float(s[5]) for s in allst)),
               "per_rank_bases": [int(float(s[6])) for s in allst], "per_rank_s": [round(float(s[0]), 3) for s in allst],
               "rank0_contigs": per}
        # honest end-to-end figure against the aggregate HBM roofline: compulsory bytes only (SURVEY 8d:
        # index 20.2 B/base; Tier 1 and the strict scan read the text once each, 1 B/base + rows)
        from bench import measured_peak
        peak, src = measured_peak()
        per_base = 20.2 + (0.0 if args.no_scans else 2.0)
        out["compulsory_bytes_per_base"] = per_base
        out["compulsory_fraction_of_aggregate_hbm"] = round(per_base * total / wall / 1e9 / (peak * world), 5)
        out["hbm_peak_gbs_per_gpu"] = [peak, src]
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
