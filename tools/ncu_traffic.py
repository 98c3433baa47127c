"""ncu launch lists (csv) of `tools/ncu_step.py --profile index|scans|fm` -> profiles/r2_ncu_traffic.json:
per kernel of the measured index build the launches, device time and DRAM bytes (read + write), the whole
step's traffic fraction (sum of DRAM bytes / sum of durations / measured copy peak over EVERY
launch of the step), and the same per-kernel table with the L2 figures for the scans and the FM search.

    python tools/ncu_traffic.py profiles/r2_ncu_traffic.json index=gpurun_out/r2_ncu_index.csv fm=gpurun_out/r2_ncu_fm.csv
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHORT = [("bucket_sort_kernel<1>", "bucket_sort_regroup"), ("bucket_sort_kernel<0>", "bucket_sort"),
         ("bucket_sort_kernel<(bool)1>", "bucket_sort_regroup"), ("bucket_sort_kernel<true>", "bucket_sort_regroup"),
         ("bucket_sort_kernel<(bool)0>", "bucket_sort"), ("bucket_sort_kernel<false>", "bucket_sort"),
         ("onesweep_kernel<unsigned int, bwtk::rsort::PackedSuffixSource", "onesweep_u32_gen"),
         ("onesweep_kernel<unsigned int, rsort::PackedSuffixSource", "onesweep_u32_gen"),
         ("onesweep_kernel<unsigned int, PackedSuffixSource", "onesweep_u32_gen"),
         ("onesweep_kernel<unsigned int", "onesweep_u32"), ("onesweep_kernel<unsigned long", "onesweep_u64"),
         ("regroup_kernel<unsigned int", "regroup_first"), ("regroup_kernel<unsigned long", "regroup_round"),
         ("hist_kernel<unsigned long", "radix_hist_u64"), ("hist_kernel<unsigned int", "radix_hist_u32"),
         ("search_coop_kernel", "fm_search_coop_kernel"), ("search_kernel", "fm_search_thread_kernel"),
         ("bsearch_kernel", "bsearch_kernel_byte_bwt")]


def short_name(full):
    for pat, nm in SHORT:
        if pat in full:
            return nm
    m = re.search(r"([A-Za-z_0-9]+)(<.*>)?\(", full)
    return m.group(1) if m else full[:40]


def read_csv(path):
    rows = {}
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        d = rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    return [rows[i] for i in sorted(rows)]


def summarise(launches):
    per = {}
    for r in launches:
        if "at::" in r["name"] or "elementwise" in r["name"]:
            continue                       # torch's own kernels (fills, copies) are not part of the step
        k = per.setdefault(short_name(r["name"]), {"launches": 0, "ns": 0.0, "dram_read": 0.0, "dram_write": 0.0,
                                                  "lts_sectors": 0.0, "hit_w": 0.0})
        k["launches"] += 1
        k["ns"] += r.get("gpu__time_duration.sum", 0.0)
        k["dram_read"] += r.get("dram__bytes_read.sum", 0.0)
        k["dram_write"] += r.get("dram__bytes_write.sum", 0.0)
        k["lts_sectors"] += r.get("lts__t_sectors.sum", 0.0)
        k["hit_w"] += r.get("lts__t_sector_hit_rate.pct", 0.0) * r.get("lts__t_sectors.sum", 0.0)
    out = {}
    for nm, k in per.items():
        tb = k["dram_read"] + k["dram_write"]
        out[nm] = {"launches": k["launches"], "duration_us_under_ncu": round(k["ns"] / 1e3, 2),
                   "traffic_bytes_per_launch": int(tb / k["launches"]), "dram_read_bytes": int(k["dram_read"]),
                   "dram_write_bytes": int(k["dram_write"]), "dram_gbs": round(tb / k["ns"], 1) if k["ns"] else None,
                   "lts_sectors": int(k["lts_sectors"]),
                   "l2_hit_rate_pct": round(k["hit_w"] / k["lts_sectors"], 2) if k["lts_sectors"] else None}
    tot_ns = sum(k["ns"] for k in per.values())
    tot_b = sum(k["dram_read"] + k["dram_write"] for k in per.values())
    n_l = sum(k["launches"] for k in per.values())
    return out, tot_ns, tot_b, n_l


def main():
    out_path = sys.argv[1]
    srcs = dict(a.split("=", 1) for a in sys.argv[2:])
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
    except Exception:
        peak = 6650.0
    result = {"_note": "`ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
                       "gpu__time_duration.sum,lts__t_sectors.sum,lts__t_sector_hit_rate.pct --clock-control none` over "
                       "`python tools/ncu_step.py --profile <phase>` on a B200 (cold-cache, serialised launches: shares, "
                       "not absolutes); bench.py copies traffic_bytes_per_launch of the dominant kernel into "
                       "roofline.traffic", "peak_gbs": peak}
    if "index" in srcs:
        idx, ns, by, n_l = summarise(read_csv(srcs["index"]))
        result.update(idx)
        result["traffic_fraction"] = {
            "launches": n_l, "sum_duration_us": round(ns / 1e3, 1), "sum_dram_bytes": int(by),
            "dram_gbs_time_weighted": round(by / ns, 1) if ns else None, "frac": round(by / ns / peak, 4) if ns else None,
            "note": "EVERY launch of one index build (SA + BWT + Occ + LCP) of the chr21-sized contig"}
    for ph in ("scans", "fm"):
        if ph in srcs:
            d, ns, by, n_l = summarise(read_csv(srcs[ph]))
            result["phase_" + ph] = {"launches": n_l, "sum_duration_us": round(ns / 1e3, 1), "sum_dram_bytes": int(by),
                                     "kernels": d}
    with open(out_path, "w") as f:
        json.dump(result, f, indent=1)
    print(json.dumps(result.get("traffic_fraction")))


if __name__ == "__main__":
    main()
