"""Where does a strict scan's wall time go, contig by contig: kernel (device-timed), pinned
allocation, download?  Prints one line per contig of the genome bench's 1-GPU order."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bwt_algorithm_b200  # noqa: F401,E402
from bwt_algorithm_b200 import detect  # noqa: E402
from tools.genome_bench import HG38, device_contig  # noqa: E402

dev = torch.device("cuda", 0)
order = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,5,6,22,7,14,20").split(",")]
orig = detect._rows_to_host
stamp = {}


def timed_rows_to_host(t, d_rows):
    torch.cuda.synchronize()
    a = time.perf_counter()
    out = orig(t, d_rows)
    stamp["download_ms"] = (time.perf_counter() - a) * 1e3
    stamp["mb"] = d_rows.numel() * 4 / 1e6
    return out


detect._rows_to_host = timed_rows_to_host
w = device_contig(torch, HG38[order[0]], 100 + order[0], dev)
detect.strict_rows(w, 1, 1000, 0, 3)
del w
for i in order:
    text = device_contig(torch, HG38[i], 100 + i, dev)
    torch.cuda.synchronize()
    a = time.perf_counter()
    rows = detect.strict_rows(text, 1, 1000, 0, 3)
    torch.cuda.synchronize()
    total = (time.perf_counter() - a) * 1e3
    print(f"contig {i:2d} {HG38[i]/1e6:6.1f} Mb rows {len(rows):9d} ({stamp['mb']:6.1f} MB) total {total:7.2f} ms "
          f"download {stamp['download_ms']:7.2f} ms rest {total - stamp['download_ms']:7.2f} ms", flush=True)
    del rows, text
