"""Per-contig timeline of the two-slot index pipeline (upload, build, download) from CUDA events:
shows which builds overlap the other slot's download."""
import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bwt_algorithm_b200
from bench import gen_contig
from bwt_algorithm_b200 import _lib
from bwt_algorithm_b200.streaming import IndexPipeline
print("asyncEngineCount", torch.cuda.get_device_properties(0))
n = 46_709_984
text = np.concatenate([gen_contig(n - 1, 21), np.frombuffer(b"$", np.uint8)])
pinned = torch.from_numpy(text).pin_memory()
pipe = IndexPipeline(n, slots=2)
L = _lib.lib()
for _ in range(3): pipe.result(pipe.submit(pinned))
pipe.drain(); torch.cuda.synchronize()
ev = []
t0 = time.perf_counter()
for i in range(4):
    slot = pipe.slots[i % 2]
    if slot.busy: slot.done.synchronize()
    slot.busy = True
    E = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.cuda.stream(slot.stream):
        E[0].record()
        slot.d_text.copy_(pinned, non_blocking=True)
        E[1].record()
        h0 = time.perf_counter()
        rc = L.bwtk_index_build(slot.d_text.data_ptr(), n, 128, slot.sa.data_ptr(), None, slot.bwt.data_ptr(), slot.occ.data_ptr(), 8,
                                slot.lcp.data_ptr(), slot.totals.ctypes.data, slot.row.ctypes.data, slot.stats.ctypes.data,
                                slot.ws.data_ptr(), slot.ws.numel(), slot.stream.cuda_stream)
        h1 = time.perf_counter()
        E[2].record()
        slot.h_sa.copy_(slot.sa, non_blocking=True); slot.h_bwt.copy_(slot.bwt, non_blocking=True); slot.h_lcp.copy_(slot.lcp, non_blocking=True)
        E[3].record(); slot.done.record()
    ev.append((E, (h0 - t0) * 1e3, (h1 - h0) * 1e3))
torch.cuda.synchronize()
base = ev[0][0][0]
for i, (E, hs, hd) in enumerate(ev):
    print(i, "host build start %.2f dur %.2f |" % (hs, hd), "gpu: h2d %.2f-%.2f build-end %.2f d2h-end %.2f" % tuple(base.elapsed_time(e) for e in E))
