import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bwt_algorithm_b200
from bench import gen_contig
from bwt_algorithm_b200.streaming import IndexPipeline
n = 46_709_984
text = np.concatenate([gen_contig(n - 1, 21), np.frombuffer(b"$", np.uint8)])
pinned = torch.from_numpy(text).pin_memory()
pipe = IndexPipeline(n, slots=2)
for _ in range(3): pipe.result(pipe.submit(pinned))
pipe.drain(); torch.cuda.synchronize()
for trial in range(2):
    ts = []
    t0 = time.perf_counter()
    for i in range(6):
        a = time.perf_counter(); tk = pipe.submit(pinned); b = time.perf_counter()
        ts.append((a - t0, b - a))
    pipe.drain(); t1 = time.perf_counter()
    print("submit start/duration ms:", [(round(x*1e3,2), round(y*1e3,2)) for x, y in ts], "total", round((t1-t0)*1e3,2))
    for s in pipe.slots:
        print("  slot start->done ms", round(s.start.elapsed_time(s.done), 2))
