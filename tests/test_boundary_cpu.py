"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/bwtk.h declares, the product refuses to run without a GPU
(no CPU fallback, no oracle on the product path), and the contig sharding +
record gather work across two ranks (gloo)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import _lib

    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    header = open(os.path.join(ROOT, "include", "bwtk.h")).read()
    declared = set(re.findall(r"\b(bwtk_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/bwtk.h but not exported by libbwtk.so"
    assert declared == set(_lib.EXPORTS), "ctypes signature table and header disagree"
    assert L.bwtk_version() >= 100
    assert L.bwtk_sa_workspace_bytes(1_000_000) > 36_000_000
    assert L.bwtk_launch_count() == 0


def test_product_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import _lib, bwt

    with pytest.raises(_lib.BwtkError):
        bwt.BWTCore("ACGTACGT$")
    with pytest.raises(_lib.BwtkError):
        bwt.Tier1STRFinder(np.frombuffer(b"ACACACACAC", np.uint8)).find_strs("c")
    from bwt_algorithm_b200.streaming import IndexPipeline

    with pytest.raises(_lib.BwtkError):
        IndexPipeline(1000)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bwt-algorithm_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower() or f == "never", f"{f} mentions the oracle"


def test_lpt_partition_balances_hg38_lengths():
    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200.sharding import lpt_partition

    lens = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
            133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616,
            64444167, 46709983, 50818468, 156040895, 57227415]
    bins = lpt_partition(lens, 8)
    assert sorted(i for b in bins for i in b) == list(range(24))
    loads = [sum(lens[i] for i in b) for b in bins]
    assert max(loads) / (sum(lens) / 8) < 1.05      # 96.5 % balance
    assert lpt_partition([5, 1], 4)[0] == [0]


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
import bwt_algorithm_b200
from bwt_algorithm_b200.sharding import gather_rows, lpt_partition
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lens = [900, 500, 400, 300, 100]
mine = lpt_partition(lens, world)[rank]
rows, ids = [], []
for c in mine:
    k = lens[c] // 100
    r = np.zeros((k, 8), np.int32)
    r[:, 0] = np.arange(k) * 10 + c
    r[:, 1] = r[:, 0] + 6
    r[:, 2] = c
    rows.append(r); ids.append(np.full(k, c, np.int32))
rows = np.concatenate(rows) if rows else np.zeros((0, 8), np.int32)
ids = np.concatenate(ids) if ids else np.zeros(0, np.int32)
out = gather_rows(rows, ids)
if rank == 0:
    allrows, allids = out
    assert allrows.shape == (22, 8), allrows.shape
    for c, L in enumerate(lens):
        sel = allrows[allids == c]
        assert sel.shape[0] == L // 100 and (sel[:, 2] == c).all()
    print("GATHER_OK", flush=True)
else:
    assert out is None
dist.destroy_process_group()
"""


def test_two_rank_record_gather_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517", CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "GATHER_OK" in res.stdout


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) on a tiny sample."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "40000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "SA+BWT+LCP Gbases/s" and line["unit"] == "Gbases/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
