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


def test_start_up_path_without_torch_needs_a_gpu_too():
    """lean.py (the CLI's detector call on the library's own buffers while torch is not loaded) is a start-up path,
    not a fallback: without a device it reports zero devices and the worker goes on to the torch path, which raises;
    its own calls fail loudly; and it steps aside as soon as torch is in the process or BWTK_LEAN=0."""
    import subprocess
    import sys

    import torch

    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200 import _lib, lean, pipeline

    assert "torch" in sys.modules and not lean.enabled()
    probe = ("import sys; sys.path.insert(0, %r); import bwt_algorithm_b200; from bwt_algorithm_b200 import lean; "
             "print(lean.enabled(), lean.device_count(), 'torch' in sys.modules)" % ROOT)
    out = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "True" and out[2] == "False"
    off = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True, check=True,
                         env=dict(os.environ, BWTK_LEAN="0")).stdout.split()
    assert off[0] == "False"
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert out[1] == "0" and lean.device_count() == 0
    with pytest.raises(_lib.BwtkError):
        lean.strict_rows(np.frombuffer(b"ACACACACAC$", np.uint8), 1, 4, 0, 3)
    with pytest.raises(_lib.BwtkError):
        pipeline._detect_rows("c", "ACACACACACAC", {"sa_sample_rate": 32, "enable_tier2": True, "min_copies": 3,
                                                    "max_unit_len": 120})


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
import torch
from bwt_algorithm_b200.sharding import gather_rows, gather_rows_to_rank0, lpt_partition
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
else:
    assert out is None
# the exact-size point-to-point form (rows stay tensors; NCCL between GPUs, gloo here)
table, first = [], 0
for c in mine:
    table.append((c, 1, first, lens[c] // 100)); first += lens[c] // 100
got, merged = gather_rows_to_rank0(torch.from_numpy(rows), table)
if rank == 0:
    assert got.shape == (22, 8) and sorted(e[0] for e in merged) == [0, 1, 2, 3, 4]
    for cid, kind, a, cnt in merged:
        blk = got[a:a + cnt].numpy()
        assert kind == 1 and cnt == lens[cid] // 100 and (blk[:, 2] == cid).all()
        assert (blk[:, 0] == np.arange(cnt) * 10 + cid).all()
    print("GATHER_OK", flush=True)
else:
    assert got is None and merged is None
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


_SEARCH_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import bwt_algorithm_b200
from bwt_algorithm_b200.sharding import broadcast_fm, sharded_search, split_range
from oracle import oracle as orc
from tests.util import gen_contig
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
assert split_range(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)] and split_range(2, 4)[2:] == [(2, 2), (2, 2)]
assert split_range(0, 3) == [(0, 0)] * 3


class Built:                      # what rank 0's DeviceIndex exposes, filled from the oracle
    def __init__(self, text):
        oi = orc.OracleIndex(text)
        codes = [c for c in range(256) if oi.totals[c] > 0]
        self.n, self.occ_rate, self.totals = oi.n, 128, oi.totals
        self.row_of_code = np.full(256, -1, np.int32)
        self.row_of_code[codes] = np.arange(len(codes))
        self.bwt = torch.from_numpy(oi.bwt.copy())
        self.occ = torch.from_numpy(np.ascontiguousarray(oi.occ_matrix()[codes]))


text = gen_contig(30_000, 77).tobytes() + b"$"
fm = broadcast_fm(Built(text) if rank == 0 else None, src=0)
assert fm["n"] == len(text) and fm["bwt"].shape == (len(text),) and fm["occ"].shape[0] == 5

# a CPU search bound to the replica arrays only (the GPU ranks bind FMReplica.search_block)
totals = fm["totals"]
counts = np.concatenate(([0], np.cumsum(totals)[:-1])).astype(np.int64)
occ256 = np.zeros((256, fm["occ"].shape[1]), np.int32)
for code in range(256):
    if fm["row"][code] >= 0:
        occ256[code] = fm["occ"][fm["row"][code]].numpy()
bwt = fm["bwt"].numpy()


def search(pats, lens):
    pats = np.ascontiguousarray(pats.numpy()); lens = np.ascontiguousarray(lens.numpy())
    sp = np.empty(lens.size, np.int64); ep = np.empty(lens.size, np.int64)
    orc.lib().orc_backward_search_batch(orc._ptr(bwt), fm["n"], fm["occ_rate"], orc._ptr(occ256), occ256.shape[1],
                                        orc._ptr(totals), orc._ptr(counts), orc._ptr(pats), pats.shape[1],
                                        orc._ptr(lens), lens.size, orc._ptr(sp), orc._ptr(ep))
    return torch.from_numpy(sp.astype(np.int32)), torch.from_numpy(ep.astype(np.int32))


rng = np.random.default_rng(5)
for nq in (0, 1, 2, 3, 1001):
    lens = rng.integers(0, 11, nq).astype(np.int32)
    pats = np.frombuffer(b"ACGTN", np.uint8)[rng.integers(0, 5, (nq, 10))]
    sp, ep = sharded_search(search, torch.from_numpy(pats), torch.from_numpy(lens))
    full = orc.OracleIndex(text)
    wsp, wep = full.backward_search_batch(pats, lens) if nq else (np.zeros(0), np.zeros(0))
    assert np.array_equal(sp.numpy(), wsp) and np.array_equal(ep.numpy(), wep), nq
    if nq > 100:
        assert (sp.numpy() >= 0).sum() > 100 and (sp.numpy() < 0).sum() > 10
dist.barrier()
if rank == 0:
    print("SEARCH_OK", flush=True)
dist.destroy_process_group()
"""


def test_two_rank_sharded_search_gloo(tmp_path):
    """One contig, two ranks: the FM arrays are broadcast from the rank that built them, the query batch is
    split, and every rank ends up with the (sp, ep) of every query in query order."""
    script = tmp_path / "worker.py"
    script.write_text(_SEARCH_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29519", CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29519", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "SEARCH_OK" in res.stdout


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) on a tiny sample."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-window", "3000", "--cpu-cores", "4", "--no-python-ref"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "SA+BWT+LCP Gbases/s" and line["unit"] == "Gbases/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Gbases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
