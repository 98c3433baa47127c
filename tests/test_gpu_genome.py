"""genome.GenomeScanner: the per-contig unit (index + Tier 1 + Tier 2 rows) over a list of contigs, with the
uploads / downloads overlapped -- against the one-call-at-a-time launchers, the oracle, and (two GPUs) against
itself after an NCCL gather (bwt.py:3040-3141, 3863-3899)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.util import gen_contig

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _contigs():
    s = gen_contig(70_000, 7)
    s[20_000:27_000] = ord("N")
    return [gen_contig(50_000, 3).tobytes() + b"$", gen_contig(5_000, 4).tobytes() + b"$", s.tobytes() + b"$",
            b"ACGTRYKMSWNacgtn" * 300 + b"$", gen_contig(24_000, 9).tobytes() + b"$", b"AC" * 4000 + b"$"]


_CACHE = {}


def _expected(orc, text):
    if text not in _CACHE:
        _CACHE[text] = _expected_rows(orc, text)
    return _CACHE[text]


def _expected_rows(orc, text):
    sa = orc.suffix_array(text)
    lcp = orc.kasai_lcp(text, sa)
    t1 = orc.tier1_scan(text)
    mask = np.zeros(len(text) - 1, np.uint8)
    for a, b in t1[:, :2].tolist():
        mask[a:b] = 1
    n_seq = len(text) - 1
    unit_cap = max(120, min(n_seq // 3, 1000))
    return {"tier1": t1, "strict": orc.strict_scan(text, 1, unit_cap, 0, 3),
            "plateaus": orc.lcp_plateaus(text, sa, lcp, 1, 1000, 3)[0],
            "period": orc.period_scan(text, tier1_mask=mask)[0]}


@pytest.mark.parametrize("where", ["host", "device"])
def test_genome_scanner_rows_match_oracle(oracle, where):
    import torch

    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200.genome import KINDS, GenomeScanner

    texts = _contigs()
    if where == "host":
        ins = [torch.from_numpy(np.frombuffer(t, np.uint8).copy()).pin_memory() for t in texts]
    else:
        ins = [torch.from_numpy(np.frombuffer(t, np.uint8).copy()).cuda() for t in texts]
    sc = GenomeScanner(max(len(t) for t in texts), arena_rows=800_000)
    for rep in range(2):                                  # a second pass reuses every buffer and event
        res = sc.scan(ins, ids=list(range(10, 10 + len(texts))), download=(where == "host"))
        assert [r.contig for r in res] == list(range(10, 10 + len(texts)))
        assert sc.elapsed_ms > 0
        for r, t in zip(res, texts):
            want = _expected(oracle, t)
            for k in KINDS:
                got = sc.host_rows(r, k) if where == "host" else sc.device_rows(r, k).cpu().numpy()
                assert np.array_equal(got, want[k]), f"contig of {len(t)} symbols: {k} rows differ"
        if where == "host":
            assert sc.h2d_bytes == sum(len(t) for t in texts) and sc.d2h_bytes == sc.used * 32
    with pytest.raises(Exception):
        GenomeScanner(1000, arena_rows=10).scan([torch.from_numpy(np.frombuffer(texts[0], np.uint8).copy())])
    small = GenomeScanner(len(texts[0]), arena_rows=16)
    with pytest.raises(Exception, match="arena"):
        small.scan([ins[0]], download=False)


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import bwt_algorithm_b200
from bwt_algorithm_b200 import sharding
from bwt_algorithm_b200.genome import KINDS, GenomeScanner
from tests.test_gpu_genome import _contigs
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
texts = _contigs()
bins = sharding.lpt_partition([len(t) for t in texts], world)
mine = bins[rank]
pin = lambda t: torch.from_numpy(np.frombuffer(t, np.uint8).copy()).pin_memory()
sc = GenomeScanner(max(len(t) for t in texts), arena_rows=800_000, device=dev)
res = sc.scan([pin(texts[i]) for i in mine], ids=mine)
table = [(r.contig, KINDS.index(k), a, c) for r in res for k, (a, c) in r.span.items()]
rows, merged = sharding.gather_rows_to_rank0(sc.arena[: sc.used], table)
if rank == 0:
    rows = rows.cpu().numpy()
    got = {{(cid, kind): rows[a:a + c] for cid, kind, a, c in merged}}
    res1 = sc.scan([pin(t) for t in texts], ids=list(range(len(texts))))       # the whole genome on one GPU
    assert len(got) == len(texts) * len(KINDS)
    total = 0
    for r in res1:
        for k in KINDS:
            want = sc.host_rows(r, k)
            assert np.array_equal(got[(r.contig, KINDS.index(k))], want), (r.contig, k)
            total += len(want)
    assert total == rows.shape[0] and total > 1000
    print("GENOME_GATHER_OK", total, flush=True)
else:
    assert rows is None
dist.barrier()
dist.destroy_process_group()
"""


def test_two_gpu_genome_gather_nccl(tmp_path):
    """Contigs dealt to two ranks by LPT, rows gathered to rank 0 over NCCL == the single-GPU rows."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "GENOME_GATHER_OK" in res.stdout


def test_bench_genome_smoke(tmp_path):
    """bench.py end to end on a 1/2000-scale genome: one JSON line with the contract's keys."""
    import json

    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--scale", "0.0005", "--steps", "1",
                          "--warmup", "1", "--fm-log2-queries", "16", "--no-cpu"], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "e2e", "roofline", "gpu_launches", "clocks", "config"):
        assert key in line, key
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] > 0 and line["rows"]["strict"] > 0
    assert line["fm_search_chr1"]["random_10mers"]["queries"] == 1 << 16
