"""One contig on several GPUs: the FM index built on rank 0 is replicated over NCCL and a query
batch is split across the ranks.  The single-GPU part (an ``FMReplica`` made from a ``DeviceIndex``'s arrays
answers like the index itself) runs on any box; the two-rank part needs two GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.util import gen_contig

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _queries(rng, nq):
    lens = rng.integers(0, 13, nq).astype(np.int32)
    pats = np.frombuffer(b"ACGTN", np.uint8)[rng.integers(0, 5, (nq, 12))]
    pats[: nq // 2][pats[: nq // 2] == ord("N")] = ord("A")     # half of the batch is pure ACGT
    return np.ascontiguousarray(pats), lens


def test_replica_answers_like_the_index(oracle):
    import torch

    import bwt_algorithm_b200  # noqa: F401
    from bwt_algorithm_b200.device_index import DeviceIndex, FMReplica

    text = gen_contig(200_000, 5).tobytes() + b"$"
    ix = DeviceIndex(text, build_kmer=False)
    rep = FMReplica(ix.bwt.clone(), ix.occ.clone(), ix.totals, ix.row_of_code, ix.n, ix.occ_rate)
    pats, lens = _queries(np.random.default_rng(3), 20_001)
    d_p, d_l = torch.from_numpy(pats).to(ix.device), torch.from_numpy(lens).to(ix.device)
    want_sp, want_ep = oracle.OracleIndex(text).backward_search_batch(pats, lens)
    for lo, hi in ((0, 20_001), (7, 7), (5, 6), (3333, 20_001)):     # blocks start at odd byte offsets
        sp, ep = rep.search_block(d_p[lo:hi], d_l[lo:hi])
        assert np.array_equal(sp.cpu().numpy(), want_sp[lo:hi]) and np.array_equal(ep.cpu().numpy(), want_ep[lo:hi])
    assert np.array_equal(rep.counts, ix.counts) and rep.codes == ix.codes


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import bwt_algorithm_b200
from bwt_algorithm_b200.device_index import DeviceIndex, FMReplica
from bwt_algorithm_b200.sharding import broadcast_fm, sharded_search
from oracle import oracle as orc
from tests.util import gen_contig
from tests.test_gpu_sharded_search import _queries
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank()
text = gen_contig(300_000, 9).tobytes() + b"$"
built = DeviceIndex(text, build_kmer=False) if rank == 0 else None
rep = FMReplica(device=dev, **broadcast_fm(built, src=0, device=dev))
oi = orc.OracleIndex(text)
assert np.array_equal(rep.bwt.cpu().numpy(), oi.bwt)
for nq in (0, 1, 3, 50_001):
    pats, lens = _queries(np.random.default_rng(nq), nq)
    sp, ep = sharded_search(rep.search_block, torch.from_numpy(pats).to(dev), torch.from_numpy(lens).to(dev))
    if nq:
        wsp, wep = oi.backward_search_batch(pats, lens)
        assert np.array_equal(sp.cpu().numpy(), wsp) and np.array_equal(ep.cpu().numpy(), wep), nq
    else:
        assert sp.numel() == 0 and ep.numel() == 0
torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    print("SHARDED_SEARCH_OK", flush=True)
dist.destroy_process_group()
"""


def test_two_gpu_sharded_search_nccl(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "SHARDED_SEARCH_OK" in res.stdout
