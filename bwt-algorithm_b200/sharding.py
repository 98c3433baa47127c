"""Contig sharding over the GPUs of one box.

Every index array and detector is per contig (bwt.py:3040-3054), so the path
shards by dealing contigs to GPUs with longest-processing-time bin packing and
needs no data-path collective.  The only exchange is the gather of fixed-width
int32 record rows to rank 0 (``gather_rows``; NCCL over NVLink when the ranks
own GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Iterator, List, Sequence, Tuple

import numpy as np

REC_W = 8


def lpt_partition(lengths: Sequence[int], nbins: int) -> List[List[int]]:
    """Indices of `lengths` dealt to `nbins` bins, longest first into the lightest bin."""
    nbins = max(1, int(nbins))
    bins: List[List[int]] = [[] for _ in range(nbins)]
    load = [0] * nbins
    for idx in sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i)):
        b = min(range(nbins), key=lambda j: (load[j], j))
        bins[b].append(idx)
        load[b] += int(lengths[idx])
    return bins


def _gpu_count() -> int:
    import sys

    if "torch" not in sys.modules:            # the CLI's start-up path (lean.py): ask the library, not torch
        from . import lean

        if lean.enabled():
            return lean.device_count()
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _pool_init(counter, ngpu):
    import torch

    with counter.get_lock():
        wid = counter.value
        counter.value += 1
    torch.cuda.set_device(wid % ngpu)


def _run_bin(args):
    worker, tasks = args
    return [worker(t) for t in tasks]


def worker_processes(n_tasks: int, n_jobs: int) -> int:
    """How many worker processes ``run_tasks`` uses for `n_tasks` contigs (1 = in-process)."""
    return max(1, min(_gpu_count(), int(n_tasks), max(1, int(n_jobs or 1))))


def run_tasks(worker: Callable, tasks: List[Tuple], n_jobs: int) -> Iterator[list]:
    """Yields worker(task) results (any order).  One process per GPU when several
    GPUs are visible and there are several contigs; otherwise in-process."""
    ngpu = worker_processes(len(tasks), n_jobs)
    if ngpu <= 1:
        for t in tasks:
            yield worker(t)
        return
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    bins = lpt_partition([len(t[1]) for t in tasks], ngpu)
    counter = ctx.Value("i", 0)
    with ctx.Pool(ngpu, initializer=_pool_init, initargs=(counter, ngpu)) as pool:
        for results in pool.imap_unordered(_run_bin, [(worker, [tasks[i] for i in b]) for b in bins if b]):
            for r in results:
                yield r


def gather_rows(rows: np.ndarray, contig_ids: np.ndarray, group=None, device=None):
    """All ranks pass their record rows (int32[R, 8]) and the contig id of each
    row; rank 0 receives the concatenation ordered by (rank, row), the others
    None.  One all_gather of the counts, one all_gather of padded blocks."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = np.ascontiguousarray(rows, np.int32).reshape(-1, REC_W)
    ids = np.ascontiguousarray(contig_ids, np.int32).reshape(-1)
    assert ids.size == rows.shape[0]
    dev = torch.device(device) if device is not None else torch.device("cpu")
    count = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    block = torch.zeros((cap, REC_W + 1), dtype=torch.int32, device=dev)
    if rows.shape[0]:
        block[: rows.shape[0], :REC_W] = torch.from_numpy(rows).to(dev)
        block[: rows.shape[0], REC_W] = torch.from_numpy(ids).to(dev)
    blocks = [torch.zeros_like(block) for _ in range(world)]
    dist.all_gather(blocks, block, group=group)
    if rank != 0:
        return None
    parts = [b[:c].cpu().numpy() for b, c in zip(blocks, counts)]
    merged = np.concatenate(parts, axis=0) if parts else np.zeros((0, REC_W + 1), np.int32)
    return merged[:, :REC_W].copy(), merged[:, REC_W].copy()


def gather_rows_to_rank0(rows, table: Sequence[Sequence[int]], group=None, out=None):
    """Device-side form of ``gather_rows`` for a genome's worth of rows (the merge of the per-contig
    results, bwt.py:3896-3899): every rank passes ONE tensor ``rows`` (int32[R, 8], on its GPU under
    NCCL, on the CPU under gloo) and a small ``table`` of (contig id, kind, first row, count) entries
    describing it.  Rank 0 receives the rows of all ranks back to back in one tensor on its own device
    (written straight into ``out`` when given) plus the merged table with row offsets rebased; the
    other ranks get (None, None).  Exact-size point-to-point transfers (over NVLink between GPUs), no
    padding and no copy to ranks that do not need the rows: one all_gather of the counts, one object
    gather of the tables, then one batched send/recv."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rows = rows.reshape(-1, REC_W)
    dev = rows.device
    count = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    tables = [None] * world
    dist.all_gather_object(tables, [tuple(int(x) for x in e) for e in table], group=group)
    starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    if rank != 0:
        if counts[rank]:
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, rows.contiguous(), 0, group=group)]):
                w.wait()
        return None, None
    total = int(starts[-1])
    if out is None:
        out = torch.empty((total, REC_W), dtype=torch.int32, device=dev)
    elif out.shape[0] < total:
        raise ValueError(f"gather_rows_to_rank0: out holds {out.shape[0]} rows, {total} needed")
    out[: counts[0]].copy_(rows)
    ops = [dist.P2POp(dist.irecv, out[int(starts[r]): int(starts[r + 1])], r, group=group)
           for r in range(1, world) if counts[r]]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    merged = [(cid, kind, first + int(starts[r]), cnt) for r in range(world) for cid, kind, first, cnt in tables[r]]
    return out[:total], merged


# ---------------------------------------------------------------------------
# one contig, many GPUs: the motif batch is split, the index is replicated
# ---------------------------------------------------------------------------
def split_range(count: int, parts: int) -> List[Tuple[int, int]]:
    """[lo, hi) of each of `parts` equal blocks of ceil(count/parts) items (the last ones may
    be short or empty) -- equal blocks keep the result gather one fixed-size all_gather."""
    parts = max(1, int(parts))
    block = -(-int(count) // parts) if count > 0 else 0
    return [(min(count, r * block), min(count, (r + 1) * block)) for r in range(parts)]


def broadcast_fm(index, src: int = 0, group=None, device=None) -> dict:
    """Replicates the search side of a contig's FM index -- BWT bytes, Occ checkpoints, byte
    totals and the byte->row map (bwt.py:266-326) -- from rank `src` to every rank: one small
    object broadcast for the shapes, then the two arrays (NCCL over NVLink when the ranks own
    GPUs; 1.25 B/base for an ACGT$ contig at the default checkpoint rate).  `index` is a
    ``DeviceIndex`` (or any object with bwt/occ/totals/row_of_code/n/occ_rate) on `src` and
    ignored elsewhere.  Returns the keyword arguments of ``FMReplica``."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    meta = [None]
    if rank == src:
        meta = [dict(n=int(index.n), occ_rate=int(index.occ_rate), occ_shape=tuple(index.occ.shape),
                     totals=np.asarray(index.totals, np.int64).tolist(),
                     row=np.asarray(index.row_of_code, np.int32).tolist())]
    dist.broadcast_object_list(meta, src=src, group=group)
    m = meta[0]
    if rank == src:
        bwt = torch.as_tensor(index.bwt).to(dev).contiguous()
        occ = torch.as_tensor(index.occ).to(dev).contiguous()
    else:
        bwt = torch.empty(m["n"], dtype=torch.uint8, device=dev)
        occ = torch.empty(m["occ_shape"], dtype=torch.int32, device=dev)
    if m["n"]:
        dist.broadcast(bwt, src=src, group=group)
    if occ.numel():
        dist.broadcast(occ, src=src, group=group)
    return dict(bwt=bwt, occ=occ, totals=np.array(m["totals"], np.int64), row=np.array(m["row"], np.int32),
                n=m["n"], occ_rate=m["occ_rate"])


def sharded_search(search: Callable, patterns, lengths, group=None):
    """Backward search of one query batch split over the ranks of `group`.

    Every rank passes the same `patterns` (uint8[nq, stride] tensor) and `lengths` (int32[nq])
    on its own device and a `search(patterns_block, lengths_block) -> (sp, ep)` callable bound to
    its replica of the index (``FMReplica.search_block``).  Rank r answers block r of
    ``split_range(nq, world)``; one all_gather of fixed-size (2, block) int32 tiles returns the
    inclusive (sp, ep) of every query, in query order, on every rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    nq = int(lengths.shape[0])
    dev = lengths.device
    if nq == 0:
        empty = torch.zeros(0, dtype=torch.int32, device=dev)
        return empty, empty.clone()
    spans = split_range(nq, world)
    lo, hi = spans[rank]
    block = spans[0][1] - spans[0][0]
    tile = torch.full((2, block), -1, dtype=torch.int32, device=dev)
    if hi > lo:
        sp, ep = search(patterns[lo:hi], lengths[lo:hi])
        tile[0, : hi - lo] = sp
        tile[1, : hi - lo] = ep
    tiles = [torch.empty_like(tile) for _ in range(world)]
    dist.all_gather(tiles, tile, group=group)
    sp_all = torch.cat([t[0, : b - a] for t, (a, b) in zip(tiles, spans)])
    ep_all = torch.cat([t[1, : b - a] for t, (a, b) in zip(tiles, spans)])
    return sp_all, ep_all
