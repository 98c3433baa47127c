"""Contig sharding over the GPUs of one box.

Every index array and detector is per contig (bwt.py:3040-3054), so the path
shards by dealing contigs to GPUs with longest-processing-time bin packing and
needs no data-path collective.  The only exchange is the gather of fixed-width
int32 record rows to rank 0 (``gather_rows``; NCCL over NVLink when the ranks
own GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Callable, Iterable, Iterator, List, Sequence, Tuple

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


def run_tasks(worker: Callable, tasks: List[Tuple], n_jobs: int) -> Iterator[list]:
    """Yields worker(task) results (any order).  One process per GPU when several
    GPUs are visible and there are several contigs; otherwise in-process."""
    ngpu = min(_gpu_count(), len(tasks), max(1, int(n_jobs or 1)))
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
