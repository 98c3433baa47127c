"""Per-contig worker, ``TandemRepeatFinder`` orchestration and the CLI.

Host glue around the GPU path with the reference's behaviour (bwt.py:3040-4370):
FASTA loading with flank trimming, the per-contig worker (the sharding unit:
one contig -> one GPU), the post-processing chain that turns raw detector
records into the final call set, and the BED / VCF / TRF / STRfinder writers.
Where the reference fans contigs out over a ``multiprocessing.Pool`` of CPU
workers, contigs here are dealt to the visible GPUs (longest contig first).
"""
from __future__ import annotations

import argparse
import os
import time
from multiprocessing import cpu_count
from typing import Dict, List, Optional, Tuple

from .core import BWTCore
from .finders import Tier2LCPFinder
from .motifs import MotifUtils, align_region_native
from .records import TandemRepeat, _natural_sort_key

_BAR = 40


def _bar(done: int, total: int) -> str:
    filled = int(_BAR * done / total) if total else _BAR
    return "█" * filled + "░" * (_BAR - filled)


def _elapsed(t0: float) -> str:
    e = time.time() - t0
    return f"{int(e // 60)}m {int(e % 60)}s" if e >= 60 else f"{int(e)}s"


def _detect_rows(chrom: str, seq: str, config: dict):
    """One contig at row level: index build + strict adjacency scan (bwt.py:3040-3106).  Returns
    (rows int32[R, 8] or None when no detector ran, text_arr uint8 incl. '$' or None when there are no rows)."""
    import numpy as np

    from . import lean

    verbose = config.get("show_progress", False)
    rows = None
    core = None
    # A fresh CLI process in which nobody has imported torch yet: the strict scan runs on buffers of the library's own
    # (lean.py) -- same kernels, same rows, seconds of start-up saved.  Otherwise the index object and torch's allocator.
    if lean.enabled() and seq.isascii() and lean.device_count() > 0:
        text_arr = np.frombuffer((seq + "$").encode("ascii"), np.uint8)

        def scan(unit_cap, min_copies):
            return lean.strict_rows(text_arr, 1, unit_cap, 0, min_copies)
    else:
        from . import detect
        from .finders import _device_text_of

        pinned = _PINNED.get(chrom)
        if pinned is not None and pinned.numel() != len(seq) + 1:
            pinned = None                      # a different contig of the same name
        core = BWTCore(seq + "$", config["sa_sample_rate"], _pinned_text=pinned)
        text_arr = core.text_arr

        def scan(unit_cap, min_copies):
            return detect.strict_rows(_device_text_of(core), 1, unit_cap, 0, min_copies)
    try:
        if verbose:
            print(f"  [{chrom}] Building indices ({len(seq):,} bp)...")
        if config.get("enable_tier2", False):
            if len(seq) > 50_000_000 and not verbose:
                pass  # very large contigs are skipped unless --progress is given (bwt.py:3070)
            else:
                min_copies = config["min_copies"]
                unit_cap = max(config["max_unit_len"], min(len(seq) // min_copies, 1000))
                rows = scan(unit_cap, min_copies)
                if verbose:
                    print(f"  [{chrom}] Strict adjacency: {len(rows)} tandem repeats detected")
        text_arr = np.array(text_arr, copy=True) if rows is not None and len(rows) else None
    finally:
        if core is not None:
            core.clear()
    return rows, text_arr


def _process_chromosome_worker(args):
    """One contig: index build + strict adjacency scan + the short-motif
    perfection filter (bwt.py:3040-3141).  Exceptions are reported and yield []."""
    chrom, seq, config = args
    try:
        from .finders import strict_records

        rows, text_arr = _detect_rows(chrom, seq, config)
        found: List[TandemRepeat] = []
        if rows is not None and len(rows):
            found = strict_records(text_arr, rows, chrom, 0)
        kept = []
        for r in found:
            motif = r.consensus_motif or r.motif
            if len(motif) < 5 and r.copies < 30 and (r.mismatch_rate > 0 or r.max_mismatches_per_copy > 0):
                continue
            kept.append(r)
        return kept
    except Exception as exc:  # the reference swallows worker failures the same way
        print(f"ERROR processing chromosome {chrom}: {exc}")
        import traceback

        traceback.print_exc()
        return []


def _distinct_sort_names(names) -> bool:
    keys = [_natural_sort_key(n) for n in names]
    return len(set(keys)) == len(keys)


def _finish_contig(chrom: str, seq: str, config: dict, left: str, right: str, raw: List["TandemRepeat"]):
    """The post-processing chain of one contig on records, away from the parent: a finder that knows just this
    contig (trimmed sequence for the re-alignments, flanks for the coordinate restore)."""
    return _contig_finder(chrom, seq, config, left, right)._postprocess_counts(raw)


def _contig_finder(chrom: str, seq: str, config: dict, left: str, right: str) -> "TandemRepeatFinder":
    finder = TandemRepeatFinder(
        "", sa_sample_rate=config["sa_sample_rate"], show_progress=config.get("show_progress", False),
        allow_mismatches=config["allow_mismatches"], max_motif_length=config["max_motif_length"],
        min_period=config["min_period"], max_period=config["max_period"], min_copies=config["min_copies"],
        min_entropy=config["min_entropy"], flank_trim=len(left), max_unit_len=config["max_unit_len"])
    finder.sequences = {chrom: seq}
    finder.full_sequences = {chrom: left + seq + right}
    finder.trim_offsets = {chrom: len(left)}
    return finder


def _process_and_finish_contig(args, pool=None):
    """Worker of the per-contig path: detection (bwt.py:3040-3141) at row level, then this contig's share of
    the post-processing chain on the rows (rowchain.finish_rows) -- records are only built for the calls
    that survive.  Returns (final calls, raw, after suppression, after dedup).  `pool`: CPU worker processes
    that finish independent pieces of the contig's rows side by side (_run_with_segment_pool)."""
    chrom, seq, config, left, right = args
    try:
        rows, text_arr = _detect_rows(chrom, seq, config)
    except Exception as exc:  # the reference swallows worker failures the same way
        print(f"ERROR processing chromosome {chrom}: {exc}")
        import traceback

        traceback.print_exc()
        return [], 0, 0, 0
    n_raw = 0 if rows is None else len(rows)
    if not n_raw:
        return [], 0, 0, 0
    return _finish_contig_rows((chrom, seq, config, left, right, rows, text_arr), pool)


def _finish_contig_rows(args, pool=None):
    """The host half of the per-contig path: this contig's share of the post-processing chain on the detector's
    rows.  CPU only (NumPy + the host functions of libbwtk.so), so it can run in a pool of plain worker processes
    while the parent keeps the GPU busy with the next contig.  `seq` may be None for an ASCII contig: it is the
    text without its sentinel.  With a `pool` and enough rows the contig is cut into independent pieces
    (rowchain.finish_rows_segmented) that the pool's workers finish side by side."""
    chrom, seq, config, left, right, rows, text_arr = args
    n_raw = len(rows)
    try:
        from . import rowchain

        if seq is None:
            seq = text_arr[:-1].tobytes().decode("ascii")
        finder = _contig_finder(chrom, seq, config, left, right)
        if pool is not None and n_raw >= 2 * rowchain.SEGMENT_MIN_ROWS:
            try:
                return _finish_rows_in_pool(pool, finder, chrom, seq, config, left, right, rows, text_arr)
            except Exception as exc:  # a lost worker must not lose the contig: finish it here
                print(f"WARNING: chain workers failed for {chrom} ({exc}); finishing in-process")
        return rowchain.finish_rows(finder, chrom, text_arr, rows)
    except Exception as exc:  # same policy as the detection worker
        print(f"ERROR post-processing chromosome {chrom}: {exc}")
        import traceback

        traceback.print_exc()
        return [], n_raw, n_raw, n_raw


def _finish_rows_in_pool(pool, finder, chrom, seq, config, left, right, rows, text_arr):
    """One contig's rows, cut into independent pieces, finished by the workers of `pool`.  The text goes to the
    workers once, through a file in a temporary directory (they read it on their first piece); the pieces' rows and
    the finished calls travel through the pool's pipes."""
    import shutil
    import tempfile

    import numpy as np

    from . import rowchain

    tmp = tempfile.mkdtemp(prefix="bwtk_chain_")
    try:
        path = os.path.join(tmp, "text.u8")
        np.ascontiguousarray(text_arr, np.uint8).tofile(path)
        head = (chrom, None if seq.isascii() else seq, config, left, right, path)

        def run_pieces(parts):
            jobs = [pool.apply_async(_finish_piece, (head + (np.ascontiguousarray(p[:, :4]),),)) for p in parts]
            return [j.get() for j in jobs]

        return rowchain.finish_rows_segmented(finder, chrom, text_arr, rows, run_pieces,
                                              pieces=max(1, getattr(pool, "_processes", 1)))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


_PIECE_TEXT: Dict[str, tuple] = {}      # worker side: text file -> (text_arr, finder) of the contig in progress


def _finish_piece(args):
    """Pool worker of _finish_rows_in_pool: one piece of one contig's rows -> rowchain.finish_segment."""
    import numpy as np

    from . import rowchain

    chrom, seq, config, left, right, path, rows = args
    got = _PIECE_TEXT.get(path)
    if got is None:
        _PIECE_TEXT.clear()
        text_arr = np.fromfile(path, np.uint8)
        if seq is None:
            seq = text_arr[:-1].tobytes().decode("ascii")
        got = _PIECE_TEXT[path] = (text_arr, _contig_finder(chrom, seq, config, left, right))
    text_arr, finder = got
    return rowchain.finish_segment(finder, chrom, text_arr, rows)


# One large contig on one GPU: its chain is cut into pieces for this many CPU workers (--jobs) from this many
# bases on; below it the chain is faster than starting the workers.  Tests lower it.
SEGMENT_POOL_MIN_BASES = 8_000_000


def _segment_pool_size(tasks: List[Tuple], n_jobs: int) -> int:
    if not tasks or (n_jobs or 1) < 2 or max(len(t[1]) for t in tasks) < SEGMENT_POOL_MIN_BASES:
        return 0
    return max(2, min(int(n_jobs), cpu_count(), 16))      # beyond that the parent (unpickling the calls) is the limit


def _run_with_segment_pool(tasks: List[Tuple], procs: int):
    """Per-contig path for a few large contigs on ONE GPU: contig after contig, detection here, the chain of every
    contig cut into independent pieces for `procs` spawned CPU workers.  The workers start (interpreter, NumPy,
    this package) while the first contig is on the GPU."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")          # never fork a process that holds a CUDA context
    with ctx.Pool(procs) as pool:
        for task in tasks:
            yield _process_and_finish_contig(task, pool=pool)


# The chain pool (one GPU, several contigs): from this many bases on, the contigs' chains go to CPU worker
# processes; below it starting the workers costs more than the chains.  Tests set it to 0.
CHAIN_POOL_MIN_BASES = 16_000_000


def _chain_pool_size(tasks: List[Tuple], n_jobs: int) -> int:
    """Worker processes for the post-processing chains of `tasks` when detection stays in this process."""
    if len(tasks) < 2 or (n_jobs or 1) < 2 or sum(len(t[1]) for t in tasks) < CHAIN_POOL_MIN_BASES:
        return 0
    return max(1, min(int(n_jobs), len(tasks), cpu_count()))


def _run_with_chain_pool(tasks: List[Tuple], procs: int):
    """Per-contig path on ONE GPU: detection (index build + strict scan, milliseconds per contig) runs here, contig
    after contig, longest first; every contig's rows go to a pool of `procs` CPU processes for the chain, which is
    what a contig's time consists of (seconds per contig).  Yields what _process_and_finish_contig returns, in
    any order.  The reference does the same split the other way round -- its pool workers do everything
    (bwt.py:3863-3899) -- because there the detection is the expensive half."""
    import multiprocessing as mp

    import numpy as np

    ctx = mp.get_context("spawn")          # the parent holds a CUDA context: workers must not be forked from it
    with ctx.Pool(procs) as pool:
        pending = []
        for chrom, seq, config, left, right in sorted(tasks, key=lambda t: -len(t[1])):
            try:
                rows, text_arr = _detect_rows(chrom, seq, config)
            except Exception as exc:  # the reference swallows worker failures the same way
                print(f"ERROR processing chromosome {chrom}: {exc}")
                import traceback

                traceback.print_exc()
                yield [], 0, 0, 0
                continue
            if rows is None or not len(rows):
                yield [], 0, 0, 0
                continue
            # the chain reads (start, end, primitive period, copies); an ASCII contig is rebuilt from its text
            payload = (chrom, None if seq.isascii() else seq, config, left, right,
                       np.ascontiguousarray(rows[:, :4]), text_arr)
            pending.append(pool.apply_async(_finish_contig_rows, (payload,)))
            while pending and pending[0].ready():
                yield pending.pop(0).get()
        for res in pending:
            yield res.get()


# consensus motif -> (strand, composition, entropy) for _recompute_repeat (pure functions of the string)
_PER_CONSENSUS: Dict[str, tuple] = {}

# Pinned host copies (trimmed text + '$') of the contigs of the FASTA loaded last, filled by the native
# parser; the in-process worker uploads from here.  Worker processes (one per GPU) do not see it and
# encode their contig as before.
_PINNED: Dict[str, object] = {}


def _pinning_enabled() -> bool:
    """Pinned staging of the contigs pays when torch's allocator is in play; a process that has not imported torch
    (the CLI's start-up path, lean.py) uploads from ordinary memory and does not import it for this."""
    import sys

    if "torch" not in sys.modules:
        return False
    try:
        import torch

        return bool(torch.cuda.is_available())
    except Exception:
        return False


class _LazyCores(dict):
    """``bwt_cores`` of the CLI path: chrom -> BWTCore, each built on first access."""

    def __init__(self, sequences: Dict[str, str], sa_sample_rate: int):
        super().__init__()
        self._sequences = sequences
        self._rate = sa_sample_rate

    def __missing__(self, chrom):
        core = BWTCore(self._sequences[chrom] + "$", self._rate)
        self[chrom] = core
        return core

    def items(self):
        return ((c, self[c]) for c in self._sequences)

    def __iter__(self):
        return iter(self._sequences)

    def __len__(self):
        return len(self._sequences)

    def __contains__(self, chrom):
        return chrom in self._sequences


class TandemRepeatFinder:
    """Coordinates loading, detection, post-processing and output (bwt.py:3144-4198)."""

    def __init__(self, reference_file: str, sa_sample_rate: int = 32, show_progress: bool = False,
                 allow_mismatches: bool = True, max_motif_length: int = 9, min_period: int = 10,
                 max_period: int = 1000, min_copies: int = 3, min_entropy: float = 1.0, flank_trim: int = 30,
                 max_unit_len: int = 120):
        self.reference_file = reference_file
        self.sa_sample_rate = sa_sample_rate
        self.bwt_cores: Dict[str, BWTCore] = {}
        self.sequences: Dict[str, str] = {}
        self.show_progress = show_progress
        self.allow_mismatches = allow_mismatches
        self.max_motif_length = max_motif_length
        self.min_period = min_period
        self.max_period = max_period
        self.min_copies = min_copies
        self.min_entropy = min_entropy
        self.flank_trim = max(0, flank_trim)
        self.max_unit_len = max_unit_len
        self.trim_offsets: Dict[str, int] = {}
        self.full_sequences: Dict[str, str] = {}

    # ------------------------------------------------------------------ ordering
    @staticmethod
    def _repeat_sort_key(repeat: TandemRepeat):
        return (_natural_sort_key(repeat.chrom), repeat.start, repeat.end)

    # ------------------------------------------------------------------ loading
    def _register(self, sequences: Dict[str, str], name: str, chunks: List[str]) -> None:
        full = "".join(chunks)
        self.full_sequences[name] = full
        if len(full) <= 2 * self.flank_trim:
            trimmed, left = full, 0
        else:
            left = self.flank_trim
            trimmed = full[left:len(full) - self.flank_trim]
        self.trim_offsets[name] = left
        sequences[name] = trimmed

    def load_reference(self) -> Dict[str, str]:
        """FASTA -> {name: upper-cased, flank-trimmed sequence} (bwt.py:3713-3756).

        ASCII files are parsed natively (csrc/fasta.cu: one index pass, then every record's sequence is
        written upper-cased straight into its buffer -- and, when a GPU is present, the trimmed text with
        its '$' into PINNED host memory, from where the per-contig worker uploads it without another
        copy).  Anything else (non-ASCII bytes, a header without a name) goes through the reference's own
        line loop, which defines the behaviour."""
        sequences = self._load_reference_native()
        if sequences is None:
            sequences = self._load_reference_lines()
        self.sequences = sequences
        return sequences

    def _load_reference_lines(self) -> Dict[str, str]:
        sequences: Dict[str, str] = {}
        name, chunks = None, []
        with open(self.reference_file, "r") as fh:
            for raw in fh:
                line = raw.strip()
                if line.startswith(">"):
                    if name:
                        self._register(sequences, name, chunks)
                    name, chunks = line[1:].split()[0], []
                elif line:
                    chunks.append(line.upper())
        if name:
            self._register(sequences, name, chunks)
        return sequences

    def _load_reference_native(self) -> Optional[Dict[str, str]]:
        import ctypes as C

        import numpy as np

        from . import _lib

        try:
            L = _lib.lib()
        except _lib.BwtkError:
            return None
        data = np.fromfile(self.reference_file, dtype=np.uint8)
        if data.size == 0:
            return {}
        cap = 1024
        while True:
            rec = np.zeros((cap, 5), np.int64)
            count, flags = C.c_int64(0), C.c_int32(0)
            rc = L.bwtk_fasta_index(data.ctypes.data, data.size, rec.ctypes.data, cap, C.addressof(count),
                                    C.addressof(flags))
            if rc == _lib.E_OVERFLOW:
                cap = int(count.value)
                continue
            break
        if rc != 0 or (flags.value & 1):
            return None
        pin = _PINNED if _pinning_enabled() else None
        if pin is not None:
            pin.clear()
        sequences: Dict[str, str] = {}
        for name_off, name_len, body_off, body_end, seq_len in rec[: count.value].tolist():
            name = data[name_off:name_off + name_len].tobytes().decode("ascii")
            buf = np.empty(seq_len, np.uint8)
            if L.bwtk_fasta_sequence(data.ctypes.data, body_off, body_end, 0, seq_len, 0, buf.ctypes.data) != seq_len:
                return None
            full = buf.tobytes().decode("ascii")
            self.full_sequences[name] = full
            left = 0 if seq_len <= 2 * self.flank_trim else self.flank_trim
            keep = seq_len - 2 * left
            self.trim_offsets[name] = left
            sequences[name] = full[left:left + keep] if left else full
            if pin is not None and keep > 0:
                import torch

                t = torch.empty(keep + 1, dtype=torch.uint8, pin_memory=True)
                if L.bwtk_fasta_sequence(data.ctypes.data, body_off, body_end, left, keep, 1, t.data_ptr()) == keep + 1:
                    pin[name] = t
        return sequences

    def build_indices(self, sequences: Dict[str, str], lazy: bool = False):
        """One BWTCore per contig in self.bwt_cores (bwt.py:3758-3790).  With ``lazy`` (the CLI path) a
        contig's index is built when ``bwt_cores[chrom]`` is first read: the per-contig workers build
        their own index on their own GPU, so building every contig here as well would double the build
        time and hold ~10 B/base of HBM on GPU 0 for nothing."""
        items = list(sequences.items())
        print("Building BWT indices...")
        t0 = time.time()
        if lazy:
            self.bwt_cores = _LazyCores(sequences, self.sa_sample_rate)
        for idx, (chrom, seq) in enumerate(items, 1):
            pct = (idx - 1) / len(items) * 100
            print(f"\r[{_bar(idx - 1, len(items))}] {pct:.1f}% Building index for {chrom} ({len(seq):,} bp) - "
                  f"{_elapsed(t0)}", end="", flush=True)
            if not lazy:
                self.bwt_cores[chrom] = BWTCore(seq + "$", self.sa_sample_rate)
        print(f"\r[{'█' * _BAR}] 100.0% BWT indices built for {len(items)} chromosome(s) - {_elapsed(t0)}     ")
        print()

    # ------------------------------------------------------------------ detection
    def _config(self, enable_tier1: bool, enable_tier2: bool) -> dict:
        return {
            "sa_sample_rate": self.sa_sample_rate, "enable_tier1": enable_tier1, "enable_tier2": enable_tier2,
            "allow_mismatches": self.allow_mismatches, "max_motif_length": self.max_motif_length,
            "min_period": self.min_period, "max_period": self.max_period, "min_copies": self.min_copies,
            "min_entropy": self.min_entropy, "show_progress": self.show_progress,
            "max_unit_len": self.max_unit_len,
        }

    def _postprocess_counts(self, raw: List[TandemRepeat]) -> Tuple[List[TandemRepeat], int, int, int]:
        """The chain of bwt.py:3827-3846 / 3928-3952; returns the final calls and the sizes of the raw,
        nested-call-suppressed and deduplicated lists."""
        kept = self._suppress_nested_short_calls(raw, overlap_threshold=0.5)
        unique = self._deduplicate_repeats(kept)
        merged = self._merge_adjacent_repeats(unique)
        refined = self._refine_repeats(merged)
        self._restore_reference_coordinates(refined)
        refined = self._collapse_overlapping_repeats(refined)
        final = [r for r in refined if r.copies >= self.min_copies and r.length >= 6]
        final.sort(key=self._repeat_sort_key)
        return final, len(raw), len(kept), len(unique)

    def _postprocess(self, raw: List[TandemRepeat], announce) -> Tuple[List[TandemRepeat], int]:
        final, n_raw, n_kept, n_unique = self._postprocess_counts(raw)
        if n_kept < n_raw:
            announce(f"Nested call suppression: {n_raw} -> {n_kept} repeats")
        return final, n_raw - n_unique

    def find_tandem_repeats(self, enable_tier1: bool = True, enable_tier2: bool = True, enable_tier3: bool = False,
                            long_reads: Optional[List[str]] = None) -> List[TandemRepeat]:
        """Sequential discovery (bwt.py:3792-3848)."""
        raw: List[TandemRepeat] = []
        items = list(self.sequences.items())
        total = len(items)
        # Every stage of the chain works inside one contig, so with distinct contig sort keys each contig is
        # finished at row level right after its scan (rowchain.finish_rows) and the results are sorted once.
        per_contig = _distinct_sort_names(self.sequences)
        cfg = self._config(enable_tier1, enable_tier2)
        finished: List[TandemRepeat] = []
        n_raw = n_kept = 0
        for idx, (chrom, seq) in enumerate(items, 1):
            pct = (idx - 1) / total * 100 if total else 100.0
            print(f"\n[{_bar(idx - 1, total)}] {pct:.1f}% ({idx}/{total})")
            print(f"Scanning chromosome {chrom} ({len(seq):,} bp)...")
            if per_contig:
                part, a, b, _c = _process_and_finish_contig(self._finish_tasks([(chrom, seq, cfg)])[0])
                finished.extend(part)
                n_raw, n_kept = n_raw + a, n_kept + b
                print(f"  Detected {a} STR blocks")
            else:
                got = _process_chromosome_worker((chrom, seq, cfg))
                raw.extend(got)
                print(f"  Detected {len(got)} STR blocks")
        print(f"\n[{'█' * _BAR}] 100.0% ({total}/{total})")
        if per_contig:
            if n_kept < n_raw:
                print(f"Nested call suppression: {n_raw} -> {n_kept} repeats")
            finished.sort(key=self._repeat_sort_key)
            final = finished
        else:
            final, _ = self._postprocess(raw, print)
        print(f"Analysis complete! Found {len(final)} total repeats.")
        return final

    def find_tandem_repeats_parallel(self, enable_tier1: bool = True, enable_tier2: bool = True,
                                     enable_tier3: bool = False, long_reads: Optional[List[str]] = None,
                                     n_jobs: Optional[int] = None) -> List[TandemRepeat]:
        """Contig-parallel discovery (bwt.py:3850-3954): contigs are dealt to the
        visible GPUs instead of CPU pool workers; the result is order-independent
        because everything is re-sorted afterwards."""
        from . import sharding

        host_jobs = cpu_count() if n_jobs is None else n_jobs      # CPU workers for the host chain: --jobs 0 = all cores
        if n_jobs is None:
            n_jobs = min(cpu_count(), len(self.sequences))
        print(f"Parallel mode: Using {n_jobs} CPU cores for {len(self.sequences)} chromosomes")
        print()
        cfg = self._config(enable_tier1, enable_tier2)
        tasks = [(chrom, seq, cfg) for chrom, seq in self.sequences.items()]
        print(f"Processing {len(tasks)} chromosome(s)...")
        print()
        t0 = time.time()
        raw: List[TandemRepeat] = []
        done = 0
        # Every stage of the post-processing chain works inside one contig (calls of different contigs
        # never nest, merge or collapse), so when the contigs go to several worker processes each worker
        # finishes its contigs itself: the chain runs in parallel and only the final calls (a few percent
        # of the raw ones) are pickled back.  Needs contig names with distinct sort keys, so that sorting
        # the concatenation equals the chain's own sorts of the mixed list.
        per_contig = _distinct_sort_names(self.sequences) and not (enable_tier3 and long_reads)
        if per_contig:
            finished: List[TandemRepeat] = []
            n_raw = n_kept = n_unique = 0
            ftasks = self._finish_tasks(tasks)
            # one process per GPU when there are several; on one GPU the host side is spread over CPU workers:
            # whole contigs when there are at least as many as workers, pieces of a contig's chain otherwise
            one_gpu = sharding.worker_processes(len(ftasks), n_jobs) <= 1
            pool_procs = _chain_pool_size(ftasks, host_jobs) if one_gpu else 0
            seg_procs = _segment_pool_size(ftasks, host_jobs) if one_gpu else 0
            if seg_procs and (not pool_procs or len(ftasks) < seg_procs):
                results = _run_with_segment_pool(ftasks, seg_procs)
            elif pool_procs:
                results = _run_with_chain_pool(ftasks, pool_procs)
            else:
                results = sharding.run_tasks(_process_and_finish_contig, ftasks, n_jobs)
            for part, a, b, c in results:
                finished.extend(part)
                n_raw, n_kept, n_unique = n_raw + a, n_kept + b, n_unique + c
                done += 1
                print(f"\r[{_bar(done, len(tasks))}] {done / len(tasks) * 100:.1f}% ({done}/{len(tasks)}) "
                      f"chromosomes completed - {_elapsed(t0)}", end="", flush=True)
            print()
            if n_kept < n_raw:
                print("\n" + f"Nested call suppression: {n_raw} -> {n_kept} repeats")
            finished.sort(key=self._repeat_sort_key)
            final, dup = finished, n_raw - n_unique
        else:
            for got in sharding.run_tasks(_process_chromosome_worker, tasks, n_jobs):
                raw.extend(got)
                done += 1
                print(f"\r[{_bar(done, len(tasks))}] {done / len(tasks) * 100:.1f}% ({done}/{len(tasks)}) "
                      f"chromosomes completed - {_elapsed(t0)}", end="", flush=True)
            print()
            if enable_tier3 and long_reads:
                raise NotImplementedError("Tier 3 (long reads) is outside the B200 hot path")
            final, dup = self._postprocess(raw, lambda m: print("\n" + m))
        if dup > 0:
            print(f"Analysis complete! Found {len(final)} unique repeats (deduplicated {dup}).")
        else:
            print(f"Analysis complete! Found {len(final)} unique repeats.")
        return final

    def _finish_tasks(self, tasks: List[Tuple]) -> List[Tuple]:
        """(chrom, seq, config) -> (chrom, seq, config, left flank, right flank): what a worker needs to
        restore reference coordinates (bwt.py:3499-3513) without the parent's dictionaries."""
        out = []
        for chrom, seq, cfg in tasks:
            full = self.full_sequences.get(chrom) or ""
            left = self.trim_offsets.get(chrom, 0)
            out.append((chrom, seq, cfg, full[:left], full[left + len(seq):]))
        return out

    # ------------------------------------------------------------------ post-processing
    def _suppress_nested_short_calls(self, repeats: List[TandemRepeat],
                                     overlap_threshold: float = 0.5) -> List[TandemRepeat]:
        """Drop short-motif calls nested in kept longer-motif calls (bwt.py:3402-3497)."""
        if not repeats:
            return []
        by_chrom: Dict[str, List[TandemRepeat]] = {}
        for r in repeats:
            by_chrom.setdefault(r.chrom, []).append(r)
        survivors: List[TandemRepeat] = []
        for chrom_repeats in by_chrom.values():
            ordered = sorted(chrom_repeats, key=lambda r: (r.mismatch_rate > 0, -len(r.motif)))
            # The reference tests every call against every kept span (quadratic).  Only kept spans
            # that overlap the call can nest it, so the kept spans are filed under the 256 bp
            # buckets they touch and a call is tested against its own buckets only: same
            # predicate, same survivors, near-linear time on millions of calls.
            buckets: Dict[int, List[Tuple[int, int, int]]] = {}
            for r in ordered:
                span = r.end - r.start
                k = len(r.motif)
                nested = False
                first = int(r.start) >> 8
                last = int(max(r.end - 1, r.start)) >> 8
                # calls without mismatches come first, longest motif first: until the first call with
                # mismatches every bucket list is non-increasing in motif length, and the first
                # entry that is not longer than this call's motif ends the candidates
                in_order = not (r.mismatch_rate > 0)
                r_start, r_end = r.start, r.end
                for b in range(first, last + 1):
                    for o_start, o_end, o_k in buckets.get(b, ()):
                        if o_k <= k:
                            if in_order:
                                break
                            continue
                        # ov = max(0, min(r.end, o_end) - max(r.start, o_start)), spelled without calls
                        ov = (r_end if r_end < o_end else o_end) - (r_start if r_start > o_start else o_start)
                        if ov <= 0:
                            continue
                        if k == 1 and o_k > 1 and ov / span >= 0.8:
                            nested = True
                            break
                        ratio = o_k / k
                        cut = 0.1 if ratio >= 10 else 0.3 if ratio >= 5 else overlap_threshold
                        if ov / span >= cut:
                            nested = True
                            break
                    if nested:
                        break
                if not nested:
                    survivors.append(r)
                    entry = (r.start, r.end, k)
                    for b in range(first, last + 1):
                        buckets.setdefault(b, []).append(entry)
        survivors.sort(key=self._repeat_sort_key)
        return survivors

    def _deduplicate_repeats(self, repeats: List[TandemRepeat]) -> List[TandemRepeat]:
        """One record per (chrom,start,end,motif): best confidence, then lowest
        mismatch rate, then lowest tier (bwt.py:3189-3220)."""
        if not repeats:
            return []
        best: Dict[Tuple[str, int, int, str], TandemRepeat] = {}
        for r in repeats:
            key = (r.chrom, r.start, r.end, r.motif)
            old = best.get(key)
            if old is None:
                best[key] = r
            elif r.confidence > old.confidence:
                best[key] = r
            elif r.confidence == old.confidence:
                if r.mismatch_rate < old.mismatch_rate or (r.mismatch_rate == old.mismatch_rate and r.tier < old.tier):
                    best[key] = r
        out = list(best.values())
        out.sort(key=self._repeat_sort_key)
        return out

    def _merge_adjacent_repeats(self, repeats: List[TandemRepeat]) -> List[TandemRepeat]:
        if not repeats:
            return []
        out: List[TandemRepeat] = []
        cur = repeats[0]
        for nxt in repeats[1:]:
            if self._should_merge_repeats(cur, nxt):
                cur = self._merge_repeats(cur, nxt)
            else:
                out.append(cur)
                cur = nxt
        out.append(cur)
        return out

    def _should_merge_repeats(self, r1: TandemRepeat, r2: TandemRepeat) -> bool:
        if r1.chrom != r2.chrom:
            return False
        m1 = r1.consensus_motif or r1.motif
        m2 = r2.consensus_motif or r2.motif
        if not m1 or not m2:
            return False
        if MotifUtils.get_canonical_motif_stranded(m1)[0] != MotifUtils.get_canonical_motif_stranded(m2)[0]:
            return False
        shortest = min(len(m1), len(m2))
        if max(0, r2.start - r1.end) > shortest + 1:
            return False
        self._union_memo = None
        key = (r1.chrom, min(r1.start, r2.start), max(r1.end, r2.end), max(1, shortest), min(r1.tier, r2.tier))
        try:
            union = self._recompute_repeat(*key[:4], tier_hint=key[4])
        except ValueError:
            return False
        # _merge_repeats re-derives the same union whenever r1's motif is the shorter one; the
        # alignment is a pure function of these arguments, so the record is handed over once.
        self._union_memo = (key, union)
        if union.copies < self.min_copies:
            return False
        return union.mismatch_rate <= max(r1.mismatch_rate, r2.mismatch_rate, 0.01) + 0.2

    def _merge_repeats(self, r1: TandemRepeat, r2: TandemRepeat) -> TandemRepeat:
        key = (r1.chrom, min(r1.start, r2.start), max(r1.end, r2.end),
               len(r1.consensus_motif or r1.motif), min(r1.tier, r2.tier))
        memo, self._union_memo = getattr(self, "_union_memo", None), None
        if memo is not None and memo[0] == key:
            return memo[1]
        return self._recompute_repeat(*key[:4], tier_hint=key[4])

    def _refine_repeats(self, repeats: List[TandemRepeat]) -> List[TandemRepeat]:
        out: List[TandemRepeat] = []
        for r in repeats:
            if r.mismatch_rate == 0.0:
                out.append(r)
                continue
            k = len(r.consensus_motif or r.motif)
            if k <= 0:
                k = max(1, r.length // max(1, int(round(r.copies)) or 1))
            out.append(self._recompute_repeat(r.chrom, r.start, r.end, k, tier_hint=r.tier))
        out.sort(key=self._repeat_sort_key)
        return out

    def _restore_reference_coordinates(self, repeats: List[TandemRepeat]) -> None:
        for r in repeats:
            shift = self.trim_offsets.get(r.chrom, 0)
            r.start += shift
            r.end += shift
            r.length = r.end - r.start
            full = self.full_sequences.get(r.chrom)
            if full:
                r.actual_sequence = full[r.start:r.end]

    def _should_collapse_duplicates(self, r1: TandemRepeat, r2: TandemRepeat) -> bool:
        if r1.chrom != r2.chrom:
            return False
        ov = min(r1.end, r2.end) - max(r1.start, r2.start)
        if ov <= 0:
            return False
        shorter = min(r1.length, r2.length)
        if shorter <= 0 or ov / shorter < 0.8:
            return False
        if MotifUtils.get_canonical_motif_stranded(r1.motif)[0] == MotifUtils.get_canonical_motif_stranded(r2.motif)[0]:
            return True
        if (len(r1.motif) == 1 or len(r2.motif) == 1) and ov / shorter >= 0.95:
            return True
        if len(r1.motif) == len(r2.motif) and ov / shorter >= 0.9:
            return abs(r1.mismatch_rate - r2.mismatch_rate) >= 0.2
        return False

    @staticmethod
    def _prefer_repeat_entry(r1: TandemRepeat, r2: TandemRepeat) -> TandemRepeat:
        m1 = r1.consensus_motif or r1.motif
        m2 = r2.consensus_motif or r2.motif
        k1, k2 = len(m1), len(m2)
        if k1 != k2:
            if k1 == 1 and k2 > 1:
                return r2
            if k2 == 1 and k1 > 1:
                return r1
            small, big = (m1, m2) if k1 < k2 else (m2, m1)
            if len(big) % len(small) == 0 and small * (len(big) // len(small)) == big:
                return r1 if k1 < k2 else r2      # the primitive root wins over its power
            return r1 if k1 > k2 else r2          # otherwise the longer motif
        if r1.mismatch_rate != r2.mismatch_rate:
            return r1 if r1.mismatch_rate < r2.mismatch_rate else r2
        if r1.confidence != r2.confidence:
            return r1 if r1.confidence > r2.confidence else r2
        if r1.length != r2.length:
            return r1 if r1.length >= r2.length else r2
        return r1

    def _collapse_overlapping_repeats(self, repeats: List[TandemRepeat]) -> List[TandemRepeat]:
        out: List[TandemRepeat] = []
        for r in sorted(repeats, key=self._repeat_sort_key):
            if out and self._should_collapse_duplicates(out[-1], r):
                out[-1] = self._prefer_repeat_entry(out[-1], r)
            else:
                out.append(r)
        return out

    def _recompute_repeat(self, chrom: str, start: int, end: int, motif_len: int,
                          tier_hint: int = 1) -> TandemRepeat:
        """Re-derive a record from the (trimmed) sequence by per-copy alignment (bwt.py:3515-3614)."""
        sequence = self.sequences.get(chrom)
        if sequence is None:
            raise ValueError(f"Sequence for chromosome {chrom} not available for recomputation.")
        total = len(sequence)
        motif_len = max(1, motif_len)
        start = max(0, int(start))
        end = min(total, int(end)) if end > 0 else total
        if end <= start:
            end = min(total, start + motif_len)
        template = sequence[start:start + motif_len]
        if not template:
            back = max(0, start - motif_len)
            template = sequence[back:back + motif_len]
        if not template:
            template = "N" * motif_len
        summary = self._align_walk(chrom, sequence, start, end, template, max(1, self.min_copies))
        if summary is None:
            summary = self._align_walk(chrom, sequence, start, end, template, 1)
        if summary is None:
            consumed = min(total - start, max(motif_len, end - start))
            actual = sequence[start:start + consumed]
            n_copies = max(1, consumed // motif_len)
            consensus = template if template else (actual[:motif_len] or "N")
            mm_rate, worst, pct_indel, notes = 0.0, 0, 0.0, None
        else:
            s_cons, n_copies, s_consumed, mm_rate, worst, s_notes, s_ins, s_del = summary
            actual = sequence[start:start + s_consumed]
            consensus = s_cons or template
            cells = n_copies * len(template)
            pct_indel = ((s_ins + s_del) / cells if cells > 0 else 0.0) * 100.0
            notes = s_notes if s_notes else None
        span = len(actual)
        k_eff = len(consensus) if consensus else motif_len
        copies = float(n_copies)
        if span > 0 and k_eff > 0:
            frac = span / k_eff
            copies = float(round(frac)) if abs(frac - round(frac)) < 1e-6 else frac
        # strand, composition and entropy are functions of the consensus alone, and a contig's merges ask for the
        # same few consensus motifs over and over
        known = _PER_CONSENSUS.get(consensus)
        if known is None:
            if len(_PER_CONSENSUS) >= 1 << 16:
                _PER_CONSENSUS.clear()
            known = _PER_CONSENSUS[consensus] = (MotifUtils.get_canonical_motif_stranded(consensus)[1],
                                                 MotifUtils.calculate_composition(consensus),
                                                 MotifUtils.calculate_entropy(consensus))
        strand, comp, ent = known
        # positional, in the field order of records.TandemRepeat (bwt.py:429-452)
        return TandemRepeat(chrom, start, start + span, consensus, copies, span, tier_hint, max(0.3, 1.0 - mm_rate),
                            consensus, mm_rate, worst, max(1, n_copies), strand, max(0.0, 100.0 - mm_rate * 100.0),
                            pct_indel, MotifUtils.calculate_trf_score(consensus, max(1, n_copies), mm_rate, span),
                            dict(comp), ent, actual, notes)

    _NATIVE_WALK = True    # tests switch it off to compare with the Python walk

    def _align_walk(self, chrom: str, sequence: str, start: int, end: int, template: str, min_copies: int):
        """``MotifUtils.align_repeat_region(sequence, start, end, template, 0.1, min_copies=...)`` reduced to
        the fields _recompute_repeat reads: (consensus, copies, consumed_length, mismatch_rate,
        max_errors_per_copy, variations, total_insertions, total_deletions) or None.  ASCII contigs take the
        native walk over the contig's bytes (bwtk_align_repeat_region, the same walk answer for answer);
        anything else, and whatever the native walk declines, goes through the Python one."""
        if self._NATIVE_WALK:
            memo = self.__dict__.setdefault("_seq_bytes", {})
            got = memo.get(chrom)
            if got is None or got[0] is not sequence:
                got = (sequence, sequence.encode("ascii") if sequence.isascii() else None)
                memo[chrom] = got
            if got[1] is not None and template.isascii():
                res = align_region_native(got[1], start, end, template.encode("ascii"), 0.1, None, min_copies)
                if res is not NotImplemented:
                    return res
        summ = MotifUtils.align_repeat_region(sequence, start, end, template, mismatch_fraction=0.1,
                                              min_copies=min_copies)
        if summ is None:
            return None
        return (summ.consensus, summ.copies, summ.consumed_length, summ.mismatch_rate, summ.max_errors_per_copy,
                summ.variations, summ.total_insertions, summ.total_deletions)

    # ------------------------------------------------------------------ compound repeats (strfinder writer)
    @staticmethod
    def _perfect_block(chrom, start, end, motif, copies, seq, tier=1) -> TandemRepeat:
        return TandemRepeat(
            chrom=chrom, start=start, end=end, motif=motif, copies=float(copies), length=end - start, tier=tier,
            confidence=1.0, consensus_motif=motif, mismatch_rate=0.0, max_mismatches_per_copy=0,
            n_copies_evaluated=copies, strand="+", percent_matches=100.0, percent_indels=0.0, score=100.0,
            composition={"A": 0, "C": 0, "G": 0, "T": 0}, entropy=1.5, actual_sequence=seq, variations=None)

    def _simple_kmer_scan(self, chrom: str, start: int, end: int, k: int = 3,
                          use_full_seq: bool = True) -> List[TandemRepeat]:
        """Perfect k-mer runs of >= 5 copies inside [start, end) (bwt.py:3956-3993)."""
        seq = (self.full_sequences if use_full_seq else self.sequences).get(chrom, "")
        if not seq or start >= end or start < 0 or end > len(seq):
            return []
        region = seq[start:end]
        out: List[TandemRepeat] = []
        i = 0
        while i < len(region) - k:
            motif = region[i:i + k]
            copies, j = 1, i + k
            while j + k <= len(region) and region[j:j + k] == motif:
                copies += 1
                j += k
            if copies >= 5:
                out.append(self._perfect_block(chrom, start + i, start + j, motif, copies, region[i:j]))
                i = j
            else:
                i += 1
        return out

    def _detect_compound_repeats(self, repeats: List[TandemRepeat]) -> List[TandemRepeat]:
        """Mark adjacent short repeats with different motifs as compound (bwt.py:3995-4139)."""
        if not repeats:
            return []
        by_chrom: Dict[str, List[TandemRepeat]] = {}
        for r in repeats:
            by_chrom.setdefault(r.chrom, []).append(r)
        for chrom in by_chrom:
            seq = self.full_sequences.get(chrom, "")
            if not seq:
                continue
            for r in list(by_chrom[chrom]):
                if r.motif and len(r.motif) == 3:
                    lo, hi = r.end, min(len(seq), r.end + 50)
                    if lo < hi:
                        for extra in self._simple_kmer_scan(chrom, lo, hi, k=3):
                            if extra.motif != r.motif:
                                by_chrom[chrom].append(extra)
        out: List[TandemRepeat] = []
        for chrom, rs in by_chrom.items():
            rs.sort(key=lambda r: r.start)
            long_spans = [(r.start, r.end, len(r.motif)) for r in rs if len(r.motif) > 10]
            i = 0
            while i < len(rs):
                cur = rs[i]
                if len(cur.motif) == 3 and cur.copies >= 10:
                    seq = self.sequences.get(cur.chrom, "")
                    if seq:
                        body = seq[cur.start:cur.end]
                        k = len(cur.motif)
                        for cut in range(k, len(body) - k, k):
                            a, b = body[:k], body[cut:cut + k]
                            if a == b:
                                continue
                            n1 = 0
                            for j in range(0, cut, k):
                                if body[j:j + k] == a:
                                    n1 += 1
                                else:
                                    break
                            n2 = 0
                            for j in range(cut, len(body), k):
                                if body[j:j + k] == b:
                                    n2 += 1
                                else:
                                    break
                            if n1 >= 5 and n2 >= 5 and (n1 * len(a) + n2 * len(b)) >= len(body) * 0.9:
                                first = self._perfect_block(cur.chrom, cur.start, cur.start + n1 * k, a, n1,
                                                            body[:n1 * k], tier=cur.tier)
                                second = self._perfect_block(cur.chrom, cur.start + n1 * k,
                                                             cur.start + n1 * k + n2 * k, b, n2,
                                                             body[n1 * k:n1 * k + n2 * k], tier=cur.tier)
                                first.is_compound = True
                                first.compound_partner = second
                                out.append(first)
                                # (sic) the reference advances the outer index inside the split loop
                                # and keeps trying later split points (its `continue` binds to the
                                # inner for); `cur` stays the same record afterwards (bwt.py:4096-4098)
                                i += 1
                if i + 1 < len(rs):
                    nxt = rs[i + 1]
                    gap = nxt.start - cur.end
                    if (gap <= 5 and len(cur.motif) <= 4 and len(nxt.motif) <= 4 and cur.motif != nxt.motif
                            and cur.copies >= 5 and nxt.copies >= 5):
                        c_start, c_end = cur.start, nxt.end
                        covered = False
                        for l_start, l_end, _k in long_spans:
                            ov = max(0, min(c_end, l_end) - max(c_start, l_start))
                            if ov / (c_end - c_start) >= 0.8:
                                covered = True
                                break
                        if not covered:
                            cur.is_compound = True
                            cur.compound_partner = nxt
                            out.append(cur)
                            i += 2
                            continue
                out.append(cur)
                i += 1
        return out

    # ------------------------------------------------------------------ writers
    def save_results(self, repeats: List[TandemRepeat], output_file: str, format_type: str = "bed"):
        """BED / VCF / TRF table / TRF DAT / STRfinder (bwt.py:4141-4198)."""
        if format_type == "strfinder":
            repeats = self._detect_compound_repeats(repeats)
        ordered = sorted(repeats, key=self._repeat_sort_key)
        with open(output_file, "w") as fh:
            if format_type == "bed":
                fh.write("# Tandem Repeats (BED format with imperfect repeat support)\n")
                fh.write("# chrom\tstart\tend\tconsensus_motif\tcopies\ttier\tmismatch_rate\tstrand\n")
                for r in ordered:
                    fh.write(r.to_bed() + "\n")
            elif format_type == "vcf":
                fh.write("##fileformat=VCFv4.2\n")
                for ident, number, kind, desc in (
                        ("MOTIF", "1", "String", "Original seed motif"),
                        ("CONS_MOTIF", "1", "String", "Consensus motif from all copies"),
                        ("COPIES", "1", "Float", "Number of copies"),
                        ("TIER", "1", "Integer", "Detection tier (1=short, 2=medium/long, 3=very long)"),
                        ("CONF", "1", "Float", "Confidence score"),
                        ("MM_RATE", "1", "Float", "Overall mismatch rate across all copies"),
                        ("MAX_MM_PER_COPY", "1", "Integer", "Maximum mismatches in any single copy"),
                        ("N_COPIES_EVAL", "1", "Integer", "Number of copies evaluated for consensus"),
                        ("STRAND", "1", "String", "Strand of canonical motif (+/-)")):
                    fh.write(f'##INFO=<ID={ident},Number={number},Type={kind},Description="{desc}">\n')
                fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
                for i, r in enumerate(ordered):
                    fh.write(f"{r.chrom}\t{r.start + 1}\tTR{i}\t.\t<TR>\t.\tPASS\t{r.to_vcf_info()}\n")
            elif format_type == "trf_table":
                fh.write("# Tandem Repeats Finder Compatible Table Format\n")
                fh.write("# Indices\tPeriod\tCopyNumber\tConsensusSize\tPercentMatches\tPercentIndels\t")
                fh.write("Score\tA\tC\tG\tT\tEntropy\n")
                for r in ordered:
                    fh.write(r.to_trf_table() + "\n")
            elif format_type == "trf_dat":
                for r in ordered:
                    fh.write(r.to_trf_dat() + "\n")
            elif format_type == "strfinder":
                fh.write("STR_marker\tSTR_position\tSTR_motif\tSTR_genotype_structure\tSTR_genotype\t")
                fh.write("STR_core_seq\tAllele_coverage\tAlleles_ratio\tReads_Distribution(consensused)\t")
                fh.write("STR_depth\tFull_seq\tVariations\n")
                for r in ordered:
                    full = self.full_sequences.get(r.chrom, "")
                    left = full[max(0, r.start - 30):r.start] if full else ""
                    right = full[r.end:r.end + 30] if full else ""
                    fh.write(r.to_strfinder(f"STR_{r.chrom}", left, right) + "\n")


def main(argv=None):
    """CLI with the reference's flags and defaults (bwt.py:4201-4370)."""
    ap = argparse.ArgumentParser(
        description="BWT-based Tandem Repeat Finder (B200-native index-and-scan path)",
        formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("reference", help="Reference genome FASTA file")
    ap.add_argument("-o", "--output", default="repeat.tab", help="Output file (default: repeat.tab)")
    ap.add_argument("--format", choices=["bed", "vcf", "trf_table", "trf_dat", "strfinder"], default="strfinder",
                    help="Output format (default: strfinder)")
    ap.add_argument("--tier1", action="store_true", help="Enable tier 1 only (short repeats, 1-9bp)")
    ap.add_argument("--tier3", action="store_true", help="Enable tier 3 (very long repeats, kb+)")
    ap.add_argument("--long-reads", help="Long reads file for tier 3")
    ap.add_argument("--sa-sample", type=int, default=32, help="Suffix array sampling rate (default: 32)")
    ap.add_argument("--progress", action="store_true", help="Show progress bars where applicable")
    ap.add_argument("--jobs", type=int, default=4,
                    help="Number of parallel workers (default: 4, 0=use all, -1=disable parallelism)")
    ap.add_argument("--no-mismatches", action="store_true", help="Disable mismatch tolerance (exact matches only)")
    ap.add_argument("--max-motif-len", type=int, default=9, help="Maximum motif length for tier 1 (default: 9)")
    ap.add_argument("--min-period", type=int, default=10, help="Minimum period for tier 2 (default: 10)")
    ap.add_argument("--max-period", type=int, default=1000, help="Maximum period for tier 2 (default: 1000)")
    ap.add_argument("--max-unit-len", type=int, default=120,
                    help="Maximum unit length for tier 2 long repeat detection (default: 120)")
    ap.add_argument("--min-copies", type=int, default=3, help="Minimum number of copies required (default: 3)")
    ap.add_argument("--min-entropy", type=float, default=1.0,
                    help="Minimum Shannon entropy to avoid low-complexity (default: 1.0)")
    ap.add_argument("--flank-trim", type=int, default=30,
                    help="Trim N bp from each end before analysis (default: 30, use 0 to disable)")
    args = ap.parse_args(argv)

    allow_mm = not args.no_mismatches
    tier1, tier2 = (True, False) if args.tier1 else (True, True)
    tiers = "Tier 1 (short repeats)" if args.tier1 else "Tier 1 + Tier 2 (short + medium repeats)"
    if args.tier3:
        tiers += " + Tier 3 (very long repeats)"
    if args.jobs == 0:
        par = f"all {cpu_count()} CPU cores"
    elif args.jobs == -1:
        par = "disabled (sequential)"
    else:
        par = f"{args.jobs} CPU cores"
    print("BWT-based Tandem Repeat Finder")
    print("=" * 60)
    print(f"Reference:    {args.reference}")
    print(f"Output:       {args.output} ({args.format} format)")
    print(f"Tiers:        {tiers}")
    print(f"Parallelism:  {par}")
    print("")
    print("Detection Parameters:")
    print(f"  Tier 1 motif length: 1-{args.max_motif_len} bp")
    if tier2:
        print(f"  Tier 2 period range: {args.min_period}-{args.max_period} bp")
    print(f"  Min copies required: {args.min_copies}")
    print(f"  Min entropy (bits):  {args.min_entropy}")
    print(f"  Mismatch tolerance:  {'Enabled (10% of full sequence)' if allow_mm else 'Disabled (exact matches only)'}")
    print(f"  SA sampling rate:    {args.sa_sample}")
    print(f"  Flank trimming:      {args.flank_trim} bp from each end")
    print()

    finder = TandemRepeatFinder(args.reference, args.sa_sample, show_progress=args.progress,
                                allow_mismatches=allow_mm, max_motif_length=args.max_motif_len,
                                min_period=args.min_period, max_period=args.max_period, min_copies=args.min_copies,
                                min_entropy=args.min_entropy, flank_trim=args.flank_trim,
                                max_unit_len=args.max_unit_len)
    if args.min_copies < 2:
        # the strict adjacency scan counts runs of (min_copies - 1) matching unit shifts; the reference's
        # one-copy "repeats" (every position, every unit length) are not supported -- say so instead of
        # letting the worker's blanket except turn an argument error into an empty result
        raise SystemExit("error: --min-copies must be at least 2 on the B200 path")
    sequences = finder.load_reference()
    too_long = [c for c, s in sequences.items() if len(s) + 1 >= (1 << 30)]
    if too_long:
        raise SystemExit(f"error: contig(s) {', '.join(too_long)} exceed 2^30 symbols "
                         "(int32 suffix arrays, as in the reference)")
    finder.build_indices(sequences, lazy=True)
    if args.long_reads and args.tier3:
        raise SystemExit("Tier 3 (long reads) is outside the B200 hot path and is not provided")
    if args.jobs != -1:
        repeats = finder.find_tandem_repeats_parallel(enable_tier1=tier1, enable_tier2=tier2, enable_tier3=args.tier3,
                                                      long_reads=None,
                                                      n_jobs=None if args.jobs == 0 else args.jobs)
    else:
        repeats = finder.find_tandem_repeats(enable_tier1=tier1, enable_tier2=tier2, enable_tier3=args.tier3,
                                             long_reads=None)
    finder.save_results(repeats, args.output, args.format)
    print(f"\n{'=' * 60}")
    print(f"Completed! Found {len(repeats)} total tandem repeats.")
    if allow_mm and repeats:
        print(f"Average mismatch rate: {sum(r.mismatch_rate for r in repeats) / len(repeats):.3f}")
        imperfect = sum(1 for r in repeats if r.mismatch_rate > 0)
        print(f"Imperfect repeats: {imperfect} ({100 * imperfect / len(repeats):.1f}%)")
    print(f"Results saved to {args.output}")


if __name__ == "__main__":
    main()
