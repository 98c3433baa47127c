"""Result record types of the ``bwt`` module surface.

``TandemRepeat`` mirrors the reference dataclass (bwt.py:429-452) field for
field, and its five line formatters produce the same bytes as the reference's
(bwt.py:454-641).  Kernels return integer rows (start, end, period, copies,
mismatch counts); everything floating point in a record is derived on the host
with the reference's expression order.
"""
from __future__ import annotations

import functools
import math
import re
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple


def _natural_sort_key_uncached(value):
    if value is None:
        return ()
    key = []
    for chunk in re.split(r"(\d+)", str(value)):
        if chunk == "":
            continue
        key.append((0, int(chunk)) if chunk.isdigit() else (1, chunk.lower()))
    return tuple(key)


_natural_sort_key_cached = functools.lru_cache(maxsize=1 << 16)(_natural_sort_key_uncached)


def _natural_sort_key(value):
    """Natural ordering key: digit runs compare as integers, the rest
    case-insensitively (bwt.py:22-36).  Memoised: the record sorts call it once per record
    with a handful of distinct contig names."""
    try:
        return _natural_sort_key_cached(value)
    except TypeError:      # unhashable value
        return _natural_sort_key_uncached(value)


_DEFAULT_COMP = {"A": 25.0, "C": 25.0, "G": 25.0, "T": 25.0}


@dataclass
class TandemRepeat:
    """One tandem-repeat call (bwt.py:429-452)."""
    chrom: str
    start: int
    end: int
    motif: str
    copies: float
    length: int
    tier: int
    confidence: float = 1.0
    consensus_motif: Optional[str] = None
    mismatch_rate: float = 0.0
    max_mismatches_per_copy: int = 0
    n_copies_evaluated: int = 0
    strand: str = "+"
    percent_matches: float = 0.0
    percent_indels: float = 0.0
    score: int = 0
    composition: Optional[Dict[str, float]] = None
    entropy: float = 0.0
    actual_sequence: Optional[str] = None
    variations: Optional[List[str]] = None

    # -- helpers ---------------------------------------------------------
    def _cons(self) -> str:
        return self.consensus_motif or self.motif

    def _comp(self) -> Dict[str, float]:
        return self.composition or _DEFAULT_COMP

    # -- formatters (bwt.py:454-641) --------------------------------------
    def to_bed(self) -> str:
        return "\t".join((
            str(self.chrom), str(self.start), str(self.end), self._cons(), f"{self.copies:.1f}",
            str(self.tier), f"{self.mismatch_rate:.3f}", self.strand))

    def to_vcf_info(self) -> str:
        fields = (
            ("MOTIF", self.motif), ("CONS_MOTIF", self._cons()), ("COPIES", f"{self.copies:.1f}"),
            ("TIER", self.tier), ("CONF", f"{self.confidence:.2f}"), ("MM_RATE", f"{self.mismatch_rate:.3f}"),
            ("MAX_MM_PER_COPY", self.max_mismatches_per_copy), ("N_COPIES_EVAL", self.n_copies_evaluated),
            ("STRAND", self.strand))
        return ";".join(f"{k}={v}" for k, v in fields)

    def _trf_numbers(self) -> List[str]:
        cons, comp = self._cons(), self._comp()
        return [str(len(cons)), f"{self.copies:.1f}", str(len(cons)), f"{self.percent_matches:.0f}",
                f"{self.percent_indels:.0f}", str(self.score), f"{comp['A']:.0f}", f"{comp['C']:.0f}",
                f"{comp['G']:.0f}", f"{comp['T']:.0f}", f"{self.entropy:.2f}"]

    def to_trf_table(self) -> str:
        return "\t".join([f"{self.start}--{self.end}"] + self._trf_numbers())

    def to_trf_dat(self) -> str:
        cons = self._cons()
        seq = self.actual_sequence or (cons * int(self.copies))
        return " ".join([str(self.start), str(self.end)] + self._trf_numbers() + [cons, seq])

    def to_strfinder(self, marker_name: Optional[str] = None, flanking_left: str = "",
                     flanking_right: str = "") -> str:
        marker = marker_name or f"STR_{self.chrom}_{self.start}"
        compound = bool(getattr(self, "is_compound", False)) and hasattr(self, "compound_partner")
        if compound:
            other = self.compound_partner
            c1, c2 = self._cons(), (other.consensus_motif or other.motif)
            n1, n2 = int(round(self.copies)), int(round(other.copies))
            core = (self.actual_sequence or c1 * n1) + (other.actual_sequence or c2 * n2)
            full = (flanking_left + core + flanking_right) if (flanking_left or flanking_right) else core
            cols = (marker, f"{self.chrom}:{self.start + 1}-{other.end}", f"[{c1}]n+[{c2}]n",
                    f"{len(c1)}[{c1}]{n1};{len(c2)}[{c2}]{n2},0", f"{n1}/{n2}", core, "100%", "-",
                    f"{n1}:{n2}", str(n1 + n2), full, "-")
            return "\t".join(cols)

        cons = self._cons()
        whole = int(math.floor(self.copies + 1e-6))
        leftover = (self.end - self.start) - len(cons) * whole
        if abs(self.copies - round(self.copies)) < 1e-6:
            genotype = str(int(round(self.copies)))
        else:
            genotype = f"{self.copies:.2f}".rstrip("0").rstrip(".")
        core_full = self.actual_sequence if self.actual_sequence else cons * int(self.copies)
        core = f"{core_full[:70]}... (x{whole})" if len(core_full) > 150 else core_full
        if getattr(self, "percent_matches", None) is not None:
            coverage = f"{self.percent_matches:.0f}%"
        else:
            coverage = f"{self.confidence * 100:.0f}%"
        full_all = (flanking_left + core_full + flanking_right) if (flanking_left or flanking_right) else core_full
        full = f"{full_all[:250]}...{full_all[-200:]}" if len(full_all) > 500 else full_all
        cols = (marker, f"{self.chrom}:{self.start + 1}-{self.end}", f"[{cons}]n",
                f"{len(cons)}[{cons}]{whole},{leftover}", genotype, core, coverage, "-",
                f"{whole}:{self.n_copies_evaluated}", str(self.n_copies_evaluated), full,
                ";".join(self.variations) if self.variations else "-")
        return "\t".join(cols)


@dataclass
class AlignmentResult:
    """Outcome of aligning one copy against the motif template (bwt.py:644-658)."""
    consumed: int
    unit_sequence: str
    mismatch_count: int
    insertion_length: int
    deletion_length: int
    operations: List[Tuple]
    observed_bases: List[Tuple[int, str]]
    edit_distance: int

    @property
    def error_count(self) -> int:
        return self.mismatch_count + self.insertion_length + self.deletion_length


@dataclass
class RepeatAlignmentSummary:
    """Aggregate of the per-copy alignments of one repeat block (bwt.py:661-674)."""
    consensus: str
    motif_len: int
    copies: int
    consumed_length: int
    mismatch_rate: float
    max_errors_per_copy: int
    variations: List[str]
    copy_sequences: List[str]
    total_insertions: int
    total_deletions: int
    error_counts: List[int]
