"""``BWTCore``: the FM index of one contig, built and queried on the GPU.

Same constructor, attributes and methods as the reference class
(bwt.py:98-427); the arrays live in HBM (``DeviceIndex``) and the NumPy
attributes the reference exposes are materialised from device memory on first
access.  There is no CPU construction path.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Tuple, Union

import numpy as np

from .device_index import DeviceIndex

_MISSING = object()


class KmerTable:
    """Read-only mapping ``code -> [positions]`` over the device 8-mer index;
    behaves like the reference's ``kmer_hash`` dict (bwt.py:138-171)."""

    def __init__(self, index: DeviceIndex):
        self._ix = index
        self._off = None
        self._cache: Dict[int, List[int]] = {}

    def _offsets(self) -> np.ndarray:
        if self._off is None:
            self._off = self._ix.kmer_off.cpu().numpy()
        return self._off

    def __len__(self) -> int:
        if self._ix.kmer_count == 0:
            return 0
        return int(np.count_nonzero(np.diff(self._offsets())))

    def __bool__(self) -> bool:
        return self._ix.kmer_count > 0

    def __contains__(self, w) -> bool:
        try:
            w = int(w)
        except (TypeError, ValueError):
            return False
        if w < 0 or w > 65535:
            return False
        off = self._offsets()
        return bool(off[w + 1] > off[w])

    def __getitem__(self, w) -> List[int]:
        got = self.get(w, _MISSING)
        if got is _MISSING:
            raise KeyError(w)
        return got

    def get(self, w, default=None):
        if w not in self:
            return default
        w = int(w)
        if w not in self._cache:
            self._cache[w] = self._ix.kmer_bucket(w).tolist()
        return self._cache[w]

    def keys(self) -> Iterator[int]:
        off = self._offsets()
        return iter(np.flatnonzero(np.diff(off)).tolist())

    __iter__ = keys

    def items(self):
        return ((w, self.get(w)) for w in self.keys())

    def values(self):
        return (self.get(w) for w in self.keys())


class BWTCore:
    """FM index with the reference's surface (bwt.py:98-427)."""

    BASE_TO_BITS = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 0}
    BITS_TO_BASE = {0: "A", 1: "C", 2: "G", 3: "T"}

    def __init__(self, text: str, sa_sample_rate: int = 32, occ_sample_rate: int = 128, device=None,
                 _index: DeviceIndex = None, _pinned_text=None):
        self.text: str = text
        self.n = len(text)
        self.sa_sample_rate = sa_sample_rate
        self.occ_sample_rate = occ_sample_rate
        if _pinned_text is not None and _index is None and int(_pinned_text.numel()) == len(text):
            # the FASTA parser already wrote these bytes into pinned host memory: upload from there
            self._host: Dict[str, object] = {"text_arr": _pinned_text.numpy()}
            import torch

            dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
            self._dev = DeviceIndex(_pinned_text.to(dev, non_blocking=True), occ_rate=int(occ_sample_rate),
                                    device=dev, text_is_device=True)
        else:
            raw = text.encode("utf-8")
            self._host = {"text_arr": np.frombuffer(raw, dtype=np.uint8)}
            # one H2D copy of the text, then everything is built in HBM
            self._dev = _index if _index is not None else DeviceIndex(raw, occ_rate=int(occ_sample_rate),
                                                                      device=device)
        self.alphabet = sorted(set(text))
        self.char_to_code = {c: ord(c) for c in self.alphabet}
        self.code_to_char = {ord(c): c for c in self.alphabet}
        tot, cnt = self._dev.totals, self._dev.counts
        self.char_totals = {c: int(tot[ord(c)]) if ord(c) < 256 else 0 for c in self.alphabet}
        self.char_counts = {c: int(cnt[ord(c)]) if ord(c) < 256 else 0 for c in self.alphabet}
        self.char_counts_code = {ord(k): v for k, v in self.char_counts.items()}
        self.char_totals_code = {ord(k): v for k, v in self.char_totals.items()}

    # ---- persistence (no counterpart in the reference) -----------
    def save_index(self, path, with_lcp: bool = True) -> None:
        """Stores the device index (text, SA, BWT, Occ, C, optionally LCP) in one .npz file."""
        self.device_index.save(path, with_lcp=with_lcp)

    @classmethod
    def load_index(cls, path, sa_sample_rate: int = 32, device=None) -> "BWTCore":
        """A BWTCore over a stored index: the arrays are uploaded, nothing is sorted again."""
        ix = DeviceIndex.load(path, device=device)
        text = ix.text.cpu().numpy().tobytes().decode("utf-8")
        return cls(text, sa_sample_rate, ix.occ_rate, device=device, _index=ix)

    # ---- lazily materialised NumPy attributes ------------------------------
    def _lazy(self, name, make):
        if name not in self._host:
            self._host[name] = make()
        return self._host[name]

    @property
    def device_index(self) -> DeviceIndex:
        if self._dev is None:
            raise RuntimeError("BWTCore.clear() released the device index")
        return self._dev

    @property
    def text_arr(self) -> np.ndarray:
        return self._host["text_arr"]

    @text_arr.setter
    def text_arr(self, v):
        self._host["text_arr"] = v

    @property
    def suffix_array(self) -> np.ndarray:
        return self._lazy("suffix_array", lambda: self.device_index.sa.cpu().numpy())

    @suffix_array.setter
    def suffix_array(self, v):
        self._host["suffix_array"] = v

    @property
    def bwt_arr(self) -> np.ndarray:
        return self._lazy("bwt_arr", lambda: self.device_index.bwt.cpu().numpy())

    @bwt_arr.setter
    def bwt_arr(self, v):
        self._host["bwt_arr"] = v

    @property
    def occ_checkpoints(self) -> Dict[int, np.ndarray]:
        def make():
            d = self.device_index
            if d.n == 0:
                return {}
            occ = d.occ.cpu().numpy()
            return {int(code): occ[r, : d.ncp].copy() for r, code in enumerate(d.codes)}
        return self._lazy("occ_checkpoints", make)

    @occ_checkpoints.setter
    def occ_checkpoints(self, v):
        self._host["occ_checkpoints"] = v

    @property
    def sampled_sa(self) -> Dict[int, int]:
        def make():
            sa = self.suffix_array
            return {i: sa[i] for i in range(0, self.n, self.sa_sample_rate)}
        return self._lazy("sampled_sa", make)

    @sampled_sa.setter
    def sampled_sa(self, v):
        self._host["sampled_sa"] = v

    @property
    def kmer_hash(self):
        return self._lazy("kmer_hash", lambda: KmerTable(self.device_index))

    @kmer_hash.setter
    def kmer_hash(self, v):
        self._host["kmer_hash"] = v

    # ---- the reference's private builders, kept callable ---------------------
    def _build_kmer_hash(self, k: int = 8):
        if k != 8:
            raise NotImplementedError("the device k-mer index is built for k = 8 (bwt.py:138)")
        self._host.pop("kmer_hash", None)
        self.device_index.build_kmer()

    def _build_suffix_array(self) -> np.ndarray:
        return self.suffix_array

    def _build_bwt_array(self) -> np.ndarray:
        return self.bwt_arr

    def _build_char_counts(self) -> Tuple[Dict[str, int], Dict[str, int]]:
        return dict(self.char_counts), dict(self.char_totals)

    def _build_occurrence_checkpoints(self) -> Dict[int, np.ndarray]:
        return self.occ_checkpoints

    def _sample_suffix_array(self) -> Dict[int, int]:
        return self.sampled_sa

    # ---- queries -----------------------------------------------------------
    def get_kmer_positions(self, kmer: str) -> List[int]:
        """bwt.py:173-193, including the unpadded encoding of motifs shorter
        than 8 (they alias the 'A'-left-padded 8-mer)."""
        if len(kmer) > 8 or not self.kmer_hash:
            return self.locate_positions(kmer)
        w = 0
        for base in kmer.upper():
            if base not in self.BASE_TO_BITS:
                return []
            w = (w << 2) | self.BASE_TO_BITS[base]
        return self.kmer_hash.get(w, [])

    def clear(self):
        """bwt.py:195-210: drop the heavy structures (device memory included)."""
        self.text = ""
        self._dev = None
        self._host = {
            "text_arr": np.array([], dtype=np.uint8), "bwt_arr": np.array([], dtype=np.uint8),
            "suffix_array": np.array([], dtype=np.int32), "sampled_sa": {}, "occ_checkpoints": {},
        }
        self.char_counts = {}
        self.char_totals = {}
        self.alphabet = []
        self.char_to_code = {}
        self.code_to_char = {}
        self.char_counts_code = {}
        self.char_totals_code = {}

    def rank(self, char: Union[str, int], pos: int) -> int:
        """# of `char` in bwt[0:pos] (bwt.py:335-357)."""
        if pos <= 0 or self.n == 0:
            return 0
        code = ord(char) if isinstance(char, str) else int(char)
        if code < 0 or code > 255:
            return 0
        return int(self.device_index.rank_batch([code], [int(pos)])[0])

    def backward_search(self, pattern: str) -> Tuple[int, int]:
        """Inclusive SA interval of `pattern`, (-1,-1) if absent (bwt.py:359-389)."""
        if not pattern:
            return (0, self.n - 1)
        if self.n == 0 or any(ord(c) > 255 for c in pattern):
            return (-1, -1)
        sp, ep = self.device_index.backward_search_batch([pattern.encode("latin1")])
        return (int(sp[0]), int(ep[0]))

    def backward_search_many(self, patterns) -> Tuple[np.ndarray, np.ndarray]:
        """Batched form of ``backward_search`` (one kernel launch)."""
        patterns = list(patterns)
        if self.n == 0:      # empty text: every non-empty pattern is absent (bwt.py:371-389)
            sp = np.array([0 if not p else -1 for p in patterns], np.int32)
            return sp, np.where(sp == 0, self.n - 1, -1).astype(np.int32)
        return self.device_index.backward_search_batch(patterns)

    def count_occurrences(self, pattern: str) -> int:
        sp, ep = self.backward_search(pattern)
        return 0 if sp == -1 else ep - sp + 1

    def locate_positions(self, pattern: str) -> List[int]:
        """Sorted text positions of `pattern` (bwt.py:398-410: slice of the full SA)."""
        sp, ep = self.backward_search(pattern)
        if sp == -1:
            return []
        out = self.device_index.sa[sp:ep + 1].cpu().numpy().tolist()
        out.sort()
        return out

    def _get_suffix_position(self, sa_index: int) -> int:
        """LF walk to the nearest sampled SA entry (bwt.py:412-426)."""
        sampled = self.sampled_sa
        if sa_index in sampled:
            return sampled[sa_index]
        steps, cur = 0, sa_index
        bwt = self.bwt_arr
        while cur not in sampled:
            code = int(bwt[cur])
            cur = self.char_counts_code[code] + self.rank(code, cur)
            steps += 1
        return (sampled[cur] + steps) % self.n
