"""AttnRanges — an ordered list of AttnRange (reference: magi_attention/common/ranges.py:101)."""
from __future__ import annotations

from typing import Any, Iterator, Sequence

import torch

from .range import AttnRange


class AttnRanges:
    def __init__(self) -> None:
        self._ranges: list[AttnRange] = []

    # ---- construction -------------------------------------------------------------
    @staticmethod
    def from_ranges(ranges: "AttnRanges | Sequence[Sequence[int]] | Sequence[AttnRange]", check: bool = False) -> "AttnRanges":
        out = AttnRanges()
        for r in ranges:
            out._ranges.append(AttnRange.from_range(r))
        return out

    @staticmethod
    def from_cu_seqlens(cu_seqlens: Sequence[int], seq_len: int | None = None) -> "AttnRanges":
        out = AttnRanges()
        for a, b in zip(cu_seqlens[:-1], cu_seqlens[1:]):
            out._ranges.append(AttnRange(int(a), int(b)))
        return out

    def append(self, r: AttnRange, check: bool = False) -> None:
        self._ranges.append(r)

    def extend(self, other: "AttnRanges", check: bool = False) -> None:
        self._ranges.extend(other._ranges)

    def clone(self) -> "AttnRanges":
        return AttnRanges.from_ranges(self._ranges)

    # ---- views ----------------------------------------------------------------------
    def to_naive_ranges(self) -> list[tuple[int, int]]:
        return [r.to_naive_range() for r in self._ranges]

    def to_tensor(self, device: str | torch.device = "cpu") -> torch.Tensor:
        if not self._ranges:
            return torch.empty(0, 2, dtype=torch.int32, device=device)
        return torch.tensor(self.to_naive_ranges(), dtype=torch.int32, device=device)

    def to_cu_seqlens(self, seq_len: int) -> list[int]:
        assert self.is_cu_seqlens(seq_len)
        return [0] + [r.end for r in self._ranges]

    def is_cu_seqlens(self, seq_len: int) -> bool:
        pos = 0
        for r in self._ranges:
            if r.start != pos:
                return False
            pos = r.end
        return pos == seq_len

    # ---- algebra ----------------------------------------------------------------------
    def sort(self) -> "AttnRanges":
        return AttnRanges.from_ranges(sorted(self._ranges, key=lambda r: (r.start, r.end)))

    def merge(self) -> "AttnRanges":
        """Sorted union of the spans (adjacent / overlapping spans fused)."""
        out = AttnRanges()
        for r in self.sort():
            if r.is_empty():
                continue
            if out._ranges and r.start <= out._ranges[-1].end:
                last = out._ranges[-1]
                out._ranges[-1] = AttnRange(last.start, max(last.end, r.end))
            else:
                out._ranges.append(r.clone())
        return out

    def is_sorted(self) -> bool:
        return all(a.start <= b.start for a, b in zip(self._ranges, self._ranges[1:]))

    def is_non_overlap(self) -> bool:
        s = self.sort()._ranges
        return all(a.end <= b.start for a, b in zip(s, s[1:]))

    def is_merged(self) -> bool:
        s = self._ranges
        return self.is_sorted() and all(a.end < b.start for a, b in zip(s, s[1:]))

    def make_range_local(self, r: AttnRange) -> AttnRange:
        """Map global span `r` (contained in one span of this MERGED list) to the packed-local axis."""
        off = 0
        for host in self._ranges:
            if r.is_subrange_of(host):
                return AttnRange(off + r.start - host.start, off + r.end - host.start)
            off += host.seqlen
        raise ValueError(f"{r} is not inside any of {self}")

    def make_ranges_local(self, other: "AttnRanges") -> "AttnRanges":
        return AttnRanges.from_ranges([self.make_range_local(r) for r in other])

    def find_overlap_ranges(self, other: "AttnRanges") -> "AttnRanges":
        a, b = self.merge()._ranges, other.merge()._ranges
        out = AttnRanges()
        i = j = 0
        while i < len(a) and j < len(b):
            s, e = max(a[i].start, b[j].start), min(a[i].end, b[j].end)
            if s < e:
                out._ranges.append(AttnRange(s, e))
            if a[i].end < b[j].end:
                i += 1
            else:
                j += 1
        return out

    def find_hole_ranges(self, other: "AttnRanges") -> "AttnRanges":
        """Parts of self (merged) not covered by other."""
        out = AttnRanges()
        cov = other.merge()._ranges
        for r in self.merge()._ranges:
            pos = r.start
            for c in cov:
                if c.end <= pos:
                    continue
                if c.start >= r.end:
                    break
                if c.start > pos:
                    out._ranges.append(AttnRange(pos, c.start))
                pos = max(pos, c.end)
            if pos < r.end:
                out._ranges.append(AttnRange(pos, r.end))
        return out

    @property
    def total_seqlen(self) -> int:
        return sum(r.seqlen for r in self._ranges)

    @property
    def max_seqlen(self) -> int:
        return max((r.seqlen for r in self._ranges), default=0)

    @property
    def start(self) -> int:
        return min(r.start for r in self._ranges)

    @property
    def end(self) -> int:
        return max(r.end for r in self._ranges)

    @property
    def size(self) -> int:
        return len(self._ranges)

    def is_empty(self) -> bool:
        return not self._ranges

    def __len__(self) -> int:
        return len(self._ranges)

    def __getitem__(self, idx: int | slice):
        if isinstance(idx, slice):
            return AttnRanges.from_ranges(self._ranges[idx])
        return self._ranges[idx]

    def __iter__(self) -> Iterator[AttnRange]:
        return iter(self._ranges)

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, AttnRanges) and self._ranges == other._ranges

    def __hash__(self) -> int:
        return hash(tuple(self._ranges))

    def __repr__(self) -> str:
        return "[" + ", ".join(repr(r) for r in self._ranges) + "]"
