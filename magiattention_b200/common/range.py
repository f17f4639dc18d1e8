"""AttnRange — a half-open token span [start, end) (reference: magi_attention/common/range.py:24)."""
from __future__ import annotations

from typing import Any, Sequence


class RangeError(Exception):
    pass


class AttnRange:
    __slots__ = ("_start", "_end")

    def __init__(self, start: int, end: int) -> None:
        if not (0 <= start <= end):
            raise RangeError(f"invalid range [{start}, {end})")
        self._start, self._end = int(start), int(end)

    start = property(lambda self: self._start)
    end = property(lambda self: self._end)

    @property
    def seqlen(self) -> int:
        return self._end - self._start

    def __len__(self) -> int:
        return self.seqlen

    def to_naive_range(self) -> tuple[int, int]:
        return (self._start, self._end)

    @staticmethod
    def from_range(r: "AttnRange | Sequence[int]") -> "AttnRange":
        return r.clone() if isinstance(r, AttnRange) else AttnRange(r[0], r[1])

    def clone(self) -> "AttnRange":
        return AttnRange(self._start, self._end)

    def offset(self, off: int) -> "AttnRange":
        return AttnRange(self._start + off, self._end + off)

    def truncate(self, start: int | None = None, end: int | None = None) -> "AttnRange":
        s = self._start if start is None else max(self._start, start)
        e = self._end if end is None else min(self._end, end)
        return AttnRange(min(s, e), e) if s <= e else AttnRange(e, e)

    def intersect(self, other: "AttnRange") -> "AttnRange":
        s, e = max(self._start, other._start), min(self._end, other._end)
        return AttnRange(s, e) if s < e else AttnRange(s if s <= e else e, s if s <= e else e)

    def intersect_size(self, other: "AttnRange") -> int:
        return max(0, min(self._end, other._end) - max(self._start, other._start))

    def is_subrange_of(self, other: "AttnRange") -> bool:
        return other._start <= self._start and self._end <= other._end

    def is_overlap_with(self, other: "AttnRange") -> bool:
        return self.intersect_size(other) > 0

    def is_empty(self) -> bool:
        return self._start == self._end

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, AttnRange) and (self._start, self._end) == (other._start, other._end)

    def __hash__(self) -> int:
        return hash((self._start, self._end))

    def __repr__(self) -> str:
        return f"[{self._start}, {self._end})"
