"""Async work handles of the comm layer (reference: magi_attention/comm/work.py:29-114).

``WorkWithPostProcessFn.wait_post_process(*bufs)`` = make the CURRENT stream wait for the
communication, then run the post-process function exactly once and return its result.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable

import torch
import torch.distributed as dist


@dataclass
class GeneralWork:
    """Anything that can be waited on from the current stream: dist.Work(s), a CUDA event or a stream."""

    works: Any = None

    def wait(self) -> None:
        w = self.works
        if w is None:
            return
        items = w if isinstance(w, (list, tuple)) else [w]
        for it in items:
            if it is None:
                continue
            if isinstance(it, torch.cuda.Event):
                torch.cuda.current_stream().wait_event(it)
            elif isinstance(it, torch.cuda.Stream):
                torch.cuda.current_stream().wait_stream(it)
            else:  # dist.Work
                it.wait()
        self.works = None


def _identity(*args):
    return args[0] if len(args) == 1 else args


@dataclass
class WorkWithPostProcessFn:
    work: GeneralWork | None = None
    post_process_fn: Callable = _identity
    async_op: bool = False
    _done: bool = field(default=False, init=False)

    def __post_init__(self):
        if self.work is None:
            self.work = GeneralWork(None)
        elif not isinstance(self.work, GeneralWork):
            self.work = GeneralWork(self.work)
        if not self.async_op:
            # synchronous op: the communication is already ordered before the caller's next kernels
            self.work.wait()

    def wait_post_process(self, *args, **kwargs):
        if self._done:
            raise RuntimeError("wait_post_process may only be called once")
        self.work.wait()
        self._done = True
        return self.post_process_fn(*args, **kwargs)
