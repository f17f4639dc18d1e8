"""Peer-memory KV fetch: the stage exchange as copy-engine pulls over NVLink instead of NCCL P2P kernels.

EXPERIMENTAL / OPT-IN (``MAGI_B200_KV_EXCHANGE=peer``), written after the round-1 GPU budget was spent and NOT YET RUN ON
HARDWARE; the default engine is ``group_cast`` over NCCL (comm/group_collective.py).  Design:

  * every rank PUBLISHES its K,V shard once per pass in a symmetric-memory buffer (torch.distributed._symmetric_memory:
    same virtual layout on every rank, peers mapped over NVLink);
  * a consumer PULLS the row spans of each stage straight from the owners' published buffers into the stage buffer with
    device-to-device copies on a side stream — no communication kernel occupies SMs next to the attention grid, and the
    spans land in the order the stage's ``k_ranges`` index them, exactly like the receive side of ``group_cast``;
  * two symmetric-memory barriers per pass order things across ranks: "everybody has finished pulling from my buffer"
    before it is overwritten, and "everybody has published" before anyone pulls.  Both carry a timeout, so a protocol
    error ends in an exception instead of a hung box.

The pull tables are derived on the host from the planner's transfer tables of ALL ranks (the planner is deterministic, every
rank holds every table): ``make_pull_plan`` is pure Python and is unit-tested on CPU against the naive world-level group_cast.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .work import GeneralWork, WorkWithPostProcessFn

__all__ = ["make_pull_plan", "make_reduce_pull_plan", "PeerKVFetcher"]


def make_pull_plan(cast_args_all_ranks: list, rank: int) -> list[tuple[int, int, int, int]]:
    """Pull table of ``rank`` for one stage: ``(src_rank, src_row0, dst_row0, nrows)`` in receive-buffer order.

    ``cast_args_all_ranks[r]`` is rank r's GroupCollectiveArg of that stage.  The j-th output split of the receiver that comes
    from source s is the j-th input split of s whose destination list contains the receiver (splits of one source keep the
    source's order — the group_cast contract), so its rows are found by walking s's split table."""
    mine = cast_args_all_ranks[rank]
    # per source: row offsets of its input splits that go to `rank`, in order
    offsets: dict[int, list[tuple[int, int]]] = {}
    for s, arg in enumerate(cast_args_all_ranks):
        off, lst = 0, []
        for sz, dsts in zip(arg.input_split_size_list, arg.dst_indices_list):
            if sz > 0 and rank in dsts:
                lst.append((off, sz))
            off += sz
        offsets[s] = lst
    cursor = {s: 0 for s in offsets}
    plan: list[tuple[int, int, int, int]] = []
    dst = 0
    for sz, s in zip(mine.output_split_size_list, mine.src_index_list):
        if sz > 0:
            s0, ssz = offsets[s][cursor[s]]
            cursor[s] += 1
            assert ssz == sz, "transfer tables of sender and receiver disagree"
            if plan and plan[-1][0] == s and plan[-1][1] + plan[-1][3] == s0 and plan[-1][2] + plan[-1][3] == dst:
                plan[-1] = (s, plan[-1][1], plan[-1][2], plan[-1][3] + sz)  # contiguous on both sides: one copy
            else:
                plan.append((s, s0, dst, sz))
        dst += sz
    return plan


def make_reduce_pull_plan(cast_args_all_ranks: list, owner: int) -> list[tuple[int, int, int, int]]:
    """The return path (partial dK/dV of one stage) as pulls by the OWNER: ``(consumer_rank, src_row0 in the consumer's stage
    buffer, dst_row0 in the owner's local rows, nrows)``; the owner adds every listed span into its dk/dv (the fused row-span
    add can read the peer rows directly = reduce-on-receive).  Transpose of :func:`make_pull_plan`: entries with the same
    destination span are adjacent, which is what ``b200_range_reduce_sum`` needs to fold them in one race-free pass."""
    mine = cast_args_all_ranks[owner]
    # per consumer: offsets (in ITS stage buffer) of the output splits that came from `owner`, in order
    offsets: dict[int, list[tuple[int, int]]] = {}
    for d, arg in enumerate(cast_args_all_ranks):
        off, lst = 0, []
        for sz, s in zip(arg.output_split_size_list, arg.src_index_list):
            if sz > 0 and s == owner:
                lst.append((off, sz))
            off += sz
        offsets[d] = lst
    cursor = {d: 0 for d in offsets}
    plan: list[tuple[int, int, int, int]] = []
    loc = 0
    for sz, dsts in zip(mine.input_split_size_list, mine.dst_indices_list):
        if sz > 0:
            for d in dsts:
                s0, ssz = offsets[d][cursor[d]]
                cursor[d] += 1
                assert ssz == sz, "transfer tables of sender and receiver disagree"
                plan.append((d, s0, loc, sz))
        loc += sz
    return plan


class PeerKVFetcher:
    """Publishes this rank's (K, V) shard and pulls every stage's remote rows from the peers' published shards."""

    BARRIER_TIMEOUT_MS = 20_000

    def __init__(self, group: dist.ProcessGroup, rank: int, cast_args_per_stage_all_ranks: list[list], shard_rows: int,
                 num_heads_kv: int, head_dim: int, dtype: torch.dtype, device: torch.device):
        import torch.distributed._symmetric_memory as symm_mem

        self.group, self.rank, self.world = group, rank, dist.get_world_size(group)
        self.shape = (2, shard_rows, num_heads_kv, head_dim)
        self.dtype, self.device = dtype, device
        self.plans = [make_pull_plan(stage_args, rank) for stage_args in cast_args_per_stage_all_ranks]
        self.rows = [sum(p[3] for p in plan) for plan in self.plans]
        self.buf = symm_mem.empty(*self.shape, dtype=dtype, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.peers = [self.hdl.get_buffer(r, self.shape, dtype) for r in range(self.world)]
        self.side = torch.cuda.Stream(device=device, priority=-1)

    def fetch(self, k: torch.Tensor, v: torch.Tensor) -> list[WorkWithPostProcessFn | None]:
        """Same contract as DistAttnRuntime.fetch_remote_kv: one work per stage whose ``wait_post_process()`` orders the current
        stream behind that stage's pulls and returns ``(k_stage, v_stage)``."""
        cur = torch.cuda.current_stream(self.device)
        # (1) nobody is still pulling last pass's rows out of my buffer; (2) publish; (3) everybody has published
        self.hdl.barrier(channel=0, timeout_ms=self.BARRIER_TIMEOUT_MS)
        self.buf[0].copy_(k)
        self.buf[1].copy_(v)
        self.hdl.barrier(channel=1, timeout_ms=self.BARRIER_TIMEOUT_MS)
        stage_bufs = [(k.new_empty((n, *k.shape[1:])), v.new_empty((n, *v.shape[1:]))) if n else None for n in self.rows]
        self.side.wait_stream(cur)
        works: list[WorkWithPostProcessFn | None] = []
        with torch.cuda.stream(self.side):
            for plan, bufs in zip(self.plans, stage_bufs):
                if bufs is None:
                    works.append(None)
                    continue
                rk, rv = bufs
                for (src, s0, d0, n) in plan:
                    rk[d0:d0 + n].copy_(self.peers[src][0, s0:s0 + n], non_blocking=True)
                    rv[d0:d0 + n].copy_(self.peers[src][1, s0:s0 + n], non_blocking=True)
                rk.record_stream(self.side)
                rv.record_stream(self.side)
                ev = self.side.record_event()
                works.append(WorkWithPostProcessFn(GeneralWork(ev), (lambda b: (lambda *a, **kw: b))(bufs), async_op=True))
        return works
