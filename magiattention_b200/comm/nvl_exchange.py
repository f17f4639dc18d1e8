"""NVLink peer-memory engine of the context-parallel KV / dKV exchange (one NVSwitch domain, one process per GPU).

What it replaces: the reference moves the remote K,V of every overlap stage with ``group_cast`` and returns the partial
dK,dV with ``group_reduce`` (magi_attention/functional/dist_attn.py:1463-1567 ``_fetch_remote_kv``, :2068-2195
``_reduce_partial_dkv``) over its a2av engine (comm/primitive/grpcoll/_a2av_grpcoll_impl.py:69-340): a range_gather pass,
one NCCL ``all_to_all_single`` and a range_reduce pass per call, the NCCL kernels competing with the attention grid for SMs
(hence ``sm_margin``, env/comm.py:44-69).  Here, on one NVSwitch domain:

  * every rank owns one IPC-exported ARENA (``b200_nvl_alloc``) that all peers map (``b200_nvl_import``); it holds
    epoch flags, two slots of the rank's published K,V shard and two slots of its per-stage partial dK,dV;
  * bulk rows move ONLY with copy engines: a consumer PULLS the row spans of a stage straight out of the owners' published
    shards into its stage buffer (``b200_nvl_copy_spans``), an owner PULLS the partial dK,dV spans of a stage out of the
    consumers' arenas and folds them into its fp32 dk/dv with the fused row-span add (``b200_range_reduce_sum``).  No
    communication kernel occupies an SM next to the attention grid, there is no rendezvous between ranks and no
    duplication of multi-destination splits;
  * ordering across ranks is three families of monotonically increasing 32-bit flags, written into the PEER's arena by a
    one-thread kernel after the producing work of a stream (``b200_nvl_signal``) and awaited by a one-warp kernel with a
    timeout (``b200_nvl_wait``):
        kv_ready[s]       rank s has published its K,V for pass p
        part_ready[s][st] rank s has finished the stage-st backward kernel of pass p (its partial dK,dV are complete)
        pulls_done[s]     rank s has finished every pull of pass p (K,V and partials) from anybody
    a pass p (one forward or one backward of one calc_attn) uses slot p & 1 and may overwrite it once
    ``pulls_done[*] >= p - 2``.  Every rank runs the same sequence of passes (SPMD), so p needs no negotiation.

The pull tables come from the planner's transfer tables of ALL ranks (the planner is deterministic: every rank holds every
table); ``make_pull_plan`` / ``make_reduce_pull_plan`` are pure Python and unit-tested on CPU against the naive world-level
group_cast / group_reduce.  The NCCL engine (comm/group_collective.py) stays as the portable alternative
(``MAGI_B200_KV_EXCHANGE=nccl``) and as the reference point in the GPU tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .work import GeneralWork, WorkWithPostProcessFn

__all__ = ["make_pull_plan", "make_reduce_pull_plan", "NvlDomain", "NvlKVExchange", "nvl_available"]


# --------------------------------------------------------------------------------------------------------------------
# host-side pull tables (pure Python)
# --------------------------------------------------------------------------------------------------------------------
def make_pull_plan(cast_args_all_ranks: list, rank: int) -> list[tuple[int, int, int, int]]:
    """Pull table of ``rank`` for one stage: ``(src_rank, src_row0, dst_row0, nrows)`` in receive-buffer order.

    ``cast_args_all_ranks[r]`` is rank r's GroupCollectiveArg of that stage.  The j-th output split of the receiver that comes
    from source s is the j-th input split of s whose destination list contains the receiver (splits of one source keep the
    source's order — the group_cast contract), so its rows are found by walking s's split table."""
    mine = cast_args_all_ranks[rank]
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
    buffer, dst_row0 in the owner's local rows, nrows)``; the owner adds every listed span into its dk/dv.  Transpose of
    :func:`make_pull_plan`: entries with the same destination span are adjacent, which is what ``b200_range_reduce_sum``
    needs to fold them in one race-free pass."""
    mine = cast_args_all_ranks[owner]
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


# --------------------------------------------------------------------------------------------------------------------
# the arena domain: one per process group
# --------------------------------------------------------------------------------------------------------------------
_FLAG_BYTES = 16384
_MAX_STAGES = 32
_W_KV_READY, _W_PULLS_DONE, _W_PART_READY = 0, 64, 128
_ALIGN = 1 << 20


def _up(x: int, a: int = _ALIGN) -> int:
    return (x + a - 1) // a * a


class _RawCuda:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view it without owning it."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2,
                                         "strides": None}
        self._owner = owner


_avail_cache: dict = {}


def nvl_available(group) -> bool:
    """True when every rank of ``group`` can reach every other over CUDA peer access from one host (collective on first
    use per group, cached afterwards)."""
    key = (id(group), os.environ.get("MAGI_B200_KV_EXCHANGE", "nvl"))
    if key not in _avail_cache:
        _avail_cache[key] = _nvl_probe(group)
    return _avail_cache[key]


def _nvl_probe(group) -> bool:
    if not torch.cuda.is_available() or group is None or dist.get_backend(group) != "nccl":
        return False
    if os.environ.get("MAGI_B200_KV_EXCHANGE", "nvl") != "nvl":
        return False
    world = dist.get_world_size(group)
    if world == 1 or world > 64:
        return False
    dev = torch.cuda.current_device()
    import socket

    info = [None] * world
    dist.all_gather_object(info, (socket.gethostname(), dev), group=group)
    if len({h for h, _ in info}) != 1 or len({d for _, d in info}) != world:
        return False
    ok = all(d == dev or torch.cuda.can_device_access_peer(dev, d) for _, d in info)
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok), group=group)
    return all(flags)


class NvlDomain:
    """This rank's arena + the mapped arenas of its peers, the side streams and the pass counter of one group."""

    _domains: dict = {}
    TIMEOUT_NS = int(float(os.environ.get("MAGI_B200_NVL_TIMEOUT_S", "30")) * 1e9)

    @classmethod
    def get(cls, group) -> "NvlDomain":
        key = id(group)
        d = cls._domains.get(key)
        if d is None:
            d = cls(group)
            cls._domains[key] = d
        return d

    def __init__(self, group):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.kv_slot_bytes = 0
        self.part_slot_bytes = 0
        self.base = 0
        self.peer_base: list[int] = []
        self.pass_id = 0
        self.generation = 0  # bumped whenever the arenas are re-created (cached flag / pointer tables key on it)
        self.view: torch.Tensor | None = None
        self.fetch_stream = torch.cuda.Stream(device=self.device, priority=-1)
        self.reduce_stream = torch.cuda.Stream(device=self.device, priority=-1)
        # One stream feeds one copy engine at a time (~460 GB/s measured for a single peer-to-peer copy); the span copies of
        # a stage are therefore dealt out over a few helper streams so that several copy engines / NVLink ports work at once.
        n_aux = int(os.environ.get("MAGI_B200_NVL_COPY_STREAMS", "4")) - 1
        self.aux_streams = [torch.cuda.Stream(device=self.device, priority=-1) for _ in range(max(0, n_aux))]
        self.err = torch.zeros(4, dtype=torch.int32).pin_memory()
        self.bytes_pulled = 0  # accounting: bytes this rank pulled over NVLink (for exposed-comm / bandwidth reports)

    # ---- layout -------------------------------------------------------------------------------------------------
    def copy_spans(self, spans: list[tuple[int, int, int]], main: "torch.cuda.Stream") -> int:
        """Enqueue ``(dst, src, nbytes)`` span copies behind everything already on ``main`` and make ``main`` wait for all of
        them; the spans are balanced by bytes over ``main`` + the helper streams.  Returns the bytes moved."""
        from .. import _lib

        lanes: list[list] = [[] for _ in range(1 + len(self.aux_streams))]
        load = [0] * len(lanes)
        for sp in sorted(spans, key=lambda x: -x[2]):
            i = load.index(min(load))
            lanes[i].append(sp)
            load[i] += sp[2]
        used = [i for i in range(1, len(lanes)) if lanes[i]]
        if used:
            start = main.record_event()
        for i in used:
            st = self.aux_streams[i - 1]
            st.wait_event(start)
            _lib.NvlCopyList(lanes[i]).launch(st.cuda_stream)
        if lanes[0]:
            _lib.NvlCopyList(lanes[0]).launch(main.cuda_stream)
        for i in used:
            main.wait_event(self.aux_streams[i - 1].record_event())
        return sum(load)

    def kv_off(self, slot: int) -> int:
        return _FLAG_BYTES + slot * self.kv_slot_bytes

    def part_off(self, slot: int) -> int:
        return _FLAG_BYTES + 2 * self.kv_slot_bytes + slot * self.part_slot_bytes

    def flag_addr(self, rank: int, word: int) -> int:
        return self.peer_base[rank] + 4 * word

    def ensure(self, kv_bytes: int, part_bytes: int) -> None:
        """Grow the arenas (collectively: every rank calls this with the same arguments) if a slot is too small."""
        if kv_bytes <= self.kv_slot_bytes and part_bytes <= self.part_slot_bytes:
            return
        from .. import _lib

        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self._release()
        self.kv_slot_bytes = _up(max(kv_bytes, self.kv_slot_bytes))
        self.part_slot_bytes = _up(max(part_bytes, self.part_slot_bytes, 1))
        total = _FLAG_BYTES + 2 * self.kv_slot_bytes + 2 * self.part_slot_bytes
        self.base = _lib.nvl_alloc(total)
        self.total_bytes = total
        handles = [None] * self.world
        dist.all_gather_object(handles, _lib.nvl_export(self.base), group=self.group)
        self.peer_base = [self.base if r == self.rank else _lib.nvl_import(h) for r, h in enumerate(handles)]
        self.view = self._map_view(total)
        self.pass_id = 0
        self.generation += 1
        dist.barrier(group=self.group)

    def _map_view(self, total: int) -> torch.Tensor:
        """This rank's arena as a uint8 tensor (no ownership: the arena outlives it)."""
        v = torch.as_tensor(_RawCuda(self.base, total, self), device=self.device)
        assert v.data_ptr() == self.base and v.device == self.device
        return v

    def _release(self) -> None:
        if not self.base:
            return
        from .. import _lib

        self.view = None
        for r, p in enumerate(self.peer_base):
            if r != self.rank:
                _lib.nvl_release(p)
        _lib.nvl_free(self.base)
        self.base, self.peer_base = 0, []

    def check(self) -> None:
        if int(self.err[0]) != 0:
            raise RuntimeError("NVLink exchange: a flag wait timed out (a peer did not reach the matching pass); "
                               "set MAGI_B200_KV_EXCHANGE=nccl to fall back to the NCCL engine")

    @classmethod
    def shutdown(cls) -> None:
        for d in cls._domains.values():
            try:
                torch.cuda.synchronize(d.device)
                d._release()
            except Exception:
                pass
        cls._domains.clear()


# --------------------------------------------------------------------------------------------------------------------
# per-key exchange: static pull tables + the per-pass protocol
# --------------------------------------------------------------------------------------------------------------------
class NvlKVExchange:
    """The stage exchange of one DistAttnRuntime over an :class:`NvlDomain`.

    ``all_comm_metas[r]`` is rank r's CommMeta (every rank derives all of them locally)."""

    def __init__(self, group, all_comm_metas: list, overlap_degree: int, shard_rows: int):
        self.dom = NvlDomain.get(group)
        self.rank, self.world = self.dom.rank, self.dom.world
        self.degree = overlap_degree
        assert overlap_degree <= _MAX_STAGES
        self.shard_rows = shard_rows
        per_stage = [[cm.kv_group_collective_args_list[st] for cm in all_comm_metas] for st in range(overlap_degree)]
        self.pull = [make_pull_plan(a, self.rank) for a in per_stage]            # K,V rows I pull per stage
        self.rpull = [make_reduce_pull_plan(a, self.rank) for a in per_stage]     # partial rows I (owner) pull per stage
        self.rows = [sum(p[3] for p in plan) for plan in self.pull]
        # stage-buffer rows of EVERY rank (their partial regions are laid out stage after stage)
        self.rows_all = [[int(cm.num_remote_kv_tokens_per_stage[st]) for st in range(overlap_degree)] for cm in all_comm_metas]
        self.max_part_rows = max(sum(r) for r in self.rows_all)
        self.kv_sources = [sorted({p[0] for p in plan}) for plan in self.pull]
        self.part_sources = [sorted({p[0] for p in plan}) for plan in self.rpull]
        self.owners = self.kv_sources  # the owners of the rows of my stage st = the ranks I pulled them from
        self.peers = [r for r in range(self.world) if r != self.rank]
        self._flag_cache: dict = {}
        self._reduce_meta: dict = {}
        self._cur = None

    # ---- small helpers ------------------------------------------------------------------------------------------
    def _flags(self, key, addrs_fn):
        from .. import _lib

        k = (key, self.dom.generation)
        f = self._flag_cache.get(k)
        if f is None:
            f = _lib.NvlFlagList(addrs_fn())
            self._flag_cache[k] = f
        return f

    def _signal(self, key, addrs_fn, value, stream):
        self._flags(key, addrs_fn).signal(value, stream.cuda_stream)

    def _wait(self, key, addrs_fn, value, stream):
        self._flags(key, addrs_fn).wait(value, self.dom.err.data_ptr(), self.dom.TIMEOUT_NS, stream.cuda_stream)

    # ---- a pass -------------------------------------------------------------------------------------------------
    def fetch(self, k: torch.Tensor, v: torch.Tensor, wire_esize: int | None = None,
              backward: bool = False) -> list[WorkWithPostProcessFn | None]:
        """Start a pass: publish (k, v), pull every stage's remote rows.  Returns one work per stage (None when this rank
        pulls nothing in it) whose ``wait_post_process()`` orders the current stream behind the stage's pulls and yields
        ``(k_stage, v_stage)`` — the contract of DistAttnRuntime.fetch_remote_kv."""
        from .. import _lib

        dom = self.dom
        dom.check()
        assert k.is_contiguous() and v.is_contiguous() and k.shape == v.shape and k.shape[0] == self.shard_rows
        row_bytes = k.shape[1] * k.shape[2] * k.element_size()
        kv_bytes = 2 * self.shard_rows * row_bytes
        part_row_bytes = k.shape[1] * k.shape[2] * (wire_esize or k.element_size())
        dom.ensure(_up(kv_bytes), _up(2 * self.max_part_rows * part_row_bytes))
        dom.pass_id += 1
        p, slot = dom.pass_id, dom.pass_id & 1
        cur = torch.cuda.current_stream(dom.device)
        fs = dom.fetch_stream
        fs.wait_stream(cur)
        # slot p&1 (published K,V and partial dK,dV) is free once every peer has finished its pulls of pass p-2
        self._wait("pd", lambda: [dom.flag_addr(self.rank, _W_PULLS_DONE + s) for s in self.peers], p - 2, fs)
        slot_free = fs.record_event()
        koff = dom.kv_off(slot)
        voff = koff + self.shard_rows * row_bytes
        _lib.NvlCopyList([(dom.base + koff, k.data_ptr(), self.shard_rows * row_bytes),
                          (dom.base + voff, v.data_ptr(), self.shard_rows * row_bytes)]).launch(fs.cuda_stream)
        self._signal("kr", lambda: [dom.flag_addr(s, _W_KV_READY + self.rank) for s in self.peers], p, fs)
        works: list[WorkWithPostProcessFn | None] = []
        for st in range(self.degree):
            n = self.rows[st]
            if n == 0:
                works.append(None)
                continue
            rk, rv = k.new_empty((n, *k.shape[1:])), v.new_empty((n, *v.shape[1:]))
            srcs = self.kv_sources[st]
            self._wait(("kr", st), lambda: [dom.flag_addr(self.rank, _W_KV_READY + s) for s in srcs], p, fs)
            spans = []
            for (s, s0, d0, nr) in self.pull[st]:
                pb = dom.peer_base[s]
                spans.append((rk.data_ptr() + d0 * row_bytes, pb + koff + s0 * row_bytes, nr * row_bytes))
                spans.append((rv.data_ptr() + d0 * row_bytes, pb + voff + s0 * row_bytes, nr * row_bytes))
            dom.bytes_pulled += dom.copy_spans(spans, fs)
            ev = fs.record_event()
            works.append(WorkWithPostProcessFn(GeneralWork(ev), (lambda b: (lambda *a, **kw: b))((rk, rv)), async_op=True))
        self._cur = dict(p=p, slot=slot, slot_free=slot_free, part_row_bytes=part_row_bytes, waited=False,
                         hk=k.shape[1], d=k.shape[2], kv_done=fs.record_event())
        if not backward:
            # a forward pass pulls nothing else: this rank's pulls of pass p are complete once fs gets here
            self._signal("pd_out", lambda: [dom.flag_addr(s, _W_PULLS_DONE + self.rank) for s in self.peers], p, fs)
        return works

    # ---- backward only --------------------------------------------------------------------------------------------
    def _part_stage_off(self, rank: int, st: int, part_row_bytes: int) -> int:
        return 2 * part_row_bytes * sum(self.rows_all[rank][:st])

    def partial_buffers(self, st: int, dtype: torch.dtype) -> tuple[torch.Tensor, torch.Tensor]:
        """(dk_part, dv_part) of stage ``st`` for the current pass: views of this rank's arena, so the backward kernel
        writes the partial gradients where their owners pull them from.  Orders the current stream behind 'slot free'."""
        c, dom = self._cur, self.dom
        if not c["waited"]:
            c["waited"] = True
            torch.cuda.current_stream(dom.device).wait_event(c["slot_free"])
        n, hk, d = self.rows[st], c["hk"], c["d"]
        esz = torch.empty((), dtype=dtype).element_size()
        assert esz * hk * d == c["part_row_bytes"]
        off = dom.part_off(c["slot"]) + self._part_stage_off(self.rank, st, c["part_row_bytes"])
        nb = n * c["part_row_bytes"]
        dk_p = dom.view[off:off + nb].view(dtype).view(n, hk, d)
        dv_p = dom.view[off + nb:off + 2 * nb].view(dtype).view(n, hk, d)
        return dk_p, dv_p

    def partial_ready(self, st: int) -> None:
        """Call on the compute stream right after the stage-``st`` backward kernel: tell the owners their rows are ready."""
        c, dom = self._cur, self.dom
        owners = self.owners[st]
        self._signal(("pr_out", st), lambda: [dom.flag_addr(o, _W_PART_READY + self.rank * _MAX_STAGES + st) for o in owners],
                     c["p"], torch.cuda.current_stream(dom.device))

    def mark_local_done(self) -> None:
        """Call right after the local (host-stage) backward has been enqueued: the folds must not start before it."""
        self._cur["local_done"] = torch.cuda.current_stream(self.dom.device).record_event()

    def begin_reduce(self, dk: torch.Tensor, dv: torch.Tensor, wire_dtype: torch.dtype) -> None:
        """Owner side of the dKV return path: for every stage wait for the consumers' partials, pull them (copy engines)
        and fold them into dk / dv, all on the reduce stream, ordered only behind the local backward.

        HOST ORDER MATTERS: call this AFTER every ``partial_ready`` of the pass has been enqueued.  A wait enqueued before
        this rank's own signals can deadlock two ranks against each other through implicit host/device synchronisation
        (a cudaMalloc in the caching allocator orders later launches behind the pending wait kernel: each rank then holds
        back the signal the other one waits for).  With every signal already in the streams, a pending wait only ever
        depends on work the peers enqueue on their own."""
        from .. import _lib

        c, dom = self._cur, self.dom
        rs = dom.reduce_stream
        rs.wait_event(c["local_done"])
        prb = c["part_row_bytes"]
        hk, d = c["hk"], c["d"]
        keep = []
        for st in range(self.degree):
            plan = self.rpull[st]
            if not plan:
                continue
            n = sum(x[3] for x in plan)
            with torch.cuda.stream(rs):  # staging lives in the reduce stream's allocator pool: written and read only there
                sk = torch.empty((n, hk, d), dtype=wire_dtype, device=dom.device)
                sv = torch.empty((n, hk, d), dtype=wire_dtype, device=dom.device)
            keep += [sk, sv]
            cons = self.part_sources[st]
            self._wait(("pr_in", st), lambda: [dom.flag_addr(self.rank, _W_PART_READY + s * _MAX_STAGES + st) for s in cons],
                       c["p"], rs)
            spans, dst_spans, src_spans, pos = [], [], [], 0
            for (cr, s0, d0, nr) in plan:
                src = dom.peer_base[cr] + dom.part_off(c["slot"]) + self._part_stage_off(cr, st, prb)
                nbytes_stage = self.rows_all[cr][st] * prb
                spans.append((sk.data_ptr() + pos * prb, src + s0 * prb, nr * prb))
                spans.append((sv.data_ptr() + pos * prb, src + nbytes_stage + s0 * prb, nr * prb))
                dst_spans.append((d0, d0 + nr))
                src_spans.append((pos, pos + nr))
                pos += nr
            dom.bytes_pulled += dom.copy_spans(spans, rs)
            meta = self._reduce_meta.get(st)
            if meta is None:
                meta = (torch.tensor(dst_spans, dtype=torch.int32, device=dom.device),
                        torch.tensor(src_spans, dtype=torch.int32, device=dom.device))
                self._reduce_meta[st] = meta
            with torch.cuda.stream(rs):
                _lib.range_reduce_sum(dk, sk, meta[0], meta[1])
                _lib.range_reduce_sum(dv, sv, meta[0], meta[1])
        rs.wait_event(c["kv_done"])
        self._signal("pd_out", lambda: [dom.flag_addr(s, _W_PULLS_DONE + self.rank) for s in self.peers], c["p"], rs)
        c["keep"] = keep

    def finish_reduce(self) -> None:
        torch.cuda.current_stream(self.dom.device).wait_stream(self.dom.reduce_stream)
        if self._cur is not None:
            self._cur.pop("keep", None)
