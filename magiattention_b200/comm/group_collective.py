"""group_cast / group_reduce — the zero-redundant KV / dKV exchange of context parallelism.

Same contract as the reference front door (magi_attention/comm/primitive/grpcoll/
_group_collective.py:81-205, 255-408), different engine: the reference permutes rows into a send
buffer (Triton range_gather), runs one ``all_to_all_single`` and un-permutes / range-reduces on the
receive side (grpcoll/_a2av_grpcoll_impl.py, utils.py:361-659), duplicating every multi-destination
split once per destination.  Here every split is a CONTIGUOUS row span, so on one NVSwitch domain each
(split, destination) pair is a single NCCL P2P message sent straight out of the input rows and
received straight into its final rows of the output: one grouped ncclSend/ncclRecv batch per call on
NCCL's own (high-priority) stream, no pre-permute pass, no post-permute pass, no duplication in HBM.
group_reduce receives into a staging buffer and folds it into the output with ONE fused row-span
add kernel (b200_range_reduce_sum, fp32 accumulate).

group_cast:  rank r's ``input`` is cut by ``input_split_sizes``; split i goes to every rank of
``dst_indices[i]``.  ``output`` is the concatenation of the received splits: output split j has
``output_split_sizes[j]`` rows and comes from rank ``src_index[j]`` (splits of one source keep the
source's order).  group_reduce is the exact transpose with sum accumulation into ``output``.
"""
from __future__ import annotations

import os
from typing import Sequence

import torch
import torch.distributed as dist

from .work import GeneralWork, WorkWithPostProcessFn

__all__ = ["group_cast", "group_reduce"]

_TensorOrTuple = torch.Tensor | tuple[torch.Tensor, ...] | list[torch.Tensor]


def _as_tuple(x: _TensorOrTuple | None) -> tuple[torch.Tensor, ...]:
    if x is None:
        return ()
    return tuple(x) if isinstance(x, (tuple, list)) else (x,)


def _global_rank(group, r: int) -> int:
    return r if group is None or group is dist.group.WORLD else dist.get_global_rank(group, r)


def _launch(ops: list, async_op: bool):
    if not ops:
        return None
    works = dist.batch_isend_irecv(ops)
    return works


def group_cast(
    input: _TensorOrTuple,
    output: _TensorOrTuple | None,
    input_split_sizes: Sequence[int],
    output_split_sizes: Sequence[int],
    dst_indices: Sequence[Sequence[int]],
    src_index: Sequence[int],
    group: dist.ProcessGroup | None,
    async_op: bool = False,
    cast_lse: bool = False,
    input_lse: torch.Tensor | None = None,
    output_lse: torch.Tensor | None = None,
    **kwargs,
) -> WorkWithPostProcessFn:
    """See module docstring.  ``input`` may be one tensor or a tuple (e.g. ``(k, v)``) whose members are
    row-split identically; ``output`` mirrors it (allocated when ``None``).  Returns a work whose
    ``wait_post_process(output)`` yields the filled output after ordering the current stream behind the
    transfer."""
    ins = _as_tuple(input)
    lse_out = None
    if cast_lse:
        # the lse rows [rows, heads] travel with the same splits as one more member of the tuple
        # (reference: cast_lse, grpcoll/_group_collective.py:81-205)
        assert input_lse is not None, "cast_lse=True needs input_lse"
        assert not isinstance(input, (tuple, list)), "cast_lse takes a single input tensor"
        ins = (ins[0], input_lse)
        if output is not None:
            assert output_lse is not None, "cast_lse=True with a given output needs output_lse"
            output = (output, output_lse)
    assert len(input_split_sizes) == len(dst_indices)
    assert len(output_split_sizes) == len(src_index)
    n_out = int(sum(output_split_sizes))
    if output is None:
        outs = tuple(t.new_empty((n_out, *t.shape[1:])) for t in ins)
        output = outs if (isinstance(input, (tuple, list)) or cast_lse) else outs[0]
    else:
        outs = _as_tuple(output)
    for t in ins:
        assert t.shape[0] >= sum(input_split_sizes), "input has fewer rows than the split sizes cover"
    for t in outs:
        assert t.shape[0] >= n_out, "output has fewer rows than the split sizes cover"

    rank = dist.get_rank(group)
    ops: list[dist.P2POp] = []
    self_queue: list[tuple[int, int]] = []
    off = 0
    for sz, dsts in zip(input_split_sizes, dst_indices):
        if sz > 0:
            for d in dsts:
                if d == rank:
                    self_queue.append((off, sz))
                else:
                    peer = _global_rank(group, d)
                    for t in ins:
                        ops.append(dist.P2POp(dist.isend, t[off : off + sz], peer, group))
        off += sz
    off = 0
    self_pos = 0
    for sz, s in zip(output_split_sizes, src_index):
        if sz > 0:
            if s == rank:
                so, ssz = self_queue[self_pos]
                self_pos += 1
                assert ssz == sz
                for ti, to in zip(ins, outs):
                    to[off : off + sz].copy_(ti[so : so + ssz])
            else:
                peer = _global_rank(group, s)
                for t in outs:
                    ops.append(dist.P2POp(dist.irecv, t[off : off + sz], peer, group))
        off += sz
    works = _launch(ops, async_op)
    out_ref = output
    return WorkWithPostProcessFn(GeneralWork(works), lambda *a, **k: (a[0] if a else out_ref), async_op=async_op)


# ---------------------------------------------------------------------------------------------
_reduce_meta_cache: dict = {}


def _range_add(dst: torch.Tensor, src: torch.Tensor, dst_spans: list[tuple[int, int]], src_spans: list[tuple[int, int]],
               cache_key) -> None:
    """dst[dst_span_i] += src[src_span_i]; CUDA tensors go through the library kernel."""
    if not dst_spans:
        return
    if dst.is_cuda:
        from .. import _lib

        key = (cache_key, dst.device)
        meta = _reduce_meta_cache.get(key)
        if meta is None:
            meta = (torch.tensor(dst_spans, dtype=torch.int32, device=dst.device),
                    torch.tensor(src_spans, dtype=torch.int32, device=dst.device))
            _reduce_meta_cache[key] = meta
        _lib.range_reduce_sum(dst, src, meta[0], meta[1])
    else:
        # host tensors only occur in the gloo unit tests of the exchange logic
        if os.environ.get("MAGI_B200_ALLOW_HOST_COMM_TEST", "0") != "1":
            raise RuntimeError("group_reduce on host tensors is test-only (set MAGI_B200_ALLOW_HOST_COMM_TEST=1)")
        for (d0, d1), (s0, s1) in zip(dst_spans, src_spans):
            dst[d0:d1] += src[s0:s1].to(dst.dtype)


def _lse_merge_rows(out: torch.Tensor, lse: torch.Tensor, out2: torch.Tensor, lse2: torch.Tensor) -> None:
    """(out, lse) <- LSE-merge with (out2, lse2), in place on contiguous row spans (correct_attn_out_lse semantics)."""
    if out.shape[0] == 0:
        return
    if out.is_cuda:
        from .. import _lib

        _lib.lse_merge(out, lse, out2.contiguous(), lse2.contiguous())
        return
    if os.environ.get("MAGI_B200_ALLOW_HOST_COMM_TEST", "0") != "1":
        raise RuntimeError("group_reduce on host tensors is test-only (set MAGI_B200_ALLOW_HOST_COMM_TEST=1)")
    l = torch.logaddexp(lse, lse2)
    w1 = torch.exp(lse - l).nan_to_num(0.0).unsqueeze(-1)
    w2 = torch.exp(lse2 - l).nan_to_num(0.0).unsqueeze(-1)
    out.copy_((w1 * out.float() + w2 * out2.float()).to(out.dtype))
    lse.copy_(l)


def group_reduce(
    input: _TensorOrTuple,
    output: _TensorOrTuple,
    input_split_sizes: Sequence[int],
    output_split_sizes: Sequence[int],
    dst_index: Sequence[int],
    src_indices: Sequence[Sequence[int]],
    group: dist.ProcessGroup | None,
    async_op: bool = False,
    reduce_op: str = "sum",
    acc_reduce: bool = True,
    comm_dtype: torch.dtype | None = None,
    input_lse: torch.Tensor | None = None,
    output_lse: torch.Tensor | None = None,
    **kwargs,
) -> WorkWithPostProcessFn:
    """Transpose of group_cast with accumulation: input split i goes to rank ``dst_index[i]``; output split
    j (+)= the splits received from every rank of ``src_indices[j]``.  ``acc_reduce=False`` zeroes the output first.
    ``reduce_op``: "sum" (the dK/dV hot path), "avg" (mean; an accumulated old value counts as one contribution) or "lse"
    (log-sum-exp weighted merge of partial attention outputs with ``input_lse`` / ``output_lse`` fp32 [rows, heads] —
    the reduction of the reference's q/o communication, common/range_op/_range_reduce.py:239-357)."""
    if reduce_op not in ("sum", "avg", "lse"):
        raise ValueError(f"Invalid reduce_op {reduce_op!r}: expected 'sum', 'avg' or 'lse'")
    ins = _as_tuple(input)
    outs = _as_tuple(output)
    assert len(ins) == len(outs)
    if reduce_op == "lse":
        # log-sum-exp weighted merge of partial attention outputs (q/o communication): the lse rows travel as a second tensor
        assert len(ins) == 1 and input_lse is not None and output_lse is not None, "reduce_op='lse' needs input_lse / output_lse"
        assert input_lse.dtype == torch.float32 and output_lse.dtype == torch.float32, "lse tensors must be fp32"
    assert len(input_split_sizes) == len(dst_index)
    assert len(output_split_sizes) == len(src_indices)
    if comm_dtype is not None:
        ins = tuple(t.to(comm_dtype) for t in ins)
    send = ins + ((input_lse,) if reduce_op == "lse" else ())
    rank = dist.get_rank(group)
    if not acc_reduce:
        for t in outs:
            t.zero_()
        if reduce_op == "lse":
            output_lse.fill_(float("-inf"))

    n_recv = int(sum(sz * len(srcs) for sz, srcs in zip(output_split_sizes, src_indices)))
    stages = tuple(t.new_empty((n_recv, *t.shape[1:])) for t in send)
    ops: list[dist.P2POp] = []
    self_queue: list[tuple[int, int]] = []
    off = 0
    for sz, d in zip(input_split_sizes, dst_index):
        if sz > 0:
            if d == rank:
                self_queue.append((off, sz))
            else:
                peer = _global_rank(group, d)
                for t in send:
                    ops.append(dist.P2POp(dist.isend, t[off : off + sz], peer, group))
        off += sz
    dst_spans: list[tuple[int, int]] = []
    src_spans: list[tuple[int, int]] = []
    self_adds: list[tuple[int, int, int]] = []
    counts: list[tuple[int, int, int]] = []  # (out offset, rows, number of contributions) per output split
    off = 0
    spos = 0
    self_pos = 0
    for sz, srcs in zip(output_split_sizes, src_indices):
        if sz > 0:
            for s in srcs:
                if s == rank:
                    so, ssz = self_queue[self_pos]
                    self_pos += 1
                    assert ssz == sz
                    self_adds.append((off, so, sz))
                else:
                    peer = _global_rank(group, s)
                    for st in stages:
                        ops.append(dist.P2POp(dist.irecv, st[spos : spos + sz], peer, group))
                    dst_spans.append((off, off + sz))
                    src_spans.append((spos, spos + sz))
                    spos += sz
            counts.append((off, sz, len(srcs)))
        off += sz
    works = _launch(ops, async_op)
    cache_key = (tuple(output_split_sizes), tuple(tuple(s) for s in src_indices), rank)

    def post(*a, **k):
        if reduce_op == "lse":
            t_out, t_in, st, st_lse = outs[0], ins[0], stages[0], stages[1]
            for (o, so, sz) in self_adds:
                _lse_merge_rows(t_out[o : o + sz], output_lse[o : o + sz], t_in[so : so + sz], input_lse[so : so + sz])
            for (d0, d1), (s0, s1) in zip(dst_spans, src_spans):
                _lse_merge_rows(t_out[d0:d1], output_lse[d0:d1], st[s0:s1], st_lse[s0:s1])
            return a[0] if a else output
        for t_out, t_in, st in zip(outs, ins, stages):
            for (o, so, sz) in self_adds:
                t_out[o : o + sz] += t_in[so : so + sz].to(t_out.dtype)
            _range_add(t_out, st, dst_spans, src_spans, cache_key)
            if reduce_op == "avg":
                # reference range_avg_reduce_kernel (common/range_op/_range_reduce.py:152-231): the accumulated old value
                # counts as one contribution; without accumulation the mean is over the received splits only
                for (o, sz, n) in counts:
                    div = n + 1 if acc_reduce else n
                    if div > 1:
                        t_out[o : o + sz].div_(div)
        return a[0] if a else output

    return WorkWithPostProcessFn(GeneralWork(works), post, async_op=async_op)
