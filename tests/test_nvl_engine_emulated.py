"""CPU (gloo, world 2 and 4) tests of the NVLink peer-memory engine's HOST logic — the real comm/nvl_exchange.py and the real
DistAttnFunc stage loop — over the emulated primitives of tests/nvl_emulation.py (shared file mappings for the arenas,
memmove for the copy engines, polled epoch words for the flags).  The attention kernels are the oracle fakes of
test_dist_runtime_gloo.py.  What this pins without a GPU: arena layout and slot rotation over many passes, pull tables against
real data, the three flag families, and the host order of signals and waits (a wrong order deadlocks here, see the last test).
The GPU counterpart is tests/test_dist_attn_gpu.py::test_real_cp_nccl[*-nvl] / test_real_cp_nvl_multi_pass."""
import pytest
import torch
import torch.distributed as dist

from test_comm_gloo import _run


def _emu_body(rank, world, shared_dir, cases, skew_s):
    import nvl_emulation as emu
    from test_dist_runtime_gloo import _runtime_body

    dom = emu.install(shared_dir, dist.group.WORLD, skew_s=skew_s * (1 + rank % 2))
    passes = 0
    for name, degree, hp_reduce in cases:
        gen = dom.generation
        _runtime_body(rank, world, name, degree, None, hp_reduce)
        passes = passes + 2 if dom.generation == gen else 2  # a re-created arena restarts the pass counter
        assert dom.pass_id == passes, (dom.pass_id, passes)
    log = emu.op_log()
    kinds = {op[0] for op in log}
    assert kinds == {"publish", "copy", "signal", "wait"} and dom.bytes_pulled > 0, "the emulated NVLink engine did not carry the exchange"
    assert int(dom.err[0]) == 0
    # slot rotation: every publish of (K, V) for pass p into slot p & 1 comes right after the wait for pulls_done >= p - 2 of
    # every peer and right before kv_ready = p goes out to every peer (the in-order emulation cannot provoke the overwrite a
    # missing wait would allow, so the order itself is pinned)
    pubs = [i for i, op in enumerate(log) if op[0] == "publish"]
    assert len(pubs) == 2 * len(cases)
    for i in pubs:
        p = log[i + 1][2]
        assert log[i][1] == 2 and log[i + 1] == ("signal", world - 1, p) and log[i - 1] == ("wait", world - 1, p - 2), log[i - 1:i + 2]


@pytest.mark.parametrize("world,cases", [
    (2, [("causal", 1, False), ("sliding_window_causal", 2, False), ("causal", 2, True), ("sparse_attn_2k", 1, False)]),
    (4, [("causal", 2, False), ("sparse_attn_2k", 3, False), ("sliding_window_causal", 2, True), ("causal", 3, False)]),
])
def test_nvl_engine_runtime_emulated(world, cases, tmp_path):
    """Several keys back to back on one arena domain: 8 passes, both slots reused, 16-bit-style and fp32 wire."""
    _run(world, _emu_body, str(tmp_path), cases, 0.0)


def test_nvl_engine_runtime_emulated_with_rank_skew(tmp_path):
    """Skewed ranks (odd ranks' copy engines are slower, with jitter): different cross-rank interleavings, same exact results."""
    _run(4, _emu_body, str(tmp_path), [("causal", 2, False), ("varlen_block_causal_2k", 2, False), ("causal", 2, False)], 0.004)


def _emu_sink_body(rank, world, shared_dir):
    import nvl_emulation as emu
    from test_dist_runtime_gloo import _sink_body

    dom = emu.install(shared_dir, dist.group.WORLD)
    _sink_body(rank, world, "causal", 2)
    _sink_body(rank, world, "sliding_window_causal", 1)
    assert dom.bytes_pulled > 0 and int(dom.err[0]) == 0


def test_nvl_engine_sink_and_max_logits_emulated(tmp_path):
    """Attention sink (applied once, on the host stage; partial dsink summed over the group) and max_logits on the NVLink engine."""
    _run(2, _emu_sink_body, str(tmp_path))


def _wrong_order_body(rank, world, shared_dir):
    import nvl_emulation as emu
    from magiattention_b200.comm import nvl_exchange as nx
    from magiattention_b200.meta.planner import make_attn_meta, make_dispatch_meta

    emu.install(shared_dir, dist.group.WORLD, timeout_s=1.5)
    T, chunk = 64, 16
    sl = [(0, T, 0, T, 1)]
    dm = make_dispatch_meta(sl, T, chunk, world)
    comm, _ = make_attn_meta(sl, dm, 1)
    ex = nx.NvlKVExchange(dist.group.WORLD, comm, 1, dm.shard_seqlen)
    assert ex.rpull[0] and ex.pull[0], "causal cp=2: every rank both consumes and owns remote rows"
    k, v = (torch.randn(dm.shard_seqlen, 1, 8, dtype=torch.float64) for _ in range(2))
    works = ex.fetch(k, v, backward=True)
    works[0].wait_post_process()
    ex.mark_local_done()
    dk, dv = (torch.zeros(k.shape, dtype=torch.float32) for _ in range(2))
    # owner-side waits BEFORE this rank's own partial-ready signals: with an in-order device both ranks now wait for each other
    with pytest.raises(TimeoutError, match="never reached"):
        ex.begin_reduce(dk, dv, k.dtype)


def test_host_order_rule_is_what_prevents_the_deadlock(tmp_path):
    """NvlKVExchange.begin_reduce documents: enqueue it after every partial_ready of the pass.  Doing it first is the bug the
    first GPU run of the engine hit (flag-wait timeouts on both ranks); the emulation reproduces it deterministically."""
    _run(2, _wrong_order_body, str(tmp_path))
