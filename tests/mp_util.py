"""Multi-process test helper: spawn N ranks, never hang the suite."""
import os
import signal
import socket
import time

import torch.multiprocessing as mp


def free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spawn_once(worker, world, args_after_port, deadline_s):
    ctx = mp.start_processes(worker, args=(world, free_port(), *args_after_port), nprocs=world, join=False, start_method="spawn")
    t_end = time.time() + deadline_s
    try:
        while not ctx.join(timeout=5):
            if time.time() > t_end:
                raise TimeoutError(f"workers still running after {deadline_s}s")
    except TimeoutError:
        for pid in ctx.pids():  # exactly the children started above
            try:
                os.kill(pid, signal.SIGKILL)
            except ProcessLookupError:
                pass
        raise


def spawn_with_deadline(worker, world, args_after_port=(), deadline_s=240):
    """Run ``worker(rank, world, port, *args_after_port)`` in ``world`` spawned processes.  A run that exceeds the deadline
    (stuck rendezvous: the probed port was taken by someone else, ...) has its children killed and is retried ONCE on a fresh
    port; a real deadlock therefore still fails, a worker exception propagates immediately (ProcessRaisedException)."""
    try:
        _spawn_once(worker, world, args_after_port, deadline_s)
    except TimeoutError:
        _spawn_once(worker, world, args_after_port, deadline_s)
