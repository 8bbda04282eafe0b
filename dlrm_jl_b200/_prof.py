"""Optional per-launch CUDA-event timing of the library's kernels (used by bench.py).

Disabled by default (zero overhead beyond one attribute test).  When enabled, every call into
libdlrm_b200.so made through the host mirror is bracketed by a pair of events recorded on the
stream the kernels are launched on, so durations are device time of exactly those kernels.

Two modes:
  * eager: a fresh event pair per call; `summary()` after a device synchronize.
  * graph (`enable(True, external=True)` while a CUDA graph is being captured): the pairs become
    event-record NODES of the graph (cudaEventRecordExternal), so every replay re-stamps them with
    the kernels running back to back exactly as in the replayed step; call `read_replay()` after
    each replay + synchronize.
"""
from __future__ import annotations

from collections import defaultdict
from contextlib import contextmanager
from typing import Dict, List, Tuple

import builtins

import torch

_builtin_range = builtins.range   # this module defines its own `range` (the profiling context manager)
_enabled = False
_external = False
_events: Dict[str, List[Tuple[torch.cuda.Event, torch.cuda.Event]]] = defaultdict(list)


def enable(flag: bool = True, external: bool = False) -> None:
    global _enabled, _external
    _enabled = flag
    if flag:
        _external = external
        _events.clear()


@contextmanager
def range(name: str):  # noqa: A001 - mirrors nvtx.range
    if not _enabled:
        yield
        return
    a = torch.cuda.Event(enable_timing=True, external=_external)
    b = torch.cuda.Event(enable_timing=True, external=_external)
    a.record()
    try:
        yield
    finally:
        b.record()
        _events[name].append((a, b))


def summary() -> Dict[str, Dict[str, float]]:
    """name -> {count, total_ms, avg_ms}; call after a device synchronize."""
    out = {}
    for name, pairs in _events.items():
        ms = [a.elapsed_time(b) for a, b in pairs]
        out[name] = {"count": len(ms), "total_ms": float(sum(ms)), "avg_ms": float(sum(ms) / max(1, len(ms)))}
    return out


def read_replay() -> Dict[str, List[float]]:
    """Graph mode: name -> milliseconds of every captured call, as stamped by the last replay."""
    return {name: [a.elapsed_time(b) for a, b in pairs] for name, pairs in _events.items()}


def read_replay_timeline() -> Dict[str, List[Tuple[float, float]]]:
    """Graph mode: name -> [(start, end)] in microseconds from the earliest start stamp of the last replay: where
    every bracketed call sits in the step (event-record nodes add a few microseconds of latency each, so this
    is a map of the step, not a stopwatch)."""
    firsts = [pairs[0][0] for pairs in _events.values() if pairs]
    if not firsts:
        return {}
    origin = firsts[0]
    for ev in firsts[1:]:
        if ev.elapsed_time(origin) > 0:      # ev is earlier than the current origin
            origin = ev
    return {name: [(1e3 * origin.elapsed_time(a), 1e3 * origin.elapsed_time(b)) for a, b in pairs]
            for name, pairs in _events.items()}


def time_launches(fn, nb: int, iters: int = 10, use_graph: bool = True) -> float:
    """Microseconds per launch of `fn(i)`, i = 0..nb-1 (each i a different input batch): the nb launches
    are captured into one CUDA graph and the replays are timed with CUDA events, so the figure is
    device time per launch with the launches back to back on one stream (no host gaps)."""
    torch.cuda.synchronize()
    if use_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn(0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in _builtin_range(nb):
                fn(i)
        for _ in _builtin_range(3):
            g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in _builtin_range(iters):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        return 1e3 * a.elapsed_time(b) / (iters * nb)
    for i in _builtin_range(min(nb, 2)):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in _builtin_range(iters):
        for i in _builtin_range(nb):
            fn(i)
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / (iters * nb)
