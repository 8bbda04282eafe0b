"""Optional per-launch CUDA-event timing of the library's kernels (used by bench.py).

Disabled by default (zero overhead beyond one attribute test).  When enabled, every call into
libdlrm_b200.so made through the host mirror is bracketed by a pair of events recorded on the
stream the kernels are launched on, so durations are device time of exactly those kernels.
"""
from __future__ import annotations

from collections import defaultdict
from contextlib import contextmanager
from typing import Dict, List, Tuple

import torch

_enabled = False
_events: Dict[str, List[Tuple[torch.cuda.Event, torch.cuda.Event]]] = defaultdict(list)


def enable(flag: bool = True) -> None:
    global _enabled
    _enabled = flag
    if flag:
        _events.clear()


@contextmanager
def range(name: str):  # noqa: A001 - mirrors nvtx.range
    if not _enabled:
        yield
        return
    a = torch.cuda.Event(enable_timing=True)
    b = torch.cuda.Event(enable_timing=True)
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
