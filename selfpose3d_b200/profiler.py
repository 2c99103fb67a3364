"""Per-launch CUDA-event timing of the C-ABI kernels (used by bench.py inside its timed region).

When enabled, every ``_lib.call`` is bracketed by two events on the launching stream and tagged
with a kernel family and its algorithmic work (FLOPs for the convolutions, bytes for the
HBM-bound kernels), so the roofline fractions come from device time, not from a profiler run.
"""
from __future__ import annotations

import torch

_enabled = False
_records = []   # (kind, start_event, end_event, work)


def enable():
    global _enabled
    _records.clear()
    _enabled = True


def disable():
    global _enabled
    _enabled = False


def active():
    return _enabled


def begin():
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def end(kind, start, work, detail=None):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    _records.append((kind, start, ev, float(work), detail))


def summary():
    """{kind: {"ms": total device ms, "work": total algorithmic work, "launches": n}} (call after a sync)."""
    out = {}
    for kind, e0, e1, work, _ in _records:
        d = out.setdefault(kind, {"ms": 0.0, "work": 0.0, "launches": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["work"] += work
        d["launches"] += 1
    return out


def detail_summary():
    """Like ``summary`` but keyed by the per-launch detail string (layer shape); launches without one are skipped."""
    out = {}
    for _, e0, e1, work, detail in _records:
        if detail is None:
            continue
        d = out.setdefault(detail, {"ms": 0.0, "work": 0.0, "launches": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["work"] += work
        d["launches"] += 1
    return out
