"""Sketches as instrumentation sinks: the host-side mirrors of the reference's
``HyperLogLog`` / ``CountMinSketch`` / ``SketchCollector`` whose ``add()`` runs on the device.

Reference: happysimulator/sketching/hyperloglog.py:57-250, sketching/count_min_sketch.py:52-328,
components/sketching/sketch_collector.py:24-104.  The device keeps one sketch state per replica (uint8
registers / uint32 counters in HBM, csrc/hs_sketch.h), updated by the SKETCH row's handler with the
per-key hash results the host computed once (lowering.hll_table / cms_table); after a run the states are
written back into these objects, and ``merge()`` -- register max / counter sum, the reference's contracts
-- is what ``Engine.read_sketches`` does over the replicas on the device and what
``distributed.allreduce_sketches`` does across GPUs.
"""
from __future__ import annotations

import math

import numpy as np

from . import lowering


class KeyExtractor:
    """value_extractor for SketchCollector: "the request's routing key" (the reference idiom
    ``lambda e: e.context.get("metadata", {}).get("client_id")``), the one per-request value that exists on
    the device.  Calling it on a host-side event works too."""

    routing_key = True

    def __init__(self, field: str = "client_id"):
        self.field = field

    def __call__(self, event):
        return event.context.get("metadata", {}).get(self.field)


class HyperLogLog:
    """sketching/hyperloglog.py:57: cardinality estimate from 2^precision max-run-length registers."""

    _ALPHA = {4: 0.673, 5: 0.697, 6: 0.709}

    def __init__(self, precision: int = 14, seed: int | None = None):
        if not 4 <= precision <= 16:
            raise ValueError(f"precision must be in [4, 16], got {precision}")
        self._precision = precision
        self._num_registers = 1 << precision
        self._registers = np.zeros(self._num_registers, np.uint8)
        self._seed = seed if seed is not None else 0
        self._total_count = 0
        self._tab = None

    precision = property(lambda self: self._precision)
    num_registers = property(lambda self: self._num_registers)
    item_count = property(lambda self: self._total_count)

    def _key_table(self, upto: int):
        if self._tab is None or self._tab.shape[1] < upto:
            self._tab = lowering.hll_table(self._precision, self._seed, max(upto, 64))
        return self._tab

    def add(self, item: int, count: int = 1) -> None:
        """Host-side add for a non-negative int item (the same table the device indexes)."""
        if count < 0:
            raise ValueError(f"count must be non-negative, got {count}")
        if count == 0:
            return
        t = self._key_table(int(item) + 1)
        self._total_count += count
        i, run = int(t[0, item]), int(t[1, item])
        if self._registers[i] < run:
            self._registers[i] = run

    def cardinality(self) -> int:
        """hyperloglog.py:167-192: alpha m^2 / sum 2^-r with the small/large range corrections."""
        m = self._num_registers
        alpha = self._ALPHA.get(self._precision, 0.7213 / (1 + 1.079 / m))
        est = alpha * m * m / sum(2.0 ** (-int(r)) for r in self._registers)
        if est <= 2.5 * m:
            zeros = int((self._registers == 0).sum())
            if zeros > 0:
                est = m * math.log(m / zeros)
        elif est > (1 << 32) / 30:
            est = -(1 << 32) * math.log(1 - est / (1 << 32))
        return int(est)

    def standard_error(self) -> float:
        return 1.04 / math.sqrt(self._num_registers)

    def merge(self, other: "HyperLogLog") -> None:
        if not isinstance(other, HyperLogLog):
            raise TypeError(f"Can only merge with HyperLogLog, got {type(other).__name__}")
        if other._precision != self._precision:
            raise ValueError(f"Cannot merge: precision differs ({self._precision} vs {other._precision})")
        np.maximum(self._registers, other._registers, out=self._registers)
        self._total_count += other._total_count

    def clear(self) -> None:
        self._registers[:] = 0
        self._total_count = 0

    def _load_device_state(self, registers: np.ndarray, item_count: int) -> None:
        self._registers = np.array(registers, dtype=np.uint8)
        self._total_count = int(item_count)


class CountMinSketch:
    """sketching/count_min_sketch.py:52: depth x width counters, estimate = min over the rows."""

    def __init__(self, width: int, depth: int, seed: int | None = None):
        if width <= 0:
            raise ValueError(f"width must be positive, got {width}")
        if depth <= 0:
            raise ValueError(f"depth must be positive, got {depth}")
        self._width, self._depth = width, depth
        self._seed = seed if seed is not None else 0
        self._counters = np.zeros((depth, width), np.uint64)
        self._total_count = 0
        self._tab = None

    @classmethod
    def from_error_rate(cls, epsilon: float, delta: float, seed: int | None = None) -> "CountMinSketch":
        """count_min_sketch.py:107-134: width = ceil(e / epsilon), depth = ceil(ln(1 / delta))."""
        if not 0 < epsilon < 1:
            raise ValueError(f"epsilon must be in (0, 1), got {epsilon}")
        if not 0 < delta < 1:
            raise ValueError(f"delta must be in (0, 1), got {delta}")
        return cls(width=math.ceil(math.e / epsilon), depth=math.ceil(math.log(1.0 / delta)), seed=seed)

    width = property(lambda self: self._width)
    depth = property(lambda self: self._depth)
    epsilon = property(lambda self: math.e / self._width)
    delta = property(lambda self: math.exp(-self._depth))
    item_count = property(lambda self: self._total_count)

    def _key_table(self, upto: int):
        if self._tab is None or self._tab.shape[1] < upto:
            self._tab = lowering.cms_table(self._width, self._depth, self._seed, max(upto, 64))
        return self._tab

    def add(self, item: int, count: int = 1) -> None:
        if count < 0:
            raise ValueError(f"count must be non-negative, got {count}")
        if count == 0:
            return
        t = self._key_table(int(item) + 1)
        self._total_count += count
        for row in range(self._depth):
            self._counters[row, t[row, item]] += count

    def estimate(self, item: int) -> int:
        t = self._key_table(int(item) + 1)
        return int(min(int(self._counters[row, t[row, item]]) for row in range(self._depth)))

    def merge(self, other: "CountMinSketch") -> None:
        if not isinstance(other, CountMinSketch):
            raise TypeError(f"Can only merge with CountMinSketch, got {type(other).__name__}")
        if self._width != other._width or self._depth != other._depth:
            raise ValueError(f"Cannot merge: dimensions differ ({self._width}x{self._depth} vs "
                             f"{other._width}x{other._depth})")
        if self._seed != other._seed:
            raise ValueError(f"Cannot merge: seeds differ ({self._seed} vs {other._seed})")
        self._counters += other._counters
        self._total_count += other._total_count

    def clear(self) -> None:
        self._counters[:] = 0
        self._total_count = 0

    def _load_device_state(self, counters: np.ndarray, item_count: int) -> None:
        self._counters = np.array(counters, dtype=np.uint64).reshape(self._depth, self._width)
        self._total_count = int(item_count)


class SketchCollector:
    """components/sketching/sketch_collector.py:24: an entity that feeds every event's value to a sketch
    and consumes the event.  ``value_extractor`` must be a ``KeyExtractor`` to run on the device."""

    def __init__(self, name: str, sketch, value_extractor=None, weight_extractor=None):
        self.name = name
        self._sketch = sketch
        self._value_extractor = value_extractor if value_extractor is not None else KeyExtractor()
        self._weight_extractor = weight_extractor
        self._events_processed = 0

    sketch = property(lambda self: self._sketch)
    events_processed = property(lambda self: self._events_processed)

    def clear(self) -> None:
        self._sketch.clear()
        self._events_processed = 0
