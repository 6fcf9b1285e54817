"""Host-side mirrors of the reference's sample containers and collector entities.

``Data`` / ``BucketedData`` (instrumentation/data.py:19-270) are analysis utilities over
(time_s, value) samples; ``LatencyTracker`` / ``ThroughputTracker``
(instrumentation/collectors.py:18-89) are Sink-like entities whose ``handle_event`` the
device runs as HS_EV_REQ_SINK; ``Simulation.run()`` fills their ``data`` from the recorder
samples.  Aggregations follow the reference's expressions (CPython float ``sum`` etc.)."""
from __future__ import annotations

import math
import statistics
from collections import defaultdict
from typing import Any


def _percentile_sorted(ordered, p: float) -> float:
    """Linear-interpolation percentile of an ascending sequence, p in [0, 1]: the arithmetic of the reference's
    helper (instrumentation/data.py:197-210) restated -- rank = p (n - 1), then ``v[k] * (1 - f) + v[k + 1] * f`` with
    k = floor(rank), f = rank - k, in exactly that order of operations, so every percentile the mirrors report
    (Sink.latency_stats, Data.percentile, BucketedData p50 / p99) equals the reference's float for float."""
    count = len(ordered)
    if count == 0:
        return 0.0
    if p <= 0 or count == 1:
        return float(ordered[0])
    if p >= 1:
        return float(ordered[count - 1])
    rank = p * (count - 1)
    k = int(rank)
    f = rank - k
    upper = ordered[k + 1] if k + 1 < count else ordered[count - 1]
    return float(ordered[k] * (1.0 - f) + upper * f)


class BucketedData:
    """instrumentation/data.py:213-270"""

    def __init__(self) -> None:
        self._times, self._means, self._counts = [], [], []
        self._maxes, self._sums, self._p50s, self._p99s = [], [], [], []

    def times(self): return self._times
    def means(self): return self._means
    def counts(self): return self._counts
    def maxes(self): return self._maxes
    def sums(self): return self._sums
    def p50s(self): return self._p50s
    def p99s(self): return self._p99s

    def to_dict(self) -> dict[str, list]:
        return {"time_s": list(self._times), "mean": list(self._means), "p50": list(self._p50s),
                "p99": list(self._p99s), "max": list(self._maxes), "count": list(self._counts),
                "sum": list(self._sums)}

    def __len__(self): return len(self._times)
    def __bool__(self): return len(self._times) > 0


class Data:
    """instrumentation/data.py:19-195"""

    def __init__(self) -> None:
        self._samples: list[tuple[float, Any]] = []

    def add_stat(self, value: Any, time) -> None:
        self._samples.append((time.to_seconds(), value))

    def clear(self) -> None:
        self._samples.clear()

    @property
    def values(self):
        return self._samples

    def between(self, start_s: float, end_s: float) -> "Data":
        r = Data()
        r._samples = [(t, v) for t, v in self._samples if start_s <= t < end_s]
        return r

    def mean(self) -> float:
        vals = [v for _, v in self._samples]
        return sum(vals) / len(vals) if vals else 0.0

    def min(self) -> float:
        vals = [v for _, v in self._samples]
        return min(vals) if vals else 0.0

    def max(self) -> float:
        vals = [v for _, v in self._samples]
        return max(vals) if vals else 0.0

    def percentile(self, p: float) -> float:
        return _percentile_sorted(sorted(v for _, v in self._samples), p)

    def count(self) -> int:
        return len(self._samples)

    def sum(self) -> float:
        return sum(v for _, v in self._samples)

    def std(self) -> float:
        vals = [v for _, v in self._samples]
        return statistics.pstdev(vals) if len(vals) >= 2 else 0.0

    def bucket(self, window_s: float = 1.0) -> BucketedData:
        buckets: dict[int, list[float]] = defaultdict(list)
        for t, v in self._samples:
            buckets[math.floor(t / window_s)].append(float(v))
        res = BucketedData()
        for key in sorted(buckets):
            vals = buckets[key]
            vs = sorted(vals)
            res._times.append(key * window_s)
            res._means.append(sum(vals) / len(vals))
            res._counts.append(len(vals))
            res._maxes.append(max(vals))
            res._sums.append(sum(vals))
            res._p50s.append(_percentile_sorted(vs, 0.50))
            res._p99s.append(_percentile_sorted(vs, 0.99))
        return res

    def times(self): return [t for t, _ in self._samples]
    def raw_values(self): return [v for _, v in self._samples]

    def rate(self, window_s: float = 1.0) -> "Data":
        b = self.bucket(window_s)
        r = Data()
        for t, c in zip(b.times(), b.counts()):
            r._samples.append((t, c / window_s))
        return r

    def __len__(self): return len(self._samples)
    def __bool__(self): return len(self._samples) > 0


class _Collector:
    def __init__(self, name: str) -> None:
        self.name = name
        self.data = Data()
        self.count: int = 0


class LatencyTracker(_Collector):
    """instrumentation/collectors.py:18-60: (completion_time_s, latency_s) per event."""
    _sample_value = "latency"

    def __init__(self, name: str = "LatencyTracker") -> None:
        super().__init__(name)

    def p50(self) -> float: return self.data.percentile(0.50)
    def p99(self) -> float: return self.data.percentile(0.99)
    def mean_latency(self) -> float: return self.data.mean()
    def summary(self, window_s: float = 1.0) -> BucketedData: return self.data.bucket(window_s)


class ThroughputTracker(_Collector):
    """instrumentation/collectors.py:63-89: one (time_s, 1.0) sample per event."""
    _sample_value = "one"

    def __init__(self, name: str = "ThroughputTracker") -> None:
        super().__init__(name)

    def throughput(self, window_s: float = 1.0) -> BucketedData: return self.data.bucket(window_s)


class _ProbeProfile:
    """instrumentation/probe.py:25-35 -- constant rate, but NOT a ConstantRateProfile: probe ticks take the
    general (adaptive Simpson + Brent) arrival path, e.g. interval 0.1 s -> 100000000, 200000000,
    299999999, ... ns."""

    def __init__(self, interval_seconds: float):
        if interval_seconds <= 0:
            raise ValueError("Probe interval must be positive.")
        self.rate = 1.0 / interval_seconds
        self._interval = interval_seconds

    def get_rate(self, time) -> float:
        return self.rate


class _ProbeEventProvider:
    """instrumentation/probe.py:38-78"""

    def __init__(self, target, metric: str, data_sink: Data):
        self.target = target
        self.metric = metric
        self.data_sink = data_sink


class Probe:
    """instrumentation/probe.py:81-164 -- periodic sampler of ``getattr(target, metric)`` into a Data."""

    def __init__(self, target, metric: str, data: Data, interval: float = 1.0, start_time=None):
        from .api import ConstantArrivalTimeProvider, Instant
        if start_time is not None and start_time.nanoseconds != 0:
            raise NotImplementedError("probe start_time must be Instant.Epoch on the device engine")
        self.target, self.metric, self.data_sink = target, metric, data
        self.name = f"Probe_{target.name}_{metric}"
        self._event_provider = _ProbeEventProvider(target, metric, data)
        self._time_provider = ConstantArrivalTimeProvider(_ProbeProfile(interval), start_time=Instant.Epoch)
        self._generated_count = 0

    @property
    def generated_count(self) -> int:
        return self._generated_count

    @classmethod
    def on(cls, target, metric: str, interval: float = 1.0):
        data = Data()
        return cls(target=target, metric=metric, data=data, interval=interval), data

    @classmethod
    def on_many(cls, target, metrics, interval: float = 1.0):
        probes, data = [], {}
        for m in metrics:
            p, d = cls.on(target, m, interval=interval)
            probes.append(p)
            data[m] = d
        return probes, data
