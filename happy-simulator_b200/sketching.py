"""Sketches as instrumentation sinks: the host-side mirrors of the reference's
``HyperLogLog`` / ``CountMinSketch`` / ``BloomFilter`` / ``TopK`` and of ``SketchCollector`` /
``TopKCollector``, whose ``add()`` runs on the device.

Reference: happysimulator/sketching/hyperloglog.py:57-250, sketching/count_min_sketch.py:52-328,
sketching/bloom_filter.py:57-307, sketching/topk.py:37-285, components/sketching/sketch_collector.py:24-104,
components/sketching/topk_collector.py:22-150.  The device keeps one sketch state per replica (uint8
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


class BloomFilter:
    """sketching/bloom_filter.py:57: set membership over size_bits bits and num_hashes double-hashed probes."""

    def __init__(self, size_bits: int, num_hashes: int | None = None, seed: int | None = None):
        if size_bits <= 0:
            raise ValueError(f"size_bits must be positive, got {size_bits}")
        if num_hashes is not None and num_hashes <= 0:
            raise ValueError(f"num_hashes must be positive, got {num_hashes}")
        self._size_bits = size_bits
        self._num_hashes = num_hashes if num_hashes is not None else 7
        self._seed = seed if seed is not None else 0
        self._bits = np.zeros((size_bits + 63) // 64, np.uint64)
        self._total_count = 0
        self._tab = None

    @classmethod
    def from_expected_items(cls, n: int, fp_rate: float, seed: int | None = None) -> "BloomFilter":
        """bloom_filter.py:37-54,112-137: m = ceil(-n ln p / ln(2)^2) (64 for n = 0), k = max(1, round(m / n ln 2))."""
        if n < 0:
            raise ValueError(f"n must be non-negative, got {n}")
        if not 0 < fp_rate < 1:
            raise ValueError(f"fp_rate must be in (0, 1), got {fp_rate}")
        m = 64 if n == 0 else math.ceil(-n * math.log(fp_rate) / (math.log(2) ** 2))
        k = 1 if n == 0 else max(1, round((m / n) * math.log(2)))
        return cls(size_bits=m, num_hashes=k, seed=seed)

    size_bits = property(lambda self: self._size_bits)
    num_hashes = property(lambda self: self._num_hashes)
    item_count = property(lambda self: self._total_count)

    @property
    def _bits_set(self) -> int:
        return int(sum(bin(int(w)).count("1") for w in self._bits))

    fill_ratio = property(lambda self: self._bits_set / self._size_bits)

    @property
    def false_positive_rate(self) -> float:
        return 0.0 if self._bits_set == 0 else (self._bits_set / self._size_bits) ** self._num_hashes

    def _key_table(self, upto: int):
        if self._tab is None or self._tab.shape[1] < upto:
            self._tab = lowering.bloom_table(self._size_bits, self._num_hashes, self._seed, max(upto, 64))
        return self._tab

    def add(self, item: int, count: int = 1) -> None:
        if count < 0:
            raise ValueError(f"count must be non-negative, got {count}")
        if count == 0:
            return
        t = self._key_table(int(item) + 1)
        self._total_count += count
        for i in range(self._num_hashes):
            b = int(t[i, item])
            self._bits[b >> 6] |= np.uint64(1 << (b & 63))

    def contains(self, item: int) -> bool:
        t = self._key_table(int(item) + 1)
        return all((int(self._bits[int(t[i, item]) >> 6]) >> (int(t[i, item]) & 63)) & 1 for i in range(self._num_hashes))

    __contains__ = contains

    def merge(self, other: "BloomFilter") -> None:
        if not isinstance(other, BloomFilter):
            raise TypeError(f"Can only merge with BloomFilter, got {type(other).__name__}")
        if other._size_bits != self._size_bits:
            raise ValueError(f"Cannot merge: size_bits differs ({self._size_bits} vs {other._size_bits})")
        if other._num_hashes != self._num_hashes:
            raise ValueError(f"Cannot merge: num_hashes differs ({self._num_hashes} vs {other._num_hashes})")
        if other._seed != self._seed:
            raise ValueError(f"Cannot merge: seeds differ ({self._seed} vs {other._seed})")
        self._bits |= other._bits
        self._total_count += other._total_count

    def clear(self) -> None:
        self._bits[:] = 0
        self._total_count = 0

    def _load_device_state(self, words: np.ndarray, item_count: int) -> None:
        self._bits = np.array(words, dtype=np.uint64)
        self._total_count = int(item_count)


class FrequencyEstimate:
    """sketching/base.py: (item, count, error) as TopK.top() returns them."""

    __slots__ = ("item", "count", "error")

    def __init__(self, item, count, error):
        self.item, self.count, self.error = item, count, error

    def __eq__(self, o):
        return (self.item, self.count, self.error) == (o.item, o.count, o.error)

    def __repr__(self):
        return f"FrequencyEstimate(item={self.item!r}, count={self.count}, error={self.error})"


class TopK:
    """sketching/topk.py:37: Space-Saving heavy hitters over k counters; the counters live in a dict, whose
    insertion order decides which of several minimal counters is replaced (topk.py:116-128)."""

    def __init__(self, k: int, seed: int | None = None):
        if k <= 0:
            raise ValueError(f"k must be positive, got {k}")
        self._k = k
        self._counters: dict = {}          # item -> [count, error], insertion ordered
        self._total_count = 0

    k = property(lambda self: self._k)
    item_count = property(lambda self: self._total_count)
    tracked_count = property(lambda self: len(self._counters))

    def add(self, item, count: int = 1) -> None:
        if count < 0:
            raise ValueError(f"count must be non-negative, got {count}")
        if count == 0:
            return
        self._total_count += count
        c = self._counters.get(item)
        if c is not None:
            c[0] += count
        elif len(self._counters) < self._k:
            self._counters[item] = [count, 0]
        else:
            victim = min(self._counters, key=lambda it: self._counters[it][0])    # first minimum in dict order
            floor = self._counters.pop(victim)[0]
            self._counters[item] = [floor + count, floor]

    def estimate(self, item) -> int:
        c = self._counters.get(item)
        return c[0] if c is not None else 0

    def max_error(self) -> int:
        return min((c[0] for c in self._counters.values()), default=0)

    def estimate_with_error(self, item) -> FrequencyEstimate:
        c = self._counters.get(item)
        return FrequencyEstimate(item, c[0], c[1]) if c is not None else FrequencyEstimate(item, 0, self.max_error())

    def top(self, n: int | None = None) -> list:
        order = sorted(self._counters.items(), key=lambda kv: kv[1][0], reverse=True)    # stable, like the reference
        return [FrequencyEstimate(it, c[0], c[1]) for it, c in order[: len(order) if n is None else n]]

    def __contains__(self, item) -> bool:
        return item in self._counters

    def guaranteed_threshold(self) -> int:
        return self._total_count // self._k

    def merge(self, other: "TopK") -> None:
        """topk.py:216-258: tracked items add up (count and error); the others are add()-ed with their count,
        then inherit the other's error; the total grows by the other's total."""
        if not isinstance(other, TopK):
            raise TypeError(f"Can only merge with TopK, got {type(other).__name__}")
        if other._k != self._k:
            raise ValueError(f"Cannot merge TopK with k={other._k} into k={self._k}")
        before = set(self._counters)
        for item, (cnt, err) in list(other._counters.items()):
            if item in self._counters:
                self._counters[item][0] += cnt
                self._counters[item][1] += err
            else:
                self.add(item, cnt)
                if item in self._counters:
                    self._counters[item][1] += err
        self._total_count += other._total_count - sum(c[0] for it, c in other._counters.items() if it not in before)

    def clear(self) -> None:
        self._counters.clear()
        self._total_count = 0

    def _load_device_state(self, row: np.ndarray, item_count: int) -> None:
        n = int(row[0])
        self._counters = {int(row[1 + 3 * j]): [int(row[2 + 3 * j]), int(row[3 + 3 * j])] for j in range(n)}
        self._total_count = int(item_count)


class ReservoirSampler:
    """sketching/reservoir.py:30: a uniform sample of ``size`` stream items (Algorithm R), driven by the sampler's
    own ``random.Random(seed)``.  On the device the same MT19937 stream runs per replica (csrc/hs_sketch.h,
    hs_reservoir_add); the state it ends in is written back here, generator included."""

    def __init__(self, size: int, seed: int | None = None):
        import random
        if size <= 0:
            raise ValueError(f"size must be positive, got {size}")
        self._size = size
        self._reservoir: list = []
        self._total_count = 0
        self._rng = random.Random(seed)

    capacity = property(lambda self: self._size)
    item_count = property(lambda self: self._total_count)
    sample_size = property(lambda self: len(self._reservoir))
    is_full = property(lambda self: len(self._reservoir) >= self._size)

    def add(self, item, count: int = 1) -> None:
        if count < 0:
            raise ValueError(f"count must be non-negative, got {count}")
        for _ in range(count):
            self._total_count += 1
            if len(self._reservoir) < self._size:
                self._reservoir.append(item)
            else:                                         # kept with probability size / items seen
                j = self._rng.randint(0, self._total_count - 1)
                if j < self._size:
                    self._reservoir[j] = item

    def sample(self) -> list:
        return list(self._reservoir)

    def __iter__(self):
        return iter(self._reservoir)

    def __len__(self) -> int:
        return len(self._reservoir)

    def __getitem__(self, index):
        return self._reservoir[index]

    def merge(self, other: "ReservoirSampler") -> None:
        """reservoir.py:140-184: each slot of the new sample is drawn from self with probability
        seen(self) / seen(both), else from other -- with this sampler's generator."""
        if not isinstance(other, ReservoirSampler):
            raise TypeError(f"Can only merge with ReservoirSampler, got {type(other).__name__}")
        if other._size != self._size:
            raise ValueError(f"Cannot merge: capacity differs ({self._size} vs {other._size})")
        combined = self._total_count + other._total_count
        if combined == 0:
            return
        fresh = []
        for _ in range(min(self._size, combined)):
            src = self if self._rng.random() < self._total_count / combined else other
            if src._reservoir:
                fresh.append(src._reservoir[self._rng.randint(0, len(src._reservoir) - 1)])
        self._reservoir = fresh[: self._size]
        self._total_count = combined

    def clear(self) -> None:
        self._reservoir.clear()
        self._total_count = 0

    def __repr__(self):
        return f"ReservoirSampler(capacity={self._size}, sampled={len(self._reservoir)}, seen={self._total_count})"

    def _load_device_state(self, row: np.ndarray, item_count: int) -> None:
        load_reservoir_state(self, row)


def load_reservoir_state(sampler, row) -> None:
    """Fill a ReservoirSampler (this module's or the reference's: same private fields) from one replica's row of
    FlatModel.sketch_views: (items held, generator index, items seen, mt[624], sample slots).  A replica that saw no
    item never started its generator: the sampler's own state stays."""
    n, mti, total = int(row[0]), int(row[1]), int(row[2])
    sampler._reservoir = [int(x) for x in row[3 + 624: 3 + 624 + n]]
    sampler._total_count = total
    if total:
        sampler._rng.setstate((3, tuple(int(x) for x in row[3: 3 + 624]) + (mti,), None))


class TopKCollector:
    """components/sketching/topk_collector.py:22: SketchCollector specialised to TopK."""

    def __init__(self, name: str, k: int, value_extractor=None, count_extractor=None, seed: int | None = None):
        self.name = name
        self._topk = TopK(k=k, seed=seed)
        self._value_extractor = value_extractor if value_extractor is not None else KeyExtractor()
        self._count_extractor = count_extractor
        self._events_processed = 0

    k = property(lambda self: self._topk.k)
    events_processed = property(lambda self: self._events_processed)
    total_count = property(lambda self: self._topk.item_count)
    tracked_count = property(lambda self: self._topk.tracked_count)

    def top(self, n: int | None = None):
        return self._topk.top(n)

    def estimate(self, item) -> int:
        return self._topk.estimate(item)

    def __contains__(self, item) -> bool:
        return item in self._topk

    def max_error(self) -> int:
        return self._topk.max_error()

    def guaranteed_threshold(self) -> int:
        return self._topk.guaranteed_threshold()

    def clear(self) -> None:
        self._topk.clear()
        self._events_processed = 0


class LatencyExtractor:
    """value_extractor for QuantileEstimator: "the request's latency in seconds", i.e. what Sink records
    (components/common.py:39-41: (event.time - created_at).to_seconds())."""

    request_latency = True

    def __call__(self, event):
        return (event.time - event.context["created_at"]).to_seconds()


class TDigest:
    """sketching/tdigest.py:47: quantiles from centroids (mean, count) that are small near the tails.
    add() buffers int(compression * 2) values, then sorts them in as unit centroids and merges neighbours
    under max_size(q) = 4 N / (compression pi sqrt(q (1 - q)))."""

    def __init__(self, compression: float = 100.0, seed: int | None = None):
        if compression <= 0:
            raise ValueError(f"compression must be positive, got {compression}")
        self._compression = compression
        self._means: list[float] = []
        self._counts: list[int] = []
        self._total_count = 0
        self._min_value = None
        self._max_value = None
        self._buffer: list[float] = []
        self._buffer_size = int(compression * 2)

    compression = property(lambda self: self._compression)
    item_count = property(lambda self: self._total_count)
    min = property(lambda self: self._min_value)
    max = property(lambda self: self._max_value)

    @property
    def centroid_count(self) -> int:
        self._flush()
        return len(self._means)

    def _max_size(self, q: float) -> float:
        q = max(0.0001, min(0.9999, q))
        return self._total_count * 4 / (self._compression * math.pi * math.sqrt(q * (1 - q)))

    def add(self, value: float, count: int = 1) -> None:
        if count < 0:
            raise ValueError(f"count must be non-negative, got {count}")
        if count == 0:
            return
        if self._min_value is None or value < self._min_value:
            self._min_value = value
        if self._max_value is None or value > self._max_value:
            self._max_value = value
        self._total_count += count
        self._buffer.extend([value] * count)
        if len(self._buffer) >= self._buffer_size:
            self._flush()

    def _flush(self) -> None:
        if not self._buffer:
            return
        pairs = list(zip(self._means, self._counts)) + [(v, 1) for v in sorted(self._buffer)]
        self._buffer = []
        self._compress(pairs)

    def _compress(self, pairs) -> None:
        if len(pairs) > 1:
            pairs.sort(key=lambda p: p[0])                      # stable: old centroids stay ahead of equal new values
            out = [list(pairs[0])]
            running = pairs[0][1]
            for mean, cnt in pairs[1:]:
                limit = self._max_size((running + cnt / 2) / self._total_count)
                last = out[-1]
                if last[1] + cnt <= limit:
                    tot = last[1] + cnt
                    last[0] = (last[0] * last[1] + mean * cnt) / tot
                    last[1] = tot
                else:
                    out.append([mean, cnt])
                running += cnt
            pairs = out
        self._means = [p[0] for p in pairs]
        self._counts = [p[1] for p in pairs]

    def quantile(self, q: float) -> float:
        """tdigest.py:192-262: walk the centroids to the one whose weight range holds q N and interpolate
        (towards min / max in the first / last centroid, towards the previous mean elsewhere)."""
        if not 0 <= q <= 1:
            raise ValueError(f"Quantile must be in [0, 1], got {q}")
        self._flush()
        n = len(self._means)
        if n == 0:
            raise ValueError("Cannot compute quantile of empty digest")
        if q == 0:
            return self._min_value if self._min_value is not None else self._means[0]
        if q == 1:
            return self._max_value if self._max_value is not None else self._means[-1]
        target = q * self._total_count
        run = 0.0
        for i in range(n):
            mean, cnt = self._means[i], self._counts[i]
            lo = 0.0 if i == 0 else run
            hi = self._total_count if i == n - 1 else run + cnt
            if lo <= target <= hi:
                if i == 0:
                    if self._min_value is not None and target < cnt / 2:
                        return self._min_value + (target / (cnt / 2)) * (mean - self._min_value)
                    return mean
                if i == n - 1:
                    if self._max_value is not None:
                        rest = self._total_count - run
                        if target > run + rest / 2:
                            return mean + ((target - run - rest / 2) / (rest / 2)) * (self._max_value - mean)
                    return mean
                t = (target - lo) / cnt
                if t < 0.5:
                    prev = self._means[i - 1]
                    return prev + (mean - prev) * (0.5 + t)
                return mean
            run += cnt
        return self._means[-1]

    def percentile(self, p: float) -> float:
        if not 0 <= p <= 100:
            raise ValueError(f"Percentile must be in [0, 100], got {p}")
        return self.quantile(p / 100.0)

    def cdf(self, value: float) -> float:
        """tdigest.py:264-308."""
        self._flush()
        if not self._means:
            return 0.0
        if self._min_value is not None and value <= self._min_value:
            return 0.0
        if self._max_value is not None and value >= self._max_value:
            return 1.0
        below = 0.0
        for i, (mean, cnt) in enumerate(zip(self._means, self._counts)):
            if mean >= value:
                if i == 0:
                    if self._min_value is not None:
                        return ((value - self._min_value) / (mean - self._min_value)) * (cnt / 2) / self._total_count
                    return 0.0
                prev = self._means[i - 1]
                if prev < value < mean:
                    return (below + ((value - prev) / (mean - prev)) * cnt / 2) / self._total_count
                return below / self._total_count
            below += cnt
        return 1.0

    def merge(self, other: "TDigest") -> None:
        """tdigest.py:326-352: flush both, concatenate the centroids, combine count / min / max, re-compress."""
        if not isinstance(other, TDigest):
            raise TypeError(f"Can only merge with TDigest, got {type(other).__name__}")
        self._flush()
        other._flush()
        pairs = list(zip(self._means, self._counts)) + list(zip(other._means, other._counts))
        self._total_count += other._total_count
        if other._min_value is not None and (self._min_value is None or other._min_value < self._min_value):
            self._min_value = other._min_value
        if other._max_value is not None and (self._max_value is None or other._max_value > self._max_value):
            self._max_value = other._max_value
        self._compress(pairs)

    def clear(self) -> None:
        self._means, self._counts, self._buffer = [], [], []
        self._total_count, self._min_value, self._max_value = 0, None, None

    def _load_device_state(self, raw: np.ndarray, item_count: int | None = None, capacity: int | None = None) -> None:
        """raw: the row's bytes of one replica ({n_centroids, n_buffer, total, min, max}, centroids, buffer)."""
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        n_c, n_b = (int(x) for x in raw[:8].view(np.uint32))
        total = int(raw[8:16].view(np.int64)[0])
        mn, mx = (float(x) for x in raw[16:32].view(np.float64))
        cap = capacity if capacity is not None else (len(raw) - 32 - (self._buffer_size * 8 + 15) // 16 * 16) // 16
        cen = raw[32: 32 + n_c * 16]
        self._means = [float(x) for x in cen.view(np.float64)[0::2]]
        self._counts = [int(x) for x in cen.view(np.int64)[1::2]]
        off = 32 + cap * 16
        self._buffer = [float(x) for x in raw[off: off + n_b * 8].view(np.float64)]
        self._total_count = total
        self._min_value, self._max_value = (mn, mx) if total else (None, None)


class QuantileEstimator:
    """components/sketching/quantile_estimator.py:35: an entity feeding a TDigest with each event's value;
    ``value_extractor`` must be a ``LatencyExtractor`` to run on the device."""

    def __init__(self, name: str, value_extractor=None, compression: float = 100.0, seed: int | None = None):
        self.name = name
        self._tdigest = TDigest(compression=compression, seed=seed)
        self._value_extractor = value_extractor if value_extractor is not None else LatencyExtractor()
        self._events_processed = 0

    compression = property(lambda self: self._tdigest.compression)
    events_processed = property(lambda self: self._events_processed)
    sample_count = property(lambda self: self._tdigest.item_count)
    min = property(lambda self: self._tdigest.min)
    max = property(lambda self: self._tdigest.max)

    def quantile(self, q: float) -> float:
        return self._tdigest.quantile(q)

    def percentile(self, p: float) -> float:
        return self._tdigest.percentile(p)

    def cdf(self, value: float) -> float:
        return self._tdigest.cdf(value)

    def summary(self) -> dict:
        """quantile_estimator.py:157-186 (LatencyPercentiles as a dict)."""
        if self._tdigest.item_count == 0:
            return dict(p50=0.0, p75=0.0, p90=0.0, p95=0.0, p99=0.0, p999=0.0, min=None, max=None, count=0)
        return dict(p50=self.percentile(50), p75=self.percentile(75), p90=self.percentile(90), p95=self.percentile(95),
                    p99=self.percentile(99), p999=self.percentile(99.9), min=self.min, max=self.max,
                    count=self._tdigest.item_count)

    def clear(self) -> None:
        self._tdigest.clear()
        self._events_processed = 0
