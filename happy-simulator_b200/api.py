"""Host-side mirror of the reference's modelling API for the accelerated path.

Same names, argument meaning and error behaviour as the reference classes they
stand for (cited per class), so models and tests read like the reference's own;
``Simulation.run()`` lowers the object graph (lowering.py), runs it on the CUDA
engine through the C-ABI (engine.py) and writes the results back onto the entity
objects, which is where the reference's callers read them
(``sink.latencies_s``, ``server.stats``, ``source.generated_count``, ``lb.stats``).

Randomness: the reference draws from Python's and numpy's global MT19937 streams;
here every stochastic consumer owns a Philox stream keyed by ``seed`` (the
``Simulation(seed=)`` argument, default ``happysim_b200.default_seed``) -- the same
streams the Philox plug-ins inject into the unmodified reference in the parity tests.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import math
import time as _time
from typing import Any, Callable

import numpy as np

from . import _abi as A
from . import lowering
from .engine import Engine, make_params

default_seed = 0


def seed(value: int) -> None:
    """Set the default Philox key of subsequent ``Simulation`` objects (cf. random.seed)."""
    global default_seed
    default_seed = int(value)


# ----------------------------------------------------------------------------- time
class Duration:
    """core/temporal.py:24-160 -- nanosecond duration; from_seconds truncates like the reference."""
    __slots__ = ("nanoseconds",)

    def __init__(self, nanoseconds: int):
        self.nanoseconds = nanoseconds

    @classmethod
    def from_seconds(cls, seconds):
        if isinstance(seconds, int):
            return cls(seconds * 1_000_000_000)
        if isinstance(seconds, float):
            return cls(int(seconds * 1_000_000_000))
        raise TypeError("seconds must be int or float")

    def to_seconds(self) -> float:
        return float(self.nanoseconds) / 1_000_000_000

    def __eq__(self, o):
        return isinstance(o, Duration) and self.nanoseconds == o.nanoseconds

    def __lt__(self, o):
        return self.nanoseconds < o.nanoseconds

    def __hash__(self):
        return hash(self.nanoseconds)

    def __repr__(self):
        return f"Duration({self.to_seconds()}s)"


class Instant:
    """core/temporal.py:165-300 -- nanosecond time point."""
    __slots__ = ("nanoseconds",)

    def __init__(self, nanoseconds: int):
        self.nanoseconds = nanoseconds

    @classmethod
    def from_seconds(cls, seconds):
        if isinstance(seconds, int):
            return cls(seconds * 1_000_000_000)
        if isinstance(seconds, float):
            return cls(int(seconds * 1_000_000_000))
        raise TypeError("seconds must be int or float")

    def to_seconds(self) -> float:
        return float(self.nanoseconds) / 1_000_000_000

    def __add__(self, other):
        if isinstance(other, Duration):
            return Instant(self.nanoseconds + other.nanoseconds)
        if isinstance(other, (int, float)):
            return Instant(self.nanoseconds + int(other * 1_000_000_000))
        return NotImplemented

    def __sub__(self, other):
        if isinstance(other, Instant):
            return Duration(self.nanoseconds - other.nanoseconds)
        if isinstance(other, Duration):
            return Instant(self.nanoseconds - other.nanoseconds)
        if isinstance(other, (int, float)):
            return Instant(self.nanoseconds - int(other * 1_000_000_000))
        return NotImplemented

    def __eq__(self, o):
        return isinstance(o, Instant) and self.nanoseconds == o.nanoseconds

    def __lt__(self, o):
        return self.nanoseconds < o.nanoseconds

    def __le__(self, o):
        return self.nanoseconds <= o.nanoseconds

    def __gt__(self, o):
        return self.nanoseconds > o.nanoseconds

    def __ge__(self, o):
        return self.nanoseconds >= o.nanoseconds

    def __hash__(self):
        return hash(self.nanoseconds)

    def __repr__(self):
        return f"Instant({self.to_seconds()}s)"


Instant.Epoch = Instant(0)


# ----------------------------------------------------------------------------- plug-ins
@dataclass(frozen=True)
class ConstantRateProfile:
    """load/profile.py:37-47"""
    rate: float

    def get_rate(self, time) -> float:
        return self.rate


@dataclass(frozen=True)
class LinearRampProfile:
    """load/profile.py:51-74"""
    duration_s: float
    start_rate: float
    end_rate: float

    def get_rate(self, time) -> float:
        t = time.to_seconds()
        if t <= 0:
            return self.start_rate
        if t >= self.duration_s:
            return self.end_rate
        return self.start_rate + (t / self.duration_s) * (self.end_rate - self.start_rate)


@dataclass(frozen=True)
class SpikeProfile:
    """load/profile.py:77-110"""
    baseline_rate: float = 10.0
    spike_rate: float = 150.0
    warmup_s: float = 10.0
    spike_duration_s: float = 15.0

    def get_rate(self, time) -> float:
        t = time.to_seconds()
        if t < self.warmup_s:
            return self.baseline_rate
        if t < self.warmup_s + self.spike_duration_s:
            return self.spike_rate
        return self.baseline_rate


@dataclass(frozen=True)
class StepProfile:
    """A piecewise-constant rate profile (not in the reference's load/profile.py; the reference's examples define
    such profiles themselves, e.g. examples/queuing/m_m_1_queue.py:104-169): ``rates[k]`` applies from
    ``breakpoints[k - 1]`` (inclusive) to ``breakpoints[k]`` (exclusive).  Lowered as HS_PROF_STEP."""
    breakpoints: tuple = ()
    rates: tuple = (1.0,)

    def __post_init__(self):
        if len(self.rates) != len(self.breakpoints) + 1:
            raise ValueError("StepProfile: n breakpoints need n + 1 rates")
        if any(b <= a for a, b in zip(self.breakpoints, self.breakpoints[1:])):
            raise ValueError("StepProfile: breakpoints must ascend")

    def get_rate(self, time) -> float:
        import bisect
        return float(self.rates[bisect.bisect_right(self.breakpoints, time.to_seconds())])

    @classmethod
    def from_profile(cls, profile, end_s: float, scan_step_s: float | None = None) -> "StepProfile":
        """Tabulate any step-function ``Profile`` (exactly, see lowering.step_table_from_profile)."""
        breaks, rates = lowering.step_table_from_profile(profile, scan_end_s=float(end_s), scan_step_s=scan_step_s)
        return cls(tuple(breaks), tuple(rates))


class _ArrivalTimeProvider:
    """load/arrival_time_provider.py:28-47 (constant-rate profiles only on the device)."""

    def __init__(self, profile, start_time: Instant):
        self.profile = profile
        self.current_time = start_time


class ConstantArrivalTimeProvider(_ArrivalTimeProvider):
    """load/providers/constant_arrival.py:11-23"""


class PoissonArrivalTimeProvider(_ArrivalTimeProvider):
    """load/providers/poisson_arrival.py:18-31"""


class _LatencyDistribution:
    """distributions/latency_distribution.py:17-41"""

    def __init__(self, mean_latency):
        self._mean_latency = mean_latency.to_seconds() if isinstance(mean_latency, Duration) else float(mean_latency)


class ConstantLatency(_LatencyDistribution):
    """distributions/constant.py:17-35"""


class ExponentialLatency(_LatencyDistribution):
    """distributions/exponential.py:17-45"""

    def __init__(self, mean_latency):
        super().__init__(mean_latency)
        self._lambda = 1 / self._mean_latency


class FIFOQueue:
    """components/queue_policy.py:75-114"""

    def __init__(self, capacity: float = float("inf")):
        self._capacity = capacity

    @property
    def capacity(self):
        return self._capacity


class LIFOQueue(FIFOQueue):
    """components/queue_policy.py:117-156"""


class FixedConcurrency:
    """components/server/concurrency.py:66-140"""

    def __init__(self, max_concurrent: int):
        if max_concurrent < 1:
            raise ValueError(f"max_concurrent must be >= 1, got {max_concurrent}")
        self._max_concurrent = max_concurrent

    @property
    def limit(self) -> int:
        return self._max_concurrent


class RoundRobin:
    """components/load_balancer/strategies.py:50-72"""


class ConsistentHash:
    """components/load_balancer/strategies.py:336-433 (default key extraction: metadata client_id)."""

    def __init__(self, virtual_nodes: int = 100, get_key: Callable | None = None):
        if virtual_nodes < 1:
            raise ValueError(f"virtual_nodes must be >= 1, got {virtual_nodes}")
        if get_key is not None:
            raise lowering.UnsupportedModelError("custom get_key callbacks cannot run on the device")
        self._virtual_nodes = virtual_nodes
        self._get_key = None


class UniformKeyContext:
    """context_fn for SimpleEventProvider: metadata {"client_id": k}, k ~ Uniform{0..population-1}
    drawn from the Philox routing stream (cf. the reference's
    SimpleEventProvider(context_fn=...) + distributions/uniform.py:57-63)."""

    def __init__(self, population: int):
        if population < 1:
            raise ValueError("population must be >= 1")
        self.key_population = int(population)


class ZipfKeyContext(UniformKeyContext):
    """context_fn for SimpleEventProvider: metadata {"client_id": k}, k ~ ZipfDistribution(range(population), s)
    (distributions/zipf.py:27-123: inverse transform, bisect_left over the cumulative probabilities) drawn from
    the Philox routing stream; rank 0 is the hottest key."""

    def __init__(self, population: int, s: float = 1.0):
        super().__init__(population)
        if s < 0:
            raise ValueError(f"s must be non-negative, got {s}")
        self.zipf_s = float(s)


# ----------------------------------------------------------------------------- entities
class Entity:
    """core/entity.py:31-127"""

    def __init__(self, name: str):
        self.name = name


class SimpleEventProvider:
    """load/source.py:31-86"""

    def __init__(self, target: Entity, event_type: str = "Request", stop_after: Instant | None = None,
                 context_fn=None):
        self._target = target
        self._event_type = event_type
        self._stop_after = stop_after
        self._context_fn = context_fn
        self._generated = 0


class Source(Entity):
    """load/source.py:92-341"""

    def __init__(self, name: str, event_provider, arrival_time_provider):
        super().__init__(name)
        self._event_provider = event_provider
        self._time_provider = arrival_time_provider
        self._generated_count = 0

    @staticmethod
    def _resolve_stop_after(stop_after):
        if stop_after is None or isinstance(stop_after, Instant):
            return stop_after
        return Instant.from_seconds(stop_after)

    @classmethod
    def _make(cls, provider_cls, rate, target, event_type, name, stop_after, event_provider):
        if event_provider is None:
            if target is None:
                raise ValueError("Either 'target' or 'event_provider' must be provided")
            event_provider = SimpleEventProvider(target, event_type, cls._resolve_stop_after(stop_after))
        return cls(name=name, event_provider=event_provider,
                   arrival_time_provider=provider_cls(ConstantRateProfile(rate=rate), start_time=Instant.Epoch))

    @classmethod
    def constant(cls, rate, target=None, event_type="Request", *, name="Source", stop_after=None, event_provider=None):
        return cls._make(ConstantArrivalTimeProvider, rate, target, event_type, name, stop_after, event_provider)

    @classmethod
    def poisson(cls, rate, target=None, event_type="Request", *, name="Source", stop_after=None, event_provider=None):
        return cls._make(PoissonArrivalTimeProvider, rate, target, event_type, name, stop_after, event_provider)

    @classmethod
    def with_profile(cls, profile, target=None, event_type="Request", *, poisson=True, name="Source",
                     stop_after=None, event_provider=None):
        """load/source.py:271-318"""
        if event_provider is None:
            if target is None:
                raise ValueError("Either 'target' or 'event_provider' must be provided")
            event_provider = SimpleEventProvider(target, event_type, cls._resolve_stop_after(stop_after))
        provider_cls = PoissonArrivalTimeProvider if poisson else ConstantArrivalTimeProvider
        return cls(name=name, event_provider=event_provider,
                   arrival_time_provider=provider_cls(profile, start_time=Instant.Epoch))

    @property
    def generated_count(self) -> int:
        return self._generated_count


class _Queue:
    """components/queue.py:76-170 (state holder; the protocol itself runs on the device)."""

    def __init__(self, name, policy):
        self.name = name
        self.policy = policy
        self.stats_dropped = 0
        self.stats_accepted = 0


@dataclass(frozen=True)
class ServerStats:
    """components/server/server.py:34-40"""
    requests_completed: int = 0
    requests_rejected: int = 0
    total_service_time: float = 0.0


class Server(Entity):
    """components/server/server.py:43-300 (QueuedResource + FixedConcurrency + service distribution)."""

    def __init__(self, name: str, concurrency=1, service_time=None, queue_policy=None, queue_capacity=None,
                 downstream: Entity | None = None):
        super().__init__(name)
        if queue_policy is None:
            queue_policy = FIFOQueue(capacity=queue_capacity if queue_capacity is not None else float("inf"))
        self._queue = _Queue(f"{name}.queue", queue_policy)
        self._concurrency_model = FixedConcurrency(concurrency) if isinstance(concurrency, int) else concurrency
        self._service_time = service_time or ConstantLatency(0.01)
        self._downstream = downstream
        self._requests_completed = 0
        self._requests_rejected = 0
        self._total_service_time = 0.0
        self._service_times: list[float] = []

    @property
    def downstream(self):
        return self._downstream

    @downstream.setter
    def downstream(self, target):
        self._downstream = target

    @property
    def concurrency(self) -> int:
        return self._concurrency_model.limit

    @property
    def stats_accepted(self) -> int:
        return self._queue.stats_accepted

    @property
    def stats_dropped(self) -> int:
        return self._queue.stats_dropped

    @property
    def stats(self) -> ServerStats:
        return ServerStats(self._requests_completed, self._requests_rejected, self._total_service_time)

    @property
    def average_service_time(self) -> float:
        return sum(self._service_times) / len(self._service_times) if self._service_times else 0.0


@dataclass
class CachingServerStats:
    """examples/load-balancing/common.py:90-97"""
    requests_processed: int = 0
    cache_hits: int = 0
    cache_misses: int = 0


class CachingServer(Entity):
    """examples/load-balancing/common.py:100-275: a server with a local TTL cache in front of a shared datastore.
    Same constructor; ``datastore`` is accepted and unused (its read latency is ``datastore_read_latency_s``).
    The request's customer id is its routing key (``UniformKeyContext`` / ``ZipfKeyContext`` on the source).
    ``cache_capacity`` must exceed the key population: the reference class raises on its first eviction."""

    def __init__(self, name: str, server_id: int = 0, datastore=None, cache_capacity: int = 100, cache_ttl_s: float = 30.0,
                 cache_read_latency_s: float = 0.0001, datastore_read_latency_s: float = 0.005,
                 processing_latency_s: float = 0.001):
        super().__init__(name)
        if cache_ttl_s <= 0:
            raise ValueError(f"ttl must be > 0, got {cache_ttl_s}")          # eviction_policies.py:174
        self.server_id = server_id
        self._datastore = datastore
        self._cache_capacity = cache_capacity
        self._cache_ttl_s = cache_ttl_s
        self._cache_read_latency_s = cache_read_latency_s
        self._datastore_read_latency_s = datastore_read_latency_s
        self._processing_latency_s = processing_latency_s
        self._queue = _Queue(f"{name}.queue", FIFOQueue())
        self.stats = CachingServerStats()
        self._insert_times: dict[str, float] = {}     # TTLEviction._insert_times after the run: "customer:<id>" -> seconds

    @property
    def stats_accepted(self) -> int:
        return self._queue.stats_accepted

    @property
    def stats_dropped(self) -> int:
        return self._queue.stats_dropped

    @property
    def hit_rate(self) -> float:
        total = self.stats.cache_hits + self.stats.cache_misses
        return self.stats.cache_hits / total if total else 0.0

    @property
    def miss_rate(self) -> float:
        total = self.stats.cache_hits + self.stats.cache_misses
        return self.stats.cache_misses / total if total else 0.0

    @property
    def cache_size(self) -> int:
        return len(self._insert_times)

    @property
    def requests_processed(self) -> int:
        return self.stats.requests_processed


class Sink(Entity):
    """components/common.py:18-76"""

    def __init__(self, name: str = "Sink"):
        super().__init__(name)
        self.events_received = 0
        self.completion_times: list[Instant] = []
        self.latencies_s: list[float] = []
        self._latency_sum = 0.0

    def average_latency(self) -> float:
        if not self.events_received:
            return 0.0
        # the device accumulates sum(latencies_s) exactly as CPython's float sum() does
        return self._latency_sum / self.events_received

    def latency_stats(self) -> dict:
        n = len(self.latencies_s)
        if n == 0:
            return {"count": 0, "avg": 0.0, "min": 0.0, "max": 0.0, "p50": 0.0, "p99": 0.0}
        v = sorted(self.latencies_s)

        def pct(p):
            pos = p * (n - 1)
            lo = int(pos)
            hi = min(lo + 1, n - 1)
            frac = pos - lo
            return v[lo] * (1.0 - frac) + v[hi] * frac
        return {"count": n, "avg": sum(v) / n, "min": v[0], "max": v[-1], "p50": pct(0.50), "p99": pct(0.99)}


class Counter(Entity):
    """components/common.py:79-95"""

    def __init__(self, name: str = "Counter"):
        super().__init__(name)
        self.total = 0
        self.by_type: dict[str, int] = {}


@dataclass
class BackendInfo:
    """components/load_balancer/load_balancer.py (BackendInfo)"""
    backend: Entity
    weight: int = 1
    is_healthy: bool = True
    total_requests: int = 0


@dataclass(frozen=True)
class LoadBalancerStats:
    requests_received: int = 0
    requests_forwarded: int = 0
    requests_failed: int = 0
    no_backend_available: int = 0
    backends_marked_unhealthy: int = 0
    backends_marked_healthy: int = 0


class LoadBalancer(Entity):
    """components/load_balancer/load_balancer.py:60-473"""

    def __init__(self, name: str, backends=None, strategy=None, on_no_backend: str = "reject"):
        super().__init__(name)
        if on_no_backend not in ("reject", "queue"):
            raise ValueError(f"on_no_backend must be 'reject' or 'queue', got {on_no_backend}")
        self._strategy = strategy or RoundRobin()
        self._backends: dict[str, BackendInfo] = {}
        self._in_flight: dict = {}
        self._requests_received = 0
        self._requests_forwarded = 0
        for b in backends or []:
            self.add_backend(b)

    def add_backend(self, backend: Entity, weight: int = 1) -> None:
        if weight < 1:
            raise ValueError(f"weight must be >= 1, got {weight}")
        self._backends[backend.name] = BackendInfo(backend=backend, weight=weight)

    @property
    def stats(self) -> LoadBalancerStats:
        return LoadBalancerStats(requests_received=self._requests_received,
                                 requests_forwarded=self._requests_forwarded)


# ----------------------------------------------------------------------------- summary
@dataclass
class QueueStats:
    """instrumentation/summary.py:14-20"""
    peak_depth: int
    total_accepted: int
    total_dropped: int


@dataclass
class EntitySummary:
    """instrumentation/summary.py:23-44"""
    name: str
    entity_type: str
    events_handled: int
    queue_stats: QueueStats | None = None


@dataclass
class SimulationSummary:
    """instrumentation/summary.py:47-87"""
    duration_s: float
    total_events_processed: int
    events_cancelled: int = 0
    events_per_second: float = 0.0
    wall_clock_seconds: float = 0.0
    entities: dict[str, EntitySummary] = field(default_factory=dict)

    def to_dict(self) -> dict[str, Any]:
        return {"duration_s": self.duration_s, "total_events_processed": self.total_events_processed,
                "events_cancelled": self.events_cancelled, "events_per_second": self.events_per_second,
                "wall_clock_seconds": self.wall_clock_seconds,
                "entities": {k: vars(v) for k, v in self.entities.items()}}


def stock_streams(seed: int, n_replicas: int, n_draws: int, seed_stride: int = 1):
    """The reference's two process-global MT19937 streams as unit-rate exponential variates:
    row r is what ``random.seed(seed + r*seed_stride); numpy.random.seed(seed + r*seed_stride)`` yields.
    arrival: -math.log(1.0 - numpy.random.random())   (load/providers/poisson_arrival.py:31)
    service: -math.log(1.0 - random.random())         (random.expovariate, distributions/exponential.py:43)
    math.log is the host libm the reference itself calls."""
    import random as _random
    arr = np.empty((n_replicas, n_draws), np.float64)
    svc = np.empty((n_replicas, n_draws), np.float64)
    for r in range(n_replicas):
        s = seed + r * seed_stride
        u = np.random.RandomState(s).random_sample(n_draws)
        arr[r] = [-math.log(1.0 - x) for x in u]
        rnd = _random.Random(s)
        svc[r] = [-math.log(1.0 - rnd.random()) for _ in range(n_draws)]
    return arr, svc


_engines: dict[int, Engine] = {}


def _engine(device: int) -> Engine:
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]


class Simulation:
    """core/simulation.py:38-591 -- same constructor, ``run() -> SimulationSummary``.

    Extra keyword arguments (not in the reference): ``seed`` (Philox key), ``replica`` (Philox
    replica word), ``device``, ``queue_ring`` (first size of the device queue rings; they grow on overflow)."""

    def __init__(self, start_time: Instant | None = None, end_time: Instant | None = None, sources=None,
                 entities=None, probes=None, trace_recorder=None, fault_schedule=None, duration: float | None = None,
                 *, seed: int | None = None, replica: int = 0, device: int = 0, rng: str = "philox",
                 queue_ring: int | None = None):
        if duration is not None and end_time is not None:
            raise ValueError("Cannot specify both 'duration' and 'end_time'")
        if start_time is not None and start_time.nanoseconds != 0:
            raise lowering.UnsupportedModelError("start_time must be Instant.Epoch on the device engine")
        for nm, v in (("trace_recorder", trace_recorder), ("fault_schedule", fault_schedule)):
            if v:
                raise lowering.UnsupportedModelError(f"{nm}= is outside the accelerated path (SURVEY.md section 8)")
        self._start_time = Instant.Epoch
        if duration is not None:
            self._end_time = self._start_time + duration
        elif end_time is not None:
            self._end_time = end_time
        else:
            raise lowering.UnsupportedModelError("the device engine needs an explicit end_time or duration "
                                                 "(auto-termination is the reference's slow loop)")
        self._sources = list(sources or [])
        self._entities = list(entities or [])
        self._probes = list(probes or [])
        self._seed = default_seed if seed is None else int(seed)
        self._replica = int(replica)
        self._device = device
        if rng not in ("philox", "stock"):
            raise ValueError("rng must be 'philox' or 'stock'")
        self._rng = rng
        self._queue_ring = int(queue_ring) if queue_ring else 0
        self.last_run_info: dict = {}
        self._summary: SimulationSummary | None = None
        self._instant_cls = Instant
        self.model, self.objects = lowering.lower(self._sources, self._entities, probes=self._probes,
                                                  horizon_s=self._end_time.to_seconds())

    @property
    def summary(self):
        return self._summary

    # -- single run -------------------------------------------------------------
    def _rate_bound(self) -> float:
        ents = self.model.entities
        rate = 0.0
        for i in self.model.ids_of(A.HS_ENT_SOURCE):
            pi = int(ents["i3"][i])
            if pi == 0:
                rate += float(ents["d0"][i])
            else:        # non-constant profile: bound by its largest rate
                rate += lowering.profile_max_rate(self.model.profiles[pi - 1], self.model.profile_table)
        return rate

    def _caps(self, n_hint: int | None = None):
        dur = self._end_time.to_seconds()
        rate = self._rate_bound()
        req = int(rate * dur * 1.3 + 6 * math.sqrt(rate * dur + 1) + 64)
        n_srv = len(self.model.ids_of(A.HS_ENT_SERVER))
        chain = 1 if (n_srv <= 1 or self.model.ids_of(A.HS_ENT_LB)) else n_srv     # tandem: one start per stage
        return dict(record_cap=0, sample_cap=req, service_cap=req * chain)

    def _events_per_request(self) -> int:
        """Upper bound of processed events per generated request: ~7 per server stage a request can pass
        (ENQUEUE, NOTIFY, POLL, DELIVER, WORKER, CONTINUATION, completion POLL) plus tick, routing and sink."""
        n_srv = len(self.model.ids_of(A.HS_ENT_SERVER))
        stages = 1 if (n_srv <= 1 or self.model.ids_of(A.HS_ENT_LB)) else n_srv
        return 8 * stages + 8 + 2 * len(self.model.ids_of(A.HS_ENT_PROBE))

    def _queue_ring_hint(self) -> int:
        """First device ring size: the backlog an overloaded model would build over the run, bounded."""
        if self._queue_ring:
            return int(self._queue_ring)
        dur = self._end_time.to_seconds()
        ents = self.model.entities
        cap = 0.0
        for i in self.model.ids_of(A.HS_ENT_SERVER):
            mean = float(ents["d0"][i])
            cap += (max(1, int(ents["i0"][i])) / mean) if mean > 0 else float("inf")
        backlog = max(0.0, (self._rate_bound() - cap) * dur)
        ring = 256
        while ring < min(2.5 * backlog + 64, 1 << 22):
            ring *= 2
        return ring

    def run(self) -> SimulationSummary:
        return _run_many([self], seed=self._seed, seed_stride=0, rid_base=self._replica, rid_stride=0)[0]

    def _write_back(self, out, r: int) -> None:
        """Publish replica ``r`` onto the Python objects, where the reference's callers look."""
        st = out["entity_stats"][r]
        kinds = self.model.entities["kind"]
        s = out["summaries"][r]
        n_smp, n_svc = int(s["n_sink_samples"]), int(s["n_service_samples"])
        samples = out["sink_samples"][r][:n_smp] if out.get("sink_samples") is not None else None
        sinks = self.model.ids_of(A.HS_ENT_SINK) + self.model.ids_of(A.HS_ENT_PROBE)
        servers = self.model.ids_of(A.HS_ENT_SERVER)
        per_server = {i: [] for i in servers}
        per_sink = {i: None for i in sinks}
        if samples is not None:
            if len(sinks) == 1:
                per_sink[sinks[0]] = samples
            elif out.get("records") is not None:
                rec = out["records"][r][: int(s["events_processed"])]
                who = rec["entity"][(rec["kind"] == A.HS_EV_REQ_SINK) | (rec["kind"] == A.HS_EV_PROBE)][: len(samples)]
                for i in sinks:
                    per_sink[i] = samples[who == i]
        if out.get("service_samples") is not None:
            svc = out["service_samples"][r][:n_svc]
            if len(servers) == 1:
                per_server[servers[0]] = [float(x) for x in svc]
            elif out.get("records") is not None:
                rec = out["records"][r][: int(s["events_processed"])]
                who = rec["entity"][rec["kind"] == A.HS_EV_REQ_WORKER][: len(svc)]
                for ent, x in zip(who, svc):
                    per_server[int(ent)].append(float(x))
        # Probe objects: their ticking is objects[i] (a SOURCE row); the measurement row it targets
        # (kind PROBE, beyond len(objects)) carries the samples
        for i, o in enumerate(self.objects):
            if int(kinds[i]) == A.HS_ENT_SOURCE and hasattr(o, "data_sink"):
                pid = int(self.model.entities["target"][i])
                sm = per_sink.get(pid)
                if sm is not None:
                    o.data_sink._samples = [(float(int(t)) / 1_000_000_000, float(x))
                                            for t, x in zip(sm["completion_ns"], sm["latency_s"])]
        for i, o in enumerate(self.objects):
            k = int(kinds[i])
            row = st[i]
            if k == A.HS_ENT_SOURCE:
                o._generated_count = int(row["c0"])
                if hasattr(o._event_provider, "_generated"):
                    o._event_provider._generated = int(row["c1"])
            elif k == A.HS_ENT_SERVER:
                o._queue.stats_accepted, o._queue.stats_dropped = int(row["c0"]), int(row["c1"])
                o._requests_completed, o._requests_rejected = int(row["c2"]), int(row["c3"])
                o._total_service_time = float(row["f0"])
                o._service_times = per_server[i]
            elif k == A.HS_ENT_CACHE_SERVER:
                o._queue.stats_accepted, o._queue.stats_dropped = int(row["c0"]), int(row["c1"])
                o.stats.requests_processed, o.stats.cache_misses, o.stats.cache_hits = int(row["c2"]), int(row["c3"]), int(row["f0"])
                if out.get("sketches") is not None:
                    ins = self.model.cache_views(out["sketches"])[i][r]
                    K = len(ins) - 1
                    times = {("customer:unknown" if j == K else f"customer:{j}"): float(t) for j, t in enumerate(ins) if t != 0.0}
                    if hasattr(o, "_insert_times"):
                        o._insert_times = times
                    elif getattr(o, "_eviction_policy", None) is not None:      # the example's own object, already initialised
                        o._eviction_policy._insert_times = times
            elif k == A.HS_ENT_SINK and hasattr(o, "data"):          # LatencyTracker / ThroughputTracker
                o.count = int(row["c0"])
                sm = per_sink[i]
                if sm is not None:
                    one = getattr(o, "_sample_value", None) == "one" or type(o).__name__ == "ThroughputTracker"
                    o.data._samples = [(float(int(t)) / 1_000_000_000, 1.0 if one else float(x))
                                       for t, x in zip(sm["completion_ns"], sm["latency_s"])]
            elif k == A.HS_ENT_SINK:
                o.events_received = int(row["c0"])
                o._latency_sum = float(row["f0"])
                sm = per_sink[i]
                if sm is not None:
                    o.completion_times = [self._instant_cls(int(t)) for t in sm["completion_ns"]]
                    o.latencies_s = [float(x) for x in sm["latency_s"]]
            elif k == A.HS_ENT_COUNTER:
                o.total = int(row["c0"])
                o.by_type = {"Request": o.total} if o.total else {}
            elif k == A.HS_ENT_LB:
                o._requests_received, o._requests_forwarded = int(row["c0"]), int(row["c1"])
            elif k == A.HS_ENT_SKETCH:
                o._events_processed = int(row["c0"])
                sk = o._topk if hasattr(o, "_topk") else o._tdigest if hasattr(o, "_tdigest") else o._sketch
                if out.get("sketches") is not None and hasattr(sk, "_load_device_state"):
                    sk._load_device_state(self.model.sketch_views(out["sketches"])[i][r], int(row["c1"]))
                elif out.get("sketches") is not None:      # a reference sketch object: fill its own fields
                    state = self.model.sketch_views(out["sketches"])[i][r]
                    algo = int(self.model.entities["i0"][i])
                    if algo == A.HS_SK_HLL:
                        sk._registers = [int(x) for x in state]
                    elif algo == A.HS_SK_CMS:
                        sk._counters = [[int(x) for x in rowc] for rowc in state]
                    elif algo == A.HS_SK_BLOOM:
                        sk._bits = [int(x) for x in state]
                        sk._bits_set = sum(bin(w).count("1") for w in sk._bits)
                    else:                                  # TopK / TDigest: rebuild the reference's own cells
                        import sys as _sys
                        from . import sketching as _sk
                        mod = _sys.modules[type(sk).__module__]
                        if algo == A.HS_SK_RESERVOIR:
                            _sk.load_reservoir_state(sk, state)
                        elif algo == A.HS_SK_TOPK:
                            t = _sk.TopK(int(self.model.entities["i2"][i])); t._load_device_state(state, int(row["c1"]))
                            sk._counters = {it: mod._Counter(item=it, count=c[0], error=c[1]) for it, c in t._counters.items()}
                        else:
                            d = _sk.TDigest(float(self.model.entities["d0"][i])); d._load_device_state(state)
                            sk._centroids = [mod._Centroid(mean=m_, count=c_) for m_, c_ in zip(d._means, d._counts)]
                            sk._buffer = list(d._buffer)
                            sk._min_value, sk._max_value = d._min_value, d._max_value
                    sk._total_count = int(row["c1"])

    def _entity_summaries(self):
        """core/simulation.py:560-591: only objects passed as entities=, events_handled from
        count | events_received | stats_processed, queue stats for queued resources."""
        res = {}
        for o in self._entities:
            qs = None
            if hasattr(o, "_queue") and hasattr(o, "_concurrency_model"):
                qs = QueueStats(peak_depth=0, total_accepted=o.stats_accepted, total_dropped=o.stats_dropped)
            handled = 0
            for attr in ("count", "events_received", "stats_processed"):
                v = getattr(o, attr, None)
                if isinstance(v, int):
                    handled = v
                    break
            res[o.name] = EntitySummary(name=o.name, entity_type=type(o).__name__, events_handled=handled, queue_stats=qs)
        return res

    # -- ensembles ----------------------------------------------------------------
    def run_ensemble(self, n_replicas: int, *, seed: int | None = None, seed_stride: int = 0, rid_base: int = 0,
                     rid_stride: int = 1, replica_index_base: int = 0, replicas_per_cell: int = 1,
                     window_end_s: float | None = None, resume: bool = False, host: dict | None = None,
                     upload: bool = True, totals: bool = True, on_overflow: str = "grow", **caps):
        """N independent replicas of this model on the device; returns the raw per-replica arrays
        (summaries, entity_stats, optional recorder rings) and the engine's totals.

        ``window_end_s`` / ``resume`` cut one continuing run into windows exactly like the reference's
        ``Simulation._run_window`` (core/simulation.py:527-541): state stays resident in HBM between calls.
        ``host`` = caller-owned (pinned) buffers from ``Engine.alloc_host_outputs`` to read into.

        Every replica's ``status`` is checked.  A device queue ring that overflowed (the reference's queues are
        unbounded) is handled per ``on_overflow``: "grow" re-runs the ensemble with doubled rings (fresh,
        unwindowed runs only), "raise" raises, "ignore" returns the flagged statuses to the caller."""
        eng = _engine(self._device)
        if upload:
            eng.upload(self.model)
        if not resume:
            eng.set_trace(None, None)            # never inherit a stock-generator trace from an earlier run()
        end_ns = self._end_time.nanoseconds
        we = -1
        if window_end_s is not None:
            w = int(round(float(window_end_s) * 1e9))
            we = w if w < end_ns else -1
        ring = int(caps.pop("queue_ring", 0) or 0)
        bad = A.HS_ST_QUEUE_OVERFLOW | A.HS_ST_FEL_OVERFLOW | A.HS_ST_SKETCH_OVERFLOW
        for _ in range(12):
            eng.run(make_params(seed=self._seed if seed is None else seed, seed_stride=seed_stride, rid_base=rid_base,
                                rid_stride=rid_stride, end_ns=end_ns, n_replicas=n_replicas,
                                replica_index_base=replica_index_base, replicas_per_cell=replicas_per_cell,
                                window_end_ns=we, resume=int(bool(resume)), queue_ring=ring, **caps))
            out = eng.read_outputs(host)
            st = out["summaries"]["status"]
            flagged = int((st & bad != 0).sum())
            if not flagged or on_overflow == "ignore":
                break
            only_queue = not (int(np.bitwise_or.reduce(st)) & (A.HS_ST_FEL_OVERFLOW | A.HS_ST_SKETCH_OVERFLOW))
            if on_overflow == "grow" and only_queue and not resume and we < 0 and ring < (1 << 22):
                ring = max(512, 2 * (ring or 256))
                continue
            raise EnsembleStatusError(f"{flagged} of {n_replicas} replicas stopped early (status bits "
                                      f"{int(np.bitwise_or.reduce(st))}: 1 queue ring full, 2 event list full, 32 sketch "
                                      f"full); pass a larger queue_ring=", st.copy())
        if "max_events" in caps and int((st & A.HS_ST_EVENT_LIMIT != 0).sum()) and on_overflow != "ignore":
            raise EnsembleStatusError("max_events reached before end_time", st.copy())
        out["status"] = st
        out["queue_ring"] = ring
        if totals:
            out["totals"] = eng.read_totals()
        out["device_ms"] = eng.last_run_ms()
        return out


class EnsembleStatusError(RuntimeError):
    """Replicas of an ensemble stopped early; ``.status`` holds every replica's status word."""

    def __init__(self, msg, status):
        super().__init__(msg)
        self.status = status


def _same_topology(a, b) -> bool:
    """Two lowered models that differ at most in the per-cell columns: d0 (rates, mean service times) of any
    row and i0 (concurrency) of SERVER rows.  Such models run as cells of ONE launch (hs_model_desc.cell_d0/i0)."""
    ea, eb = a.entities, b.entities
    if ea.shape != eb.shape or a.n_cells or b.n_cells:
        return False
    for f in ea.dtype.names:
        if f == "d0":
            continue
        if f == "i0":
            srv = ea["kind"] == A.HS_ENT_SERVER
            if not np.array_equal(ea["i0"][~srv], eb["i0"][~srv]):
                return False
            continue
        if not np.array_equal(ea[f], eb[f]):
            return False
    for f in ("backends", "key_table", "profiles", "profile_table", "sketch_tables", "key_cdf"):
        x, y = np.asarray(getattr(a, f)), np.asarray(getattr(b, f))
        if x.shape != y.shape or x.tobytes() != y.tobytes():
            return False
    return True


def _run_many(sims, *, seed: int, seed_stride: int, rid_base: int, rid_stride: int, trace_fn=None):
    """Run ``sims`` -- Simulations of one topology (``_same_topology``), same end time and device -- as the
    replicas of ONE device launch, replica k = sims[k] with Philox key ``seed + k * seed_stride`` and replica word
    ``rid_base + k * rid_stride``; results are written back onto each Simulation's own objects.  A single
    Simulation is the n = 1 case (``Simulation.run``).  Ring capacities that turn out too small (recorder
    streams, device queues, the event limit) are doubled and the launch repeated; ``last_run_info`` records how
    many launches that took and their wall time."""
    t0 = _time.monotonic()
    lead = sims[0]
    n = len(sims)
    eng = _engine(lead._device)
    model = lead.model
    if n > 1:
        import copy
        model = copy.copy(lead.model)
        model.cell_d0 = np.stack([np.asarray(sm.model.entities["d0"], np.float64) for sm in sims])
        model.cell_i0 = np.stack([np.asarray(sm.model.entities["i0"], np.int32) for sm in sims])
    eng.upload(model)
    caps = {k: max(sm._caps()[k] for sm in sims) for k in ("record_cap", "sample_cap", "service_cap")}
    stock = getattr(lead, "_rng", "philox") == "stock" or trace_fn is not None
    if trace_fn is None:           # the reference's two MT19937 streams for Simulation(seed=, rng="stock")
        trace_fn = lambda n_draws: stock_streams(seed, n, n_draws, seed_stride)      # noqa: E731
    eng.set_trace(None, None)
    if stock:
        eng.set_trace(*trace_fn(caps["sample_cap"] * 2 + 64))
    m = lead.model
    # per-server service-time lists / per-collector samples are demultiplexed with the event records
    n_streams = len(m.ids_of(A.HS_ENT_SINK)) + len(m.ids_of(A.HS_ENT_PROBE))
    need_events = len(m.ids_of(A.HS_ENT_SERVER)) > 1 or n_streams > 1
    if m.ids_of(A.HS_ENT_PROBE):        # probe samples share the sample stream
        dur = lead._end_time.to_seconds()
        caps["sample_cap"] += max(sum(int(dur * float(sm.model.profiles[int(sm.model.entities["i3"][i]) - 1]["p"][0])) + 8
                                      for i in sm.model.ids_of(A.HS_ENT_SOURCE) if int(sm.model.entities["i3"][i]) > 0)
                                  for sm in sims)
    ring = max(sm._queue_ring_hint() for sm in sims)
    per_req = max(sm._events_per_request() for sm in sims)
    max_events = per_req * caps["sample_cap"] + 1_000_000
    launches, prev_final = 0, None
    for _ in range(24):
        kw = dict(caps)
        if need_events:
            kw["record_cap"] = kw["service_cap"] * 12
        eng.run(make_params(seed=seed, seed_stride=seed_stride, rid_base=rid_base, rid_stride=rid_stride,
                            end_ns=lead._end_time.nanoseconds, n_replicas=n, replicas_per_cell=1, flags=0,
                            max_events=max_events, queue_ring=ring, **kw))
        launches += 1
        out = eng.read_outputs()
        summ = out["summaries"]
        status = int(np.bitwise_or.reduce(summ["status"]))
        if status & A.HS_ST_TRACE_EXHAUSTED:
            caps = {k: 2 * v for k, v in caps.items()}
            max_events = per_req * caps["sample_cap"] + 1_000_000
            eng.set_trace(*trace_fn(caps["sample_cap"] * 2 + 64))
            continue
        if status & A.HS_ST_QUEUE_OVERFLOW and not (status & A.HS_ST_FEL_OVERFLOW):
            if ring >= (1 << 24):
                raise RuntimeError(f"device queue ring overflow at {ring} entries per server; the queue of this model "
                                   f"grows without bound -- pass Simulation(queue_ring=...) if that is intended")
            ring *= 4            # the reference's queues are unbounded: grow the device rings and run again
            continue
        if status & (A.HS_ST_FEL_OVERFLOW | A.HS_ST_SKETCH_OVERFLOW):
            raise RuntimeError(f"device structure overflow (status {status})")
        if status & A.HS_ST_EVENT_LIMIT:
            final = int(summ["final_time_ns"][summ["status"] & A.HS_ST_EVENT_LIMIT != 0].min())
            if prev_final is not None and final <= prev_final:
                raise RuntimeError("event limit reached: the model's clock does not advance (a source faster than "
                                   "one event per nanosecond never terminates in the reference either)")
            prev_final = final
            max_events *= 4      # a valid model denser in events than estimated: raise the safety valve
            continue
        if (int(summ["n_sink_samples"].max()) <= kw["sample_cap"] and int(summ["n_service_samples"].max()) <= kw["service_cap"]
                and (not need_events or int(summ["events_processed"].max()) <= kw["record_cap"])):
            break
        caps = dict(record_cap=0, sample_cap=2 * int(summ["n_sink_samples"].max()) + 64,
                    service_cap=2 * int(summ["n_service_samples"].max()) + 64)
        max_events = max(max_events, per_req * caps["sample_cap"] + 1_000_000)
    else:
        raise RuntimeError("could not size the device buffers for this model")
    if stock:
        eng.set_trace(None, None)            # the trace must not leak into a later ensemble on this engine
    wall = _time.monotonic() - t0
    res = []
    for k, sm in enumerate(sims):
        sm._write_back(out, k)
        s = summ[k]
        duration_s = float(int(s["final_time_ns"])) / 1_000_000_000
        ev = int(s["events_processed"])
        sm.last_run_info = {"launches": launches, "wall_s": wall, "queue_ring": ring, "batched_with": n,
                            "device_ms_last_launch": eng.last_run_ms(), "status": int(s["status"])}
        sm._summary = SimulationSummary(duration_s=duration_s, total_events_processed=ev, events_cancelled=0,
                                        events_per_second=ev / duration_s if duration_s > 0 else 0.0,
                                        wall_clock_seconds=wall, entities=sm._entity_summaries())
        res.append(sm._summary)
    return res


def _group_by_topology(sims):
    """Indices of ``sims`` grouped so that each group can run as one launch."""
    groups: list[list[int]] = []
    for i, sm in enumerate(sims):
        for g in groups:
            h = sims[g[0]]
            if (h._device == sm._device and h._end_time.nanoseconds == sm._end_time.nanoseconds
                    and getattr(h, "_rng", "philox") == getattr(sm, "_rng", "philox") and _same_topology(h.model, sm.model)):
                g.append(i)
                break
        else:
            groups.append([i])
    return groups


def run_lowered(ref_sim, model=None, objects=None, *, seed: int | None = None, replica: int = 0, device: int = 0,
                trace_fn=None):
    """Run a REFERENCE ``happysimulator.Simulation`` object on the device and write the results back
    onto its own entity objects (the hook shown in INTEGRATION.md section 3).  ``ref_sim`` only needs the
    reference's attributes ``_sources``, ``_entities``, ``_start_time``, ``_end_time``."""
    if model is None:
        model, objects = lowering.lower(ref_sim._sources, ref_sim._entities, probes=getattr(ref_sim, "_probes", None) or None,
                                        horizon_s=float(int(ref_sim._end_time.nanoseconds)) / 1e9)
    shell = Simulation.__new__(Simulation)
    shell._start_time = Instant.Epoch
    shell._end_time = Instant(int(ref_sim._end_time.nanoseconds))
    shell._sources, shell._entities = list(ref_sim._sources), list(ref_sim._entities)
    shell._seed = default_seed if seed is None else int(seed)
    shell._replica, shell._device, shell._summary = int(replica), device, None
    shell._rng, shell._queue_ring, shell.last_run_info = "philox", 0, {}
    shell._probes = []
    shell._instant_cls = type(ref_sim._start_time)
    shell.model, shell.objects = model, objects
    if trace_fn is not None:
        return _run_many([shell], seed=shell._seed, seed_stride=0, rid_base=shell._replica, rid_stride=0, trace_fn=trace_fn)[0]
    return shell.run()


# ----------------------------------------------------------------------------- parallel/runner.py
@dataclass
class RunConfig:
    """parallel/runner.py:42-54"""
    name: str
    build_fn: Callable
    seed: int | None = None


@dataclass
class ParallelResult:
    """parallel/runner.py:57-70 (+ ``status``: the replica's device status word, 0 = ran to end_time)"""
    name: str
    summary: SimulationSummary
    artifacts: dict[str, Any] = field(default_factory=dict)
    status: int = 0


class _ReplicaResults:
    """List-like view of an ensemble's ``ParallelResult``s.  Results are materialised on access (the
    write-back of a replica onto the model's Python objects is O(samples)); ``len``, indexing, slicing and
    iteration behave like the reference's list."""

    def __init__(self, sim, out, wall: float):
        self._sim, self.raw, self._wall = sim, out, wall

    def __len__(self):
        return len(self.raw["summaries"])

    def _one(self, i: int) -> ParallelResult:
        sim, out = self._sim, self.raw
        sim._write_back(out, i)
        s = out["summaries"][i]
        d = float(int(s["final_time_ns"])) / 1_000_000_000
        n = int(s["events_processed"])
        return ParallelResult(name=f"replica_{i}", status=int(s["status"]), summary=SimulationSummary(
            duration_s=d, total_events_processed=n, events_per_second=n / d if d > 0 else 0.0,
            wall_clock_seconds=self._wall, entities=sim._entity_summaries()))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._one(k) for k in range(*i.indices(len(self)))]
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        return self._one(i)

    def __iter__(self):
        return (self._one(i) for i in range(len(self)))

    @property
    def total_events_processed(self) -> int:
        return int(self.raw["summaries"]["events_processed"].sum())


class ParallelRunner:
    """parallel/runner.py:82-142 -- replicas run as one device ensemble instead of a process pool.

    Replica i uses Philox key ``base_seed + i`` (the reference seeds ``random`` with base_seed + i),
    so ``run_replicas(build, n, s)[i]`` equals ``Simulation(seed=s + i).run()``.  A sweep runs as one
    launch per topology: configurations whose lowered models differ only in rates, mean service times and
    server concurrency are the cells of one launch (hs_model_desc.cell_d0 / cell_i0)."""

    def __init__(self, max_workers: int | None = None, device: int = 0):
        self._max_workers = max_workers
        self._device = device

    def run_replicas(self, build_fn: Callable, n_replicas: int, base_seed: int = 42, **caps):
        """Returns a list-like of ParallelResult (materialised on access).  Device queue rings that overflow
        are grown and the ensemble re-run (``Simulation.run_ensemble``); every result carries ``status``."""
        sim = build_fn()
        t0 = _time.monotonic()
        out = sim.run_ensemble(n_replicas, seed=base_seed, seed_stride=1, rid_base=sim._replica, rid_stride=0,
                               queue_ring=sim._queue_ring_hint(), **caps)
        return _ReplicaResults(sim, out, _time.monotonic() - t0)

    def run_sweep(self, configs: list[RunConfig]) -> list[ParallelResult]:
        if not configs:
            return []
        sims = []
        for cfg in configs:
            sim = cfg.build_fn()
            if cfg.seed is not None:
                sim._seed = int(cfg.seed)
            sims.append(sim)
        res: list = [None] * len(sims)
        for g in _group_by_topology(sims):
            seeds = [sims[i]._seed for i in g]
            rids = [sims[i]._replica for i in g]
            ds = {b - a for a, b in zip(seeds, seeds[1:])}
            dr = {b - a for a, b in zip(rids, rids[1:])}
            if len(g) > 1 and len(ds) <= 1 and len(dr) <= 1 and min(ds | {0}) >= 0 and min(dr | {0}) >= 0:
                sums = _run_many([sims[i] for i in g], seed=seeds[0], seed_stride=(ds.pop() if ds else 0),
                                 rid_base=rids[0], rid_stride=(dr.pop() if dr else 0))
            else:            # seeds that are not an arithmetic progression: one launch each
                sums = [sims[i].run() for i in g]
            for i, sm in zip(g, sums):
                res[i] = ParallelResult(name=configs[i].name, summary=sm, status=sims[i].last_run_info.get("status", 0))
        return res
