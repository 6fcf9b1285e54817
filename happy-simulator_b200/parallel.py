"""Mirror of happysimulator/parallel for partitioned runs.

``ParallelSimulation`` without links runs every partition as its own Simulation (own heap,
own creation counter -- parallel/simulation.py:170-195) and aggregates the summaries exactly
as the reference's ``_build_summary`` does.  Partitions connected by ``PartitionLink``s run under the
windowed coordinator (parallel/coordinator.py:75-227) on the device: one engine per partition, one
``hs_run`` per window and partition, ``hs_coordinator_exchange`` at every barrier (linked.py)."""
from __future__ import annotations

import time as _time

import numpy as np
from dataclasses import dataclass, field
from typing import Any

from .api import EntitySummary, Simulation, SimulationSummary
from .lowering import UnsupportedModelError


@dataclass
class SimulationPartition:
    """parallel/partition.py:20-38"""
    name: str
    entities: list = field(default_factory=list)
    sources: list = field(default_factory=list)
    probes: list = field(default_factory=list)
    fault_schedule: Any = None
    trace_recorder: Any = None


@dataclass(frozen=True)
class PartitionLink:
    """parallel/link.py:18-79.  ``latency``: a ConstantLatency / ExponentialLatency (this package's or the
    reference's): the coordinator overrides every cross-partition event's time with send time + a sample of it
    (coordinator.py:207-210).  Without it the reference insists on ``event.time - send_time >= min_latency``
    (coordinator.py:211-221), which no stock component can satisfy -- Entity.forward stamps the event with the current
    time -- so such a link cannot carry a lowered model's events."""
    source_partition: str
    dest_partition: str
    min_latency: float
    latency: Any = None
    packet_loss: float = 0.0

    def __post_init__(self) -> None:
        if self.min_latency <= 0:
            raise ValueError(f"PartitionLink min_latency must be > 0, got {self.min_latency}")
        if not (0.0 <= self.packet_loss < 1.0):
            raise ValueError(f"PartitionLink packet_loss must be in [0, 1), got {self.packet_loss}")
        if self.source_partition == self.dest_partition:
            raise ValueError(f"PartitionLink source and dest must differ, got '{self.source_partition}'")


@dataclass
class ParallelSimulationSummary:
    """parallel/summary.py:11-86"""
    duration_s: float
    total_events_processed: int
    events_per_second: float
    wall_clock_seconds: float
    partitions: dict[str, SimulationSummary] = field(default_factory=dict)
    entities: dict[str, EntitySummary] = field(default_factory=dict)
    partition_wall_times: dict[str, float] = field(default_factory=dict)
    speedup: float = 1.0
    parallelism_efficiency: float = 1.0
    total_windows: int = 0
    total_cross_partition_events: int = 0
    window_size_s: float = 0.0
    barrier_overhead_seconds: float = 0.0
    coordination_efficiency: float = 1.0


def _events_bound(model, end_ns: int) -> float:
    """Upper estimate of the Sink samples / service starts one replica of ``model`` produces (ring sizing)."""
    from . import _abi as A
    rate = 0.0
    for i in model.ids_of(A.HS_ENT_SOURCE):
        rate += float(model.entities["d0"][i])
    return (rate * 4 + 50.0) * (end_ns / 1e9 + 1.0) + 4.0 * max(model.inbox_cap, 0)


class ParallelSimulation:
    """parallel/simulation.py:31-284"""

    def __init__(self, partitions, *, start_time=None, end_time=None, duration=None, max_workers=None,
                 links=None, window_size=None, seed: int = 42, device: int = 0):
        if not partitions:
            raise ValueError("At least one partition is required")
        if duration is not None and end_time is not None:
            raise ValueError("Cannot specify both 'duration' and 'end_time'")
        names = [p.name for p in partitions]
        if len(set(names)) != len(names):
            raise ValueError("Partition names must be unique")       # parallel/validation.py
        self._seed = seed
        self._partitions = partitions
        self._device = device
        self._links = list(links or [])
        self._simulations: dict[str, Simulation] = {}
        self._linked = None
        if self._links:
            self._init_linked(start_time, end_time, duration, window_size)
            return
        for k, p in enumerate(partitions):
            self._simulations[p.name] = Simulation(start_time=start_time, end_time=end_time, duration=duration,
                                                   sources=p.sources or None, entities=p.entities or None,
                                                   probes=p.probes or None, trace_recorder=p.trace_recorder,
                                                   fault_schedule=p.fault_schedule, seed=seed, replica=k,
                                                   device=device)

    @property
    def simulations(self) -> dict[str, Simulation]:
        return dict(self._simulations)

    # ---- partitions joined by links --------------------------------------------------------------------------
    def _init_linked(self, start_time, end_time, duration, window_size):
        """parallel/validation.py:19-110 + parallel/simulation.py:84-150: check the declarations, find the events that
        cross partitions (a Server whose downstream lives elsewhere) and lower every partition to a model of its own."""
        from . import _abi as A
        from .api import Instant
        from .linked import LinkedModel, LinkSpec
        from .lowering import _service, lower
        parts = self._partitions
        names = [p.name for p in parts]
        if start_time is not None and int(start_time.nanoseconds) != 0:
            raise UnsupportedModelError("linked partitions start at Instant.Epoch")
        if duration is not None:
            end_time = Instant.Epoch + duration
        if end_time is None:
            raise UnsupportedModelError("linked partitions need an end time (auto-termination cannot end a Source)")
        self._end_ns = int(end_time.nanoseconds)
        index = {n: k for k, n in enumerate(names)}
        for l in self._links:                                    # validation.py:73-84
            if l.source_partition not in index:
                raise ValueError(f"PartitionLink references unknown source partition '{l.source_partition}'")
            if l.dest_partition not in index:
                raise ValueError(f"PartitionLink references unknown dest partition '{l.dest_partition}'")
        min_lat = min(l.min_latency for l in self._links)
        if window_size is not None and window_size > min_lat:    # validation.py:103-110
            raise ValueError(f"window_size ({window_size}s) must be <= min(link.min_latency) ({min_lat}s)")
        window = float(window_size if window_size is not None else min_lat)
        owner: dict[int, int] = {}
        for k, p in enumerate(parts):                            # validation.py:40-51
            for e in list(p.entities) + list(p.sources) + list(p.probes):
                if id(e) in owner and owner[id(e)] != k:
                    raise ValueError(f"Entity '{getattr(e, 'name', e)}' is in partitions '{names[owner[id(e)]]}' and '{p.name}'")
                owner[id(e)] = k
        streams: dict[int, int] = {}
        out_links: list[list] = [[] for _ in parts]
        slot_of: list[dict[int, int]] = [dict() for _ in parts]    # destination partition -> link slot
        for l in self._links:
            if l.latency is None:
                raise UnsupportedModelError(
                    f"PartitionLink {l.source_partition}->{l.dest_partition} has no latency override: the reference then "
                    "requires event.time - send_time >= min_latency (coordinator.py:211-221), and every stock component "
                    "forwards with the current time, so the reference itself raises on the first cross-partition event")
            kind, mean = _service(l.latency)
            s = streams.setdefault(id(l.latency), len(streams))
            q, d = index[l.source_partition], index[l.dest_partition]
            slot_of[q][d] = len(out_links[q])
            out_links[q].append(LinkSpec(d, kind, mean, float(l.packet_loss), s))
        models, objects = [], []
        hidden = lambda ents: {id(getattr(o, a)) for o in ents for a in ("queue", "driver", "worker")
                               if hasattr(o, "_concurrency_model") and hasattr(o, a)}
        for k, p in enumerate(parts):
            skip = hidden(p.entities)       # a reference script has to list a Server's hidden parts for its router
            ents = [e for e in p.entities if id(e) not in skip]
            remote = {}
            for e in ents:                  # the only edge that may leave a partition: Server -> downstream
                t = getattr(e, "_downstream", None) if hasattr(e, "_concurrency_model") else None
                if t is not None and id(t) in owner and owner[id(t)] != k:
                    d = owner[id(t)]
                    if d not in slot_of[k]:
                        raise ValueError(f"Entity '{getattr(e, 'name', e)}' in partition '{p.name}' references entity "
                                         f"'{getattr(t, 'name', t)}' in partition '{names[d]}' without a PartitionLink "
                                         f"('{p.name}' -> '{names[d]}')")       # validation.py:160-200
                    remote[id(t)] = slot_of[k][d]
            m, objs = lower(p.sources or [], ents, probes=p.probes or None, horizon_s=self._end_ns / 1e9, remote=remote)
            models.append(m)
            objects.append(objs)
        for k, m in enumerate(models):      # REMOTE rows: the entity's id over there
            m.entities = m.entities.copy()
            for i in m.ids_of(A.HS_ENT_REMOTE):
                d = out_links[k][int(m.entities["i0"][i])].dest
                where = [j for j, o in enumerate(objects[d]) if o is objects[k][i]]
                if not where:
                    raise UnsupportedModelError(f"'{getattr(objects[k][i], 'name', '?')}' is not an entity of partition '{names[d]}'")
                m.entities["i1"][i] = where[0]
            if m.ids_of(A.HS_ENT_REMOTE):
                m.outbox_cap = self.link_buffer
        for q in range(len(parts)):
            for l in out_links[q]:
                models[l.dest].inbox_cap = self.link_buffer * max(1, sum(1 for qq in range(len(parts)) for x in out_links[qq] if x.dest == l.dest))
        self._linked = LinkedModel(models, names, out_links, window_s=window, n_streams=max(1, len(streams)), objects=objects)
        self._linked.validate()
        self._linked.window_ends(self._end_ns)      # raises if the coordinator's clock could not reach the end time

    link_buffer = 256        # cross-partition events one replica may emit per window (class default; overflow is reported)
    queue_ring = 0           # device slots per server queue of a linked run (0: the engine's default); grown and re-run on overflow

    def _run_linked(self, n_replicas: int = 1, replica_index_base: int = 0):
        from . import _abi as A
        from .api import Instant
        from .linked import LinkedRun
        lm = self._linked
        t0 = _time.monotonic()
        run = LinkedRun(lm, device=self._device)
        try:
            caps = []
            for m in lm.models:
                ev = max(64, int(_events_bound(m, self._end_ns)))
                many = len(m.ids_of(A.HS_ENT_SINK)) + len(m.ids_of(A.HS_ENT_PROBE)) > 1 or len(m.ids_of(A.HS_ENT_SERVER)) > 1
                caps.append(dict(sample_cap=ev, service_cap=ev, record_cap=8 * ev if many else 0))   # records tell the sinks / servers apart
            # The reference's queues are unbounded, the device's are rings: a replica whose ring filled up stopped early
            # (HS_ST_QUEUE_OVERFLOW).  Like Simulation.run(), grow the ring and run the whole thing again (every window
            # starts from scratch: resume = 0 at window 0, a fresh coordinator) instead of handing that to the caller.
            ring = int(getattr(self, "queue_ring", 0) or 0)
            for _attempt in range(6):
                outs, (delivered, lost, over) = run.run(seed=self._seed, end_ns=self._end_ns, n_replicas=n_replicas,
                                                        replica_index_base=replica_index_base, caps=caps, flags=0, queue_ring=ring)
                status = 0
                for o in outs:
                    status |= int(np.bitwise_or.reduce(o["summaries"]["status"])) if len(o["summaries"]) else 0
                if not (status & A.HS_ST_QUEUE_OVERFLOW) or (status & ~(A.HS_ST_QUEUE_OVERFLOW | A.HS_ST_LINK_TIE)) or over.any():
                    break
                ring = max(512, 4 * ring)
            self.last_queue_ring = ring
        finally:
            run.close()
        wall = _time.monotonic() - t0
        bad = [(lm.names[q], int(s)) for q, o in enumerate(outs) for s in o["summaries"]["status"] if int(s) & ~A.HS_ST_LINK_TIE]
        if bad or over.any():
            bits = 0
            for _, s_ in bad:
                bits |= s_
            why = []
            if bits & A.HS_ST_QUEUE_OVERFLOW:
                why.append(f"a server queue outgrew {ring} device slots (ParallelSimulation.queue_ring)")
            if (bits & A.HS_ST_LINK_OVERFLOW) or over.any():
                why.append(f"an outbox / inbox outgrew ParallelSimulation.link_buffer = {self.link_buffer}")
            if bits & ~(A.HS_ST_QUEUE_OVERFLOW | A.HS_ST_LINK_OVERFLOW):
                why.append(f"status bits {bits & ~(A.HS_ST_QUEUE_OVERFLOW | A.HS_ST_LINK_OVERFLOW):#x}")
            raise RuntimeError(f"linked run did not complete cleanly: partition status {bad[:4]}, inbox overflows {int(over.sum())}: "
                               + "; ".join(why))
        self.link_ties = int(sum(int(s) & A.HS_ST_LINK_TIE != 0 for o in outs for s in o["summaries"]["status"]))
        self.last_outputs, self.last_delivered, self.last_lost = outs, delivered, lost
        return outs, delivered, lost, wall, run.windows

    def run(self) -> ParallelSimulationSummary:
        if self._linked is not None:
            return self._summarise_linked(*self._run_linked(1))
        return self._run_independent()

    def _summarise_linked(self, outs, delivered, lost, wall, windows) -> ParallelSimulationSummary:
        """coordinator.py:123-172: per-partition summaries from the partitions' final state, the aggregate like
        _build_summary; results are written back onto the script's own objects (replica 0)."""
        from .api import Instant
        lm = self._linked
        summaries = {}
        for q, name in enumerate(lm.names):
            shell = Simulation.__new__(Simulation)
            shell.model, shell.objects, shell._instant_cls = lm.models[q], lm.objects[q], Instant
            shell._entities = [o for o in self._partitions[q].entities if any(o is x for x in lm.objects[q])]
            shell._write_back(outs[q], 0)
            s = outs[q]["summaries"][0]
            d = float(int(s["final_time_ns"])) / 1e9
            ev = int(s["events_processed"])
            summaries[name] = SimulationSummary(duration_s=d, total_events_processed=ev, events_per_second=ev / d if d > 0 else 0.0,
                                                wall_clock_seconds=wall, entities=shell._entity_summaries())
        total = sum(s.total_events_processed for s in summaries.values())
        duration_s = max((s.duration_s for s in summaries.values()), default=0.0)
        merged = {}
        for s in summaries.values():
            merged.update(s.entities)
        n = len(summaries)
        return ParallelSimulationSummary(
            duration_s=duration_s, total_events_processed=total, events_per_second=total / duration_s if duration_s > 0 else 0.0,
            wall_clock_seconds=wall, partitions=summaries, entities=merged,
            partition_wall_times={nm: wall / n for nm in summaries}, speedup=1.0, parallelism_efficiency=1.0 / n if n else 1.0,
            total_windows=windows, total_cross_partition_events=int(delivered[0]), window_size_s=lm.window_s)

    def run_ensemble(self, n_replicas: int, replica_index_base: int = 0):
        """Linked partitions only: n replicas of the whole ParallelSimulation in one set of launches.  Returns
        {partition name: per-replica outputs (Engine.read_outputs)}, delivered and lost cross-partition events per
        replica."""
        if self._linked is None:
            raise UnsupportedModelError("run_ensemble is for partitions joined by PartitionLinks")
        outs, delivered, lost, wall, windows = self._run_linked(n_replicas, replica_index_base)     # a rank's shard: base = rank * n
        return {n: o for n, o in zip(self._linked.names, outs)}, delivered, lost

    def _run_independent(self) -> ParallelSimulationSummary:
        """Independent partitions (parallel/simulation.py:170-195).  Partitions whose lowered models share a
        topology (``api._same_topology``: they differ at most in rates, mean service times and concurrency) are
        the replicas of ONE device launch, replica word = partition index as in the sequential case; the others
        get a launch each.  ``partition_wall_times`` are measured: a launch group's wall time split evenly over
        its partitions; ``speedup`` keeps the reference's definition (sum of partition times / wall time), which
        is ~1 here because launch groups run one after another -- the gain of batching shows in the wall time."""
        from .api import _group_by_topology, _run_many
        t0 = _time.monotonic()
        names = list(self._simulations)
        sims = [self._simulations[n] for n in names]
        summaries, walls = {}, {}
        self.launch_groups = []
        for g in _group_by_topology(sims):
            t1 = _time.monotonic()
            rids = [sims[i]._replica for i in g]
            dr = {b - a for a, b in zip(rids, rids[1:])}
            if len(g) > 1 and len(dr) == 1 and min(dr) > 0:
                res = _run_many([sims[i] for i in g], seed=self._seed, seed_stride=0, rid_base=rids[0], rid_stride=dr.pop())
            else:
                res = [sims[i].run() for i in g]
            dt = _time.monotonic() - t1
            self.launch_groups.append([names[i] for i in g])
            for i, r in zip(g, res):
                summaries[names[i]] = r
                walls[names[i]] = dt / len(g)
        summaries = {n: summaries[n] for n in names}
        wall = _time.monotonic() - t0
        total = sum(s.total_events_processed for s in summaries.values())
        duration_s = max((s.duration_s for s in summaries.values()), default=0.0)
        merged = {}
        for s in summaries.values():
            merged.update(s.entities)
        seq = sum(walls.values())
        speedup = seq / wall if wall > 0 else 1.0
        n = len(summaries)
        return ParallelSimulationSummary(
            duration_s=duration_s, total_events_processed=total,
            events_per_second=total / duration_s if duration_s > 0 else 0.0, wall_clock_seconds=wall,
            partitions=summaries, entities=merged, partition_wall_times=walls, speedup=speedup,
            parallelism_efficiency=speedup / n if n else 1.0)
