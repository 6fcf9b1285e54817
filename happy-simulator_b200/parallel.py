"""Mirror of happysimulator/parallel for partitioned runs.

``ParallelSimulation`` without links runs every partition as its own Simulation (own heap,
own creation counter -- parallel/simulation.py:170-195) and aggregates the summaries exactly
as the reference's ``_build_summary`` does.  Partitions connected by ``PartitionLink``s need the
windowed coordinator (parallel/coordinator.py:75-227); that logical-process mode is SURVEY.md
8(f) row 4 and is rejected here rather than approximated."""
from __future__ import annotations

import time as _time
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
    """parallel/link.py:18-79 (validation only; linked execution is not lowered yet)."""
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
        if links:
            raise UnsupportedModelError("partitions connected by PartitionLinks need the windowed coordinator "
                                        "(logical-process mode, SURVEY.md 8(f) row 4)")
        self._seed = seed
        self._partitions = partitions
        self._simulations: dict[str, Simulation] = {}
        for k, p in enumerate(partitions):
            self._simulations[p.name] = Simulation(start_time=start_time, end_time=end_time, duration=duration,
                                                   sources=p.sources or None, entities=p.entities or None,
                                                   probes=p.probes or None, trace_recorder=p.trace_recorder,
                                                   fault_schedule=p.fault_schedule, seed=seed, replica=k,
                                                   device=device)

    @property
    def simulations(self) -> dict[str, Simulation]:
        return dict(self._simulations)

    def run(self) -> ParallelSimulationSummary:
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
