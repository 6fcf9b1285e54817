"""Partitions joined by PartitionLinks (first half): the data model of a linked run.

Reference: parallel/simulation.py:31-284 (one Simulation per partition), parallel/routing.py:17-63 (the router that
moves events aimed at another partition into an outbox), parallel/coordinator.py:75-227 (windows of
min(link.min_latency), barrier, exchange with loss / latency override).

Here a partition is a FlatModel of its own in which every entity of ANOTHER partition that it sends to appears as an
HS_ENT_REMOTE row (link slot, entity id over there); the coordinator's window ends are computed on the host, in the
float arithmetic of coordinator.py:88-95, and every partition runs window after window with
``hs_run(end_ns=window end, resume=window > 0)``; ``hs_coordinator_exchange`` moves the outboxes at each barrier."""
from __future__ import annotations

from dataclasses import dataclass, field

from . import _abi as A
from .model import FlatModel


@dataclass(frozen=True)
class LinkSpec:
    """One outgoing link of a partition, as the engine sees it (hs_link_desc + the partition it ends in)."""
    dest: int                       # index of the destination partition
    latency_kind: int               # HS_SVC_CONSTANT | HS_SVC_EXPONENTIAL
    latency_mean_s: float
    packet_loss: float = 0.0
    stream: int = 0                 # id of the latency object (shared objects share their draw counter)


@dataclass
class LinkedModel:
    models: list[FlatModel]
    names: list[str]
    links: list[list[LinkSpec]]     # links[p][slot]
    window_s: float
    n_streams: int = 1
    objects: list[list] = field(default_factory=list)     # per partition: entity id -> user object (lowering)

    @property
    def n_partitions(self) -> int:
        return len(self.models)

    def window_ends(self, end_ns: int, start_ns: int = 0) -> list[int]:
        """WindowedCoordinator.run's window ends (coordinator.py:86-96): float seconds, clamped to the end time,
        converted back with Instant.from_seconds (truncation)."""
        ends, cur = [], int(start_ns)
        end_s = float(end_ns) / 1_000_000_000
        while cur < end_ns:
            w = float(cur) / 1_000_000_000 + self.window_s
            if w > end_s:
                w = end_s
            nxt = int(w * 1_000_000_000)
            if nxt <= cur:
                raise ValueError(f"window of {self.window_s} s does not advance the clock at {cur} ns")
            ends.append(nxt)
            cur = nxt
        return ends

    def link_descs(self, p: int):
        """(ctypes hs_link_desc array, destination partition indices) of partition p's outgoing links."""
        arr = (A.LinkDesc * max(1, len(self.links[p])))()
        for k, l in enumerate(self.links[p]):
            arr[k].latency_kind, arr[k].stream = int(l.latency_kind), int(l.stream)
            arr[k].latency_mean_s, arr[k].packet_loss = float(l.latency_mean_s), float(l.packet_loss)
        return arr, [int(l.dest) for l in self.links[p]]

    def validate(self) -> None:
        for p, ls in enumerate(self.links):
            for l in ls:
                if not 0 <= l.dest < self.n_partitions or l.dest == p:
                    raise ValueError("link destination out of range")
                if self.models[l.dest].inbox_cap <= 0:
                    raise ValueError(f"partition {self.names[l.dest]!r} is a link destination but has no inbox")
                if not 0 <= l.stream < self.n_streams:
                    raise ValueError("latency stream id out of range")
        for p, m in enumerate(self.models):
            for i in m.ids_of(A.HS_ENT_REMOTE):
                slot, dst_ent = int(m.entities["i0"][i]), int(m.entities["i1"][i])
                if not 0 <= slot < len(self.links[p]):
                    raise ValueError(f"partition {self.names[p]!r}: REMOTE row {i} uses link slot {slot} of {len(self.links[p])}")
                dm = self.models[self.links[p][slot].dest]
                if not 0 <= dst_ent < dm.n_entities or int(dm.entities["kind"][dst_ent]) in (A.HS_ENT_SOURCE, A.HS_ENT_PROBE, A.HS_ENT_REMOTE):
                    raise ValueError(f"partition {self.names[p]!r}: REMOTE row {i} points at entity {dst_ent} of "
                                     f"{self.names[self.links[p][slot].dest]!r}, which cannot receive requests")
            if m.ids_of(A.HS_ENT_REMOTE) and m.outbox_cap <= 0:
                raise ValueError(f"partition {self.names[p]!r} has REMOTE rows but no outbox")
