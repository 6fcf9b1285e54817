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
                # e.g. an end time whose nanoseconds do not survive the trip through float seconds: the clamped window
                # ends 1 ns short of it, for ever -- the reference's coordinator never returns from such a run
                raise ValueError(f"window of {self.window_s} s does not advance the clock at {cur} ns (end {end_ns} ns): "
                                 "the reference's WindowedCoordinator would loop for ever")
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


class LinkedRun:
    """A LinkedModel on one GPU: one engine per partition (the same replicas in each), one coordinator.

    ``run()`` is WindowedCoordinator.run (coordinator.py:75-172): for every window, every partition runs
    ``hs_run(end_ns=window end, resume=window > 0, HS_RUN_LINKED)``, then ``hs_coordinator_exchange`` drains the
    partitions' outboxes in partition order.  Replica r of partition q draws from the Philox replica word
    ``q + g * (P + 1)`` (g = global replica index), the coordinator from ``P + g * (P + 1)``."""

    def __init__(self, lm: LinkedModel, *, device: int = 0):
        from . import engine as _engine
        import torch
        lm.validate()
        self.lm, self.device = lm, device
        # one CUDA stream for every partition and the coordinator: the window loop (partitions x windows launches plus the
        # barrier kernels) is queued without a host synchronisation in between
        self._stream = torch.cuda.Stream(device=device) if torch.cuda.is_available() else None
        sp = self._stream.cuda_stream if self._stream is not None else None
        self._stream_ptr = sp
        self.engines = [_engine.Engine(device, stream=sp) for _ in lm.models]
        for e, m in zip(self.engines, lm.models):
            e.upload(m)
        self.coordinator = None
        self.windows = 0

    def close(self):
        for e in self.engines:
            e.close()
        if self.coordinator is not None:
            self.coordinator.close()

    def run(self, *, seed, end_ns, n_replicas=1, replica_index_base=0, caps=None, flags=A.HS_RUN_ORDER_HASH, queue_ring=0):
        """caps: per-partition dicts of record_cap / sample_cap / service_cap (or one dict for all).  Returns the
        per-partition outputs (Engine.read_outputs) and (delivered, lost, overflowed) per replica."""
        from . import engine as _engine
        lm, nP = self.lm, self.lm.n_partitions
        if self.coordinator is not None:
            self.coordinator.close()
        self.coordinator = _engine.Coordinator(self.device, n_replicas, lm.n_streams, seed=seed, rid_base=nP,
                                               rid_stride=nP + 1, replica_index_base=replica_index_base, stream=self._stream_ptr)
        caps = caps or {}
        link_args = [lm.link_descs(q) for q in range(nP)]
        ends = lm.window_ends(end_ns)
        for w, wend in enumerate(ends):
            for q, e in enumerate(self.engines):
                c = caps[q] if isinstance(caps, (list, tuple)) else caps
                e.run(_engine.make_params(seed=seed, end_ns=wend, n_replicas=n_replicas, rid_base=q, rid_stride=nP + 1,
                                          replica_index_base=replica_index_base, engine=3, resume=1 if w else 0,
                                          flags=flags | A.HS_RUN_LINKED, queue_ring=queue_ring, **c))
            for q, e in enumerate(self.engines):
                arr, dst = link_args[q]
                if dst:
                    self.coordinator.exchange(e, arr, [self.engines[d] for d in dst])
        self.windows = len(ends)
        outs = [e.read_outputs() for e in self.engines]
        return outs, self.coordinator.read()
