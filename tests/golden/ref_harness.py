"""Drive the UNMODIFIED reference (imported from /root/reference) on a FlatModel.

Only tests/golden/gen_golden.py (and ad-hoc checks in the build container) use
this module: /root/reference does not exist on the GPU box.  It

  * rebuilds the reference object graph (Source / Server / Sink / Counter /
    LoadBalancer) that a FlatModel describes,
  * injects the shared Philox sampler through the reference's own plug-in points
    (ArrivalTimeProvider._get_target_integral_value,
    LatencyDistribution.get_latency, SimpleEventProvider(context_fn=)),
    SURVEY.md section 8(c),
  * taps EventHeap.pop to capture the processed-event
    sequence as (time_ns, sort_index, kind, entity) records,
  * returns the same arrays the oracle / the CUDA engine produce.
"""
from __future__ import annotations

import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("HS_REFERENCE_ROOT", "/root/reference")


def _import_reference():
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import happysimulator  # noqa: F401
    return happysimulator


class _RemoteStub:
    """Placeholder for an entity of another partition while one partition's objects are built."""

    def __init__(self, row):
        self.row = row


def run_reference(model, *, seed, rid=0, end_ns, names=None, chash_vnodes=None, stock_rng=False,
                  max_records=None, sketch_seeds=None, zipf_s=None, profile_objects=None, _build_only=False):
    """Run the reference on ``model`` (a happysim_b200.FlatModel); replica word ``rid``.

    stock_rng=True leaves the reference's own MT19937 streams in place (seeded
    random.seed(seed); np.random.seed(seed)) instead of the Philox plug-ins.
    """
    _import_reference()
    import oracle_lib as O
    from happysim_b200 import _abi as A

    from happysimulator.components.common import Counter, Sink
    from happysimulator.components.load_balancer.load_balancer import LoadBalancer
    from happysimulator.components.load_balancer.strategies import ConsistentHash, RoundRobin
    from happysimulator.components.queue_policy import FIFOQueue, LIFOQueue
    from happysimulator.components.queued_resource import _QueuedResourceWorkerAdapter
    from happysimulator.components.server.server import Server
    from happysimulator.core.event import ProcessContinuation
    from happysimulator.core.simulation import Simulation
    from happysimulator.core.temporal import Duration, Instant
    from happysimulator.distributions.constant import ConstantLatency
    from happysimulator.distributions.exponential import ExponentialLatency
    from happysimulator.distributions.latency_distribution import LatencyDistribution
    from happysimulator.load.arrival_time_provider import ArrivalTimeProvider
    from happysimulator.load.profile import ConstantRateProfile, LinearRampProfile, SpikeProfile
    from happysimulator.load.providers.constant_arrival import ConstantArrivalTimeProvider
    from happysimulator.load.providers.poisson_arrival import PoissonArrivalTimeProvider
    from happysimulator.load.source import SimpleEventProvider, Source
    from happysimulator.load.source_event import SourceEvent
    from happysimulator.instrumentation.probe import Probe
    from happysimulator.instrumentation.data import Data
    from happysimulator.components.sketching.sketch_collector import SketchCollector
    from happysimulator.sketching.hyperloglog import HyperLogLog
    from happysimulator.sketching.count_min_sketch import CountMinSketch
    from happysimulator.sketching.bloom_filter import BloomFilter
    from happysimulator.sketching.topk import TopK
    from happysimulator.components.sketching.topk_collector import TopKCollector
    from happysimulator.components.sketching.quantile_estimator import QuantileEstimator

    CachingServer = KVStore = None
    if (model.entities["kind"] == A.HS_ENT_CACHE_SERVER).any():
        # the example's own class, imported from the example file (examples/load-balancing/common.py:100-275)
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_lb_common", os.path.join(REFERENCE_ROOT, "examples", "load-balancing", "common.py"))
        lbmod = importlib.util.module_from_spec(spec)
        sys.modules["ref_lb_common"] = lbmod
        spec.loader.exec_module(lbmod)
        CachingServer = lbmod.CachingServer
        from happysimulator.components.datastore.kv_store import KVStore

    L = O.lib()
    ents = model.entities
    n = model.n_entities
    names = names or model.names or [f"e{i}" for i in range(n)]

    class PhiloxPoissonArrival(ArrivalTimeProvider):
        def __init__(self, profile, start_time, sid):
            super().__init__(profile, start_time)
            self._sid, self._n = sid, 0

        def _get_target_integral_value(self):
            u = L.hs_cpu_uniform(seed, rid, A.HS_STREAM_ARRIVAL | (self._sid << 8), self._n)
            self._n += 1
            return L.hs_cpu_exp1(u)

    class PhiloxExponentialLatency(LatencyDistribution):
        def __init__(self, mean, sid):
            super().__init__(mean)
            self._lambda = 1 / self._mean_latency
            self._sid, self._n = sid, 0

        def get_latency(self, current_time):
            u = L.hs_cpu_uniform(seed, rid, A.HS_STREAM_SERVICE | (self._sid << 8), self._n)
            self._n += 1
            return Duration.from_seconds(L.hs_cpu_exp1(u) / self._lambda)

    objs = [None] * n
    # leaves first: sinks / counters, then servers, then LBs, then sources
    for i in range(n):
        k = int(ents["kind"][i])
        if k == A.HS_ENT_SINK:
            objs[i] = Sink(names[i])
        elif k == A.HS_ENT_COUNTER:
            objs[i] = Counter(names[i])
        elif k == A.HS_ENT_REMOTE:
            objs[i] = _RemoteStub(i)
        elif k == A.HS_ENT_SKETCH:
            e = ents[i]
            sk_seed = (sketch_seeds or {}).get(i)
            extract = lambda ev: ev.context.get("metadata", {}).get("client_id")   # sketch_collector.py:36-41
            if int(e["i0"]) == A.HS_SK_TOPK:
                objs[i] = TopKCollector(names[i], k=int(e["i2"]), value_extractor=extract)
                continue
            if int(e["i0"]) == A.HS_SK_TDIGEST:     # the value Sink records (common.py:39-41)
                objs[i] = QuantileEstimator(names[i], compression=float(e["d0"]),
                                            value_extractor=lambda ev: (ev.time - ev.context["created_at"]).to_seconds())
                continue
            if int(e["i0"]) == A.HS_SK_RESERVOIR:    # starts from the generator state in the row's table
                from happysimulator.sketching.reservoir import ReservoirSampler
                sk = ReservoirSampler(size=int(e["i2"]))
                words = model.sketch_tables[int(e["i1"]): int(e["i1"]) + 625].view(np.uint32)
                sk._rng.setstate((3, tuple(int(x) for x in words), None))
                objs[i] = SketchCollector(names[i], sketch=sk, value_extractor=extract)
                continue
            if int(e["i0"]) == A.HS_SK_HLL:
                sk = HyperLogLog(precision=int(e["i2"]), seed=sk_seed)
            elif int(e["i0"]) == A.HS_SK_BLOOM:
                sk = BloomFilter(size_bits=int(e["i3"]), num_hashes=int(e["i2"]), seed=sk_seed)
            else:
                sk = CountMinSketch(width=int(e["i3"]), depth=int(e["i2"]), seed=sk_seed)
            objs[i] = SketchCollector(names[i], sketch=sk, value_extractor=extract)
    datastore = None
    for i in range(n):
        if int(ents["kind"][i]) != A.HS_ENT_CACHE_SERVER:
            continue
        e = ents[i]
        if datastore is None:
            datastore = KVStore(name="SharedDatastore", read_latency=0.005, write_latency=0.010)
        lat = [int(e["i2"]) / 1e9, int(e["l0"]) / 1e9, int(e["i3"]) / 1e9]
        assert [int(x * 1e9) for x in lat] == [int(e["i2"]), int(e["l0"]), int(e["i3"])], "delays must round-trip through seconds"
        assert int(e["i1"]) == A.HS_Q_FIFO
        # capacity K + 1 > key population K: the example cannot evict (common.py:264 raises on its first eviction)
        objs[i] = CachingServer(names[i], server_id=i, datastore=datastore, cache_capacity=int(e["i0"]) + 1,
                                cache_ttl_s=float(e["d0"]), cache_read_latency_s=lat[0], datastore_read_latency_s=lat[1],
                                processing_latency_s=lat[2])
    for i in range(n):
        if int(ents["kind"][i]) != A.HS_ENT_SERVER:
            continue
        e = ents[i]
        mean = float(e["d0"])
        if int(e["i2"]) == A.HS_SVC_EXPONENTIAL:
            dist = ExponentialLatency(mean) if stock_rng else PhiloxExponentialLatency(mean, i)
        else:
            dist = ConstantLatency(mean)
        cap = int(e["l0"])
        pol_cls = LIFOQueue if int(e["i1"]) == A.HS_Q_LIFO else FIFOQueue
        policy = pol_cls(capacity=cap) if cap >= 0 else pol_cls()
        objs[i] = Server(names[i], concurrency=int(e["i0"]), service_time=dist, queue_policy=policy)
    for i in range(n):          # server downstreams (may point at servers: tandem queues)
        if int(ents["kind"][i]) == A.HS_ENT_SERVER and int(ents["target"][i]) >= 0:
            objs[i].downstream = objs[int(ents["target"][i])]
    for i in range(n):
        if int(ents["kind"][i]) != A.HS_ENT_LB:
            continue
        e = ents[i]
        be = [objs[int(b)] for b in model.backends[int(e["i1"]): int(e["i1"]) + int(e["i2"])]]
        if int(e["i0"]) == A.HS_LB_KEY_TABLE:
            strat = ConsistentHash(virtual_nodes=chash_vnodes or 100)
        else:
            strat = RoundRobin()
        objs[i] = LoadBalancer(names[i], backends=be, strategy=strat)
    sources = []
    probes = []
    metric_names = {v: k for k, v in A.METRICS.items()}
    probe_data = {}          # PROBE row id -> Data
    for i in range(n):
        if int(ents["kind"][i]) != A.HS_ENT_SOURCE:
            continue
        e = ents[i]
        if int(ents["kind"][int(e["target"])]) == A.HS_ENT_PROBE:
            prow = ents[int(e["target"])]
            rate_ = float(model.profiles[int(e["i3"]) - 1]["p"][0])
            d_ = Data()
            pr_ = Probe(target=objs[int(prow["target"])], metric=metric_names[int(prow["i0"])], data=d_,
                        interval=1.0 / rate_)
            pr_._time_provider.profile.rate = rate_       # exactly the lowered rate (1/interval may not round-trip)
            objs[i] = pr_
            objs[int(e["target"])] = d_                    # stand-in object for the measurement row
            probe_data[int(e["target"])] = d_
            probes.append(pr_)
            continue
        target = objs[int(e["target"])]
        stop = Instant(int(e["l0"])) if int(e["l0"]) >= 0 else None
        pop = int(e["i1"])
        ctx_fn = None
        if pop > 0 and int(e["i2"]) > 0:
            # the reference's own ZipfDistribution (its cum_probs, its bisect), fed with the Philox uniforms
            from happysimulator.distributions.zipf import ZipfDistribution

            class _PhiloxUniform:
                def __init__(self, sid):
                    self.sid, self.n = sid, 0

                def random(self):
                    u = L.hs_cpu_uniform(seed, rid, A.HS_STREAM_ROUTING | (self.sid << 8), self.n)
                    self.n += 1
                    return u

            zd = ZipfDistribution(range(pop), s=float((zipf_s or {})[i]))
            zd._rng = _PhiloxUniform(i)

            def ctx_fn(time, count, _zd=zd):
                return {"created_at": time, "request_id": count, "metadata": {"client_id": _zd.sample()}}
        elif pop > 0:
            def ctx_fn(time, count, _sid=i, _pop=pop):
                u = L.hs_cpu_uniform(seed, rid, A.HS_STREAM_ROUTING | (_sid << 8), count - 1)
                return {"created_at": time, "request_id": count, "metadata": {"client_id": int(u * _pop)}}
        prov = SimpleEventProvider(target, "Request", stop, ctx_fn)
        profile = ConstantRateProfile(rate=float(e["d0"]))
        if int(e["i3"]) > 0:
            pr = model.profiles[int(e["i3"]) - 1]
            pp = [float(x) for x in pr["p"]]
            if int(pr["kind"]) == A.HS_PROF_STEP:
                # the user's own Profile object (the table in the model was tabulated from it): the reference runs
                # the ORIGINAL get_rate, the oracle and the device the table
                profile = profile_objects[i]
            else:
                profile = (LinearRampProfile(pp[0], pp[1], pp[2]) if int(pr["kind"]) == A.HS_PROF_LINEAR_RAMP
                           else SpikeProfile(pp[0], pp[1], pp[2], pp[3]))
        if int(e["i0"]) == A.HS_ARR_POISSON:
            atp = (PoissonArrivalTimeProvider(profile, Instant.Epoch) if stock_rng
                   else PhiloxPoissonArrival(profile, Instant.Epoch, i))
        else:
            atp = ConstantArrivalTimeProvider(profile, Instant.Epoch)
        objs[i] = Source(names[i], prov, atp)
        sources.append(objs[i])

    if stock_rng:
        import random
        random.seed(seed)
        np.random.seed(seed)

    entities = [o for i, o in enumerate(objs) if int(ents["kind"][i]) not in (A.HS_ENT_SOURCE, A.HS_ENT_PROBE, A.HS_ENT_REMOTE)]

    def attach(sim):
        # ---- object -> entity id, for the pop tap
        oid = {}
        for i, o in enumerate(objs):
            oid[id(o)] = i
            if int(ents["kind"][i]) in (A.HS_ENT_SERVER, A.HS_ENT_CACHE_SERVER):
                oid[id(o.queue)] = i
                oid[id(o.driver)] = i
                oid[id(o.worker)] = i

        recs = []
        heap = sim._event_heap
        orig_pop = heap.pop

        def classify(ev):
            t = ev.target
            if isinstance(ev, SourceEvent):
                return A.HS_EV_SOURCE_TICK
            if isinstance(ev, ProcessContinuation):
                return A.HS_EV_CONTINUATION
            et = ev.event_type
            if et == "probe_event":
                return A.HS_EV_PROBE
            if et == "QUEUE_NOTIFY":
                return A.HS_EV_NOTIFY
            if et == "QUEUE_POLL":
                return A.HS_EV_POLL
            if et == "QUEUE_DELIVER":
                return A.HS_EV_DELIVER
            if et == "_lb_response":
                return A.HS_EV_LB_RESPONSE
            if isinstance(t, _QueuedResourceWorkerAdapter):
                return A.HS_EV_REQ_WORKER
            if isinstance(t, Server) or (CachingServer is not None and isinstance(t, CachingServer)):
                return A.HS_EV_REQ_ENQUEUE
            if isinstance(t, Sink):
                return A.HS_EV_REQ_SINK
            if isinstance(t, Counter):
                return A.HS_EV_REQ_COUNTER
            if isinstance(t, LoadBalancer):
                return A.HS_EV_REQ_LB
            if isinstance(t, (SketchCollector, TopKCollector, QuantileEstimator)):
                return A.HS_EV_REQ_SKETCH
            raise AssertionError(f"unclassified event {ev!r}")

        data_to_row = {id(d): row for row, d in probe_data.items()}

        def target_id(ev):
            if ev.event_type == "probe_event":         # Event.once -> CallbackEntity(fn=measure_callback)
                for cell in ev.target._fn.__closure__:
                    if id(cell.cell_contents) in data_to_row:
                        return data_to_row[id(cell.cell_contents)]
            return oid[id(ev.target)]

        def tap():
            ev = orig_pop()
            if ev.time < sim._clock.now:     # "time travel" (an event delivered across a partition link behind the
                return ev                    # clock): the loop skips it, uncounted (core/simulation.py:479-489)
            recs.append((ev.time.nanoseconds, ev._sort_index, classify(ev), target_id(ev)))
            return ev

        heap.pop = tap
        return recs

    def extract(sim, recs, summary):
        heap = sim._event_heap

        rec = np.zeros(len(recs), A.RECORD_DTYPE)
        if recs:
            arr = np.array(recs, dtype=np.int64)
            rec["time_ns"], rec["sort_index"], rec["kind"], rec["entity"] = arr[:, 0], arr[:, 1], arr[:, 2], arr[:, 3]
        h = 0xcbf29ce484222325
        for t, idx, kind, ent in recs:
            h = L.hs_cpu_hash_step(h, t, idx, kind, ent)

        summ = np.zeros(1, A.SUMMARY_DTYPE)
        summ["events_processed"] = summary.total_events_processed
        summ["final_time_ns"] = sim._current_time.nanoseconds
        summ["order_hash"] = h
        summ["heap_left"] = heap.size()
        stats = np.zeros(n, A.STATS_DTYPE)
        sink_samples = []
        per_server_service = {}
        for i, o in enumerate(objs):
            k = int(ents["kind"][i])
            if k == A.HS_ENT_PROBE:
                vals = [float(v) for _, v in o._samples]
                stats[i]["c0"] = len(vals)
                stats[i]["f0"] = sum(vals)
                stats[i]["f2"] = min(vals) if vals else np.inf
                stats[i]["f3"] = max(vals) if vals else -np.inf
                sink_samples.append((i, [int(round(t * 1e9)) for t, _ in o._samples], vals))
            elif k == A.HS_ENT_SOURCE:
                stats[i]["c0"] = o.generated_count
                stats[i]["c1"] = getattr(o._event_provider, "_generated", o.generated_count - (1 if False else 0))
            elif k == A.HS_ENT_SERVER:
                st = o.stats
                stats[i]["c0"], stats[i]["c1"] = o.stats_accepted, o.stats_dropped
                stats[i]["c2"], stats[i]["c3"], stats[i]["f0"] = st.requests_completed, st.requests_rejected, st.total_service_time
                per_server_service[i] = list(o._service_times)
            elif k == A.HS_ENT_CACHE_SERVER:
                stats[i]["c0"], stats[i]["c1"] = o.stats_accepted, o.stats_dropped
                stats[i]["c2"], stats[i]["c3"] = o.stats.requests_processed, o.stats.cache_misses
                stats[i]["f0"], stats[i]["f1"] = float(o.stats.cache_hits), float(o.cache_size)
            elif k == A.HS_ENT_SINK:
                stats[i]["c0"] = o.events_received
                s2 = 0.0
                for v in o.latencies_s:
                    s2 += v * v
                # exactly what Sink.average_latency() divides by n (common.py:46-50)
                stats[i]["f0"], stats[i]["f1"] = sum(o.latencies_s), s2
                stats[i]["f2"] = min(o.latencies_s) if o.latencies_s else np.inf
                stats[i]["f3"] = max(o.latencies_s) if o.latencies_s else -np.inf
                sink_samples.append((i, [t.nanoseconds for t in o.completion_times], list(o.latencies_s)))
            elif k == A.HS_ENT_COUNTER:
                stats[i]["c0"] = o.total
            elif k == A.HS_ENT_SKETCH:
                stats[i]["c0"] = o.events_processed
                stats[i]["c1"] = (o.total_count if isinstance(o, TopKCollector) else o.sample_count
                                  if isinstance(o, QuantileEstimator) else o.sketch.item_count)
            elif k == A.HS_ENT_LB:
                s = o.stats
                stats[i]["c0"], stats[i]["c1"], stats[i]["c2"] = s.requests_received, s.requests_forwarded, len(o._in_flight)
                stats[i]["c3"] = sum(1 for r in recs if r[2] == A.HS_EV_LB_RESPONSE and r[3] == i)

        # all sinks' samples merged in arrival order (= order of REQ_SINK records)
        cursors = {i: 0 for i, _, _ in sink_samples}
        by_id = {i: (ct, ls) for i, ct, ls in sink_samples}
        merged = []
        svc_cursor = {i: 0 for i in per_server_service}
        svc_merged = []
        for t, idx, kind, ent in recs:
            if kind in (A.HS_EV_REQ_SINK, A.HS_EV_PROBE):
                c = cursors[ent]
                merged.append((by_id[ent][0][c], by_id[ent][1][c]))
                cursors[ent] = c + 1
            elif kind == A.HS_EV_REQ_WORKER:
                c = svc_cursor.get(ent, 0)
                if ent in per_server_service and c < len(per_server_service[ent]):
                    svc_merged.append(per_server_service[ent][c])
                    svc_cursor[ent] = c + 1
        smp = np.zeros(len(merged), A.SAMPLE_DTYPE)
        if merged:
            smp["completion_ns"] = [m[0] for m in merged]
            smp["latency_s"] = [m[1] for m in merged]
        summ["n_sink_samples"] = len(merged)
        summ["n_service_samples"] = len(svc_merged)
        out = {
            "summaries": summ, "entity_stats": stats[None, :], "records": rec, "sink_samples": smp,
            "service_samples": np.array(svc_merged, dtype=np.float64), "objects": objs, "sim": sim,
            "summary": summary,
        }
        # SKETCH rows: the reference sketch objects' own state, laid out like hs_outputs.sketches
        per, _, total, _ = model.sketch_layout()
        if total:
            img = np.zeros(total, np.uint8)
            for i, o in enumerate(objs):
                if int(ents["kind"][i]) == A.HS_ENT_CACHE_SERVER:     # TTLEviction._insert_times, one slot per customer key
                    K = int(ents["i0"][i])
                    ins = np.zeros(K + 1, np.float64)
                    pol = o._eviction_policy
                    for key, t in (pol._insert_times.items() if pol is not None else ()):
                        cid = key.split(":", 1)[1]
                        ins[K if cid == "unknown" else int(cid)] = t
                    assert set(o._cache._cache) == set(pol._insert_times) if pol is not None else True
                    img[per[i]: per[i] + ins.size * 8] = ins.view(np.uint8)
                    continue
                if int(ents["kind"][i]) != A.HS_ENT_SKETCH:
                    continue
                algo = int(ents["i0"][i])
                if algo == A.HS_SK_HLL:
                    img[per[i]: per[i] + len(o.sketch._registers)] = np.array(o.sketch._registers, dtype=np.uint8)
                elif algo == A.HS_SK_BLOOM:
                    w = np.array(o.sketch._bits, dtype=np.uint64)
                    img[per[i]: per[i] + w.size * 8] = w.view(np.uint8)
                elif algo == A.HS_SK_TDIGEST:       # state BEFORE any query (quantile() would flush the buffer)
                    td = o._tdigest
                    cap, bsz = int(ents["i3"][i]), int(ents["i2"][i])
                    hdr = np.zeros(32, np.uint8)
                    hdr[0:8] = np.array([len(td._centroids), len(td._buffer)], dtype=np.uint32).view(np.uint8)
                    hdr[8:16] = np.array([td._total_count], dtype=np.int64).view(np.uint8)
                    hdr[16:32] = np.array([td._min_value or 0.0, td._max_value or 0.0], dtype=np.float64).view(np.uint8)
                    img[per[i]: per[i] + 32] = hdr
                    if td._centroids:
                        cen = np.zeros(len(td._centroids) * 2, np.float64)
                        cen[0::2] = [c.mean for c in td._centroids]
                        cen[1::2] = np.array([c.count for c in td._centroids], dtype=np.int64).view(np.float64)
                        img[per[i] + 32: per[i] + 32 + cen.size * 8] = cen.view(np.uint8)
                    if td._buffer:
                        bb = np.array(td._buffer, dtype=np.float64)
                        off = per[i] + 32 + cap * 16
                        img[off: off + bb.size * 8] = bb.view(np.uint8)
                elif algo == A.HS_SK_RESERVOIR:     # {items held, generator index, items seen}, mt[624], items
                    rs = o.sketch
                    st = rs._rng.getstate()[1]
                    if rs._total_count:             # a sampler that saw nothing never touched its generator on the device
                        img[per[i]: per[i] + 8] = np.array([len(rs._reservoir), st[624]], dtype=np.uint32).view(np.uint8)
                        img[per[i] + 8: per[i] + 16] = np.array([rs._total_count], dtype=np.int64).view(np.uint8)
                        img[per[i] + 16: per[i] + 2512] = np.array(st[:624], dtype=np.uint32).view(np.uint8)
                        it = np.array(rs._reservoir, dtype=np.int32)
                        img[per[i] + 2512: per[i] + 2512 + it.size * 4] = it.view(np.uint8)
                elif algo == A.HS_SK_TOPK:          # dict order = insertion order (topk.py:116-128)
                    cs = list(o._topk._counters.values())
                    hdr = np.array([len(cs), 0, 0, 0], dtype=np.uint32)
                    sl = np.array([[c.item, c.count, c.error] for c in cs], dtype=np.int32).ravel()
                    img[per[i]: per[i] + 16] = hdr.view(np.uint8)
                    img[per[i] + 16: per[i] + 16 + sl.size * 4] = sl.view(np.uint8)
                else:
                    c = np.array(o.sketch._counters, dtype=np.uint32).ravel()
                    img[per[i]: per[i] + c.size * 4] = c.view(np.uint8)
            out["sketches"] = img
            # the reference's own answers off those states (hyperloglog.py:167, count_min_sketch.py:189)
            ans = {}
            for i, o in enumerate(objs):
                if int(ents["kind"][i]) != A.HS_ENT_SKETCH:
                    continue
                algo = int(ents["i0"][i])
                if algo == A.HS_SK_HLL:
                    ans[i] = np.array([o.sketch.cardinality()], dtype=np.int64)
                elif algo == A.HS_SK_BLOOM:     # contains(k) for every key, then the bit count
                    ans[i] = np.array([int(o.sketch.contains(k)) for k in range(int(ents["l0"][i]))] + [o.sketch._bits_set],
                                      dtype=np.int64)
                elif algo == A.HS_SK_TDIGEST and o.sample_count == 0:
                    ans[i] = np.zeros(0, np.int64)      # quantile() of an empty digest raises (tdigest.py:206-207)
                elif algo == A.HS_SK_TDIGEST:   # percentiles, cdf at a few points, centroid count (as float64 bits)
                    qs = [0.0, 0.001, 0.01, 0.25, 0.5, 0.75, 0.9, 0.99, 0.999, 1.0]
                    vals = [o.quantile(q) for q in qs] + [o.cdf(v) for v in (0.0, 0.01, 0.05, 0.1, 0.3, 1.0, 5.0)] + \
                           [float(o._tdigest.centroid_count)]
                    ans[i] = np.array(vals, dtype=np.float64).view(np.int64)
                elif algo == A.HS_SK_RESERVOIR: # the sample, then what the sampler's generator hands out next
                    ans[i] = np.array(o.sketch.sample() + [o.sketch.item_count, o.sketch._rng.getrandbits(32)], dtype=np.int64)
                elif algo == A.HS_SK_TOPK:      # top(): (item, count, error) rows, then max_error and the threshold
                    ans[i] = np.array([v for fe in o.top() for v in (fe.item, fe.count, fe.error)] +
                                      [o.max_error(), o.guaranteed_threshold()], dtype=np.int64)
                else:
                    ans[i] = np.array([o.sketch.estimate(k) for k in range(int(ents["l0"][i]))], dtype=np.int64)
            out["sketch_answers"] = ans
        if max_records is not None:
            out["records"] = rec[:max_records]
        return out

    if _build_only:
        return {"objs": objs, "sources": sources, "entities": entities, "probes": probes, "attach": attach, "extract": extract}
    sim = Simulation(end_time=Instant(int(end_ns)), sources=sources, entities=entities, probes=probes or None)
    recs = attach(sim)
    summary = sim.run()
    return extract(sim, recs, summary)


def run_reference_linked(lm, *, seed, replica=0, end_ns, max_workers=None):
    """The reference's ParallelSimulation + WindowedCoordinator on a happysim_b200.linked.LinkedModel.

    Every partition's objects are built exactly as run_reference builds them (Philox replica word
    partition + replica * (n_partitions + 1)); REMOTE rows become the real objects of the other partitions.  The
    links carry a latency object with the ``sample()`` method coordinator.py:209 calls (the stock
    LatencyDistribution classes do not have one), drawing from the coordinator's LINK_LATENCY streams; the
    coordinator's own generator (packet loss, coordinator.py:71,204) is replaced by the LINK_LOSS stream.
    Returns (per-partition outputs as run_reference gives them, the ParallelSimulationSummary)."""
    _import_reference()
    import oracle_lib as O
    from happysim_b200 import _abi as A
    import happysimulator.parallel.coordinator as coord
    from happysimulator.core.temporal import Duration, Instant
    from happysimulator.parallel.link import PartitionLink
    from happysimulator.parallel.partition import SimulationPartition
    from happysimulator.parallel.simulation import ParallelSimulation

    L = O.lib()
    nP = lm.n_partitions
    stride = nP + 1
    ctxs = [run_reference(lm.models[q], seed=seed, rid=q + replica * stride, end_ns=end_ns, _build_only=True)
            for q in range(nP)]
    for q, ctx in enumerate(ctxs):                       # REMOTE rows -> the objects they stand for
        ents = lm.models[q].entities
        for i in range(lm.models[q].n_entities):
            if int(ents["kind"][i]) == A.HS_ENT_SERVER and int(ents["target"][i]) >= 0 and \
                    int(ents["kind"][int(ents["target"][i])]) == A.HS_ENT_REMOTE:
                rem = ents[int(ents["target"][i])]
                dest = lm.links[q][int(rem["i0"])].dest
                ctx["objs"][i].downstream = ctxs[dest]["objs"][int(rem["i1"])]
    crid = nP + replica * stride

    class _LinkLatency:                                  # what a user passes as PartitionLink.latency
        def __init__(self, kind, mean_s, stream):
            self.kind, self.mean_s, self.stream, self.n = kind, mean_s, stream, 0

        def sample(self):
            if self.kind == A.HS_SVC_EXPONENTIAL:        # ExponentialLatency.get_latency, exponential.py:36-43
                u = L.hs_cpu_uniform(seed, crid, A.HS_STREAM_LINK_LATENCY | (self.stream << 8), self.n)
                self.n += 1
                return Duration.from_seconds(L.hs_cpu_exp1(u) / (1.0 / self.mean_s))
            return Duration.from_seconds(self.mean_s)    # ConstantLatency.get_latency

    class _LossStream:
        def __init__(self, _seed=None):
            self.n = 0

        def random(self):
            u = L.hs_cpu_uniform(seed, crid, A.HS_STREAM_LINK_LOSS, self.n)
            self.n += 1
            return u

    class _RandomShim:
        Random = _LossStream

    lat_objs = {}
    links = []
    for q in range(nP):
        for l in lm.links[q]:
            lat = lat_objs.setdefault(l.stream, _LinkLatency(l.latency_kind, l.latency_mean_s, l.stream))
            assert (lat.kind, lat.mean_s) == (l.latency_kind, l.latency_mean_s), "links sharing a stream share the object"
            links.append(PartitionLink(source_partition=lm.names[q], dest_partition=lm.names[l.dest],
                                       min_latency=lm.window_s, latency=lat, packet_loss=l.packet_loss))
    # a stock Server schedules events for its hidden queue / driver / worker entities; the partition's router
    # (routing.py:40-61) only knows the entities the partition lists, so they have to be listed with it
    def with_hidden(objs):
        out = []
        for o in objs:
            out.append(o)
            out += [getattr(o, a) for a in ("queue", "driver", "worker") if hasattr(o, "_concurrency_model") and hasattr(o, a)]
        return out
    parts = [SimulationPartition(name=lm.names[q], entities=with_hidden(ctxs[q]["entities"]), sources=ctxs[q]["sources"],
                                 probes=ctxs[q]["probes"]) for q in range(nP)]
    saved = coord.random
    coord.random = _RandomShim
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ps = ParallelSimulation(parts, end_time=Instant(int(end_ns)), links=links, max_workers=max_workers)
        recs = [ctxs[q]["attach"](ps._simulations[lm.names[q]]) for q in range(nP)]
        sent = [dict() for _ in range(nP)]                # REMOTE row -> events routed to the outbox through it
        for q in range(nP):
            sim_q = ps._simulations[lm.names[q]]
            ents_q = lm.models[q].entities
            row_of = {id(ctxs[lm.links[q][int(ents_q["i0"][i])].dest]["objs"][int(ents_q["i1"][i])]): i
                      for i in range(lm.models[q].n_entities) if int(ents_q["kind"][i]) == A.HS_ENT_REMOTE}

            def counting(events, now, _orig=sim_q._event_router, _rows=row_of, _sent=sent[q]):
                for e in events:
                    if id(e.target) in _rows:
                        _sent[_rows[id(e.target)]] = _sent.get(_rows[id(e.target)], 0) + 1
                return _orig(events, now)
            sim_q._event_router = counting
        summary = ps.run()
    finally:
        coord.random = saved
    outs = [ctxs[q]["extract"](ps._simulations[lm.names[q]], recs[q], summary.partitions[lm.names[q]]) for q in range(nP)]
    for q in range(nP):
        for row, cnt in sent[q].items():
            outs[q]["entity_stats"][0][row]["c0"] = cnt
    return outs, summary


def ring_table_from_reference(names, vnodes, population):
    """key -> backend slot by asking the reference's ConsistentHash itself."""
    _import_reference()
    from happysimulator.components.load_balancer.strategies import ConsistentHash
    from happysimulator.core.entity import Entity

    class _B(Entity):
        def handle_event(self, event):
            return None

    bes = [_B(nm) for nm in names]
    ch = ConsistentHash(virtual_nodes=vnodes)
    for b in bes:
        ch.add_backend(b)
    ring = sorted(ch._ring)
    import bisect
    hashes = [h for h, _ in ring]
    slot = {nm: i for i, nm in enumerate(names)}
    tab = np.zeros(population, np.int32)
    for k in range(population):
        hv = ch._hash(str(k))
        j = bisect.bisect_left(hashes, hv)
        tab[k] = slot[ring[j][1]] if j < len(ring) else slot[ring[0][1]]
    return tab
