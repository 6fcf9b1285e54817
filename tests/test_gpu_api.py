"""The modelling API end to end on the GPU: models written like the reference's quick-start,
results read back off the entity objects, compared with the fixtures recorded from the
UNMODIFIED reference (tests/golden/philox_*.npz)."""
import numpy as np
import pytest

import golden_lib as G
import happysim_b200 as hs
from happysim_b200 import _abi as A

pytestmark = pytest.mark.gpu


def quickstart(seed=None, end_s=60, rate=8):
    sink = hs.Sink()
    server = hs.Server("Server", service_time=hs.ExponentialLatency(0.1), downstream=sink)
    source = hs.Source.poisson(rate=rate, target=server)
    sim = hs.Simulation(end_time=hs.Instant.from_seconds(end_s), sources=[source], entities=[server, sink], seed=seed)
    return sim, source, server, sink


def test_readme_quickstart_matches_the_reference_fixture():
    _, kw, z = G.load("philox_mm1_seed42")
    sim, source, server, sink = quickstart(seed=kw["seed"])
    summary = sim.run()
    ws, st = z["summaries"][0], z["entity_stats"][0]
    assert summary.total_events_processed == int(ws["events_processed"])
    assert summary.duration_s == float(int(ws["final_time_ns"])) / 1e9
    assert source.generated_count == int(st[0]["c0"])
    assert server.stats.requests_completed == int(st[1]["c2"])
    assert server.stats.total_service_time == float(st[1]["f0"])
    assert server.stats_accepted == int(st[1]["c0"]) and server.stats_dropped == 0
    assert sink.events_received == int(st[2]["c0"])
    assert sink.latencies_s == [float(x) for x in z["sink_samples"]["latency_s"]]
    assert [t.nanoseconds for t in sink.completion_times] == [int(t) for t in z["sink_samples"]["completion_ns"]]
    assert server._service_times == [float(x) for x in z["service_samples"]]
    assert sink.average_latency() == float(st[2]["f0"]) / int(st[2]["c0"])
    assert sink.average_latency() == sum(sink.latencies_s) / len(sink.latencies_s)   # CPython float sum()
    es = summary.entities["Server"]
    assert es.queue_stats.total_accepted == int(st[1]["c0"]) and summary.entities["Sink"].events_handled == sink.events_received
    p = sink.latency_stats()
    assert p["count"] == sink.events_received and p["min"] == float(st[2]["f2"]) and p["max"] == float(st[2]["f3"])


def test_load_balanced_farm_matches_the_reference_fixture():
    _, kw, z = G.load("philox_lb_rr8")
    sink = hs.Sink()
    servers = [hs.Server(f"S{i}", service_time=hs.ExponentialLatency(0.1)) for i in range(8)]
    for s in servers:
        s.downstream = sink
    lb = hs.LoadBalancer("LB", backends=servers, strategy=hs.RoundRobin())
    src = hs.Source.poisson(rate=64.0, target=lb)
    # entity ids of the fixture: source, S0..S7, sink, LB
    sim = hs.Simulation(end_time=hs.Instant(kw["end_ns"]), sources=[src], entities=[*servers, sink, lb],
                        seed=kw["seed"], replica=kw["rid_base"])
    summary = sim.run()
    st = z["entity_stats"][0]
    assert summary.total_events_processed == int(z["summaries"]["events_processed"][0])
    assert lb.stats.requests_received == int(st[10]["c0"]) and lb.stats.requests_forwarded == int(st[10]["c1"])
    assert [s.stats.requests_completed for s in servers] == [int(st[1 + i]["c2"]) for i in range(8)]
    assert [s.stats.total_service_time for s in servers] == [float(st[1 + i]["f0"]) for i in range(8)]
    assert sink.latencies_s == [float(x) for x in z["sink_samples"]["latency_s"]]
    # per-server service-time lists: total_service_time is their sequential sum (server.py:260)
    for s in servers:
        acc = 0.0
        for x in s._service_times[: s.stats.requests_completed]:
            acc += x
        assert acc == s.stats.total_service_time or len(s._service_times) != s.stats.requests_completed


def test_parallel_runner_replica_i_equals_single_run_with_seed_plus_i():
    def build():
        return quickstart(end_s=30)[0]
    res = hs.ParallelRunner().run_replicas(build, n_replicas=6, base_seed=42)
    assert [r.name for r in res] == [f"replica_{i}" for i in range(6)]
    for i in (0, 3, 5):
        sim = quickstart(seed=42 + i, end_s=30)[0]
        s = sim.run()
        assert res[i].summary.total_events_processed == s.total_events_processed
        assert res[i].summary.duration_s == s.duration_s
    assert len({r.summary.total_events_processed for r in res}) > 1


def test_run_ensemble_totals():
    sim = quickstart(seed=7, end_s=100)[0]
    out = sim.run_ensemble(2048, flags=0)
    t = hs.engine.totals_to_dict(out["totals"])
    assert t["replicas"] == 2048 and t["replicas_flagged"] == 0
    assert t["events_processed"] == int(out["summaries"]["events_processed"].sum())
    mean_lat = t["sum_latency"] / t["sink_events"]
    assert 0.40 < mean_lat < 0.60            # M/M/1, rho = 0.8: W = 1 / (mu - lambda) = 0.5 s


def test_overloaded_model_runs_like_the_reference_unbounded_queue():
    """rho = 500: the reference's queue is unbounded, so the run must complete; the device rings are sized
    from the backlog estimate and grown on overflow.  Checked against the oracle with a huge ring."""
    import oracle_lib as O
    from happysim_b200 import engine
    sim, source, server, sink = quickstart(seed=1, end_s=5, rate=5000)
    summary = sim.run()
    want = O.oracle_run(sim.model, O.make_params(seed=1, end_ns=5 * 10**9, n_replicas=1, flags=0, queue_ring=1 << 16))
    assert summary.total_events_processed == int(want["summaries"]["events_processed"][0])
    assert int(want["summaries"]["status"][0]) == 0 and sim.last_run_info["status"] == 0
    assert server.stats_accepted == int(want["entity_stats"][0][1]["c0"]) > 20000
    assert sink.events_received == int(want["entity_stats"][0][2]["c0"])
    # a deliberately tiny first ring is grown, not reported as an error
    sim2, _, server2, _ = quickstart(seed=1, end_s=5, rate=5000)
    sim2._queue_ring = 64
    assert sim2.run().total_events_processed == summary.total_events_processed
    assert sim2.last_run_info["launches"] > 1 and sim2.last_run_info["queue_ring"] >= 32768


def test_ensemble_status_is_checked_and_rings_grow():
    import oracle_lib as O
    sim = quickstart(seed=3, end_s=40, rate=30)[0]            # rho = 3: backlog ~ 800 requests per replica
    with pytest.raises(hs.api.EnsembleStatusError) as ei:
        sim.run_ensemble(64, flags=0, queue_ring=64, on_overflow="raise")
    assert (ei.value.status & A.HS_ST_QUEUE_OVERFLOW).all()
    out = sim.run_ensemble(64, flags=0, queue_ring=64)       # default: grow and re-run
    assert (out["status"] == 0).all() and out["queue_ring"] >= 1024
    want = O.oracle_run(sim.model, O.make_params(seed=3, end_ns=40 * 10**9, n_replicas=64, flags=0, queue_ring=1 << 14))
    assert out["summaries"].tobytes() == want["summaries"].tobytes()
    res = hs.ParallelRunner().run_replicas(lambda: quickstart(end_s=40, rate=30)[0], 16, base_seed=5)
    assert len(res) == 16 and all(r.status == 0 for r in res)
    assert res[-1].summary.total_events_processed == quickstart(seed=5 + 15, end_s=40, rate=30)[0].run().total_events_processed


def test_spike_profile_defaults_run_without_tuning():
    """SpikeProfile's own defaults (baseline 10, spike 150 for 15 s) against a 10 req/s server build a queue of
    ~2100: far beyond the default device ring, fine for the reference."""
    sink = hs.Sink()
    server = hs.Server("Server", service_time=hs.ExponentialLatency(0.1), downstream=sink)
    source = hs.Source.with_profile(hs.SpikeProfile(), target=server)
    sim = hs.Simulation(end_time=hs.Instant.from_seconds(120), sources=[source], entities=[server, sink], seed=11)
    summary = sim.run()
    assert sim.last_run_info["status"] == 0 and summary.total_events_processed > 10000
    assert server.stats_accepted - server.stats.requests_completed > 1500     # the backlog the spike left behind
    import oracle_lib as O
    want = O.oracle_run(sim.model, O.make_params(seed=11, end_ns=120 * 10**9, n_replicas=1, flags=0))
    assert summary.total_events_processed == int(want["summaries"]["events_processed"][0])


def test_tandem_event_budget_scales_with_depth():
    """12 servers in series cost ~86 events per request: the event-limit safety valve must scale with the
    topology (or be raised on retry) instead of failing a valid model."""
    sink = hs.Sink()
    servers = [hs.Server(f"S{i}", service_time=hs.ExponentialLatency(0.002)) for i in range(12)]
    for a, b in zip(servers, servers[1:]):
        a.downstream = b
    servers[-1].downstream = sink
    src = hs.Source.poisson(rate=100.0, target=servers[0])
    sim = hs.Simulation(end_time=hs.Instant.from_seconds(400), sources=[src], entities=[*servers, sink], seed=2)
    assert sim._events_per_request() >= 86
    summary = sim.run()
    assert summary.total_events_processed > 80 * sink.events_received > 3_000_000


def test_stock_trace_does_not_leak_into_later_ensembles():
    sim = quickstart(seed=42, end_s=20)[0]
    base = sim.run_ensemble(1, flags=0)["summaries"].copy()
    stock = hs.Simulation(end_time=hs.Instant.from_seconds(20), seed=42, rng="stock",
                          **dict(zip(("sources", "entities"), _qs_parts())))
    stock.run()
    again = sim.run_ensemble(1, flags=0)["summaries"]
    assert again.tobytes() == base.tobytes()
    assert sim.run_ensemble(5, flags=0)["summaries"]["events_processed"][0] == base["events_processed"][0]


def _qs_parts():
    sink = hs.Sink()
    server = hs.Server("Server", service_time=hs.ExponentialLatency(0.1), downstream=sink)
    return [hs.Source.poisson(rate=8, target=server)], [server, sink]


def test_run_sweep_is_one_launch_per_topology_and_equals_single_runs():
    grid = [(6.0, 0.1, 1), (8.0, 0.1, 1), (30.0, 0.1, 4), (9.0, 0.05, 2)]

    def build(rate, mean, c):
        def f():
            sink = hs.Sink()
            server = hs.Server("Server", concurrency=c, service_time=hs.ExponentialLatency(mean), downstream=sink)
            src = hs.Source.poisson(rate=rate, target=server)
            return hs.Simulation(end_time=hs.Instant.from_seconds(50), sources=[src], entities=[server, sink])
        return f
    cfgs = [hs.RunConfig(name=f"cfg{i}", build_fn=build(*g), seed=100 + i) for i, g in enumerate(grid)]
    # plus one configuration of another topology (a load-balanced pair)
    def build_lb():
        sink = hs.Sink()
        sv = [hs.Server(f"S{i}", service_time=hs.ExponentialLatency(0.1), downstream=sink) for i in range(2)]
        lb = hs.LoadBalancer("LB", backends=sv, strategy=hs.RoundRobin())
        return hs.Simulation(end_time=hs.Instant.from_seconds(50), sources=[hs.Source.poisson(rate=12.0, target=lb)],
                             entities=[*sv, sink, lb])
    cfgs.insert(2, hs.RunConfig(name="lb", build_fn=build_lb, seed=7))
    res = hs.ParallelRunner().run_sweep(cfgs)
    assert [r.name for r in res] == ["cfg0", "cfg1", "lb", "cfg2", "cfg3"]
    for cfg, r in zip(cfgs, res):
        one = cfg.build_fn()
        one._seed = cfg.seed
        s = one.run()
        assert r.summary.total_events_processed == s.total_events_processed, cfg.name
        assert r.summary.duration_s == s.duration_s and r.status == 0
        assert {k: v.events_handled for k, v in r.summary.entities.items()} == {k: v.events_handled for k, v in s.entities.items()}


def test_parallel_simulation_batches_partitions_of_one_topology():
    def part(name, rate):
        sink = hs.Sink(f"{name}.sink")
        server = hs.Server(f"{name}.srv", service_time=hs.ExponentialLatency(0.1), downstream=sink)
        src = hs.Source.poisson(rate=rate, target=server, name=f"{name}.src")
        return hs.SimulationPartition(name=name, entities=[server, sink], sources=[src]), sink
    parts = [part("a", 5.0), part("b", 8.0), part("c", 9.0)]
    ps = hs.ParallelSimulation([p for p, _ in parts], duration=60.0, seed=9)
    summ = ps.run()
    assert ps.launch_groups == [["a", "b", "c"]]
    for k, (p, sink) in enumerate(parts):
        sink2 = hs.Sink("x")
        server2 = hs.Server("y", service_time=hs.ExponentialLatency(0.1), downstream=sink2)
        src2 = hs.Source.poisson(rate=(5.0, 8.0, 9.0)[k], target=server2)
        one = hs.Simulation(duration=60.0, sources=[src2], entities=[server2, sink2], seed=9, replica=k)
        s = one.run()
        assert summ.partitions[p.name].total_events_processed == s.total_events_processed
        assert sink.latencies_s == sink2.latencies_s
    assert summ.total_events_processed == sum(s.total_events_processed for s in summ.partitions.values())


def test_run_ensemble_windows_equal_the_uncut_run():
    sim = quickstart(seed=13, end_s=90)[0]
    whole = sim.run_ensemble(256, flags=0)
    sim.run_ensemble(256, flags=0, window_end_s=20.0)
    sim.run_ensemble(256, flags=0, window_end_s=55.5, resume=True, upload=False)
    cut = sim.run_ensemble(256, flags=0, window_end_s=None, resume=True, upload=False)
    assert cut["summaries"].tobytes() == whole["summaries"].tobytes()
    assert cut["entity_stats"].tobytes() == whole["entity_stats"].tobytes()


def test_latency_tracker_and_throughput_tracker_collect_like_the_reference_sink():
    """instrumentation/collectors.py: LatencyTracker stores (completion_time_s, latency_s); the values
    must be the ones the reference's Sink recorded for the same run (fixture)."""
    _, kw, z = G.load("philox_mm1_seed42")
    lt = hs.LatencyTracker()
    server = hs.Server("Server", service_time=hs.ExponentialLatency(0.1), downstream=lt)
    sim = hs.Simulation(end_time=hs.Instant(kw["end_ns"]), sources=[hs.Source.poisson(rate=8, target=server)],
                        entities=[server, lt], seed=kw["seed"])
    summary = sim.run()
    assert summary.total_events_processed == int(z["summaries"]["events_processed"][0])
    assert lt.count == int(z["entity_stats"][0][2]["c0"])
    assert lt.data.raw_values() == [float(x) for x in z["sink_samples"]["latency_s"]]
    assert lt.data.times() == [float(int(t)) / 1e9 for t in z["sink_samples"]["completion_ns"]]
    assert lt.mean_latency() == sum(lt.data.raw_values()) / lt.count
    assert summary.entities["LatencyTracker"].events_handled == lt.count          # simulation.py:579 ("count")
    b = lt.summary(window_s=10.0)
    assert sum(b.counts()) == lt.count and len(b) == len({int(t // 10.0) for t in lt.data.times()})
    tt = hs.ThroughputTracker()
    server2 = hs.Server("Server", service_time=hs.ExponentialLatency(0.1), downstream=tt)
    hs.Simulation(end_time=hs.Instant(kw["end_ns"]), sources=[hs.Source.poisson(rate=8, target=server2)],
                  entities=[server2, tt], seed=kw["seed"]).run()
    assert tt.count == lt.count and tt.data.times() == lt.data.times() and set(tt.data.raw_values()) == {1.0}


def test_parallel_simulation_independent_partitions():
    """parallel/simulation.py:170-195: without links each partition is its own Simulation."""
    def part(name, rate):
        sink = hs.Sink(f"{name}.sink")
        srv = hs.Server(f"{name}.srv", service_time=hs.ExponentialLatency(0.05), downstream=sink)
        return hs.SimulationPartition(name, entities=[srv, sink], sources=[hs.Source.poisson(rate=rate, target=srv)]), sink
    (pa, sa), (pb, sb) = part("a", 10.0), part("b", 15.0)
    ps = hs.ParallelSimulation([pa, pb], duration=40.0, seed=5)
    summ = ps.run()
    assert set(summ.partitions) == {"a", "b"} and summ.total_windows == 0
    assert summ.total_events_processed == sum(s.total_events_processed for s in summ.partitions.values())
    assert summ.duration_s == max(s.duration_s for s in summ.partitions.values())
    assert set(summ.entities) == {"a.srv", "a.sink", "b.srv", "b.sink"}
    # partition k == a plain Simulation with the same seed and replica word k
    (pa2, sa2), _ = part("a", 10.0), None
    alone = hs.Simulation(duration=40.0, sources=pa2.sources, entities=pa2.entities, seed=5, replica=0).run()
    assert alone.total_events_processed == summ.partitions["a"].total_events_processed
    assert sa2.latencies_s == sa.latencies_s and sb.events_received > sa.events_received
    with pytest.raises(hs.UnsupportedModelError, match="no latency override"):      # linked runs: tests/test_gpu_linked.py
        hs.ParallelSimulation([pa, pb], duration=1.0, links=[hs.PartitionLink("a", "b", min_latency=0.1)])
    with pytest.raises(ValueError, match="min_latency must be > 0"):
        hs.PartitionLink("a", "b", min_latency=0.0)


@pytest.mark.parametrize("name", G.case_names("stock_mm1"))
def test_stock_seeded_reference_run_is_reproduced_bit_for_bit(name):
    """random.seed(s); numpy.random.seed(s) on the UNMODIFIED reference (no plug-ins at all) vs the
    device fed with those two MT19937 streams (rng="stock"): the README quick-start known answers."""
    _, kw, z = G.load(name)
    sim, source, server, sink = quickstart(seed=kw["seed"], end_s=kw["end_ns"] / 1e9)
    sim2 = hs.Simulation(end_time=hs.Instant(kw["end_ns"]), sources=sim._sources, entities=sim._entities,
                         seed=kw["seed"], rng="stock")
    summary = sim2.run()
    ws, st = z["summaries"][0], z["entity_stats"][0]
    assert summary.total_events_processed == int(ws["events_processed"])
    assert summary.duration_s == float(int(ws["final_time_ns"])) / 1e9
    assert sink.events_received == int(st[2]["c0"]) and source.generated_count == int(st[0]["c0"])
    want = [float(x) for x in z["sink_samples"]["latency_s"]]          # the fixture keeps the first 1500
    assert sink.latencies_s[: len(want)] == want and len(sink.latencies_s) == int(z["n_samples"])
    assert server.stats.total_service_time == float(st[1]["f0"])
    if name == "stock_mm1_seed42":
        assert summary.total_events_processed == 3621 and sink.average_latency() == 0.5696996189709543


def test_source_with_profile_matches_the_reference_fixture():
    """Source.with_profile(LinearRampProfile) -- arrival times from the adaptive
    Simpson + Brent path on the device -- against the fixture recorded from the reference."""
    _, kw, z = G.load("philox_ramp_poisson_mm1")
    sink = hs.Sink()
    server = hs.Server("Server", service_time=hs.ExponentialLatency(0.05), downstream=sink)
    src = hs.Source.with_profile(hs.LinearRampProfile(20.0, 2.0, 12.0), target=server)
    summary = hs.Simulation(end_time=hs.Instant(kw["end_ns"]), sources=[src], entities=[server, sink],
                            seed=kw["seed"], replica=kw["rid_base"]).run()
    assert summary.total_events_processed == int(z["summaries"]["events_processed"][0])
    want = [float(x) for x in z["sink_samples"]["latency_s"]]
    assert sink.latencies_s[: len(want)] == want


def test_probes_sample_queue_depth_like_the_reference():
    """Probe.on(server, "depth", 0.1) -- tick times come from the general arrival path
    (100000000, 200000000, 299999999, ... ns) and each tick samples the device-side queue depth."""
    _, kw, z = G.load("philox_probe_mm1")
    sink = hs.Sink()
    server = hs.Server("Server", service_time=hs.ExponentialLatency(0.1), downstream=sink)
    src = hs.Source.poisson(rate=8, target=server)
    p1, depth = hs.Probe.on(server, "depth", interval=0.1)
    p2, seen = hs.Probe.on(sink, "events_received", interval=0.5)
    summary = hs.Simulation(end_time=hs.Instant(kw["end_ns"]), sources=[src], entities=[server, sink],
                            probes=[p1, p2], seed=kw["seed"]).run()
    assert summary.total_events_processed == int(z["summaries"]["events_processed"][0])
    st = z["entity_stats"][0]
    assert depth.count() == int(st[5]["c0"]) and seen.count() == int(st[6]["c0"])
    assert depth.sum() == float(st[5]["f0"]) and depth.max() == float(st[5]["f3"])
    assert [round(t * 1e9) for t in depth.times()[:4]] == [100000000, 200000000, 299999999, 399999998]
    assert seen.raw_values() == sorted(seen.raw_values()) and seen.raw_values()[-1] <= sink.events_received
    assert p1.generated_count == int(st[1]["c0"])


@pytest.mark.parametrize("fixture,strategy", [("philox_cache_chash5", "chash"), ("philox_cache_rr5", "rr")])
def test_caching_server_farm_matches_the_reference_fixture(fixture, strategy):
    """examples/load-balancing/consistent_hashing_basics.py: consistent hashing keeps a customer on one server's TTL cache
    (hit rate ~80 %), round robin spreads it over all five (~40 %).  The fixture is the unmodified reference running
    the example's own CachingServer class; here the mirror classes run on the device."""
    _, kw, z = G.load(fixture)
    servers = [hs.CachingServer(f"Server_{i}", server_id=i, cache_capacity=100, cache_ttl_s=0.8) for i in range(5)]
    strat = hs.ConsistentHash(virtual_nodes=30) if strategy == "chash" else hs.RoundRobin()
    lb = hs.LoadBalancer("LB", backends=servers, strategy=strat)
    src = hs.Source.poisson(rate=200.0, event_provider=hs.SimpleEventProvider(lb, context_fn=hs.UniformKeyContext(40)))
    sim = hs.Simulation(end_time=hs.Instant(kw["end_ns"]), sources=[src], entities=[*servers, lb], seed=kw["seed"],
                        replica=kw["rid_base"])
    summary = sim.run()
    st = z["entity_stats"][0]
    assert summary.total_events_processed == int(z["summaries"]["events_processed"][0])
    for i, sv in enumerate(servers):
        row = st[1 + i]
        assert (sv.stats.requests_processed, sv.stats.cache_misses, sv.stats.cache_hits) == (int(row["c2"]), int(row["c3"]), int(row["f0"]))
        assert sv.cache_size == int(row["f1"]) and sv.stats_accepted == int(row["c0"])
    hit = sum(s.stats.cache_hits for s in servers) / max(1, sum(s.stats.cache_hits + s.stats.cache_misses for s in servers))
    assert (hit > 0.7) if strategy == "chash" else (hit < 0.5)
    assert lb.stats.requests_forwarded == int(st[6]["c1"])
