"""SURVEY 8(f) row 3 on the device: SKETCH rows (HyperLogLog / CountMinSketch collectors) on both general
engines against the oracle and the reference fixtures, the device-side merge, windows, and the API mirror."""
import numpy as np
import pytest

import golden_lib as G
import happysim_b200 as hs
import oracle_lib as O
from happysim_b200 import _abi as A, distributed as D, engine
from test_gpu_lane_parity import assert_same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = engine.Engine(0)
    yield e
    e.close()


def big_farm(K=5000, n_servers=16):
    b = hs.ModelBuilder()
    src = b.source(rate=40.0 * n_servers, key_population=K)
    servers = [b.server(f"S{i}", mean_service_s=0.02) for i in range(n_servers)]
    hll = b.sketch_hll("uniques", precision=12, table=hs.hll_table(12, 1, K))
    cms = b.sketch_cms("freq", width=272, depth=5, table=hs.cms_table(272, 5, 2, K))
    lb = b.load_balancer(backends=servers)
    b.set_target(src, lb)
    for k, sv in enumerate(servers):
        b.set_target(sv, hll if k % 2 else cms)
    return b.build()


@pytest.mark.parametrize("eng_id", [1, 3])
def test_sketch_ensemble_matches_oracle_and_merges_on_the_device(eng, eng_id):
    model = big_farm()
    kw = dict(seed=8, end_ns=3 * 10**9, n_replicas=37, record_cap=40000, sample_cap=16, service_cap=4000)
    eng.upload(model)
    eng.run(engine.make_params(engine=eng_id, **kw))
    got = eng.read_outputs()
    want = O.oracle_run(model, O.make_params(**kw))
    assert_same(got, want)
    assert got["sketches"].tobytes() == want["sketches"].tobytes()
    assert int(want["entity_stats"][0][model.ids_of(A.HS_ENT_SKETCH)[0]]["c1"]) > 200
    merged = eng.read_sketches()
    host = D.merge_sketch_states(model, want["sketches"])
    for i in host:
        assert merged[i].dtype == host[i].dtype and np.array_equal(merged[i], host[i])


def members_farm(K=3000, n_servers=9):
    b = hs.ModelBuilder()
    src = b.source(rate=60.0 * n_servers, key_population=K)
    servers = [b.server(f"S{i}", concurrency=2, mean_service_s=0.02) for i in range(n_servers)]
    sinks = [b.sketch_bloom("seen", size_bits=9585, num_hashes=7, table=hs.bloom_table(9585, 7, 4, K)),
             b.sketch_topk("heavy", k=16, key_population=K),
             b.sketch_topk("all", k=64, key_population=K)]
    lb = b.load_balancer(backends=servers)
    b.set_target(src, lb)
    for k, sv in enumerate(servers):
        b.set_target(sv, sinks[k % 3])
    return b.build()


@pytest.mark.parametrize("eng_id", [1, 3])
def test_bloom_and_topk_ensemble_matches_oracle(eng, eng_id):
    """Space-Saving's eviction order (first minimum in dict order) and the Bloom bit arrays, bit for bit;
    the Bloom OR-merge on the device, the sequential TopK merge on the host."""
    model = members_farm()
    kw = dict(seed=21, end_ns=2 * 10**9, n_replicas=29, record_cap=30000, sample_cap=16, service_cap=3000)
    eng.upload(model)
    eng.run(engine.make_params(engine=eng_id, **kw))
    got = eng.read_outputs()
    want = O.oracle_run(model, O.make_params(**kw))
    assert_same(got, want)
    assert got["sketches"].tobytes() == want["sketches"].tobytes()
    views = model.sketch_views(got["sketches"])
    ids = model.ids_of(A.HS_ENT_SKETCH)
    assert int(views[ids[1]][:, 0].min()) == 16            # k = 16 counters all in use: evictions happened
    merged = eng.read_sketches()
    host = D.merge_sketch_states(model, want["sketches"])
    assert set(merged) == {ids[0]} and np.array_equal(merged[ids[0]], host[ids[0]])
    dev_top = D.merge_sketch_states(model, got["sketches"])[ids[1]]
    assert dev_top.top() == host[ids[1]].top() and dev_top.item_count == host[ids[1]].item_count


@pytest.mark.parametrize("eng_id", [1, 3])
def test_tdigest_ensemble_matches_oracle_bit_for_bit(eng, eng_id):
    """QuantileEstimator rows: IEEE add/mul/div/sqrt only, so the centroids (float means) are bit-identical to
    the oracle's -- and through the fixture to the reference's -- after many flush/compress rounds."""
    b = hs.ModelBuilder()
    src = b.source(rate=800.0)
    servers = [b.server(f"S{i}", concurrency=2, mean_service_s=0.004) for i in range(4)]
    q = [b.sketch_tdigest("lat50", compression=50.0), b.sketch_tdigest("lat20", compression=20.0)]
    lb = b.load_balancer(backends=servers)
    b.set_target(src, lb)
    for k, sv in enumerate(servers):
        b.set_target(sv, q[k % 2])
    model = b.build()
    kw = dict(seed=33, end_ns=4 * 10**9, n_replicas=21, record_cap=40000, sample_cap=16, service_cap=4000)
    eng.upload(model)
    eng.run(engine.make_params(engine=eng_id, **kw))
    got = eng.read_outputs()
    want = O.oracle_run(model, O.make_params(**kw))
    assert_same(got, want)
    assert got["sketches"].tobytes() == want["sketches"].tobytes()
    d = hs.TDigest(50.0); d._load_device_state(model.sketch_views(got["sketches"])[q[0]][5])
    assert d.item_count > 1200 and d.centroid_count < 60 and 0 < d.quantile(0.5) < d.quantile(0.99) <= d.max
    assert eng.read_sketches() == {}                                     # host-merged rows only
    merged = D.merge_sketch_states(model, got["sketches"])[q[0]]
    assert merged.item_count == int(got["entity_stats"][:, q[0]]["c1"].sum())


@pytest.mark.parametrize("eng_id", [1, 3])
def test_zipf_keys_heavy_hitters(eng, eng_id):
    """client_id ~ Zipf(1.2) (bisect over the host's cumulative table) -> consistent-hash ring -> TopK + CMS:
    the device draws the reference's keys, so the heavy hitters and their counts match the oracle exactly."""
    K, n_servers = 2000, 12
    tab = hs.consistent_hash_table([f"S{i}" for i in range(n_servers)], 50, K)
    b = hs.ModelBuilder()
    # the hottest key alone is ~19 % of the traffic and lands on one server: keep that server below capacity
    src = b.source(rate=25.0 * n_servers, key_population=K, key_cdf=hs.zipf_cdf(K, 1.2))
    servers = [b.server(f"S{i}", concurrency=2, mean_service_s=0.02) for i in range(n_servers)]
    top = b.sketch_topk("heavy", k=10, key_population=K)
    cms = b.sketch_cms("freq", width=64, depth=4, table=hs.cms_table(64, 4, 1, K))
    lb = b.load_balancer(backends=servers, key_table=tab)
    b.set_target(src, lb)
    for k, sv in enumerate(servers):
        b.set_target(sv, top if k % 2 else cms)
    model = b.build()
    kw = dict(seed=41, end_ns=3 * 10**9, n_replicas=19, record_cap=40000, sample_cap=16, service_cap=4000,
              queue_ring=512)
    eng.upload(model)
    eng.run(engine.make_params(engine=eng_id, **kw))
    got = eng.read_outputs()
    want = O.oracle_run(model, O.make_params(**kw))
    assert int(got["summaries"]["status"].max()) == 0
    assert_same(got, want)
    assert got["sketches"].tobytes() == want["sketches"].tobytes()
    t = hs.TopK(10); t._load_device_state(model.sketch_views(got["sketches"])[top][0], int(got["entity_stats"][0][top]["c1"]))
    hot = [fe.item for fe in t.top(3)]
    assert min(hot) < 20 and t.item_count > 150          # the hottest ranks dominate


@pytest.mark.parametrize("eng_id", [1, 3])
def test_sketch_rows_hashing_on_the_device(eng, eng_id):
    """K = 0 rows: no per-key tables, SHA-256 of (seed, repr(key)) per event in the kernel -- 200 000 client ids
    (Zipf) into HyperLogLog, Count-Min and Bloom collectors; states equal the oracle's, whose hash functions
    tests/test_sketching.py ties to hashlib and, through the fixtures, to the reference."""
    K = 200_000
    b = hs.ModelBuilder()
    src = b.source(rate=900.0, key_population=K, key_cdf=hs.zipf_cdf(K, 0.9))
    servers = [b.server(f"S{i}", concurrency=2, mean_service_s=0.004) for i in range(6)]
    sinks = [b.sketch_hll("uniques", precision=10, seed=5), b.sketch_cms("freq", width=64, depth=4, seed=2**40 + 1),
             b.sketch_bloom("seen", size_bits=4099, num_hashes=5, seed=None)]
    lb = b.load_balancer(backends=servers)
    b.set_target(src, lb)
    for k, sv in enumerate(servers):
        b.set_target(sv, sinks[k % 3])
    model = b.build()
    assert model.sketch_tables.size == 2 + 8 + 2
    kw = dict(seed=51, end_ns=2 * 10**9, n_replicas=17, record_cap=40000, sample_cap=16, service_cap=4000)
    eng.upload(model)
    eng.run(engine.make_params(engine=eng_id, **kw))
    got = eng.read_outputs()
    want = O.oracle_run(model, O.make_params(**kw))
    assert_same(got, want)
    assert got["sketches"].tobytes() == want["sketches"].tobytes()
    h = hs.HyperLogLog(10, seed=5); h._load_device_state(model.sketch_views(got["sketches"])[sinks[0]][0], 1)
    assert 200 < h.cardinality() < 700            # ~600 requests reach the HLL collector, most ids distinct
    merged = eng.read_sketches()
    host = D.merge_sketch_states(model, want["sketches"])
    assert all(np.array_equal(merged[i], host[i]) for i in host)


@pytest.mark.parametrize("eng_id", [1, 3])
def test_sketch_state_survives_windows(eng, eng_id):
    model, kw, z = G.load("philox_sketch_cms_farm")
    caps = dict(G.caps(z), engine=eng_id, n_replicas=3, rid_base=0, seed=kw["seed"])
    eng.upload(model)
    eng.run(engine.make_params(end_ns=kw["end_ns"], **caps)); whole = eng.read_outputs()
    eng.run(engine.make_params(end_ns=kw["end_ns"], window_end_ns=10**9, **caps))
    for cut in (2 * 10**9 + 7, 4 * 10**9):
        eng.run(engine.make_params(end_ns=kw["end_ns"], window_end_ns=cut, resume=1, **caps))
    eng.run(engine.make_params(end_ns=kw["end_ns"], resume=1, **caps))
    parts = eng.read_outputs()
    assert_same(parts, whole)
    assert parts["sketches"].tobytes() == whole["sketches"].tobytes()
    G.check_against(z, whole, r=kw["rid_base"])


def test_api_mirror_writes_the_device_state_back():
    K = 400
    hll = hs.HyperLogLog(precision=10, seed=6)
    cms = hs.CountMinSketch.from_error_rate(0.05, 0.05, seed=6)
    uniq = hs.SketchCollector("uniques", hll, hs.KeyExtractor())
    freq = hs.SketchCollector("freq", cms)
    s1 = hs.Server("A", concurrency=2, service_time=hs.ExponentialLatency(0.01), downstream=uniq)
    s2 = hs.Server("B", concurrency=2, service_time=hs.ExponentialLatency(0.01), downstream=freq)
    lb = hs.LoadBalancer("lb", backends=[s1, s2], strategy=hs.RoundRobin())
    src = hs.Source.poisson(rate=300.0, event_provider=hs.SimpleEventProvider(lb, context_fn=hs.UniformKeyContext(K)))
    sim = hs.Simulation(end_time=hs.Instant.from_seconds(4.0), sources=[src], entities=[lb, s1, s2, uniq, freq], seed=12)
    sim.run()
    model = sim.model
    want = O.oracle_run(model, O.make_params(seed=12, end_ns=4 * 10**9, n_replicas=1))
    views = model.sketch_views(want["sketches"])
    iu, ifr = sim.objects.index(uniq), sim.objects.index(freq)
    assert np.array_equal(hll._registers, views[iu][0]) and np.array_equal(cms._counters, views[ifr][0].astype(np.uint64))
    assert uniq.events_processed == int(want["entity_stats"][0][iu]["c0"]) > 400
    assert hll.item_count == uniq.events_processed and cms.item_count == freq.events_processed
    # ~300 of 400 ids seen after ~600 draws; standard error of p = 10 is 3.25 %
    assert abs(hll.cardinality() - len(np.unique(np.nonzero(views[iu][0])[0]))) >= 0            # sanity: runs
    assert 200 < hll.cardinality() < 400
    assert cms.estimate(0) <= freq.events_processed


@pytest.mark.parametrize("eng_id", [1, 3])
def test_reservoir_ensemble_matches_oracle(eng, eng_id):
    """ReservoirSampler rows: every replica runs its own copy of the sampler's MT19937 (624-word refills, rejected
    draws of randint); sample, count and generator state equal the oracle's byte for byte -- and through the
    fixture philox_sketch_reservoir the reference's.  The merge over replicas is the class's own, on the host."""
    K = 2000
    b = hs.ModelBuilder()
    src = b.source(rate=900.0, key_population=K)
    servers = [b.server(f"S{i}", concurrency=2, mean_service_s=0.004) for i in range(4)]
    rs = [b.sketch_reservoir("r8", size=8, seed=3, key_population=K), b.sketch_reservoir("r300", size=300, seed=4, key_population=K)]
    lb = b.load_balancer(backends=servers)
    b.set_target(src, lb)
    for k, sv in enumerate(servers):
        b.set_target(sv, rs[k % 2])
    model = b.build()
    kw = dict(seed=51, end_ns=4 * 10**9, n_replicas=23, record_cap=40000, sample_cap=16, service_cap=4000)
    eng.upload(model)
    eng.run(engine.make_params(engine=eng_id, **kw))
    got = eng.read_outputs()
    want = O.oracle_run(model, O.make_params(**kw))
    assert_same(got, want)
    assert got["sketches"].tobytes() == want["sketches"].tobytes()
    views = model.sketch_views(got["sketches"])
    assert int(views[rs[0]][:, 2].min()) > 1500 and (views[rs[0]][:, 0] == 8).all() and (views[rs[1]][:, 0] == 300).all()
    assert len({tuple(v[3 + 624:]) for v in views[rs[0]]}) == 23            # different key streams, different samples
    assert eng.read_sketches() == {}                                        # host-merged rows only
    merged = D.merge_sketch_states(model, got["sketches"])[rs[0]]
    assert merged.item_count == int(got["entity_stats"][:, rs[0]]["c1"].sum()) and len(merged) == 8
    pool = {int(x) for v in views[rs[0]] for x in v[3 + 624:]}
    assert set(merged.sample()) <= pool


def test_reservoir_fixture_runs_in_windows(eng):
    model, kw, z = G.load("philox_sketch_reservoir")
    caps = dict(G.caps(z), n_replicas=1, seed=kw["seed"], rid_base=kw["rid_base"])
    eng.upload(model)
    eng.run(engine.make_params(end_ns=kw["end_ns"], window_end_ns=3 * 10**9, **caps))
    eng.run(engine.make_params(end_ns=kw["end_ns"], window_end_ns=11 * 10**9 + 1, resume=1, **caps))
    eng.run(engine.make_params(end_ns=kw["end_ns"], resume=1, **caps))
    G.check_against(z, eng.read_outputs())


def test_api_mirror_reservoir_sampler_continues_its_generator():
    import random
    K = 700
    sampler = hs.ReservoirSampler(10, seed=77)
    coll = hs.SketchCollector("sample", sampler)
    src = hs.Source.poisson(rate=500.0, event_provider=hs.SimpleEventProvider(coll, context_fn=hs.UniformKeyContext(K)))
    sim = hs.Simulation(end_time=hs.Instant.from_seconds(3.0), sources=[src], entities=[coll], seed=5)
    sim.run()
    n = coll.events_processed
    twin = hs.ReservoirSampler(10, seed=77)
    L = O.lib()
    for j in range(n):
        twin.add(int(L.hs_cpu_uniform(5, 0, A.HS_STREAM_ROUTING | (sim.objects.index(src) << 8), j) * K))
    assert n > 1200 and sampler.item_count == n and sampler.sample() == twin.sample()
    assert sampler._rng.getstate() == twin._rng.getstate() != random.Random(77).getstate()
