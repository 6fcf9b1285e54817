"""BASELINE.json's configurations at (or near) their full replica counts, checked through
size-independent properties -- the oracle cannot run 65 536 replicas in seconds, so:
cut-into-windows == uncut (order hash of every replica), a sample of replicas == oracle,
conservation laws, round-robin fairness, queueing-theory sanity, device totals == numpy."""
import numpy as np
import pytest

import happysim_b200 as hs
from happysim_b200 import engine, _abi as A, distributed as D
import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = engine.Engine(0)
    yield e
    e.close()


def test_config1_65536_mm1_replicas(eng):
    n, end = 65536, 200 * 10**9
    model = hs.mm1()
    eng.upload(model)
    kw = dict(seed=1234, n_replicas=n, end_ns=end)
    eng.run(engine.make_params(**kw))
    whole = eng.read_outputs()
    tot = eng.read_totals()
    # (1) windows: the same run cut at awkward points
    eng.run(engine.make_params(window_end_ns=70 * 10**9, **kw))
    for cut in (70 * 10**9 + 1, 150 * 10**9):
        eng.run(engine.make_params(window_end_ns=cut, resume=1, **kw))
    eng.run(engine.make_params(resume=1, **kw))
    cut = eng.read_outputs()
    assert whole["summaries"].tobytes() == cut["summaries"].tobytes()
    assert whole["entity_stats"].tobytes() == cut["entity_stats"].tobytes()
    # (2) a sample of replicas against the oracle (global replica ids)
    for r in (0, 1, 31, 32, 4095, 40000, 65535):
        w = O.oracle_run(model, O.make_params(seed=1234, end_ns=end, n_replicas=1, replica_index_base=r))
        assert whole["summaries"][r].tobytes() == w["summaries"][0].tobytes()
        assert whole["entity_stats"][r].tobytes() == w["entity_stats"][0].tobytes()
    # (3) conservation per replica
    s, st = whole["summaries"], whole["entity_stats"]
    src, srv, snk = st[:, 0], st[:, 1], st[:, 2]
    assert (s["status"] == 0).all() and (s["final_time_ns"] > end).all()
    assert (src["c0"] >= src["c1"]).all() and (src["c0"] - src["c1"] == 0).all()      # no stop_after
    assert (srv["c1"] == 0).all()                                                     # unbounded queue
    assert ((src["c1"] - srv["c0"] >= 0) & (src["c1"] - srv["c0"] <= 1)).all()       # at most one ENQUEUE pending
    assert ((srv["c0"] - srv["c2"]) >= 0).all()
    assert ((srv["c2"] - snk["c0"] >= 0) & (srv["c2"] - snk["c0"] <= 1)).all()
    assert (s["n_service_samples"] - srv["c2"] <= 1).all()
    # (4) queueing theory: rho = 0.8 -> W = 0.5 s, utilisation 0.8
    assert abs(float(snk["f0"].sum()) / float(snk["c0"].sum()) - 0.5) < 0.01
    assert abs(float(srv["f0"].sum()) / (n * 200.0) - 0.8) < 0.005
    # (5) the deterministic device reduction equals numpy
    want = D.totals_from_outputs(model, whole)
    assert list(tot.i) == list(want.i) and tot.fmin == want.fmin and tot.fmax == want.fmax
    assert np.allclose(list(tot.fsum), list(want.fsum), rtol=1e-12)


def test_config1_lane_and_warp_engines_agree_on_a_slice(eng):
    model = hs.mm1()
    eng.upload(model)
    kw = dict(seed=1234, n_replicas=512, end_ns=300 * 10**9, replica_index_base=12345)
    eng.run(engine.make_params(engine=2, **kw)); a = eng.read_outputs()
    eng.run(engine.make_params(engine=1, **kw)); b = eng.read_outputs()
    assert a["summaries"].tobytes() == b["summaries"].tobytes()
    assert a["entity_stats"].tobytes() == b["entity_stats"].tobytes()


def test_config2_round_robin_64_servers_16384_replicas(eng):
    n = 16384
    model = hs.lb_round_robin(64, 512.0)
    eng.upload(model)
    eng.run(engine.make_params(seed=7, n_replicas=n, end_ns=2 * 10**9))
    out = eng.read_outputs()
    s, st = out["summaries"], out["entity_stats"]
    assert (s["status"] == 0).all()
    src, servers, snk, lb = st[:, 0], st[:, 1:65], st[:, 65], st[:, 66]
    arrivals = servers["c0"] + servers["c1"]
    assert (arrivals.max(axis=1) - arrivals.min(axis=1) <= 1).all()          # RoundRobin.select fairness
    assert (lb["c1"] - arrivals.sum(axis=1) >= 0).all() and (lb["c1"] - arrivals.sum(axis=1) <= 1).all()
    assert (lb["c0"] == lb["c1"]).all() and (lb["c3"] + lb["c2"] == lb["c1"]).all()   # responses + in flight
    assert ((servers["c2"].sum(axis=1) - snk["c0"]) >= 0).all()
    for r in (0, 777, n - 1):
        w = O.oracle_run(model, O.make_params(seed=7, end_ns=2 * 10**9, n_replicas=1, replica_index_base=r))
        assert s[r].tobytes() == w["summaries"][0].tobytes()
        assert st[r].tobytes() == w["entity_stats"][0].tobytes()
    ev_per_req = float(s["events_processed"].sum()) / float(src["c1"].sum())
    assert 9.0 < ev_per_req < 11.0                                            # ~10.1 via LoadBalancer


def test_config3_consistent_hash_ring_1024_nodes(eng):
    names = [f"S{i}" for i in range(1024)]
    table = hs.consistent_hash_table(names, 100, 10000)
    assert len(set(table.tolist())) > 900                                     # nearly every node owns a key
    model = hs.lb_key_table(table, 1024, rate=8192.0)
    eng.upload(model)
    eng.run(engine.make_params(seed=3, n_replicas=96, end_ns=10**9 // 2))
    out = eng.read_outputs()
    assert (out["summaries"]["status"] == 0).all()
    w = O.oracle_run(model, O.make_params(seed=3, end_ns=10**9 // 2, n_replicas=2, replica_index_base=40))
    assert out["summaries"][40:42].tobytes() == w["summaries"].tobytes()
    assert out["entity_stats"][40:42].tobytes() == w["entity_stats"].tobytes()
    # stable routing: a key always lands on the same backend, so per-backend arrivals follow the table
    st = out["entity_stats"]
    share = np.bincount(table, minlength=1024) / len(table)
    got = (st[:, 1:1025]["c0"] + st[:, 1:1025]["c1"]).sum(axis=0)
    assert np.corrcoef(share, got / got.sum())[0, 1] > 0.9


def test_config4_mmc_sweep_cells(eng):
    cs, rhos = (1, 2, 4, 8, 16, 32), (0.5, 0.7, 0.9)
    model = hs.mmc_sweep(cs=cs, rhos=rhos)
    per_cell = 64
    n = len(cs) * len(rhos) * per_cell
    eng.upload(model)
    eng.run(engine.make_params(seed=11, n_replicas=n, replicas_per_cell=per_cell, end_ns=200 * 10**9))
    out = eng.read_outputs()
    st = out["entity_stats"].reshape(len(cs), len(rhos), per_cell, 3)
    assert (out["summaries"]["status"] == 0).all()
    for ci, c in enumerate(cs):
        for ri, rho in enumerate(rhos):
            srv = st[ci, ri, :, 1]
            util = float(srv["f0"].sum()) / (per_cell * 200.0 * c)
            assert abs(util - rho) < 0.03, (c, rho, util)
    lat = st[..., 2]["f0"].sum(axis=2) / st[..., 2]["c0"].sum(axis=2)
    assert (np.diff(lat, axis=1) > 0).all()                                   # latency grows with load
    w = O.oracle_run(model, O.make_params(seed=11, end_ns=200 * 10**9, n_replicas=1, replicas_per_cell=per_cell,
                                          replica_index_base=5 * per_cell + 3))
    assert out["summaries"][5 * per_cell + 3].tobytes() == w["summaries"][0].tobytes()


def test_latency_histograms_and_cell_totals(eng):
    """Device-side instrumentation for ensembles: per-replica 64-bin latency histograms (bit-exact vs the
    oracle on both engines) and the per-cell reduction that a sweep all-reduces (configs[4])."""
    model = hs.mmc_sweep(cs=(1, 3), rhos=(0.6, 0.9))
    per_cell, n = 50, 4 * 50
    kw = dict(seed=21, n_replicas=n, replicas_per_cell=per_cell, end_ns=100 * 10**9,
              flags=A.HS_RUN_ORDER_HASH | A.HS_RUN_HISTOGRAM)
    eng.upload(model)
    eng.run(engine.make_params(**kw))
    out = eng.read_outputs()
    want = O.oracle_run(model, O.make_params(**kw))
    assert out["histograms"].tobytes() == want["histograms"].tobytes()
    assert (out["histograms"].sum(axis=1) == out["entity_stats"][:, 2]["c0"]).all()
    cells = eng.read_cell_totals(4)
    ref = D.cell_totals_from_outputs(model, out, 4, per_cell)
    for (d, h), (t, hh) in zip(cells, ref):
        assert d["events_processed"] == t.i[0] and d["replicas"] == per_cell and d["sink_events"] == t.i[1]
        assert d["min_latency"] == t.fmin and d["max_latency"] == t.fmax
        assert np.allclose([d["sum_latency"], d["sum_latency_sq"], d["sum_service"]], list(t.fsum), rtol=1e-12)
        assert np.array_equal(h, hh)
    # the histogram is good enough for ensemble percentiles: compare with exact samples of one cell
    eng.run(engine.make_params(sample_cap=20000, **kw))
    smp = eng.read_outputs()
    lat = np.concatenate([A.unroll_ring(smp["sink_samples"][r], smp["summaries"]["n_sink_samples"][r], 20000)["latency_s"]
                          for r in range(per_cell)])
    for p in (0.5, 0.9, 0.99):
        est, exact = D.histogram_percentile(cells[0][1], p), float(np.quantile(lat, p))
        assert abs(est - exact) / exact < 0.25, (p, est, exact)
    # lane engine too
    m1 = hs.mm1()
    eng.upload(m1)
    k2 = dict(seed=3, n_replicas=96, end_ns=100 * 10**9, flags=A.HS_RUN_HISTOGRAM)
    eng.run(engine.make_params(**k2))
    assert eng.read_outputs()["histograms"].tobytes() == O.oracle_run(m1, O.make_params(**k2))["histograms"].tobytes()


def test_config3_full_size_one_gpus_share(eng):
    """BASELINE configs[3] as one GPU of the four sees it: 1 024 replicas of the 1 024-node ring (global replica ids of
    rank 2), 2 sim-s in two windows; a sample of replicas against the oracle, routing conservation on all of them."""
    n, base = 1024, 2 * 1024
    table = hs.consistent_hash_table([f"S{i}" for i in range(1024)], 100, 10000)
    model = hs.lb_key_table(table, 1024, rate=8192.0)
    end = 2 * 10**9
    kw = dict(seed=1234, n_replicas=n, replica_index_base=base, end_ns=end, flags=0)
    eng.upload(model)
    eng.run(engine.make_params(window_end_ns=end // 2, **kw))
    eng.run(engine.make_params(resume=1, **kw))
    out = eng.read_outputs()
    s, st = out["summaries"], out["entity_stats"]
    assert (s["status"] == 0).all() and (s["final_time_ns"] > end).all()
    for r in (0, 1, 511, 1023):
        w = O.oracle_run(model, O.make_params(seed=1234, end_ns=end, n_replicas=1, replica_index_base=base + r, flags=0))
        assert s[r].tobytes() == w["summaries"][0].tobytes(), r
        assert st[r].tobytes() == w["entity_stats"][0].tobytes(), r
    src, servers, snk, lb = st[:, 0], st[:, 1:1025], st[:, 1025], st[:, 1026]
    arrivals = (servers["c0"] + servers["c1"]).sum(axis=1)
    assert ((lb["c1"] - arrivals >= 0) & (lb["c1"] - arrivals <= 1)).all()          # forwarded == enqueued (+ one in flight)
    assert ((src["c1"] - lb["c0"] >= 0) & (src["c1"] - lb["c0"] <= 1)).all()
    assert ((servers["c2"].sum(axis=1) - snk["c0"]) >= 0).all()
    assert 1.5e4 < float(src["c1"].mean()) < 1.8e4                                   # 8 192 requests/s for 2 s


def test_config4_full_size_one_gpus_share(eng):
    """BASELINE configs[4] as one GPU of the eight sees it: 32 768 replicas = 256 (c, rho) cells x 128 seeds (global ids
    of rank 5), 100 sim-s; oracle sample, per-cell device reduction == numpy, utilisation per cell == rho."""
    model = hs.mmc_sweep()
    n, per_cell, base = 32768, 128, 5 * 32768
    end = 100 * 10**9
    kw = dict(seed=1234, n_replicas=n, replica_index_base=base, replicas_per_cell=per_cell, end_ns=end,
              flags=A.HS_RUN_HISTOGRAM, queue_ring=4096)
    eng.upload(model)
    eng.run(engine.make_params(**kw))
    out = eng.read_outputs()
    assert (out["summaries"]["status"] == 0).all()
    for r in (0, 127, 128, 7 * 128 + 5, 255 * 128 + 127):          # incl. the slowest cell (c = 1, rho = 0.99) and the last one
        w = O.oracle_run(model, O.make_params(**dict(kw, n_replicas=1, replica_index_base=base + r)))
        assert out["summaries"][r].tobytes() == w["summaries"][0].tobytes(), r
        assert out["entity_stats"][r].tobytes() == w["entity_stats"][0].tobytes(), r
        assert out["histograms"][r].tobytes() == w["histograms"][0].tobytes(), r
    cells = eng.read_cell_totals(256)
    ref = D.cell_totals_from_outputs(model, out, 256, per_cell, index_base=base)
    for c, ((d, h), (t, hh)) in enumerate(zip(cells, ref)):
        assert d["events_processed"] == t.i[0] and d["replicas"] == per_cell and d["sink_events"] == t.i[1], c
        assert d["min_latency"] == t.fmin and d["max_latency"] == t.fmax
        assert np.allclose([d["sum_latency"], d["sum_latency_sq"], d["sum_service"]], list(t.fsum), rtol=1e-12)
        assert np.array_equal(h, hh)
        cc, rho = model.cells[c]
        util = d["sum_service"] / (per_cell * 100.0 * cc)
        assert abs(util - rho) < 0.05, (cc, rho, util)
