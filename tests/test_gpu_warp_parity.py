"""GPU parity of the two general engines (1 = warp per replica, 3 = thread per replica; they share
the handlers in csrc/hs_handlers.inc) against the CPU oracle: M/M/c,
load-balanced server farms (round robin and consistent-hash key table), tandem
queues, multiple sources, sweeps with per-cell parameters -- bit-exact event
sequence, statistics and samples, through the C-ABI."""
import numpy as np
import pytest

import happysim_b200 as hs
from happysim_b200 import engine
import oracle_lib as O
from test_gpu_lane_parity import assert_same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = engine.Engine(0)
    yield e
    e.close()


def both(eng, model, **kw):
    eng.upload(model)
    eng.run(engine.make_params(**kw))
    got = eng.read_outputs()
    want = O.oracle_run(model, O.make_params(**kw))
    return got, want


def tandem():
    b = hs.ModelBuilder()
    src = b.source(rate=6.0)
    s1 = b.server("A", mean_service_s=0.08, capacity=3)
    s2 = b.server("B", concurrency=2, mean_service_s=0.2, lifo=True)
    c = b.counter()
    b.set_target(src, s1); b.set_target(s1, s2); b.set_target(s2, c)
    return b.build()


def two_sources():
    b = hs.ModelBuilder()
    a = b.source("A", rate=3.0)
    c = b.source("B", rate=4.0, poisson=False, stop_after_ns=20 * 10**9)
    srv = b.server(concurrency=2, mean_service_s=0.1)
    snk = b.sink()
    b.set_target(a, srv); b.set_target(c, srv); b.set_target(srv, snk)
    return b.build()


def chash(n_servers=16, pop=300):
    rng = np.random.RandomState(3)
    return hs.lb_key_table(rng.randint(0, n_servers, size=pop).astype(np.int32), n_servers, rate=8.0 * n_servers)


CASES = {
    "mm1_on_warp_engine": (lambda: hs.mm1(), 60, 40),
    "mmc4": (lambda: hs.mm1(rate=32, concurrency=4), 30, 40),
    "mmc32": (lambda: hs.mm1(rate=256, concurrency=32), 6, 24),
    "lb_rr8": (lambda: hs.lb_round_robin(8, 64.0), 10, 40),
    "lb_rr64": (lambda: hs.lb_round_robin(64, 512.0), 4, 24),
    "lb_chash16": (chash, 6, 24),
    "tandem": (tandem, 60, 40),
    "two_sources": (two_sources, 40, 40),
    "source_to_sink_only": (lambda: _src_sink(), 30, 8),
    "zero_gap_poisson_c2": (lambda: hs.mm1(rate=3e8, mean_service_s=4e-9, concurrency=2), 2e-5, 8),
}


def _src_sink():
    b = hs.ModelBuilder()
    s = b.source(rate=5.0)
    k = b.sink()
    b.set_target(s, k)
    return b.build()


@pytest.mark.parametrize("eng_id", [1, 3])
@pytest.mark.parametrize("name", sorted(CASES))
def test_warp_matches_oracle(eng, name, eng_id):
    mk, end_s, n = CASES[name]
    model = mk()
    got, want = both(eng, model, seed=99, end_ns=int(end_s * 1e9), n_replicas=n, record_cap=30000,
                     sample_cap=3000, service_cap=3000, engine=eng_id,
                     queue_ring=(1 << 16) if name.startswith("zero_gap") else 0)
    assert int(want["summaries"]["events_processed"].min()) > 100
    assert_same(got, want)


def test_warp_and_lane_engines_agree(eng):
    model = hs.mm1(rate=9.0)
    kw = dict(seed=5, end_ns=80 * 10**9, n_replicas=70, record_cap=8000, sample_cap=900, service_cap=900)
    eng.upload(model)
    eng.run(engine.make_params(engine=1, **kw)); a = eng.read_outputs()
    eng.run(engine.make_params(engine=2, **kw)); b = eng.read_outputs()
    eng.run(engine.make_params(engine=3, **kw)); c = eng.read_outputs()
    assert_same(a, b)
    assert_same(c, b)


@pytest.mark.parametrize("eng_id", [0, 1, 3])
def test_mmc_sweep_cells(eng, eng_id):
    """configs[4] in small: per-cell (c, rho) overrides; replica -> cell = index // replicas_per_cell."""
    model = hs.mmc_sweep(cs=(1, 2, 5, 32), rhos=(0.5, 0.9))
    got, want = both(eng, model, seed=4, end_ns=20 * 10**9, n_replicas=8 * 6, replicas_per_cell=6, record_cap=60000,
                     sample_cap=7000, service_cap=7000, engine=eng_id)
    assert_same(got, want)
    ev = got["summaries"]["events_processed"].reshape(8, 6).mean(axis=1)
    assert ev[-1] > 10 * ev[0]          # c = 32 cells process far more requests than c = 1


@pytest.mark.parametrize("eng_id", [1, 3])
def test_warp_windowed_resume_uses_staged_state(eng, eng_id):
    """Pause/resume: the warp engine goes through the TMA bulk store/load of the replica block, the thread
    engine continues from the block where it lies."""
    model = hs.lb_round_robin(8, 64.0)
    end = 12 * 10**9
    kw = dict(seed=17, n_replicas=33, record_cap=12000, sample_cap=1500, service_cap=1500, engine=eng_id)
    want = O.oracle_run(model, O.make_params(end_ns=end, **kw))
    eng.upload(model)
    eng.run(engine.make_params(end_ns=end, window_end_ns=3 * 10**9, **kw))
    part = eng.read_outputs()
    pw = O.oracle_run(model, O.make_params(end_ns=end, window_end_ns=3 * 10**9, **kw))
    assert part["summaries"].tobytes() == pw["summaries"].tobytes()
    for c in (3 * 10**9 + 5, 7 * 10**9, 11_999_999_999):
        eng.run(engine.make_params(end_ns=end, window_end_ns=c, resume=1, **kw))
    eng.run(engine.make_params(end_ns=end, resume=1, **kw))
    assert_same(eng.read_outputs(), want)
