"""The CUDA engine against the fixtures recorded from the UNMODIFIED reference
(tests/golden/philox_*.npz): same Philox key, same replica word -> identical processed-event
sequence, counts, statistics and samples, through the C-ABI."""
import numpy as np
import pytest

import golden_lib as G
from happysim_b200 import engine, _abi as A

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = engine.Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("eng_id", [0, 1, 3])      # auto (lane where it applies, else thread), warp, thread
@pytest.mark.parametrize("name", G.case_names("philox_"))
def test_engine_reproduces_reference_fixture(eng, name, eng_id):
    model, kw, z = G.load(name)
    eng.upload(model)
    kw["engine"] = eng_id
    # the fixture's replica sits in the middle of a small ensemble
    kw = dict(kw)
    rid = kw.pop("rid_base")
    eng.run(engine.make_params(n_replicas=5, rid_base=rid - 2 if rid >= 2 else rid, rid_stride=1, **G.caps(z), **kw))
    got = eng.read_outputs()
    G.check_against(z, got, r=2 if rid >= 2 else 0)


def test_flight_recorder_rings_keep_the_tail(eng):
    """Rings smaller than the run retain exactly the last `cap` items, in order."""
    model, kw, z = G.load("philox_mm1_rid77_long")
    eng.upload(model)
    full = dict(G.caps(z))
    eng.run(engine.make_params(n_replicas=1, **full, **kw))
    whole = eng.read_outputs()
    eng.run(engine.make_params(n_replicas=1, record_cap=1000, sample_cap=300, service_cap=77, **kw))
    ring = eng.read_outputs()
    s = ring["summaries"][0]
    n, ns, nv = int(s["events_processed"]), int(s["n_sink_samples"]), int(s["n_service_samples"])
    assert n == int(z["n_records"]) and ns == int(z["n_samples"]) and nv == int(z["n_service"])
    assert np.array_equal(A.unroll_ring(ring["records"][0], n, 1000), whole["records"][0][n - 1000:n])
    assert np.array_equal(A.unroll_ring(ring["sink_samples"][0], ns, 300), whole["sink_samples"][0][ns - 300:ns])
    assert np.array_equal(A.unroll_ring(ring["service_samples"][0], nv, 77), whole["service_samples"][0][nv - 77:nv])
    assert int(s["order_hash"]) == int(z["summaries"]["order_hash"][0])


@pytest.mark.parametrize("eng_id", [0, 1, 3])
@pytest.mark.parametrize("name", G.case_names("stock_"))
def test_engine_reproduces_stock_seeded_reference_from_its_generator_streams(eng, name, eng_id):
    """The UNMODIFIED, stock-seeded reference (no plug-ins) vs the engine fed with the two MT19937
    streams as unit-rate exponentials (hs_set_trace): lane engine for M/M/1, the general engines with the
    shared-stream cursors for several servers."""
    model, kw, z = G.load(name)
    kw = dict(kw); kw.pop("seed"); kw.pop("rid_base")
    eng.upload(model)
    eng.set_trace(z["trace_targets"][None, :], z["trace_service"][None, :])
    try:
        eng.run(engine.make_params(n_replicas=1, engine=eng_id, **G.caps(z), **kw))
        got = eng.read_outputs()
    finally:
        eng.set_trace(None, None)
    G.check_against(z, got)
