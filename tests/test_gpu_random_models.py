"""Randomised differential testing on the device: 64 seeded models mixing every lowered feature, both general
engines (and the automatic choice) against the CPU oracle -- event records, statistics, samples, sketch states."""
import pytest

import oracle_lib as O
from happysim_b200 import engine
from random_models import random_model
from test_gpu_lane_parity import assert_same
from test_random_models import SEEDS, check_against_reference

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = engine.Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("seed", SEEDS)
def test_random_model_on_every_engine(eng, seed):
    model, end_s, what = random_model(seed)
    kw = dict(seed=1000 + seed, end_ns=int(end_s * 1e9), n_replicas=5, record_cap=12000, sample_cap=1500, service_cap=1500,
              queue_ring=1024)
    want = O.oracle_run(model, O.make_params(**kw))
    kw1 = dict(kw, n_replicas=1)                 # the fixture's replica: word 0 alone, as the reference ran it
    eng.upload(model)
    eng.run(engine.make_params(**kw1))
    check_against_reference(model, eng.read_outputs(), seed)
    for eng_id in (0, 1, 3):
        eng.run(engine.make_params(engine=eng_id, **kw))
        got = eng.read_outputs()
        try:
            assert_same(got, want)
            if want.get("sketches") is not None:
                assert got["sketches"].tobytes() == want["sketches"].tobytes(), "sketch states differ"
        except AssertionError as e:
            raise AssertionError(f"{what}, engine {eng_id}: {e}") from None


@pytest.mark.parametrize("seed", SEEDS[::2])
def test_random_model_cut_into_windows(eng, seed):
    """Pause / resume on both general engines: a run cut at three arbitrary instants yields the same records,
    statistics, samples and sketch states as the uncut run (state parked in HBM between the calls)."""
    model, end_s, what = random_model(seed)
    end_ns = int(end_s * 1e9)
    kw = dict(seed=1000 + seed, n_replicas=4, record_cap=12000, sample_cap=1500, service_cap=1500, queue_ring=1024)
    want = O.oracle_run(model, O.make_params(end_ns=end_ns, **kw))
    cuts = [end_ns // 7, end_ns // 2 + 3, end_ns - 1]
    eng.upload(model)
    for eng_id in (1, 3):
        eng.run(engine.make_params(end_ns=end_ns, window_end_ns=cuts[0], engine=eng_id, **kw))
        for c in cuts[1:]:
            eng.run(engine.make_params(end_ns=end_ns, window_end_ns=c, resume=1, engine=eng_id, **kw))
        eng.run(engine.make_params(end_ns=end_ns, resume=1, engine=eng_id, **kw))
        got = eng.read_outputs()
        try:
            assert_same(got, want)
            if want.get("sketches") is not None:
                assert got["sketches"].tobytes() == want["sketches"].tobytes(), "sketch states differ"
        except AssertionError as e:
            raise AssertionError(f"{what}, engine {eng_id}, windows: {e}") from None


from random_models import random_lane_model          # noqa: E402
from test_random_models import LANE_SEEDS            # noqa: E402


@pytest.mark.parametrize("seed", LANE_SEEDS)
def test_random_single_server_model_on_lane_and_general_engines(eng, seed):
    """The register-resident lane engine (recorder and summary kernels, uncut and cut into windows) and the
    thread engine against the oracle on random Source -> Server -> Sink|Counter|- models, including exact
    tick/completion ties."""
    model, end_s, what = random_lane_model(seed)
    end_ns = int(end_s * 1e9)
    kw = dict(seed=77 + seed, n_replicas=70, queue_ring=2048)
    caps = dict(record_cap=12000, sample_cap=3000, service_cap=3000)
    want = O.oracle_run(model, O.make_params(end_ns=end_ns, **kw, **caps))
    eng.upload(model)
    try:
        for eng_id in (2, 3):
            eng.run(engine.make_params(end_ns=end_ns, engine=eng_id, **kw, **caps))
            assert_same(eng.read_outputs(), want)
        # summary kernels (no recorder): summaries and statistics only
        eng.run(engine.make_params(end_ns=end_ns, engine=2, **kw))
        got = eng.read_outputs()
        assert got["summaries"].tobytes() == want["summaries"].tobytes(), "summary kernel: summaries differ"
        assert got["entity_stats"].tobytes() == want["entity_stats"].tobytes(), "summary kernel: statistics differ"
        # the lane engine cut into windows
        cuts = [end_ns // 5, end_ns // 2 + 1, end_ns - 2]
        eng.run(engine.make_params(end_ns=end_ns, window_end_ns=cuts[0], engine=2, **kw, **caps))
        for c in cuts[1:]:
            eng.run(engine.make_params(end_ns=end_ns, window_end_ns=c, resume=1, engine=2, **kw, **caps))
        eng.run(engine.make_params(end_ns=end_ns, resume=1, engine=2, **kw, **caps))
        assert_same(eng.read_outputs(), want)
    except AssertionError as e:
        raise AssertionError(f"{what}: {e}") from None


# ---- second generator: user-defined step profiles and CachingServer farms ------------------------------------------
from random_models import random_model_v2            # noqa: E402
from test_random_models import SEEDS_V2, check_against_reference_v2   # noqa: E402


@pytest.mark.parametrize("seed", SEEDS_V2)
def test_random_step_profile_and_cache_model_on_every_engine(eng, seed):
    """Every engine that accepts the model (lane for a single ordinary server, thread, warp) against the oracle --
    records, statistics, samples, TTL-cache states -- and replica word 0 against the unmodified reference; then the
    same run cut into windows."""
    model, end_s, what = random_model_v2(seed)
    end_ns = int(end_s * 1e9)
    kw = dict(seed=2000 + seed, n_replicas=5, record_cap=16000, sample_cap=2000, service_cap=2000, queue_ring=1024)
    want = O.oracle_run(model, O.make_params(end_ns=end_ns, **kw))
    eng.upload(model)
    eng.run(engine.make_params(end_ns=end_ns, **dict(kw, n_replicas=1)))
    check_against_reference_v2(eng.read_outputs(), seed)
    for eng_id in (0, 1, 3):
        try:
            eng.run(engine.make_params(end_ns=end_ns, engine=eng_id, **kw))
            got = eng.read_outputs()
            assert_same(got, want)
            if want.get("sketches") is not None:
                assert got["sketches"].tobytes() == want["sketches"].tobytes(), "TTL cache states differ"
            cuts = [end_ns // 5, end_ns // 2 + 7]
            eng.run(engine.make_params(end_ns=end_ns, window_end_ns=cuts[0], engine=eng_id, **kw))
            eng.run(engine.make_params(end_ns=end_ns, window_end_ns=cuts[1], resume=1, engine=eng_id, **kw))
            eng.run(engine.make_params(end_ns=end_ns, resume=1, engine=eng_id, **kw))
            got = eng.read_outputs()
            assert_same(got, want)
            if want.get("sketches") is not None:
                assert got["sketches"].tobytes() == want["sketches"].tobytes(), "TTL cache states differ after windows"
        except AssertionError as e:
            raise AssertionError(f"{what}, engine {eng_id}: {e}") from None
