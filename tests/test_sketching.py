"""SURVEY 8(f) row 3 on the CPU side: sketch state layout, the host hash tables and the mirror classes
against the answers the reference's own HyperLogLog / CountMinSketch gave (tests/golden/philox_sketch_*),
merge contracts, lowering and validation."""
import ctypes as C

import numpy as np
import pytest

import golden_lib as G
import happysim_b200 as hs
import oracle_lib as O
from happysim_b200 import _abi as A, distributed as D, engine, lowering
from happysim_b200.engine import EngineError


def farm_model():
    m, kw, z = G.load("philox_sketch_cms_farm")
    return m, kw, z


def test_layout_matches_the_c_rule_in_both_libraries():
    m, _, _ = farm_model()
    per, mer, total, mtotal = m.sketch_layout()
    for L in (O.lib(), engine.load_library()):
        n = m.n_entities
        a, b = (C.c_uint64 * n)(), (C.c_uint64 * n)()
        t, mt = C.c_uint64(), C.c_uint64()
        d = m.desc()
        assert L.hs_sketch_layout(C.byref(d), a, b, C.byref(t), C.byref(mt)) == 0
        assert list(a) == per and list(b) == mer and t.value == total and mt.value == mtotal
    assert total == 3 * 16 * 4 + 16 and mtotal == (3 * 16 * 8) + 16        # CMS 3x16, HLL p=4


@pytest.mark.parametrize("name", G.case_names("philox_sketch_"))
def test_mirror_classes_give_the_reference_answers(name):
    """Oracle state == reference state (test_oracle_golden); here: our cardinality() / estimate() computed from
    that state == the reference objects' answers, and the host tables == the reference's hashing."""
    m, kw, z = G.load(name)
    out = O.oracle_run(m, O.make_params(n_replicas=1, **G.caps(z), **kw))
    G.check_against(z, out)
    views = m.sketch_views(out["sketches"])
    for i, state in views.items():
        e = m.entities[i]
        seed = int(z[f"sketch_seed_{i}"]); seed = None if seed < 0 else seed
        want = z[f"sketch_answer_{i}"]
        added = int(out["entity_stats"][0][i]["c1"])
        if int(e["i0"]) == A.HS_SK_HLL:
            sk = hs.HyperLogLog(precision=int(e["i2"]), seed=seed)
            sk._load_device_state(state[0], added)
            assert sk.cardinality() == int(want[0]) and sk.item_count == added
        else:
            sk = hs.CountMinSketch(width=int(e["i3"]), depth=int(e["i2"]), seed=seed)
            sk._load_device_state(state[0], added)
            assert [sk.estimate(k) for k in range(int(e["l0"]))] == [int(x) for x in want]
            assert int(state[0].sum()) == added * int(e["i2"])           # every add touches one cell per row


def test_host_side_add_equals_device_table_path():
    K = 60
    keys = np.random.RandomState(1).randint(0, K, size=500)
    h1, c1 = hs.HyperLogLog(precision=7, seed=3), hs.CountMinSketch(width=11, depth=4, seed=3)
    for k in keys:
        h1.add(int(k)); c1.add(int(k))
    ht, ct = hs.hll_table(7, 3, K), hs.cms_table(11, 4, 3, K)
    regs = np.zeros(128, np.uint8)
    np.maximum.at(regs, ht[0, keys], ht[1, keys].astype(np.uint8))
    cnt = np.zeros((4, 11), np.uint64)
    for row in range(4):
        np.add.at(cnt[row], ct[row, keys], 1)
    assert np.array_equal(h1._registers, regs) and np.array_equal(c1._counters, cnt)
    assert c1.estimate(int(keys[0])) >= int((keys == keys[0]).sum())     # never underestimates


def test_merge_contracts_over_replicas():
    m, kw, z = farm_model()
    out = O.oracle_run(m, O.make_params(n_replicas=5, seed=kw["seed"], end_ns=kw["end_ns"], rid_base=0))
    merged = D.merge_sketch_states(m, out["sketches"])
    views = m.sketch_views(out["sketches"])
    for i, v in views.items():
        if int(m.entities["i0"][i]) == A.HS_SK_HLL:
            acc = hs.HyperLogLog(precision=int(m.entities["i2"][i]))
            for r in range(5):
                o = hs.HyperLogLog(precision=int(m.entities["i2"][i])); o._load_device_state(v[r], 0); acc.merge(o)
            assert np.array_equal(acc._registers, merged[i])
        else:
            d, w = int(m.entities["i2"][i]), int(m.entities["i3"][i])
            acc = hs.CountMinSketch(w, d)
            for r in range(5):
                o = hs.CountMinSketch(w, d); o._load_device_state(v[r], int(out["entity_stats"][r][i]["c1"])); acc.merge(o)
            assert np.array_equal(acc._counters, merged[i])
            assert acc.item_count == int(out["entity_stats"][:, i]["c1"].sum())
    with pytest.raises(ValueError):
        hs.HyperLogLog(8).merge(hs.HyperLogLog(9))
    with pytest.raises(ValueError):
        hs.CountMinSketch(8, 2, seed=1).merge(hs.CountMinSketch(8, 2, seed=2))


def test_lowering_of_a_sketch_collector_model():
    K = 30
    hll = hs.HyperLogLog(precision=6, seed=4)
    col = hs.SketchCollector("uniques", hll, hs.KeyExtractor())
    srv = hs.Server("S", concurrency=1, service_time=hs.ExponentialLatency(0.01), downstream=col)
    src = hs.Source.poisson(rate=50.0, event_provider=hs.SimpleEventProvider(srv, context_fn=hs.UniformKeyContext(K)))
    model, objs = lowering.lower([src], [srv, col])
    i = objs.index(col)
    e = model.entities[i]
    assert int(e["kind"]) == A.HS_ENT_SKETCH and int(e["i0"]) == A.HS_SK_HLL and int(e["i2"]) == 6 and int(e["l0"]) == K
    assert np.array_equal(model.sketch_tables.reshape(2, K), hs.hll_table(6, 4, K))
    engine.validate_model(model)
    bad = hs.SketchCollector("x", hs.HyperLogLog(6), value_extractor=lambda ev: 1)
    with pytest.raises(hs.UnsupportedModelError, match="value_extractor"):
        lowering.lower([hs.Source.poisson(rate=1.0, event_provider=hs.SimpleEventProvider(bad, context_fn=hs.UniformKeyContext(K)))], [bad])


def test_validation_rejects_bad_sketch_rows():
    def model(**over):
        b = hs.ModelBuilder()
        s = b.source(rate=1.0, key_population=8)
        h = b.sketch_hll(precision=5, table=hs.hll_table(5, 0, 8))
        b.set_target(s, h)
        m = b.build()
        for k, v in over.items():
            m.entities[k][1] = v
        return m
    engine.validate_model(model())
    for over, msg in ((dict(i2=3), "precision"), (dict(i0=9), "algorithm"), (dict(l0=0), "population"), (dict(i1=5), "table")):
        with pytest.raises(EngineError, match=msg):
            engine.validate_model(model(**over))
    m = model()
    m.sketch_tables = m.sketch_tables.copy(); m.sketch_tables[3] = 32        # register index out of range for p = 5
    with pytest.raises(EngineError, match="HLL table"):
        engine.validate_model(m)
