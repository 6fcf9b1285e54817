"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container).

    python tests/golden/gen_golden.py          # needs /root/reference

Two families of fixtures, both produced by the reference's own Simulation.run():

  philox_<case>.npz   the reference driven through its plug-in points by the shared
                      Philox sampler (ref_harness.run_reference): the full processed-event
                      sequence hash, counts, per-entity statistics, the first records and
                      samples.  The oracle and the CUDA engine must reproduce these bit for bit.
  stock_<case>.npz    the reference with its STOCK generators (random.seed(s);
                      np.random.seed(s)) -- the README quick-start known answers of
                      SURVEY.md 8(c) -- together with the generators' outputs, so the oracle's
                      state machine and time arithmetic can be replayed against a run that
                      never saw our sampler (oracle_lib.oracle_run_trace).

The fixtures are small (first MAX_REC records / MAX_SMP samples; the order hash covers the
whole run) and committed; nothing on the GPU box reads /root/reference.
"""
from __future__ import annotations

import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]

import happysim_b200 as hs  # noqa: E402
import ref_harness as RH  # noqa: E402

MAX_REC = 4000
MAX_SMP = 1500


def names_for(model):
    return list(model.names)


def chash_case(n_servers, vnodes, population, rate):
    names = [f"S{i}" for i in range(n_servers)]
    tab = RH.ring_table_from_reference(names, vnodes, population)
    return hs.lb_key_table(tab, n_servers, rate), vnodes


def tandem():
    b = hs.ModelBuilder()
    src = b.source(rate=6.0)
    s1 = b.server("A", mean_service_s=0.08)
    s2 = b.server("B", concurrency=2, mean_service_s=0.2)
    snk = b.sink()
    b.set_target(src, s1); b.set_target(s1, s2); b.set_target(s2, snk)
    return b.build()


def source_to_counter():
    b = hs.ModelBuilder()
    src = b.source("PingSource", rate=1.0, poisson=False)
    c = b.counter("pingcounter")
    b.set_target(src, c)
    return b.build()


def two_sources():
    b = hs.ModelBuilder()
    a = b.source("A", rate=3.0)
    c = b.source("B", rate=4.0, poisson=False, stop_after_ns=20 * 10**9)
    srv = b.server(concurrency=2, mean_service_s=0.1)
    snk = b.sink()
    b.set_target(a, srv); b.set_target(c, srv); b.set_target(srv, snk)
    return b.build()


def profiled(profile, poisson, lb=0):
    b = hs.ModelBuilder()
    src = b.source(profile=profile, poisson=poisson)
    snk = None
    if lb:
        servers = [b.server(f"S{i}", mean_service_s=0.05) for i in range(lb)]
        snk = b.sink()
        l = b.load_balancer(backends=servers)
        b.set_target(src, l)
        for sv in servers:
            b.set_target(sv, snk)
    else:
        srv = b.server(mean_service_s=0.05)
        snk = b.sink()
        b.set_target(src, srv); b.set_target(srv, snk)
    return b.build()


def probed():
    """M/M/1 with a depth probe every 0.1 s and a Sink counter probe every 0.5 s."""
    b = hs.ModelBuilder()
    src = b.source(rate=8.0)
    srv = b.server(mean_service_s=0.1)
    snk = b.sink()
    b.set_target(src, srv); b.set_target(srv, snk)
    b.probe("Probe_Server_depth", target=srv, metric="depth", interval_s=0.1)
    b.probe("Probe_Sink_events_received", target=snk, metric="events_received", interval_s=0.5)
    m = b.build()
    # Simulation.__init__ bootstraps sources first, then probes: ids must follow that order
    return reorder_sources_first(m)


def reorder_sources_first(m):
    import numpy as np
    from happysim_b200 import _abi as A
    n = m.n_entities
    kinds = m.entities["kind"]
    is_probe_src = [int(kinds[i]) == A.HS_ENT_SOURCE and int(kinds[int(m.entities["target"][i])]) == A.HS_ENT_PROBE for i in range(n)]
    order = [i for i in range(n) if int(kinds[i]) == A.HS_ENT_SOURCE and not is_probe_src[i]] + \
            [i for i in range(n) if is_probe_src[i]] + [i for i in range(n) if int(kinds[i]) != A.HS_ENT_SOURCE]
    new_id = {old: new for new, old in enumerate(order)}
    ents = m.entities[order].copy()
    for r in ents:
        if int(r["target"]) >= 0:
            r["target"] = new_id[int(r["target"])]
    m.entities = ents
    m.names = [m.names[i] for i in order]
    m.backends = np.array([new_id[int(x)] for x in m.backends], dtype=np.int32)
    return m


def sketch_direct(K=200, precision=8, hll_seed=5):
    """Source with client ids -> SketchCollector(HyperLogLog)."""
    b = hs.ModelBuilder()
    src = b.source(rate=300.0, key_population=K)
    h = b.sketch_hll("uniques", precision=precision, table=hs.hll_table(precision, hll_seed, K))
    b.set_target(src, h)
    return b.build(), {h: hll_seed}


def sketch_farm(K=40, n_servers=4, width=16, depth=3, cms_seed=9, hll_seed=None):
    """Source -> LoadBalancer(RoundRobin) -> servers -> SketchCollector(CountMinSketch); a second source
    without keys feeds a HyperLogLog collector directly (value None: counted, not added)."""
    b = hs.ModelBuilder()
    src = b.source("Keyed", rate=64.0, key_population=K)
    plain = b.source("Plain", rate=5.0, poisson=False)
    servers = [b.server(f"S{i}", mean_service_s=0.05) for i in range(n_servers)]
    cms = b.sketch_cms("freq", width=width, depth=depth, table=hs.cms_table(width, depth, cms_seed, K))
    hll = b.sketch_hll("nokeys", precision=4, table=hs.hll_table(4, hll_seed, K))
    lb = b.load_balancer(backends=servers)
    b.set_target(src, lb); b.set_target(plain, hll)
    for sv in servers:
        b.set_target(sv, cms)
    return b.build(), {cms: cms_seed, hll: hll_seed}


def sketch_members(K=60, n_servers=3, bloom_seed=3):
    """Source -> LoadBalancer -> servers -> {BloomFilter, TopK(k=5), TopK(k=5)} collectors: membership and heavy
    hitters; k < number of distinct keys, so Space-Saving evicts (topk.py:116-128)."""
    b = hs.ModelBuilder()
    src = b.source(rate=120.0, key_population=K)
    servers = [b.server(f"S{i}", concurrency=2, mean_service_s=0.02) for i in range(n_servers)]
    bloom = b.sketch_bloom("seen", size_bits=200, num_hashes=3, table=hs.bloom_table(200, 3, bloom_seed, K))
    top = b.sketch_topk("heavy", k=5, key_population=K)
    top2 = b.sketch_topk("heavy2", k=40, key_population=K)
    lb = b.load_balancer(backends=servers)
    b.set_target(src, lb)
    for sv, dst in zip(servers, (bloom, top, top2)):
        b.set_target(sv, dst)
    return b.build(), {bloom: bloom_seed}


def sketch_quantiles():
    """M/M/2 -> QuantileEstimator(compression=20): 40-value buffer, several flush + compress rounds; a second
    estimator with the default compression behind a second server keeps a partly filled buffer."""
    b = hs.ModelBuilder()
    src = b.source(rate=30.0)
    s1 = b.server("A", concurrency=2, mean_service_s=0.05)
    s2 = b.server("B", concurrency=1, mean_service_s=0.01)
    q1 = b.sketch_tdigest("lat20", compression=20.0)
    q2 = b.sketch_tdigest("lat100", compression=100.0)
    lb = b.load_balancer(backends=[s1, s2])
    b.set_target(src, lb); b.set_target(s1, q1); b.set_target(s2, q2)
    return b.build()


def sketch_reservoirs(K=500):
    """Source -> LoadBalancer -> 3 servers -> ReservoirSampler collectors of size 16, 700 and 5: every sampler
    goes through several 624-word refills of its MT19937 and plenty of rejected draws (randint -> _randbelow);
    a key-less source feeds a fourth sampler directly (value None: counted, never sampled)."""
    import random
    b = hs.ModelBuilder()
    src = b.source("Keyed", rate=200.0, key_population=K)
    plain = b.source("Plain", rate=3.0, poisson=False)
    servers = [b.server(f"S{i}", concurrency=2, mean_service_s=0.01) for i in range(3)]
    started = random.Random(99)
    for _ in range(1000):                       # a generator that is already mid-stream when the run starts
        started.random()
    rs = [b.sketch_reservoir("r16", size=16, seed=7, key_population=K),
          b.sketch_reservoir("r700", size=700, seed=123, key_population=K),
          b.sketch_reservoir("r5", size=5, state=started.getstate()[1], key_population=K)]
    idle = b.sketch_reservoir("idle", size=4, seed=1, key_population=K)
    lb = b.load_balancer(backends=servers)
    b.set_target(src, lb); b.set_target(plain, idle)
    for sv, dst in zip(servers, rs):
        b.set_target(sv, dst)
    return b.build()


def zipf_hot_keys(K=300, s=1.1, n_servers=6):
    """client_id ~ Zipf(s) -> LoadBalancer(ConsistentHash) -> servers -> TopKCollector: hot keys pile on a few
    ring nodes and Space-Saving has real heavy hitters to find (distributions/zipf.py:27-123)."""
    names = [f"S{i}" for i in range(n_servers)]
    tab = RH.ring_table_from_reference(names, 30, K)
    b = hs.ModelBuilder()
    src = b.source(rate=150.0, key_population=K, key_cdf=hs.zipf_cdf(K, s))
    servers = [b.server(nm, concurrency=2, mean_service_s=0.02) for nm in names]
    top = b.sketch_topk("heavy", k=8, key_population=K)
    lb = b.load_balancer(backends=servers, key_table=tab)
    b.set_target(src, lb)
    for sv in servers:
        b.set_target(sv, top)
    return b.build(), {src: s}


def example_metastable_profile():
    """The profile class of the reference's own example (examples/queuing/m_m_1_queue.py:104-169), imported from
    the example file itself: a user-defined step function incl. the int((t - 65.0) / 11.0) step-down phase."""
    import importlib.util
    RH._import_reference()
    path = os.path.join(RH.REFERENCE_ROOT, "examples", "queuing", "m_m_1_queue.py")
    spec = importlib.util.spec_from_file_location("ref_example_mm1", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_example_mm1"] = mod
    spec.loader.exec_module(mod)
    return mod.MetastableLoadProfile()


def step_profiled(profile_obj, end_s, mean_service_s=0.1, stop_after_s=None):
    """Source(PoissonArrivalTimeProvider(<user step profile>)) -> Server(Exp) -> Sink, the table tabulated by
    happysim_b200.lowering.step_table_from_profile exactly as Simulation() would."""
    from happysim_b200 import lowering
    breaks, rates = lowering.step_table_from_profile(profile_obj, scan_end_s=2.0 * end_s + 120.0)
    b = hs.ModelBuilder()
    src = b.source(profile=("step", breaks, rates), poisson=True,
                   stop_after_ns=-1 if stop_after_s is None else int(stop_after_s * 1e9))
    srv = b.server(mean_service_s=mean_service_s)
    snk = b.sink()
    b.set_target(src, srv); b.set_target(srv, snk)
    return b.build(), {src: profile_obj}


def cache_farm(n_servers, K, rate, ttl_s, vnodes=None, poisson=True):
    """examples/load-balancing/consistent_hashing_basics.py:161-255: Source (customer ids ~ Uniform{0..K-1}) ->
    LoadBalancer(ConsistentHash | RoundRobin) -> n x CachingServer(TTL cache over a shared datastore)."""
    names = [f"Server_{i}" for i in range(n_servers)]
    b = hs.ModelBuilder()
    src = b.source(rate=rate, key_population=K, poisson=poisson)
    servers = [b.cache_server(nm, key_slots=K, cache_ttl_s=ttl_s) for nm in names]
    if n_servers == 1:
        b.set_target(src, servers[0])
        return b.build()
    tab = RH.ring_table_from_reference(names, vnodes, K) if vnodes else None
    lb = b.load_balancer(backends=servers, key_table=tab)
    b.set_target(src, lb)
    return b.build()


def philox_cases():
    c = {}
    c["mm1_seed0"] = (hs.mm1(), dict(seed=0, rid=0, end_s=60))
    c["mm1_seed1"] = (hs.mm1(), dict(seed=1, rid=0, end_s=60))
    c["mm1_seed42"] = (hs.mm1(), dict(seed=42, rid=0, end_s=60))
    c["mm1_rid77_long"] = (hs.mm1(rate=9.5), dict(seed=1234, rid=77, end_s=400))
    c["dd1_constant"] = (hs.mm1(poisson=False, exponential=False, rate=10, mean_service_s=0.05), dict(seed=1, rid=0, end_s=20))
    c["mm1_capacity5"] = (hs.mm1(rate=50, capacity=5), dict(seed=3, rid=2, end_s=20))
    c["mm1_lifo"] = (hs.mm1(rate=9, lifo=True), dict(seed=3, rid=5, end_s=60))
    c["mmc4"] = (hs.mm1(rate=32, concurrency=4), dict(seed=9, rid=1, end_s=30))
    c["mmc32"] = (hs.mm1(rate=256, concurrency=32), dict(seed=9, rid=3, end_s=10))
    c["lb_rr8"] = (hs.lb_round_robin(n_servers=8, rate=64.0), dict(seed=5, rid=0, end_s=10))
    c["lb_rr64"] = (hs.lb_round_robin(n_servers=64, rate=512.0), dict(seed=7, rid=9, end_s=5))
    m, v = chash_case(16, 20, 500, 128.0)
    c["lb_chash16"] = (m, dict(seed=11, rid=4, end_s=5, chash_vnodes=v))
    c["tandem"] = (tandem(), dict(seed=13, rid=0, end_s=60))
    c["source_to_counter"] = (source_to_counter(), dict(seed=0, rid=0, end_s=60))
    c["two_sources"] = (two_sources(), dict(seed=21, rid=6, end_s=40))
    # non-constant rate profiles (adaptive Simpson + Brent arrival path)
    c["ramp_poisson_mm1"] = (profiled(("linear_ramp", 20.0, 2.0, 12.0), True), dict(seed=31, rid=2, end_s=30))
    c["spike_poisson_mm1"] = (profiled(("spike", 5.0, 40.0, 4.0, 3.0), True), dict(seed=32, rid=0, end_s=12))
    c["ramp_down_constant"] = (profiled(("linear_ramp", 5.0, 20.0, 1.0), False), dict(seed=0, rid=0, end_s=20))
    c["probe_mm1"] = (probed(), dict(seed=42, rid=0, end_s=20))
    c["spike_constant_lb4"] = (profiled(("spike", 10.0, 100.0, 2.0, 1.0), False, lb=4), dict(seed=3, rid=1, end_s=6))
    # sketches as instrumentation sinks
    m, seeds = sketch_direct()
    c["sketch_hll_direct"] = (m, dict(seed=17, rid=3, end_s=2, sketch_seeds=seeds))
    m, seeds = sketch_farm()
    c["sketch_cms_farm"] = (m, dict(seed=19, rid=1, end_s=6, sketch_seeds=seeds))
    m, seeds = sketch_members()
    c["sketch_bloom_topk"] = (m, dict(seed=23, rid=2, end_s=5, sketch_seeds=seeds))
    c["sketch_tdigest"] = (sketch_quantiles(), dict(seed=29, rid=0, end_s=40))
    c["sketch_reservoir"] = (sketch_reservoirs(), dict(seed=43, rid=1, end_s=20))
    m, zs = zipf_hot_keys()
    c["zipf_chash_topk"] = (m, dict(seed=37, rid=5, end_s=6, chash_vnodes=30, zipf_s=zs))
    # SURVEY 8(f) row 4, second half: CachingServer / TTL cache behind the two load-balancing strategies
    c["cache_direct_ttl"] = (cache_farm(1, 3, 10.0, 0.5, poisson=False), dict(seed=5, rid=0, end_s=6))
    c["cache_chash5"] = (cache_farm(5, 40, 200.0, 0.8, vnodes=30), dict(seed=41, rid=2, end_s=6, chash_vnodes=30))
    c["cache_rr5"] = (cache_farm(5, 40, 200.0, 0.8), dict(seed=41, rid=2, end_s=6))
    # a user-defined step profile: the example's MetastableLoadProfile, its whole 130 s scenario
    m, po = step_profiled(example_metastable_profile(), 130.0, stop_after_s=120.0)
    c["step_metastable_mm1"] = (m, dict(seed=42, rid=0, end_s=130, profile_objects=po))
    return c


def save_case(path, model, ref, meta):
    rec = ref["records"]
    np.savez_compressed(
        path,
        entities=model.entities, backends=model.backends, key_table=model.key_table, profiles=model.profiles,
        profile_table=model.profile_table,
        names=np.array(model.names), meta=np.array([meta["seed"], meta["rid"], int(meta["end_s"] * 1e9)], dtype=np.int64),
        summaries=ref["summaries"], entity_stats=ref["entity_stats"],
        n_records=np.int64(len(rec)), records=rec[:MAX_REC],
        n_samples=np.int64(len(ref["sink_samples"])), sink_samples=ref["sink_samples"][:MAX_SMP],
        n_service=np.int64(len(ref["service_samples"])), service_samples=ref["service_samples"][:MAX_SMP],
        case_name=np.array(os.path.basename(path)[:-4]), key_cdf=model.key_cdf, sketch_tables=model.sketch_tables, sketch_state=ref.get("sketches", np.zeros(0, np.uint8)),
        **{f"sketch_answer_{i}": a for i, a in ref.get("sketch_answers", {}).items()},
        **{f"sketch_seed_{i}": np.int64(-1 if sd is None else sd) for i, sd in (meta.get("sketch_seeds") or {}).items()},
        **{k: v for k, v in meta.items() if isinstance(v, np.ndarray)},
    )


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else ""          # e.g. "sketch": regenerate matching cases only
    for name, (model, kw) in philox_cases().items():
        if only not in name:
            continue
        ref = RH.run_reference(model, seed=kw["seed"], rid=kw["rid"], end_ns=int(kw["end_s"] * 1e9),
                               chash_vnodes=kw.get("chash_vnodes"), sketch_seeds=kw.get("sketch_seeds"),
                               zipf_s=kw.get("zipf_s"), profile_objects=kw.get("profile_objects"))
        save_case(os.path.join(HERE, f"philox_{name}.npz"), model, ref, kw)
        print(f"philox_{name}: {len(ref['records'])} events, hash {int(ref['summaries']['order_hash'][0]):#x}")

    # ---- stock generators: the README quick-start known answers ----
    stock = [("mm1_seed42", hs.mm1(), 42, 60), ("mm1_seed7", hs.mm1(), 7, 200),
             # several servers share Python's global stream in simulation order
             ("lb_rr8_seed5", hs.lb_round_robin(n_servers=8, rate=64.0), 5, 8),
             ("mmc4_seed9", hs.mm1(rate=32, concurrency=4), 9, 20)]
    for sname, model, seed, end_s in stock:
        if only not in "stock_" + sname:
            continue
        ref = RH.run_reference(model, seed=seed, end_ns=int(end_s * 1e9), stock_rng=True)
        n_draw = len(ref["records"])
        u = np.random.RandomState(seed).random_sample(n_draw)          # numpy legacy global stream
        targets = np.array([-math.log(1.0 - x) for x in u])            # poisson_arrival.py:31
        rnd = random.Random(seed)
        # random.expovariate(l) = -log(1.0 - random()) / l (exponential.py:36,43): store -log(1 - U)
        service = np.array([-math.log(1.0 - rnd.random()) for _ in range(n_draw)])
        meta = dict(seed=seed, rid=0, end_s=end_s, trace_targets=targets, trace_service=service)
        save_case(os.path.join(HERE, f"stock_{sname}.npz"), model, ref, meta)
        sink = [o for o in ref["objects"] if type(o).__name__ == "Sink"][0]
        print(f"stock_{sname}: events={ref['summary'].total_events_processed} "
              f"sink={sink.events_received} avg_latency={sink.average_latency()!r} "
              f"final_ns={int(ref['summaries']['final_time_ns'][0])} heap_left={int(ref['summaries']['heap_left'][0])}")


if __name__ == "__main__":
    main()
