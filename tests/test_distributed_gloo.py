"""N>1 host logic on CPU: world_size-2 gloo.  Each rank simulates its shard of a replica
ensemble (with the CPU oracle standing in for the device, as the checker) using the GLOBAL
replica ids, then the ranks all-reduce the fixed-layout totals vector; the result must equal
the single-process totals over the whole ensemble."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import happysim_b200 as hs
    from happysim_b200 import distributed as D
    import oracle_lib as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = hs.lb_round_robin(4, 32.0)
    lo, hi = D.shard_range(n_total, rank, world)
    out = O.oracle_run(model, O.make_params(seed=77, end_ns=5 * 10**9, n_replicas=hi - lo, replica_index_base=lo))
    t = D.allreduce_totals(D.totals_from_outputs(model, out))
    q.put((rank, lo, hi, list(t.i), list(t.fsum), t.fmin, t.fmax,
           out["summaries"]["order_hash"].tolist()))
    dist.destroy_process_group()


def test_shard_range_covers_everything_once():
    from happysim_b200.distributed import shard_range
    for n, w in ((10, 3), (65536, 8), (7, 8), (262144, 8)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_two_rank_allreduce_equals_single_process():
    import happysim_b200 as hs
    from happysim_b200 import distributed as D
    import oracle_lib as O
    n_total, world, port = 23, 2, 29000 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = hs.lb_round_robin(4, 32.0)
    whole = O.oracle_run(model, O.make_params(seed=77, end_ns=5 * 10**9, n_replicas=n_total))
    want = D.totals_from_outputs(model, whole)
    for rank, lo, hi, ti, tf, fmin, fmax, hashes in got:
        assert ti == list(want.i)
        assert np.allclose(tf, list(want.fsum), rtol=1e-12)
        assert fmin == want.fmin and fmax == want.fmax
        # global replica ids: a shard reproduces exactly its slice of the whole ensemble
        assert hashes == whole["summaries"]["order_hash"][lo:hi].tolist()
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == n_total


def _sketch_worker(rank, world, port, n_total, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    from happysim_b200 import distributed as D
    import golden_lib as G
    import oracle_lib as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = {}
    for name in ("philox_sketch_cms_farm", "philox_sketch_bloom_topk"):
        model, kw, _ = G.load(name)
        lo, hi = D.shard_range(n_total, rank, world)
        out = O.oracle_run(model, O.make_params(seed=kw["seed"], end_ns=kw["end_ns"], n_replicas=hi - lo, replica_index_base=lo))
        merged = D.allreduce_sketches(model, D.merge_sketch_states(model, out["sketches"]))
        res[name] = {i: v.tolist() for i, v in merged.items() if hasattr(v, "tolist")}      # TopK stays rank-local
    q.put((rank, res))
    dist.destroy_process_group()


def test_two_rank_sketch_merge_equals_single_process():
    """SKETCH rows across ranks: MAX all-reduce (HLL registers), SUM all-reduce (CMS counters) and the Bloom
    filters' OR as a MAX over unpacked bits, of the per-rank merged images == the merge over the whole ensemble."""
    from happysim_b200 import distributed as D
    import golden_lib as G
    import oracle_lib as O
    n_total, world, port = 9, 2, 31000 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sketch_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for name in ("philox_sketch_cms_farm", "philox_sketch_bloom_topk"):
        model, kw, _ = G.load(name)
        whole = O.oracle_run(model, O.make_params(seed=kw["seed"], end_ns=kw["end_ns"], n_replicas=n_total))
        want = {i: v for i, v in D.merge_sketch_states(model, whole["sketches"]).items() if hasattr(v, "tolist")}
        assert want
        for rank, res in got:
            merged = res[name]
            assert set(merged) == set(want)
            for i in want:       # HLL max, CMS sum, Bloom OR over both ranks == over the whole ensemble
                assert np.array_equal(np.asarray(merged[i], dtype=np.uint64), want[i].astype(np.uint64))


def _cell_worker(rank, world, port, per_rank, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import happysim_b200 as hs
    from happysim_b200 import distributed as D, engine, _abi as A
    import oracle_lib as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = hs.mmc_sweep(cs=(1, 2, 5), rhos=(0.5, 0.9))            # 6 cells
    rpc = 3
    out = O.oracle_run(model, O.make_params(seed=5, end_ns=20 * 10**9, n_replicas=per_rank, replica_index_base=rank * per_rank,
                                            replicas_per_cell=rpc, flags=A.HS_RUN_HISTOGRAM))
    cells = [(engine.totals_to_dict(t), h) for t, h in D.cell_totals_from_outputs(model, out, model.n_cells, rpc, rank * per_rank)]
    red = D.allreduce_cell_totals(cells)
    q.put((rank, [(d, h.tolist()) for d, h in red]))
    dist.destroy_process_group()


def test_two_rank_cell_allreduce_equals_single_process():
    """configs[4]'s aggregation path: per-cell totals + latency histograms, one collective, equal to the
    single-process reduction over the whole sweep (integers exactly, float sums in rank order)."""
    import happysim_b200 as hs
    from happysim_b200 import distributed as D, engine, _abi as A
    import oracle_lib as O
    per_rank, world, port = 18, 2, 33000 + os.getpid() % 2000      # 18 = one pass over 6 cells x 3 replicas per rank
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cell_worker, args=(r, world, port, per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = hs.mmc_sweep(cs=(1, 2, 5), rhos=(0.5, 0.9))
    whole = O.oracle_run(model, O.make_params(seed=5, end_ns=20 * 10**9, n_replicas=per_rank * world, replicas_per_cell=3,
                                              flags=A.HS_RUN_HISTOGRAM))
    want = [(engine.totals_to_dict(t), h) for t, h in D.cell_totals_from_outputs(model, whole, model.n_cells, 3, 0)]
    assert D.CELL_ALLREDUCE_CALLS == 1
    for rank, red in got:
        assert len(red) == len(want) == 6
        for (d, h), (wd, wh) in zip(red, want):
            for k, v in wd.items():
                if isinstance(v, float):
                    assert v == d[k] or abs(v - d[k]) <= 1e-12 * abs(v), k
                else:
                    assert v == d[k], k
            assert h == wh.tolist()
            assert d["replicas"] == 6 and d["events_processed"] > 0
    assert got[0][1] == got[1][1]            # identical on every rank, bit for bit


def test_pack_and_reduce_cell_words_roundtrip():
    from happysim_b200 import distributed as D
    cells = [({"events_processed": 10 + c, "sink_events": 3, "server_completions": 3, "source_ticks": 4, "dropped": 0,
               "replicas": 2, "replicas_flagged": 0, "sum_final_time_us": 99, "sum_latency": 0.1 * c, "sum_latency_sq": 0.3,
               "sum_service": 1.5, "min_latency": 0.01 * (c + 1), "max_latency": 2.0 + c},
              np.arange(64, dtype=np.uint64) * c) for c in range(4)]
    w = D.pack_cell_totals(cells)
    red = D.reduce_cell_words(np.stack([w, w]))
    for (d, h), (d1, h1) in zip(cells, red):
        assert d1["events_processed"] == 2 * d["events_processed"] and d1["sum_latency"] == d["sum_latency"] + d["sum_latency"]
        assert d1["min_latency"] == d["min_latency"] and d1["max_latency"] == d["max_latency"]
        assert (h1 == 2 * h).all()
