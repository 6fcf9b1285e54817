"""Multi-GPU plumbing: replicas are independent, so the path shards with no data-path
collective.  Rank r runs the contiguous global replica range
``shard_range(n, r, world)``; Philox counters carry the GLOBAL replica index
(``hs_run_params.replica_index_base``), so results do not depend on the number of GPUs.
After the run ONE all-reduce aggregates the fixed-layout ``hs_totals`` vector: sums for the
int64 / float64 accumulators, min / max for the latency extrema.  Backend: NCCL on GPUs,
gloo in the CPU tests (tests/test_distributed_gloo.py)."""
from __future__ import annotations

import numpy as np

from . import _abi as A


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split r in [rank*n/world, (rank+1)*n/world) (remainder to the low ranks)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def totals_from_outputs(model, out) -> A.Totals:
    """numpy restatement of csrc/hs_totals.cuh (host-side check of the device reduction)."""
    t = A.Totals()
    s, st = out["summaries"], out["entity_stats"]
    kinds = model.entities["kind"]
    t.i[0] = int(s["events_processed"].sum())
    t.i[5] = len(s)
    t.i[6] = int((s["status"] != 0).sum())
    t.i[7] = int((s["final_time_ns"] // 1000).sum())
    t.fmin, t.fmax = float("inf"), float("-inf")
    for e in range(model.n_entities):
        k, col = int(kinds[e]), st[:, e]
        if k == A.HS_ENT_SINK:
            t.i[1] += int(col["c0"].sum()); t.fsum[0] += float(col["f0"].sum()); t.fsum[1] += float(col["f1"].sum())
            t.fmin = min(t.fmin, float(col["f2"].min())); t.fmax = max(t.fmax, float(col["f3"].max()))
        elif k == A.HS_ENT_SERVER:
            t.i[2] += int(col["c2"].sum()); t.i[4] += int(col["c1"].sum()); t.fsum[2] += float(col["f0"].sum())
        elif k == A.HS_ENT_SOURCE:
            t.i[3] += int(col["c0"].sum())
    return t


def _gather_words(words: np.ndarray, device, group):
    """ONE collective: every rank contributes its packed int64 vector (float64 values travel as their bit
    patterns) and receives all of them, [world, n_words].  The reduction is then done locally in rank order,
    which makes the float sums deterministic and identical on every rank (an NCCL SUM all-reduce would also
    need separate MIN / MAX calls for the extrema)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    mine = torch.from_numpy(np.ascontiguousarray(words, dtype=np.int64)).to(dev)
    allv = torch.empty(world * mine.numel(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allv, mine, group=group)
    return allv.cpu().numpy().reshape(world, -1)


TOTALS_ALLREDUCE_CALLS = 1
CELL_ALLREDUCE_CALLS = 1


def _dist_active(group) -> bool:
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def allreduce_totals(t: A.Totals, device=None, group=None) -> A.Totals:
    """The single end-of-run collective (one all-gather of the packed 13-word vector, reduced in rank order).
    No-op when torch.distributed is not initialised."""
    if not _dist_active(group):
        return t
    ni, nf = A.HS_TOTALS_I64, A.HS_TOTALS_F64_SUM
    w = np.zeros(ni + nf + 2, np.int64)
    w[:ni] = list(t.i)
    w[ni:] = np.array(list(t.fsum) + [t.fmin, t.fmax], np.float64).view(np.int64)
    g = _gather_words(w, device, group)
    f = g[:, ni:].copy().view(np.float64)
    r = A.Totals()
    for k in range(ni):
        r.i[k] = int(g[:, k].sum())
    for k in range(nf):
        acc = 0.0
        for rank in range(g.shape[0]):
            acc += float(f[rank, k])
        r.fsum[k] = acc
    r.fmin, r.fmax = float(f[:, nf].min()), float(f[:, nf + 1].max())
    return r


def run_sharded(engine, model, params_fn, n_total: int, rank: int, world: int, group=None):
    """Run this rank's shard of an ``n_total``-replica ensemble and return (local outputs,
    all-reduced totals).  ``params_fn(n_local, index_base)`` builds the hs_run_params."""
    lo, hi = shard_range(n_total, rank, world)
    engine.upload(model)
    engine.run(params_fn(hi - lo, lo))
    out = engine.read_outputs()
    return out, allreduce_totals(engine.read_totals(), group=group)


def cell_totals_from_outputs(model, out, n_cells: int, replicas_per_cell: int, index_base: int = 0):
    """numpy restatement of hs_cell_totals_kernel: [(Totals, uint64[64])] per cell."""
    n = len(out["summaries"])
    cells = ((index_base + np.arange(n)) // replicas_per_cell) % n_cells
    res = []
    for c in range(n_cells):
        sel = cells == c
        sub = {"summaries": out["summaries"][sel], "entity_stats": out["entity_stats"][sel]}
        t = totals_from_outputs(model, sub) if sel.any() else A.Totals()
        if not sel.any():
            t.fmin, t.fmax = float("inf"), float("-inf")
        h = out["histograms"][sel].sum(axis=0, dtype=np.uint64) if out.get("histograms") is not None else np.zeros(64, np.uint64)
        res.append((t, h))
    return res


_CELL_INTS = ["events_processed", "sink_events", "server_completions", "source_ticks", "dropped", "replicas",
              "replicas_flagged", "sum_final_time_us"]
_CELL_FLOATS = ["sum_latency", "sum_latency_sq", "sum_service", "min_latency", "max_latency"]


def pack_cell_totals(cells) -> np.ndarray:
    """[(totals dict, uint64[64] histogram)] per cell -> int64[n_cells, 8 + 64 + 5] (floats as bit patterns)."""
    ni, nh, nf = len(_CELL_INTS), A.HS_HISTOGRAM_BINS, len(_CELL_FLOATS)
    w = np.zeros((len(cells), ni + nh + nf), np.int64)
    for c, (d, h) in enumerate(cells):
        w[c, :ni] = [d[k] for k in _CELL_INTS]
        w[c, ni:ni + nh] = np.asarray(h, np.uint64).astype(np.int64)
        w[c, ni + nh:] = np.array([d[k] for k in _CELL_FLOATS], np.float64).view(np.int64)
    return w


def reduce_cell_words(g: np.ndarray):
    """[world, n_cells, words] -> the all-reduced cells, reduced in rank order (deterministic float sums)."""
    ni, nh = len(_CELL_INTS), A.HS_HISTOGRAM_BINS
    ints = g[:, :, :ni + nh].sum(axis=0)
    f = np.ascontiguousarray(g[:, :, ni + nh:]).view(np.float64)
    out = []
    for c in range(g.shape[1]):
        d = {k: int(ints[c, j]) for j, k in enumerate(_CELL_INTS)}
        for j, k in enumerate(_CELL_FLOATS[:3]):
            acc = 0.0
            for rank in range(g.shape[0]):
                acc += float(f[rank, c, j])
            d[k] = acc
        d["min_latency"], d["max_latency"] = float(f[:, c, 3].min()), float(f[:, c, 4].max())
        out.append((d, ints[c, ni:].astype(np.uint64)))
    return out


def allreduce_cell_totals(cells, device=None, group=None):
    """One collective for a sweep (configs[4]): every cell's totals vector and latency histogram, packed into
    one int64 buffer, all-gathered once and reduced in rank order on every rank."""
    if not _dist_active(group):
        return cells
    w = pack_cell_totals(cells)
    g = _gather_words(w.ravel(), device, group).reshape(-1, *w.shape)
    return reduce_cell_words(g)


def histogram_percentile(hist, p: float) -> float:
    """Latency (seconds) below which a fraction p of the histogram's samples fall; linear inside a bin."""
    hist = np.asarray(hist, dtype=np.float64)
    total = hist.sum()
    if total == 0:
        return 0.0
    edges = A.histogram_bin_edges_ns().astype(np.float64)
    upper = np.append(edges[1:], edges[-1] * 1.5)
    cum = np.cumsum(hist)
    k = int(np.searchsorted(cum, p * total, side="left"))
    k = min(k, 63)
    below = cum[k - 1] if k > 0 else 0.0
    frac = (p * total - below) / hist[k] if hist[k] > 0 else 0.0
    return float(edges[k] + frac * (upper[k] - edges[k])) / 1e9


def merge_sketch_states(model, raw: np.ndarray) -> dict:
    """Host restatement of Engine.read_sketches for per-replica states (hs_outputs.sketches of any
    party): HyperLogLog.merge = register max, CountMinSketch.merge = counter sum, BloomFilter.merge = OR
    over the replicas; TOPK rows (no device image) become a ``TopK`` object merged in replica order."""
    from .sketching import ReservoirSampler, TDigest, TopK
    out = {}
    for i, v in model.sketch_views(raw).items():
        algo = int(model.entities["i0"][i])
        if algo == A.HS_SK_HLL:
            out[i] = v.max(axis=0).astype(np.uint8)
        elif algo == A.HS_SK_BLOOM:
            out[i] = np.bitwise_or.reduce(v, axis=0)
        elif algo == A.HS_SK_TDIGEST:        # TDigest.merge re-compresses: sequential, replica 0, 1, 2 ...
            acc = TDigest(float(model.entities["d0"][i]))
            for r in range(v.shape[0]):
                o = TDigest(acc.compression)
                o._load_device_state(v[r])
                acc.merge(o)
            out[i] = acc
        elif algo == A.HS_SK_RESERVOIR:      # ReservoirSampler.merge draws from the accumulator's generator: replica
            acc = ReservoirSampler(int(model.entities["i2"][i]))      # order, starting from replica 0's own state
            for r in range(v.shape[0]):
                o = ReservoirSampler(acc.capacity)
                o._load_device_state(v[r], 0)
                if r == 0:
                    acc = o
                else:
                    acc.merge(o)
            out[i] = acc
        elif algo == A.HS_SK_TOPK:           # TopK.merge is sequential and order dependent: replica 0, 1, 2 ...
            acc = TopK(int(model.entities["i2"][i]))
            for r in range(v.shape[0]):
                o = TopK(acc.k)
                o._load_device_state(v[r], int(v[r][2:2 + 3 * int(v[r][0]):3].sum()))
                acc.merge(o)
            out[i] = acc
        else:
            out[i] = v.astype(np.uint64).sum(axis=0)
    return out


def allreduce_sketches(model, merged: dict, device=None, group=None) -> dict:
    """Cross-GPU merge of the per-rank merged sketches: one MAX all-reduce over the HLL registers, one SUM
    all-reduce over the CMS counters and one MAX all-reduce over the Bloom filters' unpacked bits (= bitwise OR;
    NCCL has no bitwise reductions) (the reference's
    merge() contracts, hyperloglog.py:203-226, count_min_sketch.py:276-301, bloom_filter.py:262-291).
    TopK.merge / TDigest.merge are sequential and order dependent: those entries stay rank-local here (gather the per-replica
    states and merge them in global replica order if a cross-rank TopK is wanted)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merged
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    host_only = (A.HS_SK_TOPK, A.HS_SK_TDIGEST, A.HS_SK_RESERVOIR)
    ids = sorted(i for i in merged if int(model.entities["i0"][i]) not in host_only)
    hll = [i for i in ids if int(model.entities["i0"][i]) == A.HS_SK_HLL]
    blm = [i for i in ids if int(model.entities["i0"][i]) == A.HS_SK_BLOOM]
    cms = [i for i in ids if i not in hll and i not in blm]
    out = {i: v for i, v in merged.items() if int(model.entities["i0"][i]) in host_only}   # rank-local (see docstring)
    if blm:      # NCCL has no bitwise reductions: OR over the bit arrays = MAX over their bits, one byte per bit
        bits = np.concatenate([np.unpackbits(merged[i].view(np.uint8), bitorder="little") for i in blm])
        t = torch.from_numpy(bits).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        a, pos = t.cpu().numpy(), 0
        for i in blm:
            n = merged[i].size * 64
            out[i] = np.packbits(a[pos: pos + n], bitorder="little").view(np.uint64).copy(); pos += n
    if hll:
        t = torch.from_numpy(np.concatenate([merged[i].astype(np.int32).ravel() for i in hll])).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        a, pos = t.cpu().numpy(), 0
        for i in hll:
            out[i] = a[pos: pos + merged[i].size].astype(np.uint8); pos += merged[i].size
    if cms:
        t = torch.from_numpy(np.concatenate([merged[i].astype(np.int64).ravel() for i in cms])).to(dev)
        dist.all_reduce(t, group=group)
        a, pos = t.cpu().numpy(), 0
        for i in cms:
            out[i] = a[pos: pos + merged[i].size].astype(np.uint64).reshape(merged[i].shape); pos += merged[i].size
    return out
