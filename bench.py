#!/usr/bin/env python
"""bench.py -- simulated events per wall-second of the hot path on B200 (BASELINE.json metric).

Workload at N=1: BASELINE.json configs[1], "65 536 independent M/M/1 replicas, 1e6 sim
seconds, 1xB200" (Source.poisson(8) -> Server(Exponential(0.1)) -> Sink).  One STEP advances
every replica by one window of simulated time (default 1e4 s of the 1e6 s horizon, about
3.96e10 events at 65 536 replicas): consecutive steps are consecutive windows of ONE
continuing run (state resident in HBM, resumed by the kernel), exactly the reference's
Simulation._run_window slicing; --steps 100 --window-s 1e4 covers the whole horizon.
At N>1 (torchrun, one rank per GPU) every rank runs its own 65 536 replicas (weak scaling,
global replica ids keep the Philox streams disjoint); the only collective is one NCCL
all-reduce of the fixed-layout summary vector after the run.

Modes (--mode):
  record   (default, the headline) flight-recorder ON: every processed event writes its 16 B
           record, every Sink sample 16 B and every service start 8 B to per-replica rings in
           HBM (19.2 B per event for M/M/1).  This is the mode the
           event-record HBM roofline of BASELINE.json refers to.
  summary  per-replica statistics only (negligible HBM traffic; issue-bound).
The JSON line reports the headline mode in value/roofline and the other mode under "other_mode".

--impl reference times the reference's CPU path: the reference is pure Python and cannot
travel to the GPU box, so this arm is the oracle port (oracle/hs_oracle.c, the C
restatement that is pinned event-by-event against the reference) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

METRIC = "simulated_events_per_second"
UNIT = "events/s"
RATE, MEAN = 8.0, 0.1
BYTES_EVENT, BYTES_SAMPLE, BYTES_SERVICE = 16, 16, 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="record", choices=["record", "summary"])
    ap.add_argument("--replicas", type=int, default=65536, help="replicas per GPU")
    ap.add_argument("--window-s", type=float, default=1.0e4, help="simulated seconds per step")
    ap.add_argument("--horizon-s", type=float, default=1.0e6)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-mode", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.p, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=2)
            except Exception:
                self.p.kill()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 8 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) > 8 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        pw = sorted(float(r[3]) for r in self.rows if len(r) > 8 and r[3].replace(".", "", 1).isdigit())
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_median": pw[len(pw) // 2] if pw else None}


# --------------------------------------------------------------------------- CPU arm

def cpu_oracle_throughput(budget_s, window_s, seed, threads=None):
    """Time the oracle port on the host cores over a bounded sample of the same workload
    (M/M/1 replicas of BASELINE configs[1], each simulated for `window_s` s).  Threads pull
    replicas from a shared counter until the time budget is spent, so the wall time is bounded
    whatever the box's core count or load."""
    import happysim_b200 as hs
    import oracle_lib as O
    from concurrent.futures import ThreadPoolExecutor
    import ctypes as C
    import itertools

    cores = threads or (os.cpu_count() or 1)
    model = hs.mm1(RATE, MEAN)
    d = model.desc()
    cap = 1 << 20
    p = O.make_params(seed=seed, end_ns=int(window_s * 1e9), n_replicas=cap, flags=0)
    import numpy as np
    summ = np.zeros(cap, O.A.SUMMARY_DTYPE)
    o = O.A.Outputs()
    o.summaries = summ.ctypes.data_as(C.POINTER(O.A.ReplicaSummary))
    counter = itertools.count()
    deadline = time.perf_counter() + budget_s
    done = []

    def work(_):
        n = 0
        while time.perf_counter() < deadline:
            k = next(counter)
            if k >= cap:
                break
            O.lib().hs_oracle_run_range(C.byref(d), C.byref(p), C.byref(o), k, k + 1)
            n += 1
        done.append(n)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(work, range(cores)))
    dt = time.perf_counter() - t0
    ev = int(summ["events_processed"].sum())
    n_rep = int((summ["events_processed"] > 0).sum())
    return {"value": ev / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_rep} M/M/1 replicas x {window_s:g} sim-s ({ev:.3e} events, {dt:.1f} s wall) on the "
                      f"oracle port (oracle/hs_oracle.c), {cores} threads"}, ev, dt


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    budget = 4.0
    win = min(a.window_s, 500.0)
    for _ in range(a.warmup):
        cpu_oracle_throughput(0.5, win, a.seed)
    tot_ev, tot_t, last = 0, 0.0, None
    for _ in range(a.steps):
        last, ev, dt = cpu_oracle_throughput(budget, win, a.seed)
        tot_ev += ev; tot_t += dt
    value = tot_ev / tot_t
    cb = dict(last); cb["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * tot_t / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a, 1),
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference is pure Python and cannot travel to the GPU box; this arm is its C "
                    "restatement (oracle port, pinned event-by-event against the reference) on all host cores, "
                    "each step a bounded sample of the same workload. In the build container the unmodified "
                    "Python reference ran this model at 1.7e5 events/s per core (BASELINE.md)."}
    print(json.dumps(line), flush=True)


def workload_config(a, world):
    return {"workload": "BASELINE configs[1]: 65 536 independent M/M/1 replicas (Source.poisson(8) -> "
                        "Server(ExponentialLatency(0.1)) -> Sink), 1e6 sim-s horizon, per GPU",
            "replicas_per_gpu": a.replicas, "replicas_total": a.replicas * world, "window_s": a.window_s,
            "horizon_s": a.horizon_s, "step": "one window of simulated time for every replica (resumed state)",
            "mode": a.mode, "parallelism": f"replicas sharded over {world} GPU(s), no data-path collective",
            "l2": "working set (replica state + queue rings + recorder rings) > 126 MB L2; no explicit flush"}


# --------------------------------------------------------------------------- GPU arm

def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference_arm(a, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import happysim_b200 as hs
    from happysim_b200 import engine, _abi as A

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.Stream()
    eng = engine.Engine(local, stream=stream.cuda_stream)
    model = hs.mm1(RATE, MEAN)
    eng.upload(model)
    n = a.replicas
    end_ns = int(a.horizon_s * 1e9)
    win_ns = int(a.window_s * 1e9)
    caps_by_mode = {"record": dict(record_cap=1024, sample_cap=128, service_cap=128), "summary": {}}

    def params(mode, k, resume):
        we = min((k + 1) * win_ns, end_ns)
        return engine.make_params(seed=a.seed, end_ns=end_ns, window_end_ns=(we if we < end_ns else -1),
                                  n_replicas=n, replica_index_base=rank * n, resume=resume, flags=0,
                                  **caps_by_mode[mode])

    def totals():
        return engine.totals_to_dict(eng.read_totals())

    def counters():
        o = eng.read_outputs()
        s = o["summaries"]
        return int(s["events_processed"].sum()), int(s["n_sink_samples"].sum()), int(s["n_service_samples"].sum()), \
            int((s["status"] != 0).sum())

    def timed_run(mode, steps, warmup):
        """W untimed + K timed consecutive windows of one continuing run, device-timed."""
        k = 0
        for _ in range(warmup):
            eng.run(params(mode, k, resume=int(k > 0))); k += 1
        eng.sync()
        c0 = counters()
        l0 = eng.launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launch_ms = []
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            eng.run(params(mode, k, resume=int(k > 0))); k += 1
            launch_ms.append(None)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        l1 = eng.launch_count()
        c1 = counters()
        return {"ms": ms, "wall_s": wall, "events": c1[0] - c0[0], "sink": c1[1] - c0[1], "service": c1[2] - c0[2],
                "flagged": c1[3], "launches": l1 - l0, "last_launch_ms": eng.last_run_ms()}

    # ---- headline mode, device-timed ---------------------------------------
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    res = timed_run(a.mode, a.steps, a.warmup)
    clocks = clk.stop() if rank == 0 else None

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t)

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t)

    ms = reduce_max(res["ms"])
    events = reduce_sum(float(res["events"]))
    value = events / (ms * 1e-3)

    # ---- the single end-of-run NCCL all-reduce of the summary vector -----------
    from happysim_b200 import distributed as D
    t = engine.totals_to_dict(D.allreduce_totals(eng.read_totals(), device="cuda"))
    agg = {"events_processed": t["events_processed"], "sink_events": t["sink_events"], "replicas": t["replicas"],
           "replicas_flagged": t["replicas_flagged"],
           "mean_latency_s": t["sum_latency"] / max(1, t["sink_events"]), "min_latency_s": t["min_latency"],
           "max_latency_s": t["max_latency"]}

    # ---- roofline of the dominant kernel (this rank's launches) -------------
    peak, peak_src = peaks()
    kernel = "hs_lane_kernel"
    per_launch_ms = res["ms"] / a.steps
    state_bytes = n * (2 * 512 + 56 + 3 * 64)
    if a.mode == "record":
        algo_bytes = (res["events"] * BYTES_EVENT + res["sink"] * BYTES_SAMPLE + res["service"] * BYTES_SERVICE) / a.steps + state_bytes
    else:
        algo_bytes = state_bytes
    achieved = algo_bytes / (per_launch_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:        # DRAM bytes per event measured by one `ncu --set full` capture of this kernel (profiles/)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)[a.mode]
        traffic = tj["dram_bytes"] / tj["events"] * res["events"] / a.steps
        traffic_src = tj["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": kernel, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes,
                "bytes_per_event": algo_bytes * a.steps / max(1, res["events"]),
                "note": ("record mode: 16 B/event + 16 B/Sink sample + 8 B/service start (SURVEY.md 8(d)) + replica "
                         "state load/store" if a.mode == "record" else
                         "summary mode writes only per-replica state/statistics: the kernel is issue-bound, "
                         "not HBM-bound, so this fraction is small by construction")}

    # ---- the other mode, for context ------------------------------------------
    other = None
    if not a.no_other_mode:
        om = "summary" if a.mode == "record" else "record"
        eng.upload(model)
        r2 = timed_run(om, max(2, min(a.steps, 5)), 3)
        ms2 = reduce_max(r2["ms"]); ev2 = reduce_sum(float(r2["events"]))
        other = {"mode": om, "value": ev2 / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / max(2, min(a.steps, 5))}

    # ---- end to end through the C-ABI with host buffers ------------------------
    eng.upload(model)
    host = None
    e2e_steps = max(2, min(a.steps, 5))
    p_e2e = params(a.mode, 0, resume=0)
    d2h = h2d = 0
    for it in range(2 + e2e_steps):
        if it == 2:
            barrier(); t0 = time.perf_counter(); ev_e2e = 0
        eng.upload(model)                               # H2D: the model table (entities, cells)
        eng.run(p_e2e)                                  # fresh first window
        if host is None:
            host = eng.alloc_host_outputs(p_e2e, pinned=True)
            host["records"] = host["sink_samples"] = host["service_samples"] = None   # rings stay on the device
        eng.read_outputs(host)                          # D2H: per-replica summaries + entity statistics
        if it >= 2:
            ev_e2e += int(host["summaries"]["events_processed"].sum())
    barrier()
    e2e_s = reduce_max(time.perf_counter() - t0)
    ev_e2e = reduce_sum(float(ev_e2e))
    h2d = model.entities.nbytes + model.backends.nbytes + model.key_table.nbytes
    d2h = host["summaries"].nbytes + host["entity_stats"].nbytes
    e2e = {"value": ev_e2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": e2e_steps, "what": "hs_model_upload + hs_run(first window, fresh) + hs_read_outputs(summaries, "
                                       "entity stats -> pinned host) per step, wall clock"}

    # ---- the other BASELINE configs (parity-test cases, not bench lines): device throughput for context
    others = None
    if world == 1 and not a.no_other_configs:
        others = {}
        def ctx(name, mdl, replicas, sim_s, what, **kw):
            e2 = engine.Engine(local, stream=stream.cuda_stream)
            e2.upload(mdl)
            best = None
            for _ in range(2):
                e2.run(engine.make_params(seed=a.seed, end_ns=int(sim_s * 1e9), n_replicas=replicas, flags=0, **kw))
                e2.sync()
                best = e2.last_run_ms() if best is None else min(best, e2.last_run_ms())
            o = e2.read_outputs()
            ev = int(o["summaries"]["events_processed"].sum())
            others[name] = {"value": ev / (best * 1e-3), "unit": UNIT, "replicas": replicas, "sim_seconds": sim_s,
                            "device_ms": best, "replicas_flagged": int((o["summaries"]["status"] != 0).sum()),
                            "what": what}
            e2.close()
        ctx("configs[2]", hs.lb_round_robin(64, 512.0), 16384, 10.0,
            "Source(512/s) -> LoadBalancer(RoundRobin) -> 64 x Server -> Sink; thread engine; 10 s slice of the 100 s run")
        tab = hs.consistent_hash_table([f"S{i}" for i in range(1024)], 100, 10000)
        ctx("configs[3]", hs.lb_key_table(tab, 1024, rate=8192.0), 1024, 2.0,
            "Source(8192/s, 10000 client ids) -> LoadBalancer(ConsistentHash, 100 vnodes) -> 1024 x Server -> Sink; "
            "thread engine; one GPU's 1024 replicas, 2 s slice of the 10 s run")
        ctx("configs[4]", hs.mmc_sweep(), 32768, 100.0,
            "M/M/c sweep, 256 (c, rho) cells x 128 seeds on one GPU; lane engine; 100 s slice of the 1000 s run",
            replicas_per_cell=128, queue_ring=4096)

    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(a, world),
                "roofline": roofline, "e2e": e2e, "gpu_launches": res["launches"], "clocks": clocks,
                "events_timed": events, "replicas_flagged": agg["replicas_flagged"], "aggregate": agg,
                "other_mode": other, "other_configs": others, "wall_s_timed_region": res["wall_s"]}
        if world == 1 and not a.no_cpu_baseline:
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
            cb, _, _ = cpu_oracle_throughput(12.0, min(a.window_s, 500.0), a.seed)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
