/* hs_lane_engine.cuh -- "lane engine": one THREAD per replica for the
 * single-server topology  Source -> Server(concurrency 1) -> Sink|Counter|nothing
 * (BASELINE.json configs[0] and configs[1], the headline M/M/1 ensemble).
 *
 * Why a lane and not a warp per replica: the whole future-event list of this
 * topology is {next SourceEvent, at most one ProcessContinuation, a handful of
 * same-timestamp protocol events}, i.e. a few registers.  Giving each replica a
 * lane keeps all 32 lanes of a warp doing useful event work; a warp per replica
 * would leave 31 lanes idle in every handler.  (The general "warp engine",
 * hs_warp_engine.cuh, covers models whose state does not fit a lane.)
 *
 * Exactness.  The loop below is the reference's pop-invoke-push loop
 * (happysimulator/core/simulation.py:449-505) with the future-event list held as
 *   T   the pending SourceEvent                     (time tT, sort index iT)
 *   C   the pending ProcessContinuation, if any     (time tC, sort index iC)
 *   nowq  events created at the current timestamp   (sort index, kind, payload)
 * Two execution paths produce the SAME processed-event sequence:
 *   - generic_step(): pops the (time, sort_index)-minimum of T, C and nowq and
 *     runs that one handler, exactly like the oracle;
 *   - the fused arrival / completion chains: when nowq is empty and neither T
 *     nor C shares the timestamp being processed, every event a handler creates
 *     at `now` is provably the next pop (ties between them are resolved by the
 *     creation order, which the straight-line code follows; the re-pushed payload
 *     keeps its OLD index and therefore sorts first, queue_driver.py:86-90), so the
 *     chain TICK -> ENQUEUE -> NOTIFY -> POLL -> DELIVER -> WORKER or
 *     CONTINUATION -> SINK -> POLL -> DELIVER -> WORKER is executed inline with the
 *     same counters, indices, hash and records.
 * Same-nanosecond ties fall back to generic_step().
 *
 * Reference handlers restated (paths under /root/reference/happysimulator):
 *   load/source.py:142-180, load/arrival_time_provider.py:66-82,
 *   components/queue.py:122-166, components/queue_driver.py:66-99,
 *   components/server/server.py:202-273, components/server/concurrency.py:100-128,
 *   components/common.py:36-44,92-95, core/event.py:277-325,465-533.
 */
#ifndef HS_LANE_ENGINE_CUH
#define HS_LANE_ENGINE_CUH

#include "hs_sampler.h"
#include "../../include/hs_b200.h"

#define HS_NOW_CAP 8
#define HS_LF_HASH 1      /* maintain the order hash                         */
#define HS_LF_REC 2       /* write event records / sink / service samples    */

struct hs_now_ev {        /* an event created at the current timestamp       */
    uint64_t idx;         /* Event._sort_index                               */
    int64_t created;      /* context["created_at"]                           */
    uint64_t payload_idx; /* DELIVER: sort index of the payload it carries   */
    int32_t kind;         /* HS_EV_*                                         */
    int32_t pad;
};

struct hs_ring_entry {    /* one queued request (FIFOQueue/LIFOQueue item)   */
    int64_t created;      /* context["created_at"]                           */
    uint64_t idx;         /* the queued Event's _sort_index                  */
};

struct __align__(16) hs_lane_state {   /* persisted between windows (512 B)  */
    int64_t now; uint64_t ctr; int64_t processed; uint64_t hash;
    int64_t tT; uint64_t iT; uint64_t arr_draws; int64_t gen_count; int64_t prov_count;
    int64_t tC; uint64_t iC; double svc_s; int64_t c_created;
    uint64_t svc_draws; int64_t accepted, dropped, completed, rejected; double total_service;
    int64_t received; double sum, sumsq, mn, mx;
    uint32_t q_head, q_len; int32_t active; uint32_t status;
    int64_t n_smp, n_svc; int32_t now_n; int32_t has_c;
    int32_t done; uint32_t rec_pos; double comp; uint32_t smp_pos, svc_pos; int64_t pad1;
    hs_now_ev nowq[HS_NOW_CAP];
};

struct hs_lane_model {
    int32_t src_id, srv_id, dst_id;   /* entity ids (dst_id < 0: no downstream) */
    int32_t dst_kind;                 /* HS_ENT_SINK / HS_ENT_COUNTER / 0       */
    int32_t arr_kind, svc_kind, policy, n_entities;
    int64_t capacity, stop_after;
    double rate, mean;
    uint32_t n_cells, pad;
    const double *cell_d0;            /* device pointers or NULL                */
};

struct hs_lane_run {
    uint64_t seed, seed_stride;
    uint32_t rid_base, rid_stride;
    int64_t end_ns, window_end_ns;
    uint32_t n_replicas, index_base, replicas_per_cell;
    uint32_t record_cap, sample_cap, service_cap, ring, resume;
};

struct hs_lane_out {
    hs_replica_summary *summaries;
    hs_entity_stats *stats;
    hs_event_record *records;
    hs_sink_sample *samples;
    double *service;
};

template <int FLAGS>
__global__ void __launch_bounds__(64)
hs_lane_kernel(hs_lane_model M, hs_lane_run P, hs_lane_state *__restrict__ states,
               hs_ring_entry *__restrict__ rings, hs_lane_out O)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.n_replicas) return;

    const uint32_t gidx = P.index_base + r;
    const uint64_t seed = P.seed + (uint64_t)gidx * P.seed_stride;
    const uint32_t rid = P.rid_base + gidx * P.rid_stride;
    const uint32_t sid_arr = HS_STREAM_ARRIVAL | ((uint32_t)M.src_id << 8);
    const uint32_t sid_svc = HS_STREAM_SERVICE | ((uint32_t)M.srv_id << 8);

    double rate = M.rate, mean = M.mean;
    if (M.n_cells) {
        uint32_t cell = (gidx / P.replicas_per_cell) % M.n_cells;
        rate = M.cell_d0[(size_t)cell * M.n_entities + M.src_id];
        mean = M.cell_d0[(size_t)cell * M.n_entities + M.srv_id];
    }
    const double lambda = HS_DIV(1.0, mean);          /* exponential.py:36 */
    const int64_t const_dur_ns = hs_seconds_to_ns(mean); /* constant.py:33-35 */
    const bool poisson = (M.arr_kind == HS_ARR_POISSON);
    const bool expo = (M.svc_kind == HS_SVC_EXPONENTIAL);
    const bool lifo = (M.policy == HS_Q_LIFO);
    const int64_t cap = M.capacity;
    const int dst_ev = (M.dst_kind == HS_ENT_SINK) ? HS_EV_REQ_SINK : HS_EV_REQ_COUNTER;
    const uint32_t ring_mask = P.ring - 1u;
    hs_ring_entry *ring = rings + (size_t)r * P.ring;
    const bool windowed = (P.window_end_ns >= 0 && P.window_end_ns < P.end_ns);

    hs_event_record *rec = (FLAGS & HS_LF_REC) && O.records ? O.records + (size_t)r * P.record_cap : nullptr;
    hs_sink_sample *smp = (FLAGS & HS_LF_REC) && O.samples ? O.samples + (size_t)r * P.sample_cap : nullptr;
    double *svc_out = (FLAGS & HS_LF_REC) && O.service ? O.service + (size_t)r * P.service_cap : nullptr;

    /* ---- replica state (registers; nowq in local memory, cold) ---------- */
    int64_t now, processed, tT, tC, c_created, gen_count, prov_count;
    uint64_t ctr, hash, iT, iC, arr_draws, svc_draws;
    int64_t accepted, dropped, completed, rejected, received;
    double svc_s, total_service, sum, comp, sumsq, mn, mx;
    uint32_t q_head, q_len, status, rec_pos, smp_pos, svc_pos;
    int64_t n_smp, n_svc;
    int32_t active, now_n, has_c;
    hs_now_ev nowq[HS_NOW_CAP];
    double arr_cache = 0.0, svc_cache = 0.0;

    hs_lane_state *S = states + r;
    if (P.resume) {
        if (S->done) return;
        now = S->now; ctr = S->ctr; processed = S->processed; hash = S->hash;
        tT = S->tT; iT = S->iT; arr_draws = S->arr_draws; gen_count = S->gen_count; prov_count = S->prov_count;
        tC = S->tC; iC = S->iC; svc_s = S->svc_s; c_created = S->c_created;
        svc_draws = S->svc_draws; accepted = S->accepted; dropped = S->dropped; completed = S->completed;
        rejected = S->rejected; total_service = S->total_service;
        received = S->received; sum = S->sum; comp = S->comp; sumsq = S->sumsq; mn = S->mn; mx = S->mx;
        q_head = S->q_head; q_len = S->q_len; active = S->active; status = S->status;
        n_smp = S->n_smp; n_svc = S->n_svc; now_n = S->now_n; has_c = S->has_c;
        rec_pos = S->rec_pos; smp_pos = S->smp_pos; svc_pos = S->svc_pos;
        for (int i = 0; i < HS_NOW_CAP; ++i) nowq[i] = S->nowq[i];
        double u0, u1;
        if (arr_draws & 1) { hs_uniform_pair(seed, rid, sid_arr, arr_draws >> 1, &u0, &u1); arr_cache = u1; }
        if (svc_draws & 1) { hs_uniform_pair(seed, rid, sid_svc, svc_draws >> 1, &u0, &u1); svc_cache = u1; }
    } else {
        now = 0; processed = 0; hash = HS_HASH_INIT;
        arr_draws = 0; svc_draws = 0; gen_count = 0; prov_count = 0;
        tC = 0; iC = 0; svc_s = 0.0; c_created = 0;
        accepted = dropped = completed = rejected = received = 0;
        total_service = 0.0; sum = 0.0; comp = 0.0; sumsq = 0.0;
        mn = __longlong_as_double(0x7ff0000000000000LL); mx = __longlong_as_double(0xfff0000000000000LL);
        q_head = 0; q_len = 0; active = 0; status = 0; n_smp = 0; n_svc = 0; now_n = 0; has_c = 0;
        rec_pos = 0; smp_pos = 0; svc_pos = 0;
        for (int i = 0; i < HS_NOW_CAP; ++i) { nowq[i].idx = 0; nowq[i].created = 0; nowq[i].payload_idx = 0; nowq[i].kind = 0; nowq[i].pad = 0; }
        /* Simulation.__init__: source.start() draws the first arrival and the
         * SourceEvent takes index 0 of the GLOBAL counter (simulation.py:77,145-154);
         * run() then restarts the per-heap counter at 0 (event_heap.py:48).        */
        double target = 1.0;
        if (poisson) {
            double u0, u1; hs_uniform_pair(seed, rid, sid_arr, 0, &u0, &u1);
            arr_cache = u1; arr_draws = 1; target = hs_exp1(u0);
        }
        tT = hs_next_arrival_ns(0, target, rate);
        iT = 0; ctr = 0;
    }

#define HS_EMIT(KIND, IDX, ENT)                                                              \
    do {                                                                                     \
        if (FLAGS & HS_LF_HASH) hash = hs_hash_step(hash, now, hs_record_word1((IDX), (KIND), (uint32_t)(ENT))); \
        if ((FLAGS & HS_LF_REC) && rec) {                                                    \
            hs_event_record rc_; rc_.time_ns = now; rc_.sort_index = (uint32_t)(IDX);        \
            rc_.kind = (uint8_t)(KIND); rc_.pad = 0; rc_.entity = (uint16_t)(ENT);           \
            rec[rec_pos] = rc_; rec_pos = (rec_pos + 1 == P.record_cap) ? 0u : rec_pos + 1;  \
        }                                                                                    \
        processed++;                                                                         \
    } while (0)

#define HS_DRAW(U, SID, N, CACHE)                                                            \
    do {                                                                                     \
        if ((N) & 1) { (U) = (CACHE); }                                                      \
        else { double u1_; hs_uniform_pair(seed, rid, (SID), (N) >> 1, &(U), &u1_); (CACHE) = u1_; } \
        (N)++;                                                                               \
    } while (0)

    /* Source.handle_event's arrival part: next SourceEvent time (source.py:166-170). */
#define HS_NEXT_TICK()                                                                       \
    do {                                                                                     \
        double target_ = 1.0;                                                                \
        if (poisson) { double u_; HS_DRAW(u_, sid_arr, arr_draws, arr_cache); target_ = hs_exp1(u_); } \
        tT = hs_next_arrival_ns(tT, target_, rate);                                          \
        iT = ctr++;                                                                          \
    } while (0)

    /* Server.handle_queued_event up to its yield, for the payload (CREATED):
     * inline ProcessContinuation index, acquire (the caller has checked
     * active < 1), sample, schedule resume
     * (server.py:217-253, event.py:314-325,499-508).                            */
#define HS_SERVICE_START(CREATED)                                                            \
    do {                                                                                     \
        ctr++;                                                                               \
        active++;                                                                            \
        int64_t dur_;                                                                        \
        if (expo) { double u_; HS_DRAW(u_, sid_svc, svc_draws, svc_cache); dur_ = hs_exp_latency_ns(u_, lambda); } \
        else dur_ = const_dur_ns;                                                            \
        svc_s = hs_ns_to_seconds(dur_);                                                      \
        if ((FLAGS & HS_LF_REC) && svc_out) { svc_out[svc_pos] = svc_s; svc_pos = (svc_pos + 1 == P.service_cap) ? 0u : svc_pos + 1; } \
        n_svc++;                                                                             \
        tC = hs_resume_ns(now, svc_s); iC = ctr++; c_created = (CREATED); has_c = 1;         \
    } while (0)

#define HS_SINK(CREATED)                                                                     \
    do {                                                                                     \
        received++;                                                                          \
        if (M.dst_kind == HS_ENT_SINK) {                                                     \
            double lat_ = hs_ns_to_seconds(now - (CREATED));                                 \
            hs_neumaier_add(&sum, &comp, lat_); sumsq = HS_ADD(sumsq, HS_MUL(lat_, lat_));              \
            if (lat_ < mn) mn = lat_;                                                        \
            if (lat_ > mx) mx = lat_;                                                        \
            if ((FLAGS & HS_LF_REC) && smp) { hs_sink_sample q_; q_.completion_ns = now; q_.latency_s = lat_; smp[smp_pos] = q_; \
                smp_pos = (smp_pos + 1 == P.sample_cap) ? 0u : smp_pos + 1; }                \
            n_smp++;                                                                         \
        }                                                                                    \
    } while (0)

#define HS_PUSH_NOW(KIND, IDX, CREATED, PIDX)                                                \
    do {                                                                                     \
        if (now_n >= HS_NOW_CAP) { status |= HS_ST_FEL_OVERFLOW; }                           \
        else { nowq[now_n].idx = (IDX); nowq[now_n].created = (CREATED); nowq[now_n].payload_idx = (PIDX); \
               nowq[now_n].kind = (KIND); now_n++; }                                         \
    } while (0)

    bool paused = false;
    while (true) {
        if (!(now <= P.end_ns)) break;                 /* simulation.py:472 */
        if (status & (HS_ST_QUEUE_OVERFLOW | HS_ST_FEL_OVERFLOW)) break;

        if (now_n == 0 && tT == INT64_MAX && !has_c) break;      /* heap exhausted */
        if (now_n == 0) {
            const bool pickC = has_c && (tC < tT || (tC == tT && iC < iT));
            const int64_t tn = pickC ? tC : tT;
            if (windowed && tn > P.window_end_ns) { paused = true; break; }
            const bool slow = (has_c && tC == tT) || (tn > P.end_ns) || (tn < now);
            if (!slow) {
                now = tn;
                if (!pickC) {
                    /* ===== fused arrival chain ================================== */
                    HS_EMIT(HS_EV_SOURCE_TICK, iT, M.src_id);
                    const bool payload = !(M.stop_after >= 0 && now > M.stop_after);  /* source.py:68 */
                    uint64_t idxP = 0;
                    if (payload) { prov_count++; idxP = ctr++; }
                    gen_count++;
                    HS_NEXT_TICK();
                    if (!payload) continue;
                    if (tT <= now) {          /* zero inter-arrival: the new tick ties with the chain */
                        HS_PUSH_NOW(HS_EV_REQ_ENQUEUE, idxP, now, 0);
                        continue;
                    }
                    /* Queue._handle_enqueue (queue.py:122-147) */
                    HS_EMIT(HS_EV_REQ_ENQUEUE, idxP, M.srv_id);
                    const bool was_empty = (q_len == 0);
                    if (cap >= 0 && (int64_t)q_len >= cap) { dropped++; continue; }
                    if (q_len >= P.ring) { status |= HS_ST_QUEUE_OVERFLOW; continue; }
                    accepted++;
                    if (!was_empty || active >= 1) {
                        /* request waits in the buffer */
                        hs_ring_entry e; e.created = now; e.idx = idxP;
                        ring[(q_head + q_len) & ring_mask] = e; q_len++;
                        if (was_empty) {      /* notify, but the worker is busy: no poll (queue_driver.py:92-96) */
                            uint64_t idxN = ctr++;
                            HS_EMIT(HS_EV_NOTIFY, idxN, M.srv_id);
                        }
                        continue;
                    }
                    /* buffer was empty and the worker is idle: NOTIFY -> POLL -> DELIVER -> WORKER;
                     * the item is pushed and popped again at once (FIFO and LIFO agree).          */
                    { uint64_t i_ = ctr++; HS_EMIT(HS_EV_NOTIFY, i_, M.srv_id); }
                    { uint64_t i_ = ctr++; HS_EMIT(HS_EV_POLL, i_, M.srv_id); }
                    { uint64_t i_ = ctr++; HS_EMIT(HS_EV_DELIVER, i_, M.srv_id); }
                    HS_EMIT(HS_EV_REQ_WORKER, idxP, M.srv_id);
                    HS_SERVICE_START(now);
                    continue;
                } else {
                    /* ===== fused completion chain =============================== */
                    HS_EMIT(HS_EV_CONTINUATION, iC, M.srv_id);
                    has_c = 0;
                    active = active > 0 ? active - 1 : 0;      /* FixedConcurrency.release */
                    completed++;
                    total_service = HS_ADD(total_service, svc_s);
                    uint64_t idxF = 0;
                    if (M.dst_id >= 0) idxF = ctr++;           /* Entity.forward */
                    uint64_t idxPoll = 0; bool poll = (active < 1);
                    if (poll) idxPoll = ctr++;                 /* schedule_poll hook */
                    if (M.dst_id >= 0) { HS_EMIT(dst_ev, idxF, M.dst_id); HS_SINK(c_created); }
                    if (!poll) continue;
                    HS_EMIT(HS_EV_POLL, idxPoll, M.srv_id);
                    if (q_len == 0) continue;                  /* Queue._handle_poll: empty */
                    hs_ring_entry it;
                    if (lifo) { it = ring[(q_head + q_len - 1) & ring_mask]; }
                    else { it = ring[q_head & ring_mask]; q_head++; }
                    q_len--;
                    { uint64_t i_ = ctr++; HS_EMIT(HS_EV_DELIVER, i_, M.srv_id); }
                    HS_EMIT(HS_EV_REQ_WORKER, it.idx, M.srv_id);
                    HS_SERVICE_START(it.created);
                    continue;
                }
            }
        }

        /* ===== generic single-event step (ties, run end, leftovers) ========= */
        {
            /* pop the (time, sort_index) minimum of T, C and nowq (event.py:337-344) */
            int which = -1;                 /* -1 T, -2 C, >=0 nowq slot */
            int64_t bt = tT; uint64_t bi = iT;
            if (has_c && (tC < bt || (tC == bt && iC < bi))) { which = -2; bt = tC; bi = iC; }
            for (int i = 0; i < now_n; ++i) {
                if (now < bt || (now == bt && nowq[i].idx < bi)) { which = i; bt = now; bi = nowq[i].idx; }
            }
            if (windowed && bt > P.window_end_ns) { paused = true; break; }
            if (bt < now) {                 /* "time travel": popped and skipped, not processed
                                               (simulation.py:479-489); only a SourceEvent can do it */
                tT = INT64_MAX; continue;
            }
            int kind; int64_t e_created = 0; uint64_t e_pidx = 0;
            if (which == -1) kind = HS_EV_SOURCE_TICK;
            else if (which == -2) kind = HS_EV_CONTINUATION;
            else {
                kind = nowq[which].kind; e_created = nowq[which].created; e_pidx = nowq[which].payload_idx;
                now_n--; nowq[which] = nowq[now_n];
            }
            now = bt;
            switch (kind) {
            case HS_EV_SOURCE_TICK: {
                HS_EMIT(HS_EV_SOURCE_TICK, bi, M.src_id);
                const bool payload = !(M.stop_after >= 0 && now > M.stop_after);
                uint64_t idxP = 0;
                if (payload) { prov_count++; idxP = ctr++; }
                gen_count++;
                HS_NEXT_TICK();
                if (payload) HS_PUSH_NOW(HS_EV_REQ_ENQUEUE, idxP, now, 0);
                break;
            }
            case HS_EV_REQ_ENQUEUE: {
                HS_EMIT(HS_EV_REQ_ENQUEUE, bi, M.srv_id);
                const bool was_empty = (q_len == 0);
                if (cap >= 0 && (int64_t)q_len >= cap) { dropped++; break; }
                if (q_len >= P.ring) { status |= HS_ST_QUEUE_OVERFLOW; break; }
                hs_ring_entry e; e.created = e_created; e.idx = bi;
                ring[(q_head + q_len) & ring_mask] = e; q_len++;
                accepted++;
                if (was_empty) { uint64_t i_ = ctr++; HS_PUSH_NOW(HS_EV_NOTIFY, i_, 0, 0); }
                break;
            }
            case HS_EV_NOTIFY:
                HS_EMIT(HS_EV_NOTIFY, bi, M.srv_id);
                if (active < 1) { uint64_t i_ = ctr++; HS_PUSH_NOW(HS_EV_POLL, i_, 0, 0); }
                break;
            case HS_EV_POLL:
                HS_EMIT(HS_EV_POLL, bi, M.srv_id);
                if (q_len > 0) {
                    hs_ring_entry it;
                    if (lifo) { it = ring[(q_head + q_len - 1) & ring_mask]; }
                    else { it = ring[q_head & ring_mask]; q_head++; }
                    q_len--;
                    uint64_t i_ = ctr++;
                    HS_PUSH_NOW(HS_EV_DELIVER, i_, it.created, it.idx);
                }
                break;
            case HS_EV_DELIVER:
                HS_EMIT(HS_EV_DELIVER, bi, M.srv_id);
                HS_PUSH_NOW(HS_EV_REQ_WORKER, e_pidx, e_created, 0);   /* payload keeps its old index */
                break;
            case HS_EV_REQ_WORKER:
                HS_EMIT(HS_EV_REQ_WORKER, bi, M.srv_id);
                if (active >= 1) {          /* acquire failed (server.py:223-234): hooks still run */
                    ctr++; rejected++; status |= HS_ST_REJECT_PATH;
                    if (active < 1) { uint64_t i_ = ctr++; HS_PUSH_NOW(HS_EV_POLL, i_, 0, 0); }
                } else {
                    HS_SERVICE_START(e_created);
                }
                break;
            case HS_EV_CONTINUATION: {
                HS_EMIT(HS_EV_CONTINUATION, bi, M.srv_id);
                has_c = 0;
                active = active > 0 ? active - 1 : 0;
                completed++;
                total_service = HS_ADD(total_service, svc_s);
                if (M.dst_id >= 0) { uint64_t i_ = ctr++; HS_PUSH_NOW(dst_ev, i_, c_created, 0); }
                if (active < 1) { uint64_t i_ = ctr++; HS_PUSH_NOW(HS_EV_POLL, i_, 0, 0); }
                break;
            }
            case HS_EV_REQ_SINK:
            case HS_EV_REQ_COUNTER:
                HS_EMIT(kind, bi, M.dst_id);
                HS_SINK(e_created);
                break;
            default: break;
            }
        }
    }

    /* ---- persist / publish --------------------------------------------- */
    S->now = now; S->ctr = ctr; S->processed = processed; S->hash = hash;
    S->tT = tT; S->iT = iT; S->arr_draws = arr_draws; S->gen_count = gen_count; S->prov_count = prov_count;
    S->tC = tC; S->iC = iC; S->svc_s = svc_s; S->c_created = c_created;
    S->svc_draws = svc_draws; S->accepted = accepted; S->dropped = dropped; S->completed = completed;
    S->rejected = rejected; S->total_service = total_service;
    S->received = received; S->sum = sum; S->comp = comp; S->sumsq = sumsq; S->mn = mn; S->mx = mx;
    S->q_head = q_head; S->q_len = q_len; S->active = active; S->status = status;
    S->n_smp = n_smp; S->n_svc = n_svc; S->now_n = now_n; S->has_c = has_c;
    S->rec_pos = rec_pos; S->smp_pos = smp_pos; S->svc_pos = svc_pos;
    S->done = paused ? 0 : 1;
    for (int i = 0; i < HS_NOW_CAP; ++i) S->nowq[i] = nowq[i];

    if (O.summaries) {
        hs_replica_summary s;
        s.events_processed = processed; s.final_time_ns = now;
        s.order_hash = (FLAGS & HS_LF_HASH) ? hash : 0ULL;
        s.next_sort_index = ctr; s.n_sink_samples = n_smp; s.n_service_samples = n_svc; s.heap_left = (tT != INT64_MAX) + has_c + now_n; s.status = status;
        O.summaries[r] = s;
    }
    if (O.stats) {
        hs_entity_stats *st = O.stats + (size_t)r * M.n_entities;
        hs_entity_stats a; a.c0 = gen_count; a.c1 = prov_count; a.c2 = 0; a.c3 = 0; a.f0 = a.f1 = a.f2 = a.f3 = 0.0;
        st[M.src_id] = a;
        a.c0 = accepted; a.c1 = dropped; a.c2 = completed; a.c3 = rejected; a.f0 = total_service;
        st[M.srv_id] = a;
        if (M.dst_id >= 0) {
            a.c0 = received; a.c1 = a.c2 = a.c3 = 0;
            if (M.dst_kind == HS_ENT_SINK) { a.f0 = hs_neumaier_result(sum, comp); a.f1 = sumsq; a.f2 = mn; a.f3 = mx; }
            else { a.f0 = a.f1 = a.f2 = a.f3 = 0.0; }
            st[M.dst_id] = a;
        }
    }
#undef HS_EMIT
#undef HS_DRAW
#undef HS_NEXT_TICK
#undef HS_SERVICE_START
#undef HS_SINK
#undef HS_PUSH_NOW
}

#endif /* HS_LANE_ENGINE_CUH */
