/* hs_thread_engine.cuh -- "thread engine": one THREAD per replica for ANY lowered model
 * (load-balanced farms, tandem queues, several sources, probes ...).
 *
 * The warp engine gives a replica a whole warp but its handlers are scalar control flow, so 31
 * lanes idle; here every lane runs its own replica.  The replica's state cannot live in registers
 * (a 64-server farm is ~11 KB), so it stays in a contiguous per-replica block in HBM
 *     [ header 128 B | entity state n x 96 B | future heap, S x 48 B | now tier, 24 x 48 B ]
 * that the lane reads and writes through L1/L2 (a block is touched by exactly one thread, so there
 * is nothing to stage or synchronise, and a paused window resumes from the very same bytes).
 * Pending events are kept in two tiers that together order exactly like the reference's heap:
 *   now tier     events created at the current timestamp, a small array scanned by sort index;
 *   future heap  a binary min-heap on (time, sort_index) for SourceEvents / ProcessContinuations
 *                (two pops and two pushes per request on a Source -> Server path).
 * The handlers are the shared restatement in hs_handlers.inc.
 *
 * Bound: L2/HBM latency of a dependent chain per event, hidden by running one replica per lane
 * on as many lanes as the ensemble provides.
 */
#ifndef HS_THREAD_ENGINE_CUH
#define HS_THREAD_ENGINE_CUH

#include "hs_warp_engine.cuh"       /* hs_warp_hdr, hs_went, hs_wnow, hs_wring_entry, model/run/out structs */

#define HS_THREAD_BLOCK 64

template <int FLAGS>
__global__ void __launch_bounds__(HS_THREAD_BLOCK)
hs_thread_kernel(hs_warp_model M, hs_warp_run P, unsigned char *__restrict__ blocks,
                 hs_wring_entry *__restrict__ rings, hs_warp_out O)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.n_replicas) return;
    const uint32_t S = M.fel_slots;                      /* future-heap capacity */
    const uint32_t ne = M.n_entities;
    const hs_entity_desc *ENTS = M.ents;
    const int32_t *SRVIDX = M.srv_index, *BACKENDS = M.backends;

    unsigned char *blk = blocks + (size_t)r * M.block_bytes;
    hs_warp_hdr *Hg = (hs_warp_hdr *)blk;
    hs_went *E = (hs_went *)(blk + sizeof(hs_warp_hdr));
    hs_wnow *heap = (hs_wnow *)(blk + sizeof(hs_warp_hdr) + (size_t)ne * sizeof(hs_went));
    hs_wnow *N = heap + S;

    const uint32_t gidx = P.index_base + r;
    const uint64_t seed = P.seed + (uint64_t)gidx * P.seed_stride;
    const uint32_t rid = P.rid_base + gidx * P.rid_stride;
    hs_wring_entry *ring0 = rings + (size_t)r * M.n_servers * P.ring;
    const uint32_t ring_mask = P.ring - 1u;
    const bool windowed = (P.window_end_ns >= 0 && P.window_end_ns < P.end_ns);

    hs_warp_hdr hdr;                                     /* working copy of the header */
    hs_warp_hdr *H = &hdr;
    if (P.resume) {
        hdr = *Hg;
        if (hdr.done) return;
    } else {
        for (uint32_t i = 0; i < M.block_bytes / 16; ++i) ((uint4 *)blk)[i] = make_uint4(0u, 0u, 0u, 0u);
        memset(&hdr, 0, sizeof hdr);
        const uint32_t cell = M.n_cells ? (gidx / P.replicas_per_cell) % M.n_cells : 0u;
        for (uint32_t i = 0; i < ne; ++i) {
            const hs_entity_desc d = ENTS[i];
            hs_went *e = &E[i];
            e->d0 = M.n_cells ? M.cell_d0[(size_t)cell * ne + i] : d.d0;
            e->i0 = M.n_cells ? M.cell_i0[(size_t)cell * ne + i] : d.i0;
            e->lambda = (d.kind == HS_ENT_SERVER && d.i2 == HS_SVC_EXPONENTIAL) ? HS_DIV(1.0, e->d0) : 0.0;
            if (d.kind == HS_ENT_SINK || d.kind == HS_ENT_PROBE) {
                e->u.snk.mn = __longlong_as_double(0x7ff0000000000000LL);
                e->u.snk.mx = __longlong_as_double(0xfff0000000000000LL);
            }
        }
        hdr.hash = HS_HASH_INIT;
        /* Simulation.__init__: source.start() in order; bootstrap indices come from the global
         * counter (simulation.py:77,145-154), run() restarts the per-heap one at 0. */
        uint64_t boot = 0;
        for (uint32_t i = 0; i < ne; ++i) {
            if (ENTS[i].kind != HS_ENT_SOURCE) continue;
            hs_went *e = &E[i];
            double target = 1.0;
            if (e->i0 == HS_ARR_POISSON && P.trace_arr) {
                if (hdr.np_cursor >= P.n_trace_arr) { hdr.status |= HS_ST_TRACE_EXHAUSTED; break; }
                target = P.trace_arr[(size_t)r * P.n_trace_arr + hdr.np_cursor++]; e->u.src.arr_draws++;
            } else if (e->i0 == HS_ARR_POISSON) {
                target = hs_exp1(hs_uniform(seed, rid, HS_STREAM_ARRIVAL | (i << 8), e->u.src.arr_draws++));
            }
            const int32_t pi = ENTS[i].i3;
            int64_t first;
            if ((FLAGS & HS_WF_PROFILE) && pi > 0) first = hs_next_arrival_profile_ns(&M.profiles[pi - 1], 0, target);
            else first = hs_next_arrival_ns(0, target, e->d0);
            if (first == HS_T_EXHAUSTED) continue;
            e->u.src.cur_ns = first;
            if ((uint32_t)hdr.fel_n >= S) { hdr.status |= HS_ST_FEL_OVERFLOW; break; }
            hs_wnow n_; n_.time = first; n_.idx = boot++; n_.created = 0; n_.aux = 0ull;
            n_.m0 = HS_EV_SOURCE_TICK | (i << 8); n_.key = -1; n_.hook = 0u; n_.pad = 0u;
            /* sift up */
            int k = hdr.fel_n;                           /* fel_n counts both tiers; only the heap is filled here */
            while (k > 0) { const int p = (k - 1) >> 1; const hs_wnow q = heap[p];
                            if (!(n_.time < q.time || (n_.time == q.time && n_.idx < q.idx))) break; heap[k] = q; k = p; }
            heap[k] = n_;
            hdr.fel_n++;
        }
        hdr.free_top = (uint32_t)hdr.fel_n;              /* free_top doubles as the heap size in this engine */
        hdr.ctr = 0;
    }

    hs_event_record *rec = (FLAGS & HS_WF_REC) && O.records ? O.records + (size_t)r * P.record_cap : nullptr;
    hs_sink_sample *smp = (FLAGS & HS_WF_REC) && O.samples ? O.samples + (size_t)r * P.sample_cap : nullptr;
    double *svc_out = (FLAGS & HS_WF_REC) && O.service ? O.service + (size_t)r * P.service_cap : nullptr;

#define HS_T_LT(T1, I1, T2, I2) ((T1) < (T2) || ((T1) == (T2) && (I1) < (I2)))
    uint64_t ctr = hdr.ctr;
    int now_n = hdr.now_n;
    uint32_t heap_n = hdr.free_top;
    bool paused = false;
    while (true) {
        const int64_t now0 = hdr.now;
        if (!(now0 <= P.end_ns) || (hdr.status & (HS_ST_QUEUE_OVERFLOW | HS_ST_FEL_OVERFLOW | HS_ST_TRACE_EXHAUSTED))) break;
        if (hdr.processed >= P.max_events) { hdr.status |= HS_ST_EVENT_LIMIT; break; }
        /* next event: the now tier's minimum, unless the heap's minimum sorts first */
        int nb = -1; int64_t nt = HS_W_EMPTY; uint64_t ni = ~0ull;
        for (int k = 0; k < now_n; ++k) {
            const int64_t t = N[k].time; const uint64_t ix = N[k].idx;
            if (HS_T_LT(t, ix, nt, ni)) { nt = t; ni = ix; nb = k; }
        }
        hs_wnow ev;
        if (heap_n > 0 && (nb < 0 || HS_T_LT(heap[0].time, heap[0].idx, nt, ni))) {
            ev = heap[0];
            if (windowed && ev.time > P.window_end_ns) { paused = true; break; }
            /* pop: move the last element to the root and sift it down */
            heap_n--;
            if (heap_n > 0) {
                const hs_wnow last = heap[heap_n];
                uint32_t k = 0;
                while (true) {
                    uint32_t ch = 2 * k + 1;
                    if (ch >= heap_n) break;
                    hs_wnow a = heap[ch];
                    if (ch + 1 < heap_n) { const hs_wnow b = heap[ch + 1]; if (HS_T_LT(b.time, b.idx, a.time, a.idx)) { a = b; ch++; } }
                    if (!HS_T_LT(a.time, a.idx, last.time, last.idx)) break;
                    heap[k] = a; k = ch;
                }
                heap[k] = last;
            }
        } else if (nb >= 0) {
            if (windowed && nt > P.window_end_ns) { paused = true; break; }
            ev = N[nb];
            now_n--; N[nb] = N[now_n];
        } else break;                                    /* heap exhausted */
        hdr.fel_n--;
        if (ev.time < now0) continue;                    /* "time travel": skipped (simulation.py:479-489) */

        const int64_t now = ev.time;
        const uint64_t bi = ev.idx;
        const int kind = (int)(ev.m0 & 0xffu);
        const uint32_t ent = ev.m0 >> 8;
        const int64_t e_created = ev.created;
        const uint64_t e_aux = ev.aux;
        const int32_t e_key = ev.key;
        const uint32_t e_hook = ev.hook;
        hdr.now = now;
        if (FLAGS & HS_WF_HASH) hdr.hash = hs_hash_step(hdr.hash, now, hs_record_word1(bi, (uint32_t)kind, ent));
        if ((FLAGS & HS_WF_REC) && rec) {
            hs_event_record rc; rc.time_ns = now; rc.sort_index = (uint32_t)bi; rc.kind = (uint8_t)kind;
            rc.pad = 0; rc.entity = (uint16_t)ent;
            rec[hdr.rec_pos] = rc; hdr.rec_pos = (hdr.rec_pos + 1 == P.record_cap) ? 0u : hdr.rec_pos + 1;
        }
        hdr.processed++;
        hs_went *X = &E[ent];

        /* push: an event at (or before) `now` joins the now tier, a later one the heap */
#define HS_W_PUSH(TIME, IDX, KIND, ENT, CREATED, AUX, KEY, HOOK)                                         \
    do {                                                                                                 \
        hs_wnow n_; n_.time = (TIME); n_.idx = (IDX); n_.created = (CREATED); n_.aux = (AUX);            \
        n_.m0 = (uint32_t)(KIND) | ((uint32_t)(ENT) << 8); n_.key = (KEY); n_.hook = (HOOK); n_.pad = 0; \
        if (n_.time <= now) {                                                                            \
            if (now_n >= HS_W_NCAP) hdr.status |= HS_ST_FEL_OVERFLOW;                                    \
            else { N[now_n++] = n_; hdr.fel_n++; }                                                       \
        } else if (heap_n >= S) hdr.status |= HS_ST_FEL_OVERFLOW;                                        \
        else {                                                                                           \
            uint32_t k_ = heap_n++;                                                                      \
            while (k_ > 0) { const uint32_t p_ = (k_ - 1) >> 1; const hs_wnow q_ = heap[p_];             \
                             if (!HS_T_LT(n_.time, n_.idx, q_.time, q_.idx)) break; heap[k_] = q_; k_ = p_; } \
            heap[k_] = n_; hdr.fel_n++;                                                                  \
        }                                                                                                \
    } while (0)
#include "hs_handlers.inc"
#undef HS_W_PUSH
    }
#undef HS_T_LT

    /* ---- publish ------------------------------------------------------------ */
    hdr.ctr = ctr; hdr.now_n = now_n; hdr.free_top = heap_n;
    hdr.done = paused ? 0 : 1;
    *Hg = hdr;
    if (O.summaries) {
        hs_replica_summary s;
        s.events_processed = hdr.processed; s.final_time_ns = hdr.now;
        s.order_hash = (FLAGS & HS_WF_HASH) ? hdr.hash : 0ull;
        s.next_sort_index = hdr.ctr; s.n_sink_samples = hdr.n_smp; s.n_service_samples = hdr.n_svc;
        s.heap_left = hdr.fel_n; s.status = hdr.status;
        O.summaries[r] = s;
    }
    if (O.stats) {
        for (uint32_t i = 0; i < ne; ++i) {
            const hs_went *e = &E[i];
            hs_entity_stats a; a.c0 = a.c1 = a.c2 = a.c3 = 0; a.f0 = a.f1 = a.f2 = a.f3 = 0.0;
            switch (ENTS[i].kind) {
            case HS_ENT_SOURCE: a.c0 = e->u.src.generated; a.c1 = e->u.src.provider; break;
            case HS_ENT_SERVER: a.c0 = e->u.srv.accepted; a.c1 = e->u.srv.dropped; a.c2 = e->u.srv.completed;
                a.c3 = e->u.srv.rejected; a.f0 = e->u.srv.total_service; break;
            case HS_ENT_SINK: a.c0 = e->u.snk.received; a.f0 = hs_neumaier_result(e->u.snk.sum, e->u.snk.comp);
                a.f1 = e->u.snk.sumsq; a.f2 = e->u.snk.mn; a.f3 = e->u.snk.mx; break;
            case HS_ENT_COUNTER: a.c0 = e->u.snk.received; break;
            case HS_ENT_PROBE: a.c0 = e->u.snk.received; a.f0 = hs_neumaier_result(e->u.snk.sum, e->u.snk.comp);
                a.f2 = e->u.snk.mn; a.f3 = e->u.snk.mx; break;
            case HS_ENT_LB: a.c0 = e->u.lb.received; a.c1 = e->u.lb.forwarded; a.c2 = e->u.lb.in_flight;
                a.c3 = e->u.lb.responses; break;
            }
            O.stats[(size_t)r * ne + i] = a;
        }
    }
}

#endif /* HS_THREAD_ENGINE_CUH */
