/* hs_sketch.h -- SKETCH rows: state layout and the add() step, shared by the
 * kernels, the C-ABI host code and the oracle.
 *
 * Restates  SketchCollector.handle_event   components/sketching/sketch_collector.py:79-98
 *           HyperLogLog.add                sketching/hyperloglog.py:137-165
 *           CountMinSketch.add             sketching/count_min_sketch.py:168-187
 * with the SHA-256 evaluations (hyperloglog.py:128-135, count_min_sketch.py:136-155) taken from the
 * per-key tables the host built: the items are the routing keys 0..K-1.
 */
#ifndef HS_SKETCH_H
#define HS_SKETCH_H

#include "hs_sampler.h"
#include "../../include/hs_b200.h"

/* bytes of one replica's state of a SKETCH row (multiple of 16) */
static inline uint64_t hs_sketch_row_bytes(const hs_entity_desc *d)
{
    if (d->kind != HS_ENT_SKETCH) return 0;
    if (d->i0 == HS_SK_HLL) return (uint64_t)1 << d->i2;                       /* uint8 registers[2^p], p >= 4 */
    return ((uint64_t)d->i2 * (uint64_t)d->i3 * 4u + 15u) / 16u * 16u;         /* uint32 counters[depth][width] */
}

/* bytes of the row in the merged image: CMS sums are widened to uint64 */
static inline uint64_t hs_sketch_row_merged_bytes(const hs_entity_desc *d)
{
    if (d->kind != HS_ENT_SKETCH) return 0;
    if (d->i0 == HS_SK_HLL) return (uint64_t)1 << d->i2;
    return ((uint64_t)d->i2 * (uint64_t)d->i3 * 8u + 15u) / 16u * 16u;
}

static inline void hs_sketch_layout_impl(const hs_model_desc *m, uint64_t *per_replica, uint64_t *merged,
                                         uint64_t *total, uint64_t *merged_total)
{
    uint64_t a = 0, b = 0;
    for (uint32_t i = 0; i < m->n_entities; ++i) {
        const hs_entity_desc *d = &m->entities[i];
        if (per_replica) per_replica[i] = d->kind == HS_ENT_SKETCH ? a : 0;
        if (merged) merged[i] = d->kind == HS_ENT_SKETCH ? b : 0;
        a += hs_sketch_row_bytes(d);
        b += hs_sketch_row_merged_bytes(d);
    }
    if (total) *total = a;
    if (merged_total) *merged_total = b;
}

/* sketch.add(key): `state` is this replica's state of the row, `tab` the row's table (stride K) */
HS_HD void hs_sketch_add(uint8_t *state, const int32_t *tab, int32_t algo, int32_t p_or_depth, int32_t width,
                         int64_t K, int32_t key)
{
    if (algo == HS_SK_HLL) {                      /* registers[idx] = max(registers[idx], run_length) */
        const int32_t idx = tab[key], run = tab[K + key];
        if ((int32_t)state[idx] < run) state[idx] = (uint8_t)run;
    } else {                                      /* for row in range(depth): counters[row][col] += 1 */
        uint32_t *c = (uint32_t *)state;
        for (int32_t row = 0; row < p_or_depth; ++row) c[(int64_t)row * width + tab[(int64_t)row * K + key]] += 1u;
    }
}

#endif /* HS_SKETCH_H */
