/* hs_sketch.h -- SKETCH rows: state layout and the add() step, shared by the
 * kernels, the C-ABI host code and the oracle.
 *
 * Restates  SketchCollector.handle_event   components/sketching/sketch_collector.py:79-98
 *           HyperLogLog.add                sketching/hyperloglog.py:137-165
 *           CountMinSketch.add             sketching/count_min_sketch.py:168-187
 *           BloomFilter.add                sketching/bloom_filter.py:178-199
 *           TopK.add (Space-Saving)        sketching/topk.py:90-128
 *           TDigest.add/_flush/_compress   sketching/tdigest.py:114-190 (IEEE add/mul/div/sqrt only: bit exact)
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
    if (d->i0 == HS_SK_BLOOM) return (((uint64_t)d->i3 + 63u) / 64u * 8u + 15u) / 16u * 16u;   /* uint64 words */
    if (d->i0 == HS_SK_TOPK) return (16u + (uint64_t)d->i2 * 12u + 15u) / 16u * 16u;           /* n, pad, k slots */
    if (d->i0 == HS_SK_TDIGEST) return 32u + (uint64_t)d->i3 * 16u + ((uint64_t)d->i2 * 8u + 15u) / 16u * 16u;
    return ((uint64_t)d->i2 * (uint64_t)d->i3 * 4u + 15u) / 16u * 16u;         /* uint32 counters[depth][width] */
}

/* bytes of the row in the merged image: CMS sums are widened to uint64 */
static inline uint64_t hs_sketch_row_merged_bytes(const hs_entity_desc *d)
{
    if (d->kind != HS_ENT_SKETCH) return 0;
    if (d->i0 == HS_SK_HLL) return (uint64_t)1 << d->i2;
    if (d->i0 == HS_SK_BLOOM) return hs_sketch_row_bytes(d);
    if (d->i0 == HS_SK_TOPK || d->i0 == HS_SK_TDIGEST) return 0;               /* merged on the host */
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
    } else if (algo == HS_SK_CMS) {               /* for row in range(depth): counters[row][col] += 1 */
        uint32_t *c = (uint32_t *)state;
        for (int32_t row = 0; row < p_or_depth; ++row) c[(int64_t)row * width + tab[(int64_t)row * K + key]] += 1u;
    } else if (algo == HS_SK_BLOOM) {             /* for i in range(num_hashes): set bit (h1 + i h2) % size_bits */
        uint64_t *w = (uint64_t *)state;
        for (int32_t i = 0; i < p_or_depth; ++i) {
            const uint32_t bit = (uint32_t)tab[(int64_t)i * K + key];
            w[bit >> 6] |= 1ull << (bit & 63u);
        }
    } else {                                      /* Space-Saving over k counters kept in dict (insertion) order */
        uint32_t *hdr = (uint32_t *)state;
        int32_t *slot = (int32_t *)(state + 16);  /* {item, count, error} x k */
        const uint32_t n = hdr[0], k = (uint32_t)p_or_depth;
        uint32_t j = 0;
        while (j < n && slot[3 * j] != key) ++j;
        if (j < n) { slot[3 * j + 1] += 1; return; }                     /* tracked: count += 1 */
        if (n < k) { slot[3 * n] = key; slot[3 * n + 1] = 1; slot[3 * n + 2] = 0; hdr[0] = n + 1u; return; }
        /* min(self._counters.values(), key=count): the first minimum in dict order is replaced; the new
         * counter inherits its count as error and goes to the end of the dict */
        uint32_t m = 0;
        for (j = 1; j < n; ++j) if ((uint32_t)slot[3 * j + 1] < (uint32_t)slot[3 * m + 1]) m = j;
        const int32_t mc = slot[3 * m + 1];
        for (j = m; j + 1 < n; ++j) { slot[3 * j] = slot[3 * j + 3]; slot[3 * j + 1] = slot[3 * j + 4]; slot[3 * j + 2] = slot[3 * j + 5]; }
        slot[3 * (n - 1)] = key; slot[3 * (n - 1) + 1] = mc + 1; slot[3 * (n - 1) + 2] = mc;
    }
}

/* ---- TDigest ------------------------------------------------------------------------------------ */
typedef struct hs_td_hdr { uint32_t n_centroids, n_buffer; int64_t total; double mn, mx; } hs_td_hdr;   /* 32 B */
typedef struct hs_td_centroid { double mean; int64_t count; } hs_td_centroid;                             /* 16 B */

/* TDigest._max_size(q) (tdigest.py:101-112) */
HS_HD double hs_td_max_size(int64_t total, double compression, double q)
{
    q = q < 0.9999 ? q : 0.9999;                 /* max(0.0001, min(0.9999, q)) */
    q = q > 0.0001 ? q : 0.0001;
    const double den = HS_MUL(HS_MUL(compression, 3.141592653589793), HS_SQRT(HS_MUL(q, HS_SUB(1.0, q))));
    return HS_DIV(HS_LL2D(total * 4), den);
}

/* TDigest._flush + _compress (tdigest.py:139-190): sort the buffer, append its values as unit centroids, sort
 * all centroids by mean (stable: an old centroid stays in front of an equal new value) and merge neighbours
 * while the pair fits under max_size(q).  Returns 0 when the capacity was too small for the merge. */
HS_HD int hs_td_flush(hs_td_hdr *H, hs_td_centroid *C, double *B, double compression, uint32_t cap)
{
    const uint32_t nb = H->n_buffer;
    if (nb == 0) return 1;
    if (H->n_centroids + nb > cap) return 0;
    for (uint32_t gap = nb / 2; gap > 0; gap = (gap == 2) ? 1 : (uint32_t)((uint64_t)gap * 5 / 11)) {   /* Shell sort */
        for (uint32_t i = gap; i < nb; ++i) {
            const double v = B[i]; uint32_t j = i;
            while (j >= gap && B[j - gap] > v) { B[j] = B[j - gap]; j -= gap; }
            B[j] = v;
        }
    }
    /* stable merge from the back: on equal means the buffered value goes behind the old centroid */
    int64_t i = (int64_t)H->n_centroids - 1, j = (int64_t)nb - 1, k = (int64_t)H->n_centroids + nb - 1;
    while (j >= 0) {
        if (i >= 0 && C[i].mean > B[j]) { C[k] = C[i]; --i; }
        else { C[k].mean = B[j]; C[k].count = 1; --j; }
        --k;
    }
    const uint32_t n = H->n_centroids + nb;
    H->n_buffer = 0;
    if (n <= 1) { H->n_centroids = n; return 1; }
    uint32_t w = 0;                              /* compressed[-1] lives at C[w] */
    int64_t running = C[0].count;
    for (uint32_t r = 1; r < n; ++r) {
        const hs_td_centroid c = C[r];
        const double q = HS_DIV(HS_ADD(HS_LL2D(running), HS_DIV(HS_LL2D(c.count), 2.0)), HS_LL2D(H->total));
        const double max_size = hs_td_max_size(H->total, compression, q);
        if (HS_LL2D(C[w].count + c.count) <= max_size) {          /* _Centroid.merge, tdigest.py:41-45 */
            const int64_t tot = C[w].count + c.count;
            C[w].mean = HS_DIV(HS_ADD(HS_MUL(C[w].mean, HS_LL2D(C[w].count)), HS_MUL(c.mean, HS_LL2D(c.count))), HS_LL2D(tot));
            C[w].count = tot;
        } else { ++w; C[w] = c; }
        running += c.count;
    }
    H->n_centroids = w + 1;
    return 1;
}

/* TDigest.add(value) (tdigest.py:114-137); `state` is this replica's state of the row */
HS_HD int hs_tdigest_add(uint8_t *state, double compression, uint32_t buf_size, uint32_t cap, double value)
{
    hs_td_hdr *H = (hs_td_hdr *)state;
    hs_td_centroid *C = (hs_td_centroid *)(state + 32);
    double *B = (double *)(state + 32 + (size_t)cap * 16u);
    if (H->n_buffer >= buf_size) return 0;       /* an earlier flush did not fit: the digest is frozen */
    if (H->total == 0 || value < H->mn) H->mn = value;
    if (H->total == 0 || value > H->mx) H->mx = value;
    H->total += 1;
    B[H->n_buffer++] = value;
    if (H->n_buffer >= buf_size) return hs_td_flush(H, C, B, compression, cap);
    return 1;
}

#endif /* HS_SKETCH_H */
