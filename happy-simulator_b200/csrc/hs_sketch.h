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
/* the two entry points of a SKETCH row (measured: as real calls on the device they cost the thread engine's hot
 * loop registers around the call site -- 264 B of spills -- for no gain, so they stay inline) */
#if defined(__CUDACC__)
#define HS_SK_FN __host__ __device__ __forceinline__
#else
#define HS_SK_FN static inline
#endif

#include "hs_sampler.h"
#include "../../include/hs_b200.h"

/* bytes of one replica's state of a SKETCH row (multiple of 16) */
static inline uint64_t hs_sketch_row_bytes(const hs_entity_desc *d)
{
    if (d->kind == HS_ENT_CACHE_SERVER) return (((uint64_t)d->i0 + 1u) * 8u + 15u) / 16u * 16u;   /* double insert_s[K + 1] */
    if (d->kind != HS_ENT_SKETCH) return 0;
    if (d->i0 == HS_SK_HLL) return (uint64_t)1 << d->i2;                       /* uint8 registers[2^p], p >= 4 */
    if (d->i0 == HS_SK_BLOOM) return (((uint64_t)d->i3 + 63u) / 64u * 8u + 15u) / 16u * 16u;   /* uint64 words */
    if (d->i0 == HS_SK_TOPK) return (16u + (uint64_t)d->i2 * 12u + 15u) / 16u * 16u;           /* n, pad, k slots */
    if (d->i0 == HS_SK_TDIGEST) return 32u + (uint64_t)d->i3 * 16u + ((uint64_t)d->i2 * 8u + 15u) / 16u * 16u;
    if (d->i0 == HS_SK_RESERVOIR) return 16u + 624u * 4u + ((uint64_t)d->i2 * 4u + 15u) / 16u * 16u;   /* hdr, mt[624], items[size] */
    return ((uint64_t)d->i2 * (uint64_t)d->i3 * 4u + 15u) / 16u * 16u;         /* uint32 counters[depth][width] */
}

/* bytes of the row in the merged image: CMS sums are widened to uint64 */
static inline uint64_t hs_sketch_row_merged_bytes(const hs_entity_desc *d)
{
    if (d->kind != HS_ENT_SKETCH) return 0;
    if (d->i0 == HS_SK_HLL) return (uint64_t)1 << d->i2;
    if (d->i0 == HS_SK_BLOOM) return hs_sketch_row_bytes(d);
    if (d->i0 == HS_SK_TOPK || d->i0 == HS_SK_TDIGEST || d->i0 == HS_SK_RESERVOIR) return 0;   /* merged on the host */
    return ((uint64_t)d->i2 * (uint64_t)d->i3 * 8u + 15u) / 16u * 16u;
}

static inline void hs_sketch_layout_impl(const hs_model_desc *m, uint64_t *per_replica, uint64_t *merged,
                                         uint64_t *total, uint64_t *merged_total)
{
    uint64_t a = 0, b = 0;
    for (uint32_t i = 0; i < m->n_entities; ++i) {
        const hs_entity_desc *d = &m->entities[i];
        if (per_replica) per_replica[i] = (d->kind == HS_ENT_SKETCH || d->kind == HS_ENT_CACHE_SERVER) ? a : 0;
        if (merged) merged[i] = d->kind == HS_ENT_SKETCH ? b : 0;
        a += hs_sketch_row_bytes(d);
        b += hs_sketch_row_merged_bytes(d);
    }
    if (total) *total = a;
    if (merged_total) *merged_total = b;
}

/* ---- SHA-256 of a short message (FIPS 180-4), one 64-byte block: everything the sketches hash is a packed
 * seed plus repr(key), at most 16 + 11 bytes.  Used when a SKETCH row carries no per-key table (i1 = -1):
 * the hashes of hyperloglog.py:128-135, count_min_sketch.py:136-155 and bloom_filter.py:147-160 are then
 * evaluated per event, so the key population needs no table in HBM. ----------------------------------- */
HS_HD uint32_t hs_rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

/* digest words 0..3 (the first 16 bytes, big endian) of sha256(msg[0..len)), len <= 55 */
HS_HD void hs_sha256_short(const uint8_t *msg, uint32_t len, uint32_t out[4])
{
    static const uint32_t Kc[64] = {
        0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
        0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
        0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
        0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
        0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
        0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
        0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
        0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
    uint32_t w[64];
    for (int i = 0; i < 16; ++i) w[i] = 0u;
    for (uint32_t i = 0; i < len; ++i) w[i >> 2] |= (uint32_t)msg[i] << (24 - 8 * (i & 3u));
    w[len >> 2] |= 0x80u << (24 - 8 * (len & 3u));
    w[15] = len * 8u;
    for (int i = 16; i < 64; ++i) {
        const uint32_t s0 = hs_rotr32(w[i - 15], 7) ^ hs_rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        const uint32_t s1 = hs_rotr32(w[i - 2], 17) ^ hs_rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = 0x6a09e667u, b = 0xbb67ae85u, c = 0x3c6ef372u, d = 0xa54ff53au,
             e = 0x510e527fu, f = 0x9b05688cu, g = 0x1f83d9abu, h = 0x5be0cd19u;
    for (int i = 0; i < 64; ++i) {
        const uint32_t S1 = hs_rotr32(e, 6) ^ hs_rotr32(e, 11) ^ hs_rotr32(e, 25);
        const uint32_t t1 = h + S1 + ((e & f) ^ (~e & g)) + Kc[i] + w[i];
        const uint32_t S0 = hs_rotr32(a, 2) ^ hs_rotr32(a, 13) ^ hs_rotr32(a, 22);
        const uint32_t t2 = S0 + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    out[0] = 0x6a09e667u + a; out[1] = 0xbb67ae85u + b; out[2] = 0x3c6ef372u + c; out[3] = 0xa54ff53au + d;
}

HS_HD uint32_t hs_put_be64(uint8_t *p, uint64_t v)          /* struct.pack(">Q", v) */
{ for (int i = 0; i < 8; ++i) p[i] = (uint8_t)(v >> (56 - 8 * i)); return 8u; }

HS_HD uint32_t hs_put_repr_int(uint8_t *p, int32_t key)     /* repr(key).encode() for key >= 0 */
{
    uint8_t tmp[12]; uint32_t n = 0; uint32_t v = (uint32_t)key;
    do { tmp[n++] = (uint8_t)('0' + v % 10u); v /= 10u; } while (v);
    for (uint32_t i = 0; i < n; ++i) p[i] = tmp[n - 1 - i];
    return n;
}

/* HyperLogLog._hash + the register split of add() (hyperloglog.py:128-165) */
HS_HD void hs_hll_hash(uint64_t seed, int32_t p, int32_t key, int32_t *idx, int32_t *run)
{
    uint8_t m[24]; uint32_t d[4];
    uint32_t n = hs_put_be64(m, seed);
    n += hs_put_repr_int(m + n, key);
    hs_sha256_short(m, n, d);
    const uint64_t h = ((uint64_t)d[0] << 32) | d[1];
    const int low = 64 - p;
    const uint64_t rest = h & ((1ull << low) - 1ull);
    *idx = (int32_t)(h >> low);
    int lz = low;                                           /* _count_leading_zeros(rest, low): low for rest == 0 */
    if (rest) { lz = 0; for (int i = low - 1; i >= 0 && !((rest >> i) & 1ull); --i) ++lz; }
    *run = lz + 1;
}

/* CountMinSketch._hash(item, row) for a non-negative int item, hash(item) == item (count_min_sketch.py:145-155);
 * row_seed = sha256(pack(">QQ", seed, row))[:8], evaluated by hs_cms_row_seed */
HS_HD uint64_t hs_cms_row_seed(uint64_t seed, int32_t row)
{
    uint8_t m[16]; uint32_t d[4];
    hs_put_be64(m, seed); hs_put_be64(m + 8, (uint64_t)row);
    hs_sha256_short(m, 16u, d);
    return ((uint64_t)d[0] << 32) | d[1];
}
HS_HD int32_t hs_cms_col(uint64_t row_seed, int32_t width, int32_t key)
{
    uint8_t m[8]; uint32_t d[4];
    hs_put_be64(m, (uint64_t)(int64_t)key ^ row_seed);
    hs_sha256_short(m, 8u, d);
    return (int32_t)((((uint64_t)d[0] << 32) | d[1]) % (uint64_t)width);
}

/* BloomFilter._hash(item, i) (bloom_filter.py:147-160): (h1 + i h2) mod size_bits over Python's unbounded ints;
 * size_bits < 2^31, so the residues multiply without overflow */
HS_HD int32_t hs_bloom_bit(uint64_t seed, int32_t i, int32_t size_bits, int32_t key)
{
    uint8_t m[32]; uint32_t d[4];
    uint32_t n = hs_put_be64(m, seed);
    n += hs_put_be64(m + n, (uint64_t)i);
    n += hs_put_repr_int(m + n, key);
    hs_sha256_short(m, n, d);
    const uint64_t h1 = ((uint64_t)d[0] << 32) | d[1], h2 = ((uint64_t)d[2] << 32) | d[3];
    const uint64_t mm = (uint64_t)size_bits;
    return (int32_t)((h1 % mm + ((uint64_t)i % mm) * (h2 % mm)) % mm);
}

/* ---- ReservoirSampler (sketching/reservoir.py:30): Algorithm R driven by a private random.Random(seed), i.e.
 * CPython's MT19937.  State of one replica: {uint32 n_items, uint32 mti, uint64 total_count}, uint32 mt[624],
 * int32 items[size].  The row's table holds the generator state the collector was built with (mt[624], mti). */
typedef struct hs_rs_hdr { uint32_t n_items, mti; uint64_t total; } hs_rs_hdr;   /* 16 B */

/* genrand_uint32 of CPython's _randommodule.c (the MT19937 reference algorithm): all 624 words are regenerated
 * when the index runs out, then one tempered word is handed out */
HS_HD uint32_t hs_mt_next(uint32_t *mt, uint32_t *mti)
{
    if (*mti >= 624u) {
        for (uint32_t k = 0; k < 624u; ++k) {
            const uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1u < 624u ? k + 1u : 0u] & 0x7fffffffu);
            mt[k] = mt[k + 397u < 624u ? k + 397u : k - 227u] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        *mti = 0;
    }
    uint32_t y = mt[(*mti)++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}

/* ReservoirSampler._add_one (reservoir.py:100-111): the first `size` items fill the list; after that
 * j = randint(0, total - 1) = Random._randbelow_with_getrandbits(total): k = total.bit_length(),
 * r = getrandbits(k) = genrand_uint32() >> (32 - k), redrawn while r >= total; the item replaces slot j < size. */
HS_SK_FN void hs_reservoir_add(uint8_t *state, const int32_t *tab, uint32_t size, int32_t key)
{
    hs_rs_hdr *H = (hs_rs_hdr *)state;
    uint32_t *mt = (uint32_t *)(state + 16);
    int32_t *items = (int32_t *)(state + 16 + 624 * 4);
    if (H->total == 0) {                          /* first add of this replica: the collector's own generator state */
        for (uint32_t k = 0; k < 624u; ++k) mt[k] = (uint32_t)tab[k];
        H->mti = (uint32_t)tab[624];
    }
    H->total += 1;
    if (H->n_items < size) { items[H->n_items++] = key; return; }
    const uint32_t total = (uint32_t)H->total;    /* < 2^32: a replica never processes that many events */
    int bits = 0;
    for (uint32_t t = total; t; t >>= 1) ++bits;          /* total.bit_length() */
    uint32_t mti = H->mti, r;
    do r = hs_mt_next(mt, &mti) >> (32 - bits); while (r >= total);
    H->mti = mti;
    if (r < size) items[r] = key;
}

/* sketch.add(key): `state` is this replica's state of the row, `tab` the row's table (stride K) */
HS_SK_FN void hs_sketch_add(uint8_t *state, const int32_t *tab, int32_t algo, int32_t p_or_depth, int32_t width,
                         int64_t K, int32_t key)
{
    /* K > 0: `tab` is the per-key table.  K == 0: the hashes are evaluated here; `tab` then holds the sketch seed
     * (HLL, BLOOM: lo, hi) or the row seeds (CMS: depth x (lo, hi)) as int32 pairs. */
#define HS_SK_SEED64(I) ((uint64_t)(uint32_t)tab[2 * (I)] | ((uint64_t)(uint32_t)tab[2 * (I) + 1] << 32))
    if (algo == HS_SK_HLL) {                      /* registers[idx] = max(registers[idx], run_length) */
        int32_t idx, run;
        if (K) { idx = tab[key]; run = tab[K + key]; }
        else hs_hll_hash(HS_SK_SEED64(0), p_or_depth, key, &idx, &run);
        if ((int32_t)state[idx] < run) state[idx] = (uint8_t)run;
    } else if (algo == HS_SK_CMS) {               /* for row in range(depth): counters[row][col] += 1 */
        uint32_t *c = (uint32_t *)state;
        for (int32_t row = 0; row < p_or_depth; ++row) {
            const int32_t col = K ? tab[(int64_t)row * K + key] : hs_cms_col(HS_SK_SEED64(row), width, key);
            c[(int64_t)row * width + col] += 1u;
        }
    } else if (algo == HS_SK_BLOOM) {             /* for i in range(num_hashes): set bit (h1 + i h2) % size_bits */
        uint64_t *w = (uint64_t *)state;
        for (int32_t i = 0; i < p_or_depth; ++i) {
            const uint32_t bit = (uint32_t)(K ? tab[(int64_t)i * K + key] : hs_bloom_bit(HS_SK_SEED64(0), i, width, key));
            w[bit >> 6] |= 1ull << (bit & 63u);
        }
#undef HS_SK_SEED64
    } else if (algo == HS_SK_RESERVOIR) {
        hs_reservoir_add(state, tab, (uint32_t)p_or_depth, key);
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
HS_SK_FN int hs_tdigest_add(uint8_t *state, double compression, uint32_t buf_size, uint32_t cap, double value)
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
