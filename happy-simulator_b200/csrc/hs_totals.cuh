/* hs_totals.cuh -- deterministic two-stage reduction of the per-replica results
 * into the fixed-layout hs_totals vector (the payload of the single end-of-run
 * allreduce).  Stage 1: each block folds a contiguous
 * slice of replicas in index order per thread, then a fixed shuffle/shared tree;
 * stage 2: one warp folds the block partials in block order.  No atomics, so the
 * result does not depend on scheduling. */
#ifndef HS_TOTALS_CUH
#define HS_TOTALS_CUH

#include "../../include/hs_b200.h"

__device__ __forceinline__ void hs_totals_zero(hs_totals &t)
{
    for (int k = 0; k < HS_TOTALS_I64; ++k) t.i[k] = 0;
    for (int k = 0; k < HS_TOTALS_F64_SUM; ++k) t.fsum[k] = 0.0;
    t.fmin = __longlong_as_double(0x7ff0000000000000LL);
    t.fmax = __longlong_as_double(0xfff0000000000000LL);
}

__device__ __forceinline__ void hs_totals_merge(hs_totals &a, const hs_totals &b)
{
    for (int k = 0; k < HS_TOTALS_I64; ++k) a.i[k] += b.i[k];
    for (int k = 0; k < HS_TOTALS_F64_SUM; ++k) a.fsum[k] += b.fsum[k];
    a.fmin = b.fmin < a.fmin ? b.fmin : a.fmin;
    a.fmax = b.fmax > a.fmax ? b.fmax : a.fmax;
}

__device__ __forceinline__ hs_totals hs_totals_shfl_down(const hs_totals &t, int delta)
{
    hs_totals o;
    for (int k = 0; k < HS_TOTALS_I64; ++k) o.i[k] = __shfl_down_sync(0xffffffffu, t.i[k], delta);
    for (int k = 0; k < HS_TOTALS_F64_SUM; ++k) o.fsum[k] = __shfl_down_sync(0xffffffffu, t.fsum[k], delta);
    o.fmin = __shfl_down_sync(0xffffffffu, t.fmin, delta);
    o.fmax = __shfl_down_sync(0xffffffffu, t.fmax, delta);
    return o;
}

__global__ void __launch_bounds__(256)
hs_totals_partial_kernel(const hs_replica_summary *__restrict__ summ, const hs_entity_stats *__restrict__ stats,
                         const hs_entity_desc *__restrict__ ents, uint32_t n_replicas, uint32_t n_entities,
                         hs_totals *__restrict__ partials)
{
    __shared__ hs_totals sh[8];
    hs_totals t; hs_totals_zero(t);
    const uint32_t per_block = (n_replicas + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = blockIdx.x * per_block;
    const uint32_t hi = min(n_replicas, lo + per_block);
    for (uint32_t r = lo + threadIdx.x; r < hi; r += blockDim.x) {
        const hs_replica_summary s = summ[r];
        t.i[0] += s.events_processed;
        t.i[5] += 1;
        t.i[6] += (s.status != 0);
        t.i[7] += s.final_time_ns / 1000;
        for (uint32_t e = 0; e < n_entities; ++e) {
            const int kind = ents[e].kind;
            const hs_entity_stats st = stats[(size_t)r * n_entities + e];
            if (kind == HS_ENT_SINK) {
                t.i[1] += st.c0; t.fsum[0] += st.f0; t.fsum[1] += st.f1;
                t.fmin = st.f2 < t.fmin ? st.f2 : t.fmin;
                t.fmax = st.f3 > t.fmax ? st.f3 : t.fmax;
            } else if (kind == HS_ENT_SERVER) {
                t.i[2] += st.c2; t.i[4] += st.c1; t.fsum[2] += st.f0;
            } else if (kind == HS_ENT_SOURCE) {
                t.i[3] += st.c0;
            }
        }
    }
    for (int d = 16; d >= 1; d >>= 1) { hs_totals o = hs_totals_shfl_down(t, d); hs_totals_merge(t, o); }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sh[warp] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        hs_totals acc = sh[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) hs_totals_merge(acc, sh[w]);
        partials[blockIdx.x] = acc;
    }
}

__global__ void hs_totals_final_kernel(const hs_totals *__restrict__ partials, int n, hs_totals *__restrict__ out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        hs_totals acc; hs_totals_zero(acc);
        for (int b = 0; b < n; ++b) hs_totals_merge(acc, partials[b]);
        *out = acc;
    }
}


/* Per-cell reduction (BASELINE configs[4]): block c folds the replicas whose global index maps to cell
 * c (index / replicas_per_cell, modulo n_cells) in index order per thread, then a fixed tree. */
__device__ __forceinline__ void hs_totals_accumulate(hs_totals &t, const hs_replica_summary &s,
                                                     const hs_entity_stats *st, const hs_entity_desc *ents, uint32_t n_entities)
{
    t.i[0] += s.events_processed; t.i[5] += 1; t.i[6] += (s.status != 0); t.i[7] += s.final_time_ns / 1000;
    for (uint32_t e = 0; e < n_entities; ++e) {
        const int kind = ents[e].kind;
        const hs_entity_stats x = st[e];
        if (kind == HS_ENT_SINK) {
            t.i[1] += x.c0; t.fsum[0] += x.f0; t.fsum[1] += x.f1;
            t.fmin = x.f2 < t.fmin ? x.f2 : t.fmin; t.fmax = x.f3 > t.fmax ? x.f3 : t.fmax;
        } else if (kind == HS_ENT_SERVER) { t.i[2] += x.c2; t.i[4] += x.c1; t.fsum[2] += x.f0; }
        else if (kind == HS_ENT_SOURCE) t.i[3] += x.c0;
    }
}

__global__ void __launch_bounds__(128)
hs_cell_totals_kernel(const hs_replica_summary *__restrict__ summ, const hs_entity_stats *__restrict__ stats,
                      const hs_entity_desc *__restrict__ ents, const uint32_t *__restrict__ hist,
                      uint32_t n_replicas, uint32_t n_entities, uint32_t index_base, uint32_t replicas_per_cell,
                      uint32_t n_cells, hs_cell_totals *__restrict__ out)
{
    __shared__ hs_totals sh[4];
    __shared__ unsigned long long shh[HS_HISTOGRAM_BINS];
    const uint32_t c = blockIdx.x;
    if (threadIdx.x < HS_HISTOGRAM_BINS) shh[threadIdx.x] = 0ull;
    __syncthreads();
    hs_totals t; hs_totals_zero(t);
    for (uint32_t r = threadIdx.x; r < n_replicas; r += blockDim.x) {
        const uint32_t cell = ((index_base + r) / replicas_per_cell) % n_cells;
        if (cell != c) continue;
        hs_totals_accumulate(t, summ[r], stats + (size_t)r * n_entities, ents, n_entities);
        if (hist) for (int b = 0; b < HS_HISTOGRAM_BINS; ++b) {
            const uint32_t v = hist[(size_t)r * HS_HISTOGRAM_BINS + b];
            if (v) atomicAdd(&shh[b], (unsigned long long)v);       /* integer sums: order independent */
        }
    }
    for (int d = 16; d >= 1; d >>= 1) { hs_totals o = hs_totals_shfl_down(t, d); hs_totals_merge(t, o); }
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        hs_totals acc = sh[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) hs_totals_merge(acc, sh[w]);
        out[c].totals = acc;
    }
    if (threadIdx.x < HS_HISTOGRAM_BINS) out[c].histogram[threadIdx.x] = shh[threadIdx.x];
}

/* ---- sketch merge over the replicas of a run (the reference's merge() contracts) ------------
 * One thread per 4 registers / per counter walks the replicas; neighbouring threads read neighbouring
 * words of the same replica, so every load instruction is a contiguous segment. */

/* HyperLogLog.merge: element-wise max of the registers (sketching/hyperloglog.py:203-226) */
__global__ void hs_sketch_merge_hll_kernel(const uint8_t *__restrict__ per_replica, uint64_t stride, uint32_t n_replicas,
                                           uint32_t n_words, uint32_t *__restrict__ merged)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t acc = 0u;
    for (uint32_t r = 0; r < n_replicas; ++r)
        acc = __vmaxu4(acc, *(const uint32_t *)(per_replica + (size_t)r * stride + (size_t)w * 4u));
    merged[w] = acc;
}

/* BloomFilter.merge: bitwise OR of the bit arrays (sketching/bloom_filter.py:262-291) */
__global__ void hs_sketch_merge_or_kernel(const uint8_t *__restrict__ per_replica, uint64_t stride, uint32_t n_replicas,
                                          uint32_t n_words, uint32_t *__restrict__ merged)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t acc = 0u;
    for (uint32_t r = 0; r < n_replicas; ++r)
        acc |= *(const uint32_t *)(per_replica + (size_t)r * stride + (size_t)w * 4u);
    merged[w] = acc;
}

/* CountMinSketch.merge: element-wise sum of the counters (sketching/count_min_sketch.py:276-301) */
__global__ void hs_sketch_merge_cms_kernel(const uint8_t *__restrict__ per_replica, uint64_t stride, uint32_t n_replicas,
                                           uint32_t n_cells, unsigned long long *__restrict__ merged)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    unsigned long long acc = 0ull;
    for (uint32_t r = 0; r < n_replicas; ++r)
        acc += *(const uint32_t *)(per_replica + (size_t)r * stride + (size_t)c * 4u);
    merged[c] = acc;
}

#endif
