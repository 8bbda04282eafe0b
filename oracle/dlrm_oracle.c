/*
 * dlrm_oracle.c -- C restatement of DLRM.jl's CPU hot path.  TEST INFRASTRUCTURE ONLY: used by
 * tests/ as a mid-size checker and by bench.py's cpu_baseline / --impl reference legs as the
 * timed CPU arm ("port").  Nothing under dlrm_jl_b200/ links or loads it.
 *
 * It follows the reference's CPU algorithm and threading shape (paths relative to DLRM.jl):
 *   - interaction: one sample per thread iteration (Polyester @batch per=thread,
 *     src/model/interact.jl:460,477); per sample a full F x F Gram by a triple loop with SIMD
 *     reassociation over k (gemmavx!, :318-326), then the strict-triangle copy (:64-75) and the
 *     x / zero-pad copies of process_slice! (:338-362); backward builds the symmetric
 *     zero-diagonal S (:154-173), dT = T * S (:486), dx = dOut[:d] + dT[0] (:434).
 *   - lookup: per (table, sample) gather of P rows and sum (maplookup, call site
 *     src/model/model.jl:161; EmbeddingTables.jl itself is not vendored -- parity for values is
 *     pinned by the goldens, see oracle/oracle.py header).
 *   - update: per table, dictionary dedup of the indices (the SparseIndexer role,
 *     src/train/train.jl:107-115), accumulation of duplicate deltas in first-seen slot order,
 *     then row -= lr * sum (Flux.Descent), tables spread over threads
 *     (EmbeddingTables.update!(...; num_splits = 8, nthreads = 12), src/train/train.jl:283-290).
 *
 * Parity status: the interaction and lookup are pinned by the reference's golden vectors through
 * tests/test_oracle_c.py (which checks this file against oracle/oracle.py, itself pinned in
 * tests/test_oracle_golden.py).  The dedup internals are "parity unpinned" (no reference output
 * observes them).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int width_of(int F, int d, int pad_to_mul) {
    int unpadded = d + F * (F - 1) / 2;
    return (unpadded + pad_to_mul - 1) / pad_to_mul * pad_to_mul;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* out[b][slot0+k][:] = sum_p tables[k][idx[k][b][p]][:]  (p ascending; idx 0-based int64) */
void oracle_lookup(const float* const* tables, int ntab, int D, const int64_t* idx, int B, int P,
                   float* out, int slots, int slot0, int nthreads) {
#pragma omp parallel for collapse(2) schedule(static) num_threads(nthreads)
    for (int k = 0; k < ntab; ++k)
        for (int b = 0; b < B; ++b) {
            const int64_t* ip = idx + ((size_t)k * B + b) * P;
            float* o = out + ((size_t)b * slots + slot0 + k) * D;
            const float* r0 = tables[k] + (size_t)ip[0] * D;
            for (int c = 0; c < D; ++c) o[c] = r0[c];
            for (int p = 1; p < P; ++p) {
                const float* r = tables[k] + (size_t)ip[p] * D;
                for (int c = 0; c < D; ++c) o[c] += r[c];
            }
        }
}

void oracle_interaction_fwd(const float* T, int B, int F, int d, int pad_to_mul, float* out,
                            int nthreads) {
    const int width = width_of(F, d, pad_to_mul);
    const int npair = F * (F - 1) / 2;
#pragma omp parallel num_threads(nthreads)
    {
        float* scratch = (float*)malloc(sizeof(float) * (size_t)F * F); /* per-thread F x F */
#pragma omp for schedule(static)
        for (int b = 0; b < B; ++b) {
            const float* Tb = T + (size_t)b * F * d;
            float* o = out + (size_t)b * width;
            for (int c = 0; c < d; ++c) o[c] = Tb[c];
            for (int c = d + npair; c < width; ++c) o[c] = 0.0f;
            for (int n = 0; n < F; ++n)
                for (int m = 0; m < F; ++m) {
                    float acc = 0.0f;
#pragma omp simd reduction(+ : acc)
                    for (int k = 0; k < d; ++k) acc += Tb[(size_t)m * d + k] * Tb[(size_t)n * d + k];
                    scratch[(size_t)n * F + m] = acc;
                }
            int y = d;
            for (int j = 1; j < F; ++j) {
                for (int i = 0; i < j; ++i) o[y + i] = scratch[(size_t)j * F + i];
                y += j;
            }
        }
        free(scratch);
    }
}

void oracle_interaction_bwd(const float* dOut, const float* T, int B, int F, int d, int pad_to_mul,
                            float* dT, float* dx, int nthreads) {
    const int width = width_of(F, d, pad_to_mul);
#pragma omp parallel num_threads(nthreads)
    {
        float* S = (float*)malloc(sizeof(float) * (size_t)F * F);
#pragma omp for schedule(static)
        for (int b = 0; b < B; ++b) {
            const float* g = dOut + (size_t)b * width;
            const float* Tb = T + (size_t)b * F * d;
            float* dTb = dT + (size_t)b * F * d;
            for (int j = 0; j < F; ++j)
                for (int i = 0; i < F; ++i) {
                    float v = 0.0f;
                    if (i != j) {
                        int hi = i > j ? i : j, lo = i > j ? j : i;
                        v = g[d + hi * (hi - 1) / 2 + lo];
                    }
                    S[(size_t)j * F + i] = v;
                }
            for (int f = 0; f < F; ++f) {
                float* o = dTb + (size_t)f * d;
                for (int k = 0; k < d; ++k) o[k] = 0.0f;
                for (int j = 0; j < F; ++j) {
                    const float s = S[(size_t)j * F + f];
                    const float* tj = Tb + (size_t)j * d;
#pragma omp simd
                    for (int k = 0; k < d; ++k) o[k] += s * tj[k];
                }
            }
            for (int k = 0; k < d; ++k) dx[(size_t)b * d + k] = g[k] + dTb[k];
        }
        free(S);
    }
}

/* open-addressing dictionary row id -> compact slot (the SparseIndexer role) */
typedef struct {
    int64_t* keys;
    int32_t* vals;
    size_t cap;
} dict_t;

static void dict_init(dict_t* d, size_t n) {
    size_t cap = 16;
    while (cap < 2 * n) cap <<= 1;
    d->cap = cap;
    d->keys = (int64_t*)malloc(sizeof(int64_t) * cap);
    d->vals = (int32_t*)malloc(sizeof(int32_t) * cap);
    for (size_t i = 0; i < cap; ++i) d->keys[i] = -1;
}

static int32_t dict_get_or_insert(dict_t* d, int64_t key, int32_t next) {
    size_t h = ((uint64_t)key * 0x9E3779B97F4A7C15ull) >> 20;
    for (;;) {
        h &= d->cap - 1;
        if (d->keys[h] == key) return d->vals[h];
        if (d->keys[h] < 0) {
            d->keys[h] = key;
            d->vals[h] = next;
            return next;
        }
        ++h;
    }
}

/* tables[k][r] -= lr * sum_{(b,p): idx = r} dT[b][slot0+k][:]   (in place) */
void oracle_sparse_sgd(float* const* tables, int ntab, int D, const int64_t* idx, int B, int P,
                       const float* dT, int slots, int slot0, float lr, int nthreads) {
    const size_t L = (size_t)B * P;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int k = 0; k < ntab; ++k) {
        const int64_t* ik = idx + (size_t)k * L;
        dict_t dict;
        dict_init(&dict, L);
        int64_t* uniq = (int64_t*)malloc(sizeof(int64_t) * L);
        float* acc = (float*)calloc(L * (size_t)D, sizeof(float));
        int32_t n = 0;
        for (size_t i = 0; i < L; ++i) {
            int32_t s = dict_get_or_insert(&dict, ik[i], n);
            if (s == n) uniq[n++] = ik[i];
            const float* dl = dT + ((size_t)(i / P) * slots + slot0 + k) * D;
            float* a = acc + (size_t)s * D;
            for (int c = 0; c < D; ++c) a[c] += dl[c];
        }
        for (int32_t s = 0; s < n; ++s) {
            float* row = tables[k] + (size_t)uniq[s] * D;
            const float* a = acc + (size_t)s * D;
            for (int c = 0; c < D; ++c) {
                float delta = lr * a[c]; /* Flux.Descent: delta .*= eta; x .-= delta */
                row[c] -= delta;
            }
        }
        free(acc);
        free(uniq);
        free(dict.keys);
        free(dict.vals);
    }
}

/* CPU-arm setup helper (not part of the timed path): fill a [rows][D] table with U(-1/sqrt(rows),
 * 1/sqrt(rows)) values (ScaledUniform, src/model/model.jl:61-65; the reference threads its init the
 * same way, :40-44) from a counter hash, in parallel, so that tables of tens of GB are ready in
 * seconds. */
static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void oracle_init_uniform(float* table, int64_t rows, int D, uint64_t seed, int nthreads) {
    const float scale = 1.0f / sqrtf((float)rows);
    const int64_t n = rows * (int64_t)D;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (int64_t i = 0; i < n; ++i) {
        uint64_t h = mix64(seed ^ mix64((uint64_t)i));
        float u = (float)(h >> 40) * (1.0f / 16777216.0f);
        table[i] = (2.0f * u - 1.0f) * scale;
    }
}
