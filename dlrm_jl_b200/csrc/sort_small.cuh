// Shared-memory sort of one table's index list by ONE CTA (device function, used by the stand-alone
// sort kernel in sort.cu and by the sort CTAs of the fused lookup + sort launches in lookup.cu and
// p2p.cu).  Output: the keys ascending and the stable permutation (ties by flat position) -- the
// order is a total order on (key, position), so every algorithm below produces the same bits.
//
// THREADS x ITEMS keys (up to 1024 x 16 = 16384).  Two algorithms, chosen per table (block-uniform):
//
//   bucket + rank   (tables with more than 512 rows).  The keys are spread over NBK = THREADS * ITEMS
//       buckets by their top bits (bucket = key >> shift; shift = 0 when the table has fewer rows than
//       buckets, so a bucket is one row).  A shared-memory atomic histogram, one exclusive scan and
//       an atomic scatter place every (key, position) pair inside its bucket's range in arbitrary
//       order; each pair then finds its exact rank by counting the pairs of its bucket that compare
//       lower.  With uniformly drawn rows a bucket holds about one key, so the rank loop is a couple
//       of iterations: three barriers and no multi-pass radix.  Skewed batches cost O(bucket size) per
//       key; if some bucket exceeds kBucketRankMax keys (hot rows, ids clustered in a narrow range)
//       the CTA falls back to the radix passes below, whose cost does not depend on the distribution.
//
//   LSD radix       (tables with at most 512 rows, and the fallback).  Element e lives in warp
//       w = e / (32 * ITEMS), item i, lane l (e = w*32*ITEMS + i*32 + l), so (warp, item, lane) order is
//       input order and the per-digit ranks make every pass stable.  Per pass: match_any groups the
//       lanes of a warp by digit, a per-warp digit counter in shared memory turns that into a rank
//       inside the warp's chunk, an exclusive scan over the digit totals gives the bucket starts, and
//       the keys are scattered through shared memory.  Digits are up to 9 bits wide, spread evenly
//       over ceil(bits / 9) passes: a 24-row table needs one 5-bit pass, a 40M-row table three.
#pragma once
#include "common.cuh"

namespace dlrmb {

constexpr int kSmallSortBins = 512;     // LSD digit bins
constexpr int kBucketRankMax = 192;     // largest bucket the rank-by-counting phase accepts

template <int ITEMS, int THREADS>
struct SmallSortGeom {
    static constexpr int N = THREADS * ITEMS;           // keys per CTA = buckets of the bucket + rank path
    static constexpr int NW = THREADS / 32;
    static constexpr int BPT = (kSmallSortBins + THREADS - 1) / THREADS;   // LSD: bins owned per thread in the scan
    static constexpr int AUX = (N > NW * kSmallSortBins) ? N : NW * kSmallSortBins;   // bucket ends / per-warp digit counters
    static constexpr size_t smem_bytes() {
        return sizeof(uint32_t) * ((size_t)N + (size_t)AUX + 40) + sizeof(uint16_t) * (size_t)N;
    }
    static_assert(N <= 65536, "positions are kept in 16 bits");
    static_assert((N & (N - 1)) == 0 && ITEMS % 4 == 0, "power-of-two key count, 16-byte scan loads");
};

// exclusive prefix of `v` over the THREADS threads of the CTA (thread order); wsum is 33 words of shared memory
template <int THREADS>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* wsum) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    uint32_t run = inc - v;
    for (int ww = 0; ww < THREADS / 32; ++ww)
        if (ww < w) run += wsum[ww];
    return run;
}

template <typename IdxT, int ITEMS, int THREADS>
__device__ __forceinline__ void sort_small_body(const IdxT* __restrict__ ik, int idx_base, int L, int64_t rows,
                                                uint32_t* __restrict__ ko, uint32_t* __restrict__ po,
                                                uint32_t* sort_smem) {
    using G = SmallSortGeom<ITEMS, THREADS>;
    constexpr int N = G::N, NW = G::NW, BPT = G::BPT, NB = kSmallSortBins;
    uint32_t* ksm = sort_smem;                                   // [N]   keys
    uint32_t* aux = ksm + N;                                     // [AUX] bucket ends, or [NW][NB] digit counters
    uint32_t* wsum = aux + G::AUX;                               // [40]
    uint16_t* vsm = reinterpret_cast<uint16_t*>(wsum + 40);      // [N]   positions
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    int bits = 0;
    while (bits < 32 && (1ll << bits) < rows) ++bits;

    if (bits == 0) {   // single-row table: already sorted
        for (int i = tid; i < L; i += THREADS) {
            ko[i] = 0u;
            po[i] = (uint32_t)i;
        }
        return;
    }

    const int base = w * (32 * ITEMS);
    uint32_t key[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        const int e = base + i * 32 + lane;
        key[i] = e < L ? (uint32_t)((int64_t)ik[e] - idx_base) : 0xffffffffu;
    }

    if (bits > 9) {
        // ---------------- bucket + rank ----------------
        int lg = 0;
        while ((1 << lg) < N) ++lg;
        const int shift = bits > lg ? bits - lg : 0;
        // an id outside the table (the kernels are unchecked, like the reference's @inbounds loops) must
        // not index outside the bucket array
        auto bucket_of = [shift](uint32_t k) { return min(k >> shift, (uint32_t)(N - 1)); };
        uint4* aux4 = reinterpret_cast<uint4*>(aux);
        for (int i = tid; i < N / 4; i += THREADS) aux4[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            if (base + i * 32 + lane < L) atomicAdd(&aux[bucket_of(key[i])], 1u);
        __syncthreads();
        // thread t owns buckets [t * ITEMS, (t + 1) * ITEMS): counts -> exclusive starts (16-byte accesses)
        uint32_t cnt[ITEMS];
        uint32_t mine = 0, big = 0;
#pragma unroll
        for (int j = 0; j < ITEMS / 4; ++j) {
            const uint4 q = aux4[tid * (ITEMS / 4) + j];
            cnt[4 * j] = q.x; cnt[4 * j + 1] = q.y; cnt[4 * j + 2] = q.z; cnt[4 * j + 3] = q.w;
        }
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            big |= cnt[j] > (uint32_t)kBucketRankMax;
            mine += cnt[j];
        }
        uint32_t run = block_excl_scan<THREADS>(mine, wsum);
        const int too_big = __syncthreads_or((int)big);
        if (!too_big) {
#pragma unroll
            for (int j = 0; j < ITEMS / 4; ++j) {
                uint4 q;
                q.x = run; run += cnt[4 * j];
                q.y = run; run += cnt[4 * j + 1];
                q.z = run; run += cnt[4 * j + 2];
                q.w = run; run += cnt[4 * j + 3];
                aux4[tid * (ITEMS / 4) + j] = q;
            }
            __syncthreads();
            // scatter into the bucket ranges (arbitrary order inside a bucket); afterwards aux[b] = end of bucket b
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const int e = base + i * 32 + lane;
                if (e < L) {
                    const uint32_t slot = atomicAdd(&aux[bucket_of(key[i])], 1u);
                    ksm[slot] = key[i];
                    vsm[slot] = (uint16_t)e;
                }
            }
            __syncthreads();
            // exact rank inside the bucket: pairs that compare lower on (key, position)
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const int s = tid + i * THREADS;
                if (s < L) {
                    const uint32_t k = ksm[s];
                    const uint32_t p = vsm[s];
                    const uint32_t b = bucket_of(k);
                    const uint32_t lo = b ? aux[b - 1] : 0u;
                    const uint32_t hi = aux[b];
                    uint32_t r = lo;
                    for (uint32_t j = lo; j < hi; ++j) {
                        const uint32_t kj = ksm[j];
                        const uint32_t pj = vsm[j];
                        r += (kj < k) || (kj == k && pj < p);
                    }
                    ko[r] = k;
                    po[r] = p;
                }
            }
            return;
        }
        __syncthreads();   // fall through to the radix passes (aux is re-initialised there)
    }

    // ---------------- LSD radix ----------------
    const int passes = (bits + 8) / 9;
    const int per = (bits + passes - 1) / passes;   // <= 9
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* wh = aux;                                          // [NW][NB]
    uint32_t vr[ITEMS];    // low 16 bits: original position, high 16 bits: rank inside the warp's chunk
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) vr[i] = (uint32_t)(base + i * 32 + lane);
    int shift = 0;
    for (int p = 0; p < passes; ++p) {
        const int nbits = (bits - shift) < per ? (bits - shift) : per;
        const uint32_t mask = (1u << nbits) - 1u;
        const int nb = 1 << nbits;
        // padding keys (0xffffffff) take the top digit of every pass and sit behind every real key of
        // that digit in input order, so they stay at the end of the stream
        for (int i = tid; i < NW * NB; i += THREADS)
            if ((i & (NB - 1)) < nb) wh[i] = 0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t dig = (key[i] >> shift) & mask;
            const uint32_t peers = __match_any_sync(0xffffffffu, dig);
            const uint32_t lt = peers & lt_mask;
            const uint32_t b = wh[w * NB + dig];
            __syncwarp();
            if (lt == 0) wh[w * NB + dig] = b + __popc(peers);
            __syncwarp();
            vr[i] = (vr[i] & 0xffffu) | ((b + __popc(lt)) << 16);
        }
        __syncthreads();
        {   // thread t owns digits [t * BPT, (t + 1) * BPT): bucket starts = exclusive scan of the digit totals
            uint32_t mine = 0;
#pragma unroll
            for (int j = 0; j < BPT; ++j) {
                const int d = tid * BPT + j;
                if (d < nb)
                    for (int ww = 0; ww < NW; ++ww) mine += wh[ww * NB + d];
            }
            uint32_t run = block_excl_scan<THREADS>(mine, wsum);
#pragma unroll
            for (int j = 0; j < BPT; ++j) {
                const int d = tid * BPT + j;
                if (d < nb) {
                    for (int ww = 0; ww < NW; ++ww) {
                        const uint32_t c = wh[ww * NB + d];
                        wh[ww * NB + d] = run;
                        run += c;
                    }
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t dst = wh[w * NB + ((key[i] >> shift) & mask)] + (vr[i] >> 16);
            ksm[dst] = key[i];
            vsm[dst] = (uint16_t)(vr[i] & 0xffffu);
        }
        __syncthreads();
        if (p + 1 < passes) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const int e = base + i * 32 + lane;
                key[i] = ksm[e];
                vr[i] = vsm[e];
            }
        }
        shift += nbits;
    }
    for (int i = tid; i < L; i += THREADS) {
        ko[i] = ksm[i];
        po[i] = vsm[i];
    }
}

}  // namespace dlrmb
