// Index sort / dedup for the sparse-gradient update.
//
// Role in the reference: EmbeddingTables.SparseIndexer, the per-table dictionary that
// compacts duplicate row ids before EmbeddingTables.update! (DLRM.jl src/train/train.jl:107-115,
// 276-290).  Here the dictionary is replaced by a stable sort of (row id, flat position) pairs per
// table, which makes duplicates adjacent and fixes their accumulation order (ascending flat
// position) so the update is deterministic without floating-point atomics.
//
// Two paths, chosen per call from L = B*P:
//   * L <= kSmemSortMax: one CTA per table runs the whole LSD radix sort in shared memory
//     (digits of up to 9 bits, as many passes as that table's row count needs); a single launch
//     covers all tables.  Training steps do not even pay that launch: dlrmb_embedding_fwd_sort runs
//     the same sort in extra CTAs of the lookup launch (lookup.cu), hidden behind the gather.
//   * larger L: least-significant-digit radix sort, digits of up to 9 bits, tiles of 4096 keys, all tables
//     batched through grid.y.  Per pass: per-tile digit histogram -> exclusive scan over
//     (digit, tile), one CTA per digit -> stable scatter whose in-tile ranks come from warp match_any + per-warp
//     digit counters in shared memory and whose keys are reordered in shared memory first, so each
//     digit's keys leave the CTA as one contiguous run.  The sort arrays of a whole batch fit L2 (126 MB), so the
//     passes run at L2 rather than HBM speed.
// Output: keys[sorted_buf] (ascending 0-based row ids) and pos[sorted_buf] (the stable
// permutation), both [ntab][max_lookups].
#include "common.cuh"
#include "sort_small.cuh"

namespace dlrmb {

// ---------------------------------------------------------------------------------------------
// small path: LSD radix sort held in shared memory, one CTA per table, one launch for all tables
// (body in sort_small.cuh; the fused lookup + sort launch of lookup.cu runs the same body)
// ---------------------------------------------------------------------------------------------
template <typename IdxT, int ITEMS, int THREADS>
__global__ void __launch_bounds__(THREADS)
sort_small_kernel(const IdxT* __restrict__ idx, int idx_base, int L, const TableDesc* __restrict__ desc,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ pos_out, int64_t cap,
                  unsigned long long* clk) {
    extern __shared__ uint32_t sort_smem[];
    const int k = blockIdx.x;
    clock_in(clk, blockIdx.x);
    sort_small_body<IdxT, ITEMS, THREADS>(idx + (size_t)k * L, idx_base, L, desc[k].rows,
                                          keys_out + (size_t)k * cap, pos_out + (size_t)k * cap, sort_smem);
    clock_out(clk, blockIdx.x);
}

template <typename IdxT, int ITEMS, int THREADS>
static int launch_sort_small(dlrmb_tables* t, const IdxT* idx, int idx_base, int L, cudaStream_t s) {
    constexpr size_t smem = SmallSortGeom<ITEMS, THREADS>::smem_bytes();
    static unsigned long long attr_done = 0;
    int rc = ensure_smem_attr((const void*)sort_small_kernel<IdxT, ITEMS, THREADS>, (int)smem, &attr_done);
    if (rc) return rc;
    sort_small_kernel<IdxT, ITEMS, THREADS><<<t->ntab, THREADS, smem, s>>>(idx, idx_base, L, t->d_desc, t->keys[0],
                                                                          t->pos[0], t->cap, clock_slot(CLK_SORT));
    DLRMB_LAUNCH_CHECK();
    return DLRMB_OK;
}

// ---------------------------------------------------------------------------------------------
// large path: LSD radix sort over global memory, all tables batched through grid.y
// ---------------------------------------------------------------------------------------------
constexpr int RT = 256;        // threads per CTA
constexpr int RI = 16;         // keys per thread
constexpr int RTILE = RT * RI; // keys per tile

// exclusive prefix of `v` over the 256 threads of the CTA (thread order); wsum is 8 words of smem
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* wsum, uint32_t* total_out) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    uint32_t run = inc - v, tot = 0;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) {
        const uint32_t s = wsum[ww];
        if (ww < w) run += s;
        tot += s;
    }
    if (total_out) *total_out = tot;
    __syncthreads();
    return run;
}

// The first pass reads the caller's index array directly (FIRST): no separate conversion pass.
template <typename IdxT, bool FIRST>
__device__ __forceinline__ uint32_t radix_load_key(const IdxT* __restrict__ idx, int idx_base,
                                                   const uint32_t* __restrict__ keys, int e) {
    if (FIRST) return (uint32_t)((int64_t)idx[e] - idx_base);
    return keys[e];
}

// Digits are up to 9 bits wide (NB = 512 bins max): the pass count is ceil(bits / 9) and the bits
// are spread evenly over the passes, so 1e8 rows (27 bits) take 3 passes and 1e5 rows (17 bits) 2.
constexpr int NB = 512;

template <typename IdxT, bool FIRST>
__global__ void __launch_bounds__(RT)
radix_hist_kernel(const IdxT* __restrict__ idx, int idx_base, const uint32_t* __restrict__ keys, int64_t cap,
                  int L, int shift, uint32_t mask, uint32_t* __restrict__ tile_hist, int tiles) {
    __shared__ uint32_t h[NB];
    const int k = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    h[tid] = 0;
    h[tid + RT] = 0;
    __syncthreads();
    const IdxT* ik = idx + (size_t)k * L;
    const uint32_t* kin = keys + (size_t)k * cap;
    const int base = tile * RTILE;
    uint32_t kk[RI];
#pragma unroll
    for (int i = 0; i < RI; ++i) {
        const int e = base + i * RT + tid;
        kk[i] = e < L ? radix_load_key<IdxT, FIRST>(ik, idx_base, kin, e) : 0u;
    }
#pragma unroll
    for (int i = 0; i < RI; ++i)
        if (base + i * RT + tid < L) atomicAdd(&h[(kk[i] >> shift) & mask], 1u);
    __syncthreads();
    for (uint32_t d = tid; d <= mask; d += RT) tile_hist[((size_t)k * NB + d) * tiles + tile] = h[d];
}

// One CTA per (digit, table): exclusive scan over that digit's per-tile counts (contiguous, so
// the loads coalesce), in place, plus the digit's total.
__global__ void __launch_bounds__(RT)
radix_scan_kernel(uint32_t* __restrict__ tile_hist, uint32_t* __restrict__ digit_total, int tiles) {
    __shared__ uint32_t wsum[8];
    const int d = blockIdx.x, k = blockIdx.y;
    uint32_t* h = tile_hist + ((size_t)k * NB + d) * tiles;
    uint32_t running = 0;
    for (int c = 0; c < tiles; c += RT) {
        const int i = c + threadIdx.x;
        const uint32_t v = i < tiles ? h[i] : 0u;
        uint32_t tot;
        const uint32_t ex = block_excl_scan_256(v, wsum, &tot);
        if (i < tiles) h[i] = running + ex;
        running += tot;
    }
    if (threadIdx.x == 0) digit_total[k * NB + d] = running;
}

// Scatter with a shared-memory reorder: the tile's keys are first placed in digit order in shared
// memory, then written out in that order, so the keys of one digit leave as one contiguous run
// (coalesced) instead of one 4-byte store per key.
template <typename IdxT, bool FIRST>
__global__ void __launch_bounds__(RT)
radix_scatter_kernel(const IdxT* __restrict__ idx, int idx_base, const uint32_t* __restrict__ keys_in,
                     const uint32_t* __restrict__ pos_in, uint32_t* __restrict__ keys_out,
                     uint32_t* __restrict__ pos_out, int64_t cap, int L, int shift, uint32_t mask,
                     const uint32_t* __restrict__ tile_hist, const uint32_t* __restrict__ digit_total, int tiles) {
    __shared__ uint16_t wc[RT / 32][NB];      // per-warp digit counts, then exclusive prefix over warps
    __shared__ uint32_t lstart[NB];           // first slot of each digit inside the sorted tile
    __shared__ uint32_t gdst[NB];             // global position of slot 0 of each digit, minus lstart
    __shared__ uint32_t skeys[RTILE];
    __shared__ uint32_t svals[RTILE];
    __shared__ uint32_t wsum[8];
    const int k = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int w = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (RT / 32) * NB; i += RT) (&wc[0][0])[i] = 0;
    __syncthreads();

    const IdxT* ik = idx + (size_t)k * L;
    const uint32_t* kin = keys_in + (size_t)k * cap;
    const uint32_t* pin = pos_in + (size_t)k * cap;
    const int tile_base = tile * RTILE;
    const int base = tile_base + w * (32 * RI);
    const int n_valid = min(RTILE, L - tile_base);
    uint32_t key[RI], val[RI], rank[RI];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RI; ++i) {
        const int e = base + i * 32 + lane;
        const bool valid = e < L;
        key[i] = valid ? radix_load_key<IdxT, FIRST>(ik, idx_base, kin, e) : 0xffffffffu;
        val[i] = FIRST ? (uint32_t)e : (valid ? pin[e] : 0u);
    }
#pragma unroll
    for (int i = 0; i < RI; ++i) {
        const bool valid = (base + i * 32 + lane) < L;
        const uint32_t dig = valid ? ((key[i] >> shift) & mask) : (NB + lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, dig);
        const uint32_t lt = peers & lt_mask;
        const uint32_t b = valid ? wc[w][dig] : 0u;
        __syncwarp();
        if (valid && lt == 0) wc[w][dig] = (uint16_t)(b + __popc(peers));
        __syncwarp();
        rank[i] = b + __popc(lt);
    }
    __syncthreads();
    {
        // thread t owns digits 2t and 2t+1
        const uint32_t d0 = 2 * tid, d1 = 2 * tid + 1;
        uint32_t c0 = 0, c1 = 0;
        if (d0 <= mask) {
#pragma unroll
            for (int ww = 0; ww < RT / 32; ++ww) {
                const uint32_t c = wc[ww][d0];
                wc[ww][d0] = (uint16_t)c0;
                c0 += c;
            }
        }
        if (d1 <= mask) {
#pragma unroll
            for (int ww = 0; ww < RT / 32; ++ww) {
                const uint32_t c = wc[ww][d1];
                wc[ww][d1] = (uint16_t)c1;
                c1 += c;
            }
        }
        const uint32_t lex = block_excl_scan_256(c0 + c1, wsum, nullptr);       // slots inside the tile
        const uint32_t t0 = d0 <= mask ? digit_total[k * NB + d0] : 0u;
        const uint32_t t1 = d1 <= mask ? digit_total[k * NB + d1] : 0u;
        const uint32_t gex = block_excl_scan_256(t0 + t1, wsum, nullptr);       // buckets in the stream
        if (d0 <= mask) {
            lstart[d0] = lex;
            gdst[d0] = gex + tile_hist[((size_t)k * NB + d0) * tiles + tile] - lex;
        }
        if (d1 <= mask) {
            lstart[d1] = lex + c0;
            gdst[d1] = gex + t0 + tile_hist[((size_t)k * NB + d1) * tiles + tile] - (lex + c0);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RI; ++i) {
        if ((base + i * 32 + lane) < L) {
            const uint32_t dig = (key[i] >> shift) & mask;
            const uint32_t slot = lstart[dig] + wc[w][dig] + rank[i];
            skeys[slot] = key[i];
            svals[slot] = val[i];
        }
    }
    __syncthreads();
    uint32_t* ko = keys_out + (size_t)k * cap;
    uint32_t* po = pos_out + (size_t)k * cap;
    for (int j = tid; j < n_valid; j += RT) {
        const uint32_t kk = skeys[j];
        const uint32_t dst = gdst[(kk >> shift) & mask] + (uint32_t)j;
        ko[dst] = kk;
        po[dst] = svals[j];
    }
}

static int key_bits(int64_t max_rows) {
    int bits = 1;
    while (bits < 32 && (1ll << bits) < max_rows) ++bits;
    return bits;
}

template <typename IdxT>
static int launch_sort_t(dlrmb_tables* t, const IdxT* idx, int idx_base, int B, int P, cudaStream_t s) {
    const int64_t L64 = (int64_t)B * P;
    const int L = (int)L64;
    const int64_t cap = t->cap;
    if (L <= kSmemSortMax) {
        int rc;
        if (L <= 1024) rc = launch_sort_small<IdxT, 4, 256>(t, idx, idx_base, L, s);
        else if (L <= 2048) rc = launch_sort_small<IdxT, 8, 256>(t, idx, idx_base, L, s);
        else if (L <= 4096) rc = launch_sort_small<IdxT, 16, 256>(t, idx, idx_base, L, s);
        else if (L <= 8192) rc = launch_sort_small<IdxT, 8, 1024>(t, idx, idx_base, L, s);
        else rc = launch_sort_small<IdxT, 16, 1024>(t, idx, idx_base, L, s);
        if (rc) return rc;
        t->sorted_buf = 0;
        return DLRMB_OK;
    }
    const int tiles = (int)ceil_div64(L, RTILE);
    DLRMB_REQUIRE(tiles <= t->radix_tiles_cap, "internal: radix tile capacity exceeded");
    const int bits = key_bits(t->max_rows);
    const int passes = (bits + 8) / 9;
    const int per_pass = (bits + passes - 1) / passes;     // <= 9
    int cur = 0;
    dim3 grid((unsigned)tiles, (unsigned)t->ntab);
    int shift = 0;
    for (int p = 0; p < passes; ++p) {
        const int nb = (bits - shift) < per_pass ? (bits - shift) : per_pass;
        const uint32_t mask = (1u << nb) - 1u;
        dim3 sgrid(mask + 1u, (unsigned)t->ntab);
        if (p == 0)
            radix_hist_kernel<IdxT, true><<<grid, RT, 0, s>>>(idx, idx_base, t->keys[cur], cap, L, shift, mask, t->tile_hist, tiles);
        else
            radix_hist_kernel<IdxT, false><<<grid, RT, 0, s>>>(idx, idx_base, t->keys[cur], cap, L, shift, mask, t->tile_hist, tiles);
        DLRMB_LAUNCH_CHECK();
        radix_scan_kernel<<<sgrid, RT, 0, s>>>(t->tile_hist, t->digit_total, tiles);
        DLRMB_LAUNCH_CHECK();
        if (p == 0)
            radix_scatter_kernel<IdxT, true><<<grid, RT, 0, s>>>(idx, idx_base, t->keys[cur], t->pos[cur], t->keys[cur ^ 1],
                                                                t->pos[cur ^ 1], cap, L, shift, mask, t->tile_hist, t->digit_total, tiles);
        else
            radix_scatter_kernel<IdxT, false><<<grid, RT, 0, s>>>(idx, idx_base, t->keys[cur], t->pos[cur], t->keys[cur ^ 1],
                                                                 t->pos[cur ^ 1], cap, L, shift, mask, t->tile_hist, t->digit_total, tiles);
        DLRMB_LAUNCH_CHECK();
        cur ^= 1;
        shift += nb;
    }
    t->sorted_buf = cur;
    return DLRMB_OK;
}

int launch_sort(dlrmb_tables* t, const void* idx, int idx_bytes, int idx_base, int B, int P,
                cudaStream_t s) {
    if (idx_bytes == 4) return launch_sort_t<uint32_t>(t, static_cast<const uint32_t*>(idx), idx_base, B, P, s);
    return launch_sort_t<int64_t>(t, static_cast<const int64_t*>(idx), idx_base, B, P, s);
}

// ---------------------------------------------------------------------------------------------
// parity export: unique ids + segment offsets of one table's sorted stream (single CTA)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
dedup_export_kernel(const uint32_t* __restrict__ keys, int L, int64_t* __restrict__ uniq,
                    int32_t* __restrict__ seg, int32_t* __restrict__ n_uniq) {
    __shared__ int warp_cnt[32];
    __shared__ int running;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int base = 0; base < L; base += 1024) {
        int i = base + tid;
        bool head = i < L && (i == 0 || keys[i] != keys[i - 1]);
        uint32_t bal = __ballot_sync(0xffffffffu, head);
        if (lane == 0) warp_cnt[w] = __popc(bal);
        __syncthreads();
        int before = running;
        for (int ww = 0; ww < w; ++ww) before += warp_cnt[ww];
        int total = 0;
        for (int ww = 0; ww < 32; ++ww) total += warp_cnt[ww];
        if (head) {
            int o = before + __popc(bal & ((1u << lane) - 1u));
            uniq[o] = (int64_t)keys[i];
            seg[o] = i;
        }
        __syncthreads();
        if (tid == 0) running += total;
        __syncthreads();
    }
    if (tid == 0) {
        seg[running] = L;
        *n_uniq = running;
    }
}

int launch_dedup_export(dlrmb_tables* t, int k, cudaStream_t s) {
    const int L = t->sorted_B * t->sorted_P;
    dedup_export_kernel<<<1, 1024, 0, s>>>(t->keys[t->sorted_buf] + (size_t)k * t->cap, L,
                                           t->d_uniq, t->d_seg, t->d_nuniq);
    DLRMB_LAUNCH_CHECK();
    return DLRMB_OK;
}

}  // namespace dlrmb
