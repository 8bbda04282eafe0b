// Sparse gradient scatter-add fused with the in-place SGD row update.
//
// Replaces the lookup pullback -> SparseEmbeddingUpdate(delta, indices) (DLRM.jl
// src/train/train.jl:144; densified meaning pinned by test/train/backprop.jl:148-158) and
// EmbeddingTables.update!(Flux.Descent(lr), tables, updates, indexers; num_splits, nthreads)
// (call site src/train/train.jl:283-290; Flux Descent = `delta .*= lr; x .-= delta`):
//     table_k[r] -= lr * sum_{(b,p): idx_k[b][p] = r} dT[b][slot0 + k]
//
// Input is the per-table stream sorted by row id (sort.cu).  The stream is cut into fixed
// tiles of `tile` consecutive entries; one lane group (one lane per 16-byte chunk of a row)
// walks a tile in order, accumulating gradient rows in registers while the row id stays the
// same -- a segmented reduction whose segments are the duplicate runs.  A run that lies inside
// one tile ends in exactly one read-modify-write of its table row by its own lane group: no
// atomics, fixed order.  A run that crosses tile boundaries leaves per-tile partial sums in a
// scratch buffer ("head" for the tile where the run starts, "carry" for the tiles it continues
// into) and the head tile is appended to a work list; a second small kernel gives every listed
// run one CTA, whose lane groups add the carried partial sums in a fixed strided order and
// perform the row's single read-modify-write.  Results are therefore bit-reproducible run to run, and hot rows (Zipf
// heads, tiny tables) cost a bounded, evenly spread amount of work per lane group.
//
// HBM traffic per entry: 8 B of (key, position), one D*4-byte gradient row read, and per
// distinct row one D*4-byte read + one D*4-byte write.  Loads are issued U entries ahead
// (gradient rows and the table rows of the runs that end inside the batch) before the
// dependent adds, so each lane keeps up to 2U 16-byte requests in flight.
#include "common.cuh"

namespace dlrmb {

template <int VEC> struct UV;
template <> struct UV<4> {
    using type = float4;
    static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ float4 add(float4 a, float4 b) {
        return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
    }
    // Flux.Descent: delta *= lr; x -= delta  (two roundings, never contracted into an FMA)
    static __device__ __forceinline__ float4 sgd(float4 x, float4 g, float lr) {
        return make_float4(__fsub_rn(x.x, __fmul_rn(lr, g.x)), __fsub_rn(x.y, __fmul_rn(lr, g.y)),
                           __fsub_rn(x.z, __fmul_rn(lr, g.z)), __fsub_rn(x.w, __fmul_rn(lr, g.w)));
    }
};
template <> struct UV<1> {
    using type = float;
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sgd(float x, float g, float lr) {
        return __fsub_rn(x, __fmul_rn(lr, g));
    }
};

enum : uint8_t { FLAG_CARRY_IN = 1, FLAG_CARRY_ENDS = 2, FLAG_HEAD = 4, FLAG_ABSORBED = 8 };
constexpr int kAbsorbSeg = 64;   // tiles per CTA of the absorb pass

struct UpdateGeom {
    int L, P, tiles, tile, lpr_log2, C, slots, slot0;
    int64_t cap, ptiles_cap;
    int total_groups;
};

// Batches are 4 entries: the 4 keys and 4 positions of a batch are one 16-byte load each (the
// per-table streams are 16-byte aligned and tiles start at multiples of 4), plus one look-ahead
// key.  The next batch's keys/positions are requested before the current batch's rows, so a batch
// costs one memory round trip, not two.
template <int VEC, int NCH>
__global__ void __launch_bounds__(256, (NCH <= 1) ? 3 : ((NCH <= 2) ? 2 : 1))
update_tiles_kernel(const TableDesc* __restrict__ desc, const uint32_t* __restrict__ keys,
                    const uint32_t* __restrict__ pos, const float* __restrict__ dT, float lr,
                    float* __restrict__ partial, uint8_t* __restrict__ flags,
                    uint32_t* __restrict__ head_list, uint32_t* __restrict__ head_count, UpdateGeom gm) {
    using V = typename UV<VEC>::type;
    constexpr int U = 4;
    const int lpr = 1 << gm.lpr_log2;
    const int gid = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> gm.lpr_log2);
    if (gid >= gm.total_groups) return;
    const int sl = threadIdx.x & (lpr - 1);
    const int k = gid / gm.tiles;
    const int g = gid - k * gm.tiles;
    const size_t D = (size_t)gm.C * VEC;

    const uint32_t* __restrict__ ks = keys + (size_t)k * gm.cap;
    const uint32_t* __restrict__ ps = pos + (size_t)k * gm.cap;
    float* tb = desc[k].base;
    const float* __restrict__ gbase = dT + (size_t)(gm.slot0 + k) * D;
    const size_t gstride = (size_t)gm.slots * D;

    const int e0 = g * gm.tile;
    const int e1 = min(gm.L, e0 + gm.tile);

    uint4 kq = __ldg(reinterpret_cast<const uint4*>(ks + e0));
    uint4 pq = __ldg(reinterpret_cast<const uint4*>(ps + e0));
    uint32_t kn = (e0 + U < gm.L) ? __ldg(ks + e0 + U) : 0xffffffffu;
    const uint32_t kprev = (e0 > 0) ? __ldg(ks + e0 - 1) : 0xffffffffu;

    bool chunk_ok[NCH];
#pragma unroll
    for (int m = 0; m < NCH; ++m) chunk_ok[m] = (sl + m * lpr) < gm.C;

    V acc[NCH];
#pragma unroll
    for (int m = 0; m < NCH; ++m) acc[m] = UV<VEC>::zero();
    const bool cin = e0 > 0 && kprev == kq.x;
    bool cout = false;
    bool first = cin;
    uint8_t fl = cin ? FLAG_CARRY_IN : 0;
    float* pcarry = partial + (((size_t)k * gm.ptiles_cap + g) * 2 + 0) * D;
    float* phead = partial + (((size_t)k * gm.ptiles_cap + g) * 2 + 1) * D;

    for (int e = e0; e < e1; e += U) {
        const uint32_t key[U + 1] = {kq.x, kq.y, kq.z, kq.w, kn};
        const uint32_t pp[U] = {pq.x, pq.y, pq.z, pq.w};
        if (e + U < e1) {   // request the next batch's keys / positions now
            kq = __ldg(reinterpret_cast<const uint4*>(ks + e + U));
            pq = __ldg(reinterpret_cast<const uint4*>(ps + e + U));
            kn = (e + 2 * U < gm.L) ? __ldg(ks + e + 2 * U) : 0xffffffffu;
        }
        bool valid[U], is_end[U];
        V dv[U][NCH], rv[U][NCH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int ee = e + u;
            valid[u] = ee < e1;
            is_end[u] = valid[u] && ((ee + 1 >= gm.L) || (key[u + 1] != key[u]));
            const uint32_t b = (gm.P == 1) ? pp[u] : pp[u] / (uint32_t)gm.P;
            const V* src = reinterpret_cast<const V*>(gbase + (size_t)b * gstride);
#pragma unroll
            for (int m = 0; m < NCH; ++m)
                if (valid[u] && chunk_ok[m]) dv[u][m] = __ldg(src + sl + m * lpr);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const V* row = reinterpret_cast<const V*>(tb + (size_t)key[u] * D);
#pragma unroll
            for (int m = 0; m < NCH; ++m)
                if (is_end[u] && chunk_ok[m]) rv[u][m] = row[sl + m * lpr];
        }
        if (e + U >= e1) cout = (e1 < gm.L) && (e + U == e1) && (key[U] == key[U - 1]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!valid[u]) continue;
#pragma unroll
            for (int m = 0; m < NCH; ++m)
                if (chunk_ok[m]) acc[m] = UV<VEC>::add(acc[m], dv[u][m]);
            if (is_end[u]) {
                if (first) {
                    // run started in an earlier tile: hand the partial sum to its owner
#pragma unroll
                    for (int m = 0; m < NCH; ++m)
                        if (chunk_ok[m]) reinterpret_cast<V*>(pcarry)[sl + m * lpr] = acc[m];
                    fl |= FLAG_CARRY_ENDS;
                } else {
                    V* row = reinterpret_cast<V*>(tb + (size_t)key[u] * D);
#pragma unroll
                    for (int m = 0; m < NCH; ++m)
                        if (chunk_ok[m]) row[sl + m * lpr] = UV<VEC>::sgd(rv[u][m], acc[m], lr);
                }
                first = false;
#pragma unroll
                for (int m = 0; m < NCH; ++m) acc[m] = UV<VEC>::zero();
            }
        }
    }
    if (cout) {
        float* dst = first ? pcarry : phead;
        if (!first) fl |= FLAG_HEAD;
#pragma unroll
        for (int m = 0; m < NCH; ++m)
            if (chunk_ok[m]) reinterpret_cast<V*>(dst)[sl + m * lpr] = acc[m];
    }
    if (sl == 0) {
        flags[(size_t)k * gm.ptiles_cap + g] = fl;
        if (fl & FLAG_HEAD) head_list[atomicAdd(head_count, 1u)] = (uint32_t)gid;
    }
}

// Very long runs (hot rows: Zipf heads, tables with a handful of rows) cross hundreds or
// thousands of tiles, nearly all of them "pass-through" tiles that hold nothing but that one
// run.  Before the per-run fix-up, every CTA of this pass takes kAbsorbSeg consecutive tiles and
// folds each stretch of consecutive pass-through tiles into its first tile's carry slot (ascending
// order), marking the others absorbed.  The fix-up CTA of a run then only adds one partial per
// stretch, so a run of n tiles costs n / kAbsorbSeg adds on its critical path instead of n.
template <int VEC, int NCH>
__global__ void __launch_bounds__(256)
update_absorb_kernel(float* __restrict__ partial, uint8_t* __restrict__ flags, UpdateGeom gm) {
    using V = typename UV<VEC>::type;
    __shared__ V red[256 * NCH];
    __shared__ uint8_t f[kAbsorbSeg];
    __shared__ int n_pass;
    const int k = blockIdx.y;
    const int t0 = blockIdx.x * kAbsorbSeg;
    const int t1 = min(gm.tiles, t0 + kAbsorbSeg);
    const int tid = threadIdx.x;
    uint8_t* fk = flags + (size_t)k * gm.ptiles_cap;
    if (tid == 0) n_pass = 0;
    __syncthreads();
    if (tid < kAbsorbSeg) {
        const uint8_t v = (t0 + tid < t1) ? fk[t0 + tid] : (uint8_t)0xff;
        f[tid] = v;
        if (v == FLAG_CARRY_IN) atomicAdd(&n_pass, 1);
    }
    __syncthreads();
    if (n_pass < 2) return;

    const int lpr = 1 << gm.lpr_log2;
    const int sl = tid & (lpr - 1);
    const int sub = tid >> gm.lpr_log2;
    const int nsub = 256 >> gm.lpr_log2;
    const size_t D = (size_t)gm.C * VEC;
    float* pk = partial + (size_t)k * gm.ptiles_cap * 2 * D;
    int t = 0;
    const int nt = t1 - t0;
    while (t < nt) {                     // uniform across the CTA: f[] is in shared memory
        if (f[t] != FLAG_CARRY_IN) { ++t; continue; }
        int len = 1;
        while (t + len < nt && f[t + len] == FLAG_CARRY_IN) ++len;
        if (len >= 2) {
            V acc[NCH];
#pragma unroll
            for (int m = 0; m < NCH; ++m) acc[m] = UV<VEC>::zero();
            for (int j = sub; j < len; j += nsub) {      // group `sub`: tiles t+sub, t+sub+nsub, ...
                const V* src = reinterpret_cast<const V*>(pk + (size_t)(t0 + t + j) * 2 * D);
#pragma unroll
                for (int m = 0; m < NCH; ++m)
                    if (sl + m * lpr < gm.C) acc[m] = UV<VEC>::add(acc[m], src[sl + m * lpr]);
            }
#pragma unroll
            for (int m = 0; m < NCH; ++m) red[(sub * NCH + m) * lpr + sl] = acc[m];
            __syncthreads();
            if (sub == 0) {
                const int nact = min(nsub, len);
                V* dst = reinterpret_cast<V*>(pk + (size_t)(t0 + t) * 2 * D);
#pragma unroll
                for (int m = 0; m < NCH; ++m) {
                    if (sl + m * lpr < gm.C) {
                        V total = red[m * lpr + sl];
                        for (int j = 1; j < nact; ++j) total = UV<VEC>::add(total, red[(j * NCH + m) * lpr + sl]);
                        dst[sl + m * lpr] = total;
                    }
                }
            }
            if (tid >= 1 && tid < len) fk[t0 + t + tid] = (uint8_t)(FLAG_CARRY_IN | FLAG_ABSORBED);
            __syncthreads();
        }
        t += len;
    }
}

// Runs that cross tile boundaries.  One CTA per listed head tile: the end of the run is found by
// probing the tile flags 256 at a time, lane group `sub` adds the carry partials of tiles
// g+1+sub, g+1+sub+nsub, ... (ascending, four loads in flight), the groups' sums are then added
// in group order on top of the head partial, and the row is updated once.  The order depends only
// on the geometry, so the result is bit-reproducible (the work list's order is not, but no
// arithmetic depends on it).
template <int VEC, int NCH>
__global__ void __launch_bounds__(256)
update_fixup_kernel(const TableDesc* __restrict__ desc, const uint32_t* __restrict__ keys, float lr,
                    const float* __restrict__ partial, const uint8_t* __restrict__ flags,
                    const uint32_t* __restrict__ head_list, const uint32_t* __restrict__ head_count,
                    UpdateGeom gm) {
    using V = typename UV<VEC>::type;
    __shared__ V red[256 * NCH];
    __shared__ int s_first;
    const int lpr = 1 << gm.lpr_log2;
    const int tid = threadIdx.x;
    const int sl = tid & (lpr - 1);
    const int sub = tid >> gm.lpr_log2;
    const int nsub = 256 >> gm.lpr_log2;
    const size_t D = (size_t)gm.C * VEC;
    const uint32_t n_heads = *head_count;

    for (uint32_t h = blockIdx.x; h < n_heads; h += gridDim.x) {
        const int gid = (int)head_list[h];
        const int k = gid / gm.tiles;
        const int g = gid - k * gm.tiles;
        const uint8_t* fk = flags + (size_t)k * gm.ptiles_cap;
        const float* pk = partial + (size_t)k * gm.ptiles_cap * 2 * D;
        // where the run ends: the first later tile whose carried run stops inside it.  All 256
        // threads probe one tile flag each per round (the table's last tile always stops it).
        if (tid == 0) s_first = 0x7fffffff;
        __syncthreads();
        for (int base = g + 1;; base += 256) {
            const int u = base + tid;
            const uint8_t f = (u < gm.tiles) ? fk[u] : (uint8_t)FLAG_CARRY_ENDS;
            if (f & FLAG_CARRY_ENDS) atomicMin(&s_first, u);
            __syncthreads();
            if (s_first != 0x7fffffff) break;
        }
        const int u_last = s_first;
        V acc[NCH];
#pragma unroll
        for (int m = 0; m < NCH; ++m) acc[m] = UV<VEC>::zero();
        for (int u = g + 1 + sub; u <= u_last; u += 4 * nsub) {
            V v[4][NCH];
            bool take[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int uu = u + q * nsub;
                const V* src = reinterpret_cast<const V*>(pk + (size_t)uu * 2 * D);
                take[q] = uu <= u_last && !(fk[uu] & FLAG_ABSORBED);
#pragma unroll
                for (int m = 0; m < NCH; ++m)
                    if (take[q] && sl + m * lpr < gm.C) v[q][m] = src[sl + m * lpr];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int m = 0; m < NCH; ++m)
                    if (take[q] && sl + m * lpr < gm.C) acc[m] = UV<VEC>::add(acc[m], v[q][m]);
        }
#pragma unroll
        for (int m = 0; m < NCH; ++m) red[(sub * NCH + m) * lpr + sl] = acc[m];
        __syncthreads();
        if (sub == 0) {
            const int nact = min(nsub, u_last - g);
            const uint32_t key = keys[(size_t)k * gm.cap + (size_t)(g + 1) * gm.tile - 1];
            V* row = reinterpret_cast<V*>(desc[k].base + (size_t)key * D);
            const V* hp = reinterpret_cast<const V*>(pk + ((size_t)g * 2 + 1) * D);
#pragma unroll
            for (int m = 0; m < NCH; ++m) {
                if (sl + m * lpr < gm.C) {
                    V total = hp[sl + m * lpr];
                    for (int j = 0; j < nact; ++j) total = UV<VEC>::add(total, red[(j * NCH + m) * lpr + sl]);
                    row[sl + m * lpr] = UV<VEC>::sgd(row[sl + m * lpr], total, lr);
                }
            }
        }
        __syncthreads();
    }
}

static int lanes_per_row_log2(int C) {
    int l = 0;
    while ((1 << l) < C && l < 5) ++l;
    return l;
}

// Entries per lane-group tile: as small as 4 while the whole batch still fits one wave of the
// machine (latency-bound regime of DLRM-sized batches), up to 32 for large batches.
static int choose_update_tile(int64_t total_entries, int lpr, int sm_count) {
    const int64_t capacity = (int64_t)sm_count * 2048 / lpr;
    int tile = 4;
    while (tile < 32 && total_entries / tile > capacity) tile *= 2;
    return tile;
}

int64_t update_tiles_cap(int ntab, int D, int64_t max_lookups, int sm_count) {
    const int vec = (D % 4 == 0) ? 4 : 1;
    const int lpr = 1 << lanes_per_row_log2(D / vec);
    int64_t best = 1;
    // the tile count peaks either right below a tile-size switch or at max_lookups
    for (int tile = 4; tile <= 32; tile *= 2) {
        int64_t capacity = (int64_t)sm_count * 2048 / lpr;
        int64_t lmax = (tile == 32) ? max_lookups : (capacity * tile) / ntab + 1;
        if (lmax > max_lookups) lmax = max_lookups;
        int64_t tiles = ceil_div64(lmax, tile) + 1;
        if (tiles > best) best = tiles;
    }
    return best;
}

template <int VEC, int NCH>
static int launch_update_t(dlrmb_tables* t, const float* dT, int slots, int slot0, float lr,
                           cudaStream_t s) {
    UpdateGeom gm;
    gm.L = t->sorted_B * t->sorted_P;
    gm.P = t->sorted_P;
    gm.C = t->D / VEC;
    gm.lpr_log2 = lanes_per_row_log2(gm.C);
    const int lpr = 1 << gm.lpr_log2;
    gm.tile = choose_update_tile((int64_t)t->ntab * gm.L, lpr, t->sm_count);
    gm.tiles = (gm.L + gm.tile - 1) / gm.tile;
    gm.slots = slots;
    gm.slot0 = slot0;
    gm.cap = t->cap;
    gm.ptiles_cap = t->partial_tiles_cap;
    DLRMB_REQUIRE(gm.tiles <= gm.ptiles_cap, "internal: update tile capacity exceeded (%d > %lld)",
                  gm.tiles, (long long)gm.ptiles_cap);
    const int64_t groups = (int64_t)t->ntab * gm.tiles;
    DLRMB_REQUIRE(groups < (1ll << 31) / 32, "batch too large for one update launch");
    gm.total_groups = (int)groups;
    const int groups_per_block = 256 / lpr;
    const unsigned grid = (unsigned)ceil_div64(groups, groups_per_block);
    static_assert(sizeof(uint4) == 16, "batch loads are 16 bytes");
    const uint32_t* keys = t->keys[t->sorted_buf];
    const uint32_t* pos = t->pos[t->sorted_buf];
    DLRMB_CUDA(cudaMemsetAsync(t->head_count, 0, sizeof(uint32_t), s));
    update_tiles_kernel<VEC, NCH><<<grid, 256, 0, s>>>(t->d_desc, keys, pos, dT, lr, t->partial, t->tile_flags,
                                                          t->head_list, t->head_count, gm);
    DLRMB_LAUNCH_CHECK();
    if (false && gm.tiles >= 2 * kAbsorbSeg) {   // TODO(absorb v2)
        dim3 agrid((unsigned)ceil_div64(gm.tiles, kAbsorbSeg), (unsigned)t->ntab);
        update_absorb_kernel<VEC, NCH><<<agrid, 256, 0, s>>>(t->partial, t->tile_flags, gm);
        DLRMB_LAUNCH_CHECK();
    }
    unsigned fgrid = (unsigned)(groups < (int64_t)t->sm_count * 8 ? groups : (int64_t)t->sm_count * 8);
    update_fixup_kernel<VEC, NCH><<<fgrid, 256, 0, s>>>(t->d_desc, keys, lr, t->partial, t->tile_flags,
                                                        t->head_list, t->head_count, gm);
    DLRMB_LAUNCH_CHECK();
    return DLRMB_OK;
}

int launch_update(dlrmb_tables* t, const float* dT, int slots, int slot0, float lr, cudaStream_t s) {
    const bool vec4 = (t->D % 4 == 0) && ((reinterpret_cast<uintptr_t>(dT) & 15) == 0);
    if (t->D % 4 == 0 && !vec4) {
        set_error("dT must be 16-byte aligned when D is a multiple of 4");
        return DLRMB_EINVAL;
    }
    const int C = vec4 ? t->D / 4 : t->D;
    const int lpr = 1 << lanes_per_row_log2(C);
    const int nch = (C + lpr - 1) / lpr;
    if (vec4) {
        if (nch == 1) return launch_update_t<4, 1>(t, dT, slots, slot0, lr, s);
        if (nch == 2) return launch_update_t<4, 2>(t, dT, slots, slot0, lr, s);
        if (nch <= 4) return launch_update_t<4, 4>(t, dT, slots, slot0, lr, s);
        if (nch <= 8) return launch_update_t<4, 8>(t, dT, slots, slot0, lr, s);
    } else {
        if (nch == 1) return launch_update_t<1, 1>(t, dT, slots, slot0, lr, s);
        if (nch == 2) return launch_update_t<1, 2>(t, dT, slots, slot0, lr, s);
        if (nch <= 4) return launch_update_t<1, 4>(t, dT, slots, slot0, lr, s);
        if (nch <= 8) return launch_update_t<1, 8>(t, dT, slots, slot0, lr, s);
    }
    set_error("embedding dim %d unsupported by the update kernel", t->D);
    return DLRMB_EINVAL;
}

}  // namespace dlrmb
