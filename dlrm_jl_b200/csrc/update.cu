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
// into); a second small kernel adds them in ascending tile order and performs the row's single
// read-modify-write.  Results are therefore bit-reproducible run to run, and hot rows (Zipf
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

enum : uint8_t { FLAG_CARRY_IN = 1, FLAG_CARRY_ENDS = 2, FLAG_HEAD = 4 };

struct UpdateGeom {
    int L, P, tiles, tile, lpr_log2, C, slots, slot0;
    int64_t cap, ptiles_cap;
    int total_groups;
};

template <int VEC, int NCH, int U>
__global__ void __launch_bounds__(256)
update_tiles_kernel(const TableDesc* __restrict__ desc, const uint32_t* __restrict__ keys,
                    const uint32_t* __restrict__ pos, const float* __restrict__ dT, float lr,
                    float* __restrict__ partial, uint8_t* __restrict__ flags, UpdateGeom gm) {
    using V = typename UV<VEC>::type;
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
    const bool cin = e0 > 0 && __ldg(ks + e0 - 1) == __ldg(ks + e0);
    const bool cout = e1 < gm.L && __ldg(ks + e1) == __ldg(ks + e1 - 1);

    bool chunk_ok[NCH];
#pragma unroll
    for (int m = 0; m < NCH; ++m) chunk_ok[m] = (sl + m * lpr) < gm.C;

    V acc[NCH];
#pragma unroll
    for (int m = 0; m < NCH; ++m) acc[m] = UV<VEC>::zero();
    bool first = cin;
    uint8_t fl = cin ? FLAG_CARRY_IN : 0;
    float* pcarry = partial + (((size_t)k * gm.ptiles_cap + g) * 2 + 0) * D;
    float* phead = partial + (((size_t)k * gm.ptiles_cap + g) * 2 + 1) * D;

    for (int e = e0; e < e1; e += U) {
        uint32_t key[U];
        bool valid[U], is_end[U];
        V dv[U][NCH], rv[U][NCH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int ee = e + u;
            valid[u] = ee < e1;
            key[u] = valid[u] ? __ldg(ks + ee) : 0u;
            const uint32_t p = valid[u] ? __ldg(ps + ee) : 0u;
            is_end[u] = valid[u] && ((ee + 1 >= gm.L) || (__ldg(ks + ee + 1) != key[u]));
            const uint32_t b = (gm.P == 1) ? p : p / (uint32_t)gm.P;
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
    if (sl == 0) flags[(size_t)k * gm.ptiles_cap + g] = fl;
}

// Runs that cross tile boundaries: the lane group of the tile where the run starts adds the
// carried partial sums in ascending tile order and applies the row update once.
template <int VEC, int NCH>
__global__ void __launch_bounds__(256)
update_fixup_kernel(const TableDesc* __restrict__ desc, const uint32_t* __restrict__ keys, float lr,
                    const float* __restrict__ partial, const uint8_t* __restrict__ flags,
                    UpdateGeom gm) {
    using V = typename UV<VEC>::type;
    const int lpr = 1 << gm.lpr_log2;
    const int gid = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> gm.lpr_log2);
    if (gid >= gm.total_groups) return;
    const int sl = threadIdx.x & (lpr - 1);
    const int k = gid / gm.tiles;
    const int g = gid - k * gm.tiles;
    const uint8_t* fk = flags + (size_t)k * gm.ptiles_cap;
    if (!(fk[g] & FLAG_HEAD)) return;
    const size_t D = (size_t)gm.C * VEC;
    const float* pk = partial + (size_t)k * gm.ptiles_cap * 2 * D;

    V acc[NCH];
#pragma unroll
    for (int m = 0; m < NCH; ++m) {
        acc[m] = UV<VEC>::zero();
        if (sl + m * lpr < gm.C)
            acc[m] = reinterpret_cast<const V*>(pk + ((size_t)g * 2 + 1) * D)[sl + m * lpr];
    }
    for (int u = g + 1; u < gm.tiles; ++u) {
        const uint8_t f = fk[u];
#pragma unroll
        for (int m = 0; m < NCH; ++m)
            if (sl + m * lpr < gm.C)
                acc[m] = UV<VEC>::add(acc[m], reinterpret_cast<const V*>(pk + (size_t)u * 2 * D)[sl + m * lpr]);
        if (f & FLAG_CARRY_ENDS) break;
    }
    const uint32_t key = keys[(size_t)k * gm.cap + (size_t)(g + 1) * gm.tile - 1];
    V* row = reinterpret_cast<V*>(desc[k].base + (size_t)key * D);
#pragma unroll
    for (int m = 0; m < NCH; ++m)
        if (sl + m * lpr < gm.C) row[sl + m * lpr] = UV<VEC>::sgd(row[sl + m * lpr], acc[m], lr);
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
    gm.cap = t->max_lookups;
    gm.ptiles_cap = t->partial_tiles_cap;
    DLRMB_REQUIRE(gm.tiles <= gm.ptiles_cap, "internal: update tile capacity exceeded (%d > %lld)",
                  gm.tiles, (long long)gm.ptiles_cap);
    const int64_t groups = (int64_t)t->ntab * gm.tiles;
    DLRMB_REQUIRE(groups < (1ll << 31) / 32, "batch too large for one update launch");
    gm.total_groups = (int)groups;
    const int groups_per_block = 256 / lpr;
    const unsigned grid = (unsigned)ceil_div64(groups, groups_per_block);
    constexpr int U = (NCH <= 2) ? 4 : 2;
    const uint32_t* keys = t->keys[t->sorted_buf];
    const uint32_t* pos = t->pos[t->sorted_buf];
    update_tiles_kernel<VEC, NCH, U><<<grid, 256, 0, s>>>(t->d_desc, keys, pos, dT, lr, t->partial, t->tile_flags, gm);
    DLRMB_LAUNCH_CHECK();
    update_fixup_kernel<VEC, NCH><<<grid, 256, 0, s>>>(t->d_desc, keys, lr, t->partial, t->tile_flags, gm);
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
