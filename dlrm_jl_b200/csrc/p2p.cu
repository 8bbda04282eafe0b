// Fused lookup + exchange for table-wise sharded embeddings: the owner of a table pools its rows
// for the GLOBAL batch and stores every pooled row straight into the interaction input buffer of
// the rank that owns the sample, through NVLink peer mappings (cudaIpc).  This replaces
// "lookup -> staging buffer -> NCCL all-to-all -> unpack" (dlrm_jl_b200/sharded.py) by one kernel:
// the transfer is the kernel's own store stream, posted row by row (whole 128-byte lines), so it
// overlaps the gathers that are still in flight.
//
// New functionality (DLRM.jl is single-process); BASELINE.json north_star / SURVEY.md section 8(e).
#include "common.cuh"
#include "sort_small.cuh"

struct dlrmb_xbuf {
    int device;
    void* ptr;
    size_t bytes;
};

namespace dlrmb {

constexpr int kMaxPeers = 16;
struct PeerPtrs {
    float* p[kMaxPeers];
};

// Same mapping as lookup_gather_kernel (one thread per 16-byte chunk, 4 independent index -> row
// chains), but sample b lives on rank b / B_local and table k lands in slot slotmap[k].
template <int VEC, typename RowT>
__device__ __forceinline__ typename std::conditional<VEC == 4, float4, float>::type
p2p_ldrow(const float* base, size_t r, size_t D, int c) {
    if constexpr (VEC == 4) return RowIO<RowT>::ldg4(RowIO<RowT>::row(base, r, D), c);
    else return RowIO<RowT>::ldg1(RowIO<RowT>::row(base, r, D), c);
}

// SORT_ITEMS > 0: the first `ntab` CTAs of the (then one-dimensional) grid sort the tables' index lists for
// the coming sparse update in shared memory (sort_small.cuh), as in lookup_sort_kernel of lookup.cu.
template <typename IdxT, int VEC, int U, typename RowT, int SORT_ITEMS>
__global__ void __launch_bounds__(256, (SORT_ITEMS > 0) ? 4 : 1)
lookup_p2p_kernel(const TableDesc* __restrict__ desc, const int32_t* __restrict__ slotmap,
                  const IdxT* __restrict__ idx, int idx_base, uint32_t Bg, uint32_t B_local, uint32_t P,
                  uint32_t C, PeerPtrs peers, int slots, int ntab, uint32_t bx, uint32_t* __restrict__ keys_out,
                  uint32_t* __restrict__ pos_out, int64_t cap, unsigned long long* clk) {
    using V = typename std::conditional<VEC == 4, float4, float>::type;
    int k;
    uint32_t cx, nctas;
    const unsigned clk_cta = blockIdx.y * gridDim.x + blockIdx.x;
    clock_in(clk, clk_cta);
    if constexpr (SORT_ITEMS > 0) {
        extern __shared__ uint32_t sort_smem[];
        if ((int)blockIdx.x < ntab) {
            k = blockIdx.x;
            const int L = (int)(Bg * P);
            sort_small_body<IdxT, (SORT_ITEMS > 0 ? SORT_ITEMS : 4), 256>(idx + (size_t)k * L, idx_base, L, desc[k].rows,
                                                                          keys_out + (size_t)k * cap, pos_out + (size_t)k * cap, sort_smem);
            clock_out(clk, clk_cta);
            return;
        }
        const uint32_t lin = blockIdx.x - ntab;
        k = (int)(lin / bx);
        cx = lin - (uint32_t)k * bx;
        nctas = bx;
    } else {
        k = blockIdx.y;
        cx = blockIdx.x;
        nctas = gridDim.x;
    }
    const float* __restrict__ tb = desc[k].base;
    const IdxT* __restrict__ ik = idx + (size_t)k * Bg * P;
    const uint32_t n = Bg * C;
    const uint32_t step = nctas * blockDim.x;
    const size_t D = (size_t)C * VEC;
    const size_t slot_off = (size_t)slotmap[k] * D;
    const size_t ostride = (size_t)slots * D;

    for (uint32_t t = cx * blockDim.x + threadIdx.x; t < n; t += step * U) {
        uint32_t b[U], c[U];
        bool ok[U];
        V acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t tt = t + (uint32_t)u * step;
            ok[u] = tt < n;
            b[u] = tt / C;
            c[u] = tt - b[u] * C;
        }
        if (P == 1) {
            int64_t row[U];
#pragma unroll
            for (int u = 0; u < U; ++u) row[u] = ok[u] ? (int64_t)__ldg(ik + b[u]) - idx_base : 0;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (ok[u]) acc[u] = p2p_ldrow<VEC, RowT>(tb, (size_t)row[u], D, c[u]);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!ok[u]) continue;
                const IdxT* ip = ik + (size_t)b[u] * P;
                V a = p2p_ldrow<VEC, RowT>(tb, (size_t)((int64_t)__ldg(ip) - idx_base), D, c[u]);
                for (uint32_t p = 1; p < P; ++p) {
                    V v = p2p_ldrow<VEC, RowT>(tb, (size_t)((int64_t)__ldg(ip + p) - idx_base), D, c[u]);
                    if constexpr (VEC == 4) {
                        a = make_float4(__fadd_rn(a.x, v.x), __fadd_rn(a.y, v.y), __fadd_rn(a.z, v.z), __fadd_rn(a.w, v.w));
                    } else {
                        a = __fadd_rn(a, v);
                    }
                }
                acc[u] = a;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const uint32_t h = b[u] / B_local;
            const uint32_t bl = b[u] - h * B_local;
            float* dst = peers.p[h] + (size_t)bl * ostride + slot_off;
            reinterpret_cast<V*>(dst)[c[u]] = acc[u];
        }
    }
    clock_out(clk, clk_cta);
}

int ensure_smem_attr(const void* func, int bytes, unsigned long long* done_mask);

template <typename IdxT, int VEC, typename RowT>
static int launch_p2p_t(dlrmb_tables* t, const IdxT* idx, int idx_base, int Bg, int P, const PeerPtrs& peers,
                        int B_local, int slots, bool with_sort, cudaStream_t s) {
    const uint32_t C = t->D / VEC;
    const int64_t n = (int64_t)Bg * C;
    DLRMB_REQUIRE(n < (1ll << 30), "B * D too large for one lookup launch (%lld chunks)", (long long)n);
    constexpr int U = 4;
    int64_t bx = ceil_div64(n, 256 * U);
    const int64_t cap = (int64_t)t->sm_count * 32;
    if (bx * t->ntab > cap) bx = cap / t->ntab > 0 ? cap / t->ntab : 1;
    const int64_t L = (int64_t)Bg * P;
    if constexpr (VEC == 4) {
        if (with_sort && L <= kFusedSortMax) {   // the sort of the coming update rides in the same launch
            const unsigned grid1 = (unsigned)(t->ntab + bx * t->ntab);
            if (L <= 2048) {
                constexpr size_t smem = SmallSortGeom<8, 256>::smem_bytes();
                static unsigned long long done = 0;
                int rc = ensure_smem_attr((const void*)lookup_p2p_kernel<IdxT, 4, U, RowT, 8>, (int)smem, &done);
                if (rc) return rc;
                lookup_p2p_kernel<IdxT, 4, U, RowT, 8><<<grid1, 256, smem, s>>>(t->d_desc, t->d_slotmap, idx, idx_base, Bg, B_local, P, C,
                                                                             peers, slots, t->ntab, (uint32_t)bx, t->keys[0], t->pos[0], t->cap, clock_slot(CLK_LOOKUP));
            } else {
                constexpr size_t smem = SmallSortGeom<16, 256>::smem_bytes();
                static unsigned long long done = 0;
                int rc = ensure_smem_attr((const void*)lookup_p2p_kernel<IdxT, 4, U, RowT, 16>, (int)smem, &done);
                if (rc) return rc;
                lookup_p2p_kernel<IdxT, 4, U, RowT, 16><<<grid1, 256, smem, s>>>(t->d_desc, t->d_slotmap, idx, idx_base, Bg, B_local, P, C,
                                                                              peers, slots, t->ntab, (uint32_t)bx, t->keys[0], t->pos[0], t->cap, clock_slot(CLK_LOOKUP));
            }
            DLRMB_LAUNCH_CHECK();
            t->sorted_buf = 0;
            t->sorted_valid = true;
            t->sorted_B = Bg;
            t->sorted_P = P;
            return DLRMB_OK;
        }
    }
    dim3 grid((unsigned)bx, (unsigned)t->ntab);
    lookup_p2p_kernel<IdxT, VEC, U, RowT, 0><<<grid, 256, 0, s>>>(t->d_desc, t->d_slotmap, idx, idx_base, Bg, B_local, P, C,
                                                            peers, slots, t->ntab, (uint32_t)bx, nullptr, nullptr, 0, clock_slot(CLK_LOOKUP));
    DLRMB_LAUNCH_CHECK();
    if (with_sort) {   // too many keys for the 256-thread sort CTAs: the stand-alone sort, same stream
        int rc = launch_sort(t, idx, (int)sizeof(IdxT), idx_base, Bg, P, s);
        if (rc) return rc;
        t->sorted_valid = true;
        t->sorted_B = Bg;
        t->sorted_P = P;
    }
    return DLRMB_OK;
}

}  // namespace dlrmb

using namespace dlrmb;

extern "C" {

int32_t dlrmb_xbuf_create(int32_t device, int64_t bytes, dlrmb_xbuf** out) {
    DLRMB_REQUIRE(out != nullptr && bytes > 0, "bad xbuf arguments");
    *out = nullptr;
    DeviceGuard guard(device);
    DLRMB_REQUIRE(guard.ok, "cudaSetDevice(%d) failed", device);
    void* p = nullptr;
    DLRMB_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("cudaMemset of a %lld-byte exchange buffer failed: %s", (long long)bytes, cudaGetErrorString(e));
        return DLRMB_ECUDA;
    }
    *out = new dlrmb_xbuf{device, p, (size_t)bytes};
    return DLRMB_OK;
}

int32_t dlrmb_xbuf_destroy(dlrmb_xbuf* x) {
    if (!x) return DLRMB_OK;
    DeviceGuard guard(x->device);
    cudaFree(x->ptr);
    delete x;
    return DLRMB_OK;
}

int32_t dlrmb_xbuf_ptr(dlrmb_xbuf* x, void** dev) {
    DLRMB_REQUIRE(x && dev, "null argument");
    *dev = x->ptr;
    return DLRMB_OK;
}

int32_t dlrmb_xbuf_ipc_handle(dlrmb_xbuf* x, uint8_t* handle64) {
    DLRMB_REQUIRE(x && handle64, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    cudaIpcMemHandle_t h;
    DLRMB_CUDA(cudaIpcGetMemHandle(&h, x->ptr));
    memcpy(handle64, &h, 64);
    return DLRMB_OK;
}

int32_t dlrmb_xbuf_open(int32_t device, const uint8_t* handle64, void** peer_ptr) {
    DLRMB_REQUIRE(handle64 && peer_ptr, "null argument");
    *peer_ptr = nullptr;
    DeviceGuard guard(device);
    DLRMB_REQUIRE(guard.ok, "cudaSetDevice(%d) failed", device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    DLRMB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *peer_ptr = p;
    return DLRMB_OK;
}

int32_t dlrmb_xbuf_close(int32_t device, void* peer_ptr) {
    if (!peer_ptr) return DLRMB_OK;
    DeviceGuard guard(device);
    DLRMB_REQUIRE(guard.ok, "cudaSetDevice(%d) failed", device);
    DLRMB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return DLRMB_OK;
}

int32_t dlrmb_tables_set_slot_map(dlrmb_tables* t, const int32_t* slots) {
    DLRMB_REQUIRE(t && slots, "null argument");
    DeviceGuard guard(t->device);
    DLRMB_REQUIRE(guard.ok, "cudaSetDevice(%d) failed", t->device);
    int max_slot = -1;
    for (int k = 0; k < t->ntab; ++k) {
        DLRMB_REQUIRE(slots[k] >= 0, "slot_map[%d] = %d is negative", k, slots[k]);
        if (slots[k] > max_slot) max_slot = slots[k];
    }
    if (!t->d_slotmap) DLRMB_CUDA(cudaMalloc((void**)&t->d_slotmap, sizeof(int32_t) * t->ntab));
    DLRMB_CUDA(cudaMemcpy(t->d_slotmap, slots, sizeof(int32_t) * t->ntab, cudaMemcpyHostToDevice));
    t->slotmap_max = max_slot;
    return DLRMB_OK;
}

static int32_t embedding_fwd_p2p_impl(dlrmb_tables* t, const void* idx, int32_t idx_bytes, int32_t idx_base,
                                      int32_t B_global, int32_t P, float* const* peer_T, int32_t world,
                                      int32_t B_local, int32_t slots, bool with_sort, dlrmb_stream stream) {
    DLRMB_REQUIRE(t != nullptr && idx != nullptr && peer_T != nullptr, "null argument");
    DLRMB_REQUIRE(idx_bytes == 4 || idx_bytes == 8, "idx_bytes must be 4 or 8 (got %d)", idx_bytes);
    DLRMB_REQUIRE(idx_base == 0 || idx_base == 1, "idx_base must be 0 or 1 (got %d)", idx_base);
    DLRMB_REQUIRE(P > 0, "P must be positive (got %d)", P);
    DLRMB_REQUIRE(world >= 1 && world <= kMaxPeers, "world must be in 1..%d (got %d)", kMaxPeers, world);
    DLRMB_REQUIRE(B_local > 0 && B_global == B_local * world, "B_global (%d) must equal B_local (%d) * world (%d)",
                  B_global, B_local, world);
    DLRMB_REQUIRE((int64_t)B_global * P <= t->max_lookups, "B*P = %lld exceeds max_lookups = %lld",
                  (long long)B_global * P, (long long)t->max_lookups);
    if (!t->d_slotmap) {
        set_error("dlrmb_embedding_fwd_p2p needs dlrmb_tables_set_slot_map first");
        return DLRMB_ESTATE;
    }
    DLRMB_REQUIRE(slots > t->slotmap_max, "slots = %d does not cover the slot map (largest slot %d)", slots,
                  t->slotmap_max);
    DeviceGuard guard(t->device);
    DLRMB_REQUIRE(guard.ok, "cudaSetDevice(%d) failed", t->device);
    PeerPtrs peers;
    bool aligned = (t->D % 4 == 0);
    for (int r = 0; r < kMaxPeers; ++r) {
        peers.p[r] = r < world ? peer_T[r] : nullptr;
        if (r < world) {
            DLRMB_REQUIRE(peer_T[r] != nullptr, "peer_T[%d] is null", r);
            aligned = aligned && ((reinterpret_cast<uintptr_t>(peer_T[r]) & 15) == 0);
        }
    }
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    const bool bf = t->elem_bytes == 2;
    if (idx_bytes == 4) {
        const uint32_t* ip = (const uint32_t*)idx;
        if (aligned) rc = bf ? launch_p2p_t<uint32_t, 4, __nv_bfloat16>(t, ip, idx_base, B_global, P, peers, B_local, slots, with_sort, s)
                             : launch_p2p_t<uint32_t, 4, float>(t, ip, idx_base, B_global, P, peers, B_local, slots, with_sort, s);
        else rc = bf ? launch_p2p_t<uint32_t, 1, __nv_bfloat16>(t, ip, idx_base, B_global, P, peers, B_local, slots, with_sort, s)
                     : launch_p2p_t<uint32_t, 1, float>(t, ip, idx_base, B_global, P, peers, B_local, slots, with_sort, s);
    } else {
        const int64_t* ip = (const int64_t*)idx;
        if (aligned) rc = bf ? launch_p2p_t<int64_t, 4, __nv_bfloat16>(t, ip, idx_base, B_global, P, peers, B_local, slots, with_sort, s)
                             : launch_p2p_t<int64_t, 4, float>(t, ip, idx_base, B_global, P, peers, B_local, slots, with_sort, s);
        else rc = bf ? launch_p2p_t<int64_t, 1, __nv_bfloat16>(t, ip, idx_base, B_global, P, peers, B_local, slots, with_sort, s)
                     : launch_p2p_t<int64_t, 1, float>(t, ip, idx_base, B_global, P, peers, B_local, slots, with_sort, s);
    }
    return rc;
}

int32_t dlrmb_embedding_fwd_p2p(dlrmb_tables* t, const void* idx, int32_t idx_bytes, int32_t idx_base,
                                int32_t B_global, int32_t P, float* const* peer_T, int32_t world,
                                int32_t B_local, int32_t slots, dlrmb_stream stream) {
    return embedding_fwd_p2p_impl(t, idx, idx_bytes, idx_base, B_global, P, peer_T, world, B_local, slots, false, stream);
}

int32_t dlrmb_embedding_fwd_p2p_sort(dlrmb_tables* t, const void* idx, int32_t idx_bytes, int32_t idx_base,
                                     int32_t B_global, int32_t P, float* const* peer_T, int32_t world,
                                     int32_t B_local, int32_t slots, dlrmb_stream stream) {
    if (t) t->sorted_valid = false;
    return embedding_fwd_p2p_impl(t, idx, idx_bytes, idx_base, B_global, P, peer_T, world, B_local, slots, true, stream);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Stream-ordered cross-GPU ordering without NCCL: flag barrier over the IPC-mapped buffers
// ---------------------------------------------------------------------------------------------
// Every rank owns a flag array [channels][world] of 32-bit epochs inside a dlrmb_xbuf that its peers
// map.  dlrmb_peer_barrier launches ONE tiny kernel that (1) bumps this rank's epoch of the channel,
// (2) after a system-scope fence stores it into slot [channel][rank] of every rank's array over
// NVLink, (3) waits until all `world` slots of its OWN array have reached the epoch.  Placed on the
// stream behind a kernel that stored into peers' buffers and in front of the kernels that consume
// what the peers stored, it is the ordering point of the fused exchanges: kernel completion makes the
// producer's peer stores visible before the flag store is issued, and nobody passes before every
// rank's flag -- hence every rank's data -- has arrived.  The epoch lives in device memory, so the
// call can be captured into a CUDA graph and replayed.  Each rank must run on its own GPU (the wait
// spins on flags that kernels of other processes set).
namespace dlrmb {

struct PeerFlagPtrs {
    uint32_t* p[kMaxPeers];
};

constexpr int kBarrierChannels = 8;
// state: [0 .. kBarrierChannels) epochs, [kBarrierChannels] = number of timed-out waits

__global__ void __launch_bounds__(32)
peer_barrier_kernel(PeerFlagPtrs peers, int world, int rank, int channel, uint32_t* __restrict__ state,
                    unsigned long long timeout_ns) {
    const int lane = threadIdx.x;
    uint32_t epoch = 0;
    if (lane == 0) {
        epoch = state[channel] + 1u;
        state[channel] = epoch;
    }
    epoch = __shfl_sync(0xffffffffu, epoch, 0);
    if (lane < world) {
        // the stream's earlier kernels (whose peer stores this flag publishes) have completed; the fence
        // orders their effects, as this thread observes them, before the flag store at system scope
        __threadfence_system();
        volatile uint32_t* dst = peers.p[lane] + (size_t)channel * kMaxPeers + rank;
        *dst = epoch;
        const volatile uint32_t* mine = peers.p[rank] + (size_t)channel * kMaxPeers + lane;
        unsigned long long t0 = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        // epochs only grow; compare as a signed distance so a wrap after 2^31 steps stays correct
        while ((int32_t)(*mine - epoch) < 0) {
            __nanosleep(40);
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) {       // a peer is gone: do not hang the GPU, leave a trace for the host
                atomicAdd(state + kBarrierChannels, 1u);
                break;
            }
        }
    }
    __syncwarp();
    __threadfence_system();
}

// idx_local [ntab][B_local][P] of this rank's samples -> the owners' index buffers
// [t_owner][B_global][P] (rows rank*B_local ...), one 4- or 8-byte store per index over NVLink
struct IdxDest {
    void* base;        // owner's buffer for this table: [B_global][P]
};

template <typename IdxT>
__global__ void __launch_bounds__(256)
indices_scatter_kernel(const IdxT* __restrict__ idx_local, const IdxDest* __restrict__ dests, int ntab, int B_local,
                       int P, int rank) {
    const int k = blockIdx.y;
    const int n = B_local * P;
    const IdxT* src = idx_local + (size_t)k * n;
    IdxT* dst = static_cast<IdxT*>(dests[k].base) + (size_t)rank * n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace dlrmb

extern "C" {

int32_t dlrmb_peer_barrier(int32_t device, uint32_t* const* peer_flags, int32_t world, int32_t rank, int32_t channel,
                           uint32_t* state, dlrmb_stream stream) {
    DLRMB_REQUIRE(peer_flags && state, "null argument");
    DLRMB_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "bad world / rank (%d, %d)", world, rank);
    DLRMB_REQUIRE(channel >= 0 && channel < kBarrierChannels, "channel must be in 0..%d", kBarrierChannels - 1);
    DeviceGuard guard(device);
    DLRMB_REQUIRE(guard.ok, "cudaSetDevice(%d) failed", device);
    PeerFlagPtrs peers;
    for (int r = 0; r < kMaxPeers; ++r) {
        peers.p[r] = r < world ? peer_flags[r] : nullptr;
        DLRMB_REQUIRE(r >= world || peer_flags[r] != nullptr, "peer_flags[%d] is null", r);
    }
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(peers, world, rank, channel, state, 4000000000ull);
    DLRMB_LAUNCH_CHECK();
    return DLRMB_OK;
}

int64_t dlrmb_peer_barrier_flag_bytes(void) { return (int64_t)kBarrierChannels * kMaxPeers * sizeof(uint32_t); }
int64_t dlrmb_peer_barrier_state_bytes(void) { return (int64_t)(kBarrierChannels + 1) * sizeof(uint32_t); }

int32_t dlrmb_indices_scatter_p2p(int32_t device, const void* idx_local, int32_t idx_bytes, int32_t ntab,
                                  int32_t B_local, int32_t P, const void* const* dests_dev, int32_t rank,
                                  dlrmb_stream stream) {
    DLRMB_REQUIRE(idx_local && dests_dev, "null argument");
    DLRMB_REQUIRE(idx_bytes == 4 || idx_bytes == 8, "idx_bytes must be 4 or 8 (got %d)", idx_bytes);
    DLRMB_REQUIRE(ntab > 0 && ntab <= 65535 && B_local > 0 && P > 0 && rank >= 0, "bad index scatter geometry");
    DeviceGuard guard(device);
    DLRMB_REQUIRE(guard.ok, "cudaSetDevice(%d) failed", device);
    const int n = B_local * P;
    int bx = (n + 255) / 256;
    if (bx > 64) bx = 64;
    dim3 grid((unsigned)bx, (unsigned)ntab);
    const IdxDest* d = reinterpret_cast<const IdxDest*>(dests_dev);
    if (idx_bytes == 4)
        indices_scatter_kernel<uint32_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)idx_local, d, ntab, B_local, P, rank);
    else
        indices_scatter_kernel<int64_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const int64_t*)idx_local, d, ntab, B_local, P, rank);
    DLRMB_LAUNCH_CHECK();
    return DLRMB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// One-shot all-reduce of a SMALL float buffer over the IPC-mapped exchange buffers
// ---------------------------------------------------------------------------------------------
// The data-parallel step ends with the all-reduce of the bottom MLP's gradients (0.7 MB): it sits on the
// critical path (bottom MLP backward -> all-reduce -> dense SGD) and is latency-, not bandwidth-bound.
// Every rank owns a buffer [2][world][n] (two halves, alternating by step so that a fast peer's next push
// cannot overwrite what a slow rank is still summing).  push: every rank stores its n floats into slot
// [half][rank] of EVERY rank's buffer over NVLink; flag barrier; sum: every rank adds the `world` slots of
// its own buffer in rank order -- the same order on every rank, so all ranks hold bit-identical sums.
// (world - 1) * n * 4 bytes leave each GPU: right for buffers up to ~1 MB; large buffers stay with NCCL's
// reduce-scatter + all-gather.
namespace dlrmb {

struct PeerFloatPtrs {
    float* p[kMaxPeers];
};

__global__ void __launch_bounds__(256)
allreduce_push_kernel(PeerFloatPtrs peers, int world, int rank, const float4* __restrict__ data, int n4,
                      const uint32_t* __restrict__ state, int channel) {
    const uint32_t half = state[channel] & 1u;           // epoch BEFORE this step's barrier
    const size_t off = ((size_t)half * world + rank) * n4;
    const int peer = blockIdx.y;
    float4* dst = reinterpret_cast<float4*>(peers.p[peer]) + off;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) dst[i] = data[i];
}

__global__ void __launch_bounds__(256)
allreduce_sum_kernel(const float4* __restrict__ mine, int world, float4* __restrict__ data, int n4,
                     const uint32_t* __restrict__ state, int channel) {
    const uint32_t half = (state[channel] - 1u) & 1u;    // the barrier in between has bumped the epoch
    const float4* src = mine + (size_t)half * world * n4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        float4 acc = src[i];
        for (int r = 1; r < world; ++r) {
            const float4 v = src[(size_t)r * n4 + i];
            acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
        }
        data[i] = acc;
    }
}

}  // namespace dlrmb

extern "C" {

int32_t dlrmb_peer_allreduce_f32(int32_t device, float* const* peer_bufs, uint32_t* const* peer_flags, int32_t world,
                                 int32_t rank, int32_t channel, uint32_t* state, float* data, int64_t n,
                                 dlrmb_stream stream) {
    DLRMB_REQUIRE(peer_bufs && peer_flags && state && data, "null argument");
    DLRMB_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "bad world / rank (%d, %d)", world, rank);
    DLRMB_REQUIRE(channel >= 0 && channel < kBarrierChannels, "channel must be in 0..%d", kBarrierChannels - 1);
    DLRMB_REQUIRE(n > 0 && n % 4 == 0 && n < (1ll << 28), "n must be a positive multiple of 4 (got %lld)", (long long)n);
    DLRMB_REQUIRE((reinterpret_cast<uintptr_t>(data) & 15) == 0, "data must be 16-byte aligned");
    DeviceGuard guard(device);
    DLRMB_REQUIRE(guard.ok, "cudaSetDevice(%d) failed", device);
    PeerFloatPtrs bufs;
    for (int r = 0; r < kMaxPeers; ++r) {
        bufs.p[r] = r < world ? peer_bufs[r] : nullptr;
        DLRMB_REQUIRE(r >= world || peer_bufs[r] != nullptr, "peer_bufs[%d] is null", r);
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int n4 = (int)(n / 4);
    int bx = (n4 + 255) / 256;
    if (bx > 64) bx = 64;
    allreduce_push_kernel<<<dim3((unsigned)bx, (unsigned)world), 256, 0, s>>>(bufs, world, rank, reinterpret_cast<const float4*>(data), n4,
                                                                           state, channel);
    DLRMB_LAUNCH_CHECK();
    int rc = dlrmb_peer_barrier(device, peer_flags, world, rank, channel, state, stream);
    if (rc) return rc;
    int sx = (n4 + 255) / 256;
    if (sx > 592) sx = 592;
    allreduce_sum_kernel<<<(unsigned)sx, 256, 0, s>>>(reinterpret_cast<const float4*>(bufs.p[rank]), world,
                                                     reinterpret_cast<float4*>(data), n4, state, channel);
    DLRMB_LAUNCH_CHECK();
    return DLRMB_OK;
}

}  // extern "C"
