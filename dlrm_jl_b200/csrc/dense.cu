// Dense-layer backward glue: activation pullback fused with the bias gradient.
//
// The reference's MLP layers are OneDNN.Dense(Flux.Dense(in, out, relu)) (DLRM.jl
// src/model/model.jl:72-93): one primitive computes relu(W x + b), and its pullback hands back
// dW, db and dx.  The GEMMs stay library calls here (cuBLAS, fp32); what a framework adds around
// them in the backward pass -- the relu mask (one elementwise launch), the bias gradient (a column
// reduction of dY, ~13 us per layer as a generic reduce kernel) and the accumulation of both into
// gradient buffers -- is one launch of this kernel per layer:
//     dZ[b][n] = dY[b][n] * (Y[b][n] > 0)          (Y = the layer's output; no mask when Y is NULL)
//     db[n]    = sum_b dZ[b][n]
// dY is read once, dZ written once (in place when dZ == dY).  A CTA owns 32 columns x one block of
// rows; its column sums go to scratch and the last CTA of a column block to finish (a counter, not
// a floating-point atomic) adds the row blocks' partials in ascending order, so db is
// bit-reproducible.
#include "common.cuh"

namespace dlrmb {

constexpr int kDenseRowBlocks = 16;

__global__ void __launch_bounds__(256)
dense_bwd_act_bias_kernel(const float* __restrict__ dy, const float* __restrict__ y, int B, int N,
                          float* __restrict__ dz, float* __restrict__ db, float* __restrict__ partial,
                          unsigned int* __restrict__ counters, int rows_per_block) {
    __shared__ float red[8][32];
    __shared__ bool is_last;
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + c;
    const int row0 = blockIdx.y * rows_per_block;
    const int row1 = min(B, row0 + rows_per_block);
    float acc = 0.f;
    if (col < N) {
        for (int b = row0 + r; b < row1; b += 8) {
            const size_t i = (size_t)b * N + col;
            float g = dy[i];
            if (y != nullptr && !(y[i] > 0.f)) g = 0.f;
            if (dz != dy || y != nullptr) dz[i] = g;
            acc += g;
        }
    }
    red[r][c] = acc;
    __syncthreads();
    if (r == 0) {
        float s = red[0][c];
#pragma unroll
        for (int q = 1; q < 8; ++q) s += red[q][c];
        if (col < N) partial[(size_t)blockIdx.y * N + col] = s;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        is_last = (atomicAdd(counters + blockIdx.x, 1u) == gridDim.y - 1);
        __threadfence();
    }
    __syncthreads();
    if (is_last && r == 0 && col < N) {
        float s = 0.f;
        for (unsigned q = 0; q < gridDim.y; ++q) s += __ldcg(partial + (size_t)q * N + col);   // fixed order
        db[col] = s;
    }
    if (is_last && threadIdx.x == 0) counters[blockIdx.x] = 0;   // re-armed for the next call
}

// Forward epilogue: z[b][n] = act(z[b][n] + bias[n]) in place, act = relu or identity.  (cuBLASLt's
// fp32 SIMT GEMMs run their bias / relu epilogue as a separate 8 us kernel per layer; this one moves
// the same bytes in about half the time and leaves the GEMM a plain C = A * B.)
template <bool VEC4>
__global__ void __launch_bounds__(256)
dense_fwd_bias_act_kernel(float* __restrict__ z, const float* __restrict__ bias, long long total, int N, int relu) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (VEC4) {
        const int n4 = N >> 2;
        float4* z4 = reinterpret_cast<float4*>(z);
        const float4* b4 = reinterpret_cast<const float4*>(bias);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (total >> 2); i += stride) {
            float4 v = z4[i];
            const float4 bb = __ldg(b4 + (int)(i % n4));
            v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            z4[i] = v;
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
            float v = z[i] + __ldg(bias + (int)(i % N));
            z[i] = relu ? fmaxf(v, 0.f) : v;
        }
    }
}

int launch_dense_fwd_bias_act(float* z, const float* bias, int B, int N, int relu, int sm_count, cudaStream_t s) {
    const long long total = (long long)B * N;
    const bool vec4 = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(z) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(bias) & 15) == 0);
    const long long work = vec4 ? total / 4 : total;
    long long blocks = (work + 255) / 256;
    if (blocks > (long long)sm_count * 8) blocks = (long long)sm_count * 8;
    if (blocks < 1) blocks = 1;
    if (vec4)
        dense_fwd_bias_act_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(z, bias, total, N, relu);
    else
        dense_fwd_bias_act_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(z, bias, total, N, relu);
    DLRMB_LAUNCH_CHECK();
    return DLRMB_OK;
}

int64_t dense_bwd_scratch_floats(int N) { return (int64_t)kDenseRowBlocks * N + (N + 31) / 32 + 32; }

int launch_dense_bwd_act_bias(const float* dy, const float* y, int B, int N, float* dz, float* db,
                              float* scratch, cudaStream_t s) {
    int rb = kDenseRowBlocks;
    while (rb > 1 && (B + rb - 1) / rb < 8) rb /= 2;
    const int rows_per_block = (B + rb - 1) / rb;
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)((B + rows_per_block - 1) / rows_per_block));
    unsigned int* counters = reinterpret_cast<unsigned int*>(scratch + (size_t)kDenseRowBlocks * N);
    dense_bwd_act_bias_kernel<<<grid, 256, 0, s>>>(dy, y, B, N, dz, db, scratch, counters, rows_per_block);
    DLRMB_LAUNCH_CHECK();
    return DLRMB_OK;
}

}  // namespace dlrmb
