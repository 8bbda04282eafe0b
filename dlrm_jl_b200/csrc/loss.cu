// Fused sigmoid + binary cross-entropy, forward and backward in one launch.
//
// the reference ends the top MLP with a sigmoid
// (DLRM.jl src/model/model.jl:87-90) and applies bce_loss with a hand-written pullback
// (src/train/train.jl:33-41 and :45-71).  Those are ~15 elementwise launches in a framework; here
// one CTA-parallel kernel reads the B logits and labels once and writes the mean loss and
// d(loss)/d(logit) (the sensitivity 1/B is folded in, as Zygote.sensitivity(l) = 1):
//     x  = 1 / (1 + exp(-z))
//     l  = mean( -y * max(log x, -100) + (y - 1) * max(log(1 - x), -100) )
//     dx = (1/B) * ((1 - y) / (1 - x + eps) - y / (x + eps)),   eps = eps(Float32)
//     dz = dx * x * (1 - x)
// The per-CTA partial sums are combined in a fixed order by the last CTA to finish (a counter, not
// a floating-point atomic), so the loss is bit-reproducible.
#include "common.cuh"

namespace dlrmb {

__global__ void __launch_bounds__(256)
bce_sigmoid_kernel(const float* __restrict__ z, const float* __restrict__ y, int B, float* __restrict__ prob,
                   float* __restrict__ dz, float* __restrict__ loss, float* __restrict__ block_sums,
                   unsigned int* __restrict__ counter, unsigned long long* clk) {
    __shared__ float wsum[8];
    clock_in(clk, blockIdx.x);
    __shared__ bool is_last;
    const float eps = 1.1920929e-07f;
    const float inv_b = 1.0f / (float)B;
    float local = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) {
        const float zi = z[i], yi = y[i];
        const float x = 1.0f / (1.0f + expf(-zi));
        const float lx = fmaxf(logf(x), -100.0f);
        const float l1x = fmaxf(logf(1.0f - x), -100.0f);
        local += -yi * lx + (yi - 1.0f) * l1x;
        const float dx = inv_b * ((1.0f - yi) / (1.0f - x + eps) - yi / (x + eps));
        if (prob) prob[i] = x;
        dz[i] = dx * x * (1.0f - x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += wsum[w];
        block_sums[blockIdx.x] = s;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        float s = 0.f;
        for (unsigned b = 0; b < gridDim.x; ++b) s += ((volatile float*)block_sums)[b];   // fixed order
        *loss = s * inv_b;
        *counter = 0;
    }
    clock_out(clk, blockIdx.x);
}

int launch_bce_sigmoid(const float* z, const float* y, int B, float* prob, float* dz, float* loss,
                       float* scratch /* >= 65 floats */, cudaStream_t s) {
    int blocks = (B + 255) / 256;
    if (blocks > 64) blocks = 64;
    bce_sigmoid_kernel<<<blocks, 256, 0, s>>>(z, y, B, prob, dz, loss, scratch, reinterpret_cast<unsigned int*>(scratch + 64),
                                              clock_slot(CLK_BCE));
    DLRMB_LAUNCH_CHECK();
    return DLRMB_OK;
}

}  // namespace dlrmb
