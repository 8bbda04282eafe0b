// FP32 FMA issue rate on the device: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100+).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate ffma_rate.cu && ./ffma_rate
// Each thread runs CHAINS independent accumulator chains; the kernel is pure register arithmetic, so
// TFLOP/s = 2 * FMAs / time is the pipe's rate at that occupancy.  Used to decide whether the
// interaction kernels should be written with FFMA2.
#include <cuda_runtime.h>
#include <stdio.h>

template <int CHAINS, bool PACKED>
__global__ void __launch_bounds__(256) rate_kernel(float* out, int iters, float a0, float b0) {
    float2 acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = make_float2(threadIdx.x * 1e-3f + c, c * 0.5f);
    float2 a = make_float2(a0, a0 * 0.999f), b = make_float2(b0, b0 * 1.001f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (PACKED) {
                unsigned long long ra, rb, rc;
                asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
                asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
                asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(acc[c].x), "f"(acc[c].y));
                asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rc) : "l"(ra), "l"(rb));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[c].x), "=f"(acc[c].y) : "l"(rc));
            } else {
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[c].x) : "f"(a.x), "f"(b.x));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[c].y) : "f"(a.y), "f"(b.y));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c].x + acc[c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS, bool PACKED>
static void run(const char* name, int sms, int ctas_per_sm) {
    const int iters = 4096;
    const int grid = sms * ctas_per_sm;
    float* out;
    cudaMalloc(&out, (size_t)grid * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rate_kernel<CHAINS, PACKED><<<grid, 256>>>(out, iters, 0.5f, 0.25f);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) rate_kernel<CHAINS, PACKED><<<grid, 256>>>(out, iters, 0.5f, 0.25f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = 5.0 * grid * 256.0 * iters * CHAINS * 2.0;
    printf("{\"kernel\": \"%s\", \"chains\": %d, \"warps_per_sm\": %d, \"tflops\": %.2f}\n", name, CHAINS,
           ctas_per_sm * 8, 2.0 * fmas / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

// legacy warp-level tensor-core MMA (mma.sync.m16n8k8 TF32), TILES independent accumulator tiles per warp
template <int TILES>
__global__ void __launch_bounds__(256) mma_rate_kernel(float* out, int iters, unsigned a0, unsigned b0) {
    float c[TILES][4];
#pragma unroll
    for (int i = 0; i < TILES; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    unsigned a[4] = {a0, a0 + 8192u, a0 + 16384u, a0 + 24576u};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < TILES; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b0 + 8192u));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < TILES; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int TILES>
static void run_mma(int sms, int ctas_per_sm) {
    const int iters = 4096;
    const int grid = sms * ctas_per_sm;
    float* out;
    cudaMalloc(&out, (size_t)grid * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    mma_rate_kernel<TILES><<<grid, 256>>>(out, iters, 0x3f800000u, 0x3f000000u);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) mma_rate_kernel<TILES><<<grid, 256>>>(out, iters, 0x3f800000u, 0x3f000000u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = 5.0 * grid * 8.0 * iters * TILES;
    printf("{\"kernel\": \"mma.sync.m16n8k8.tf32\", \"tiles\": %d, \"warps_per_sm\": %d, \"tflops\": %.2f, \"mma_per_clk_per_sm\": %.3f}\n",
           TILES, ctas_per_sm * 8, mmas * 2.0 * 16 * 8 * 8 / (ms * 1e-3) / 1e12, mmas / (ms * 1e-3) / sms / 1.965e9);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    run<8, false>("ffma", p.multiProcessorCount, 2);
    run<8, true>("ffma2", p.multiProcessorCount, 2);
    run<8, false>("ffma", p.multiProcessorCount, 4);
    run<8, true>("ffma2", p.multiProcessorCount, 4);
    run<16, false>("ffma", p.multiProcessorCount, 1);
    run<16, true>("ffma2", p.multiProcessorCount, 1);
    run_mma<6>(p.multiProcessorCount, 1);
    run_mma<6>(p.multiProcessorCount, 2);
    run_mma<6>(p.multiProcessorCount, 4);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
