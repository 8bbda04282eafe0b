// Issue rate of tcgen05.mma (kind::tf32, operands in shared memory, accumulator in TMEM) for the tile
// shapes a tcgen05 formulation of the per-sample Gram matrix (27 x 27 x d) would use, next to the
// legacy warp-level MMA rate measured by benchmarks/ffma_rate.cu.  The point: a DLRM sample is a
// 27-row operand, a tcgen05 tile is 64 or 128 rows, so several samples share a tile and only the
// block diagonal of the product is useful.  This program measures what one SM sustains on
//   M = 128, N = 32  (4 samples x 32 padded rows as A, one sample as B:  1/4 of the tile useful)
//   M = 64,  N = 32  (2 samples as A, one sample as B:                   1/2 useful)
//   M = 128, N = 128 / 256  (the shapes the unit is built for)
// and prints the time a 3xTF32 forward at B = 2048 would spend in the tensor pipe per SM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o benchmarks/umma_rate benchmarks/umma_rate.cu
//   ./benchmarks/umma_rate > profiles/r02_umma_rate.json
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: core matrix = 8 rows x 16 bytes, stored contiguously (128 B); the two 16-byte
// K chunks of one K = 8 (tf32) step are LBO bytes apart, consecutive 8-row groups SBO bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    return d;                           // layout type 0 = no swizzle, base offset 0
}

__global__ void __launch_bounds__(128)
umma_rate_kernel(int M, int N, int iters, unsigned long long* out_ns, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    // operands: A = M rows x 8 tf32 (two 16-byte chunks per row), B = N rows x 8 tf32; contents irrelevant, finite
    float* f = reinterpret_cast<float*>(smem);
    for (int i = tid; i < (128 + 256) * 8; i += 128) f[i] = 0.001f * (float)(i & 63);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic-proxy smem writes -> async proxy (UMMA reads)
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t d_tmem = tmem_base;
    if (tid == 0) {
        const uint32_t a_addr = smem_u32(smem);
        const uint32_t b_addr = smem_u32(smem + 128 * 32);
        const uint64_t a_desc = make_desc(a_addr, (uint32_t)(M / 8) * 128u, 128u);
        const uint64_t b_desc = make_desc(b_addr, (uint32_t)(N / 8) * 128u, 128u);
        // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (int i = 0; i < iters; ++i) {
            const uint32_t acc = i > 0 ? 1u : 0u;
            asm volatile(
                "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 2000000000ull) break;      // never hang the GPU on a mis-encoded instruction
        }
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        out_ns[blockIdx.x] = done ? (t1 - t0) : 0ull;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (warp == 0) {   // read one accumulator row back so the work is observable, then free TMEM
        uint32_t v0, v1, v2, v3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                     : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(d_tmem));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (sink) sink[blockIdx.x * 32 + (tid & 31)] = __uint_as_float(v0) + __uint_as_float(v1) + __uint_as_float(v2) + __uint_as_float(v3);
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(d_tmem), "r"(256));
    }
}

static double run(int M, int N, int iters, int ctas) {
    unsigned long long* d_ns;
    float* d_sink;
    cudaMalloc(&d_ns, sizeof(unsigned long long) * ctas);
    cudaMalloc(&d_sink, sizeof(float) * 32 * ctas);
    const size_t smem = (128 + 256) * 32 + 1024;
    cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_rate_kernel<<<ctas, 128, smem>>>(M, N, 64, d_ns, d_sink);       // warm-up
    umma_rate_kernel<<<ctas, 128, smem>>>(M, N, iters, d_ns, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        fprintf(stderr, "M=%d N=%d: %s\n", M, N, cudaGetErrorString(e));
        return -1.0;
    }
    unsigned long long* h = new unsigned long long[ctas];
    cudaMemcpy(h, d_ns, sizeof(unsigned long long) * ctas, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int i = 0; i < ctas; ++i) worst = h[i] > worst ? (double)h[i] : worst;
    delete[] h;
    cudaFree(d_ns);
    cudaFree(d_sink);
    return worst / iters;     // ns per MMA instruction on one SM (every SM busy)
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const int iters = 4096;
    const int shapes[5][2] = {{128, 32}, {64, 32}, {128, 64}, {128, 128}, {128, 256}};
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"what\": \"tcgen05.mma cta_group::1 kind::tf32, K = 8 per instruction, A and B in shared memory (K-major, no swizzle), "
           "%d back-to-back instructions per CTA, one CTA per SM, all SMs busy\", \"shapes\": [", prop.name, sms, iters);
    for (int s = 0; s < 5; ++s) {
        const int M = shapes[s][0], N = shapes[s][1];
        const double ns = run(M, N, iters, sms);
        const double flop = 2.0 * M * N * 8;
        printf("%s{\"M\": %d, \"N\": %d, \"ns_per_mma_per_sm\": %.3f, \"tflops_per_sm\": %.3f, \"tflops_gpu\": %.1f}", s ? ", " : "", M, N, ns,
               ns > 0 ? flop / ns * 1e-3 : 0.0, ns > 0 ? flop / ns * 1e-3 * sms : 0.0);
        if (s == 0 && ns > 0) {
            // 3xTF32 Gram of 4 samples (F = 27 padded to 32, d = 128): 4 B-operands x 16 K-steps x 3 products
            const double per4 = 4 * 16 * 3 * ns * 1e-3;
            fprintf(stderr, "M=128,N=32: %.3f ns/MMA -> %.2f us of tensor pipe per 4 samples, %.2f us per SM at B = 2048 (13.8 samples per SM)\n",
                    ns, per4, per4 * 13.84 / 4);
        }
        if (s == 1 && ns > 0) {
            const double per2 = 2 * 16 * 3 * ns * 1e-3;
            fprintf(stderr, "M=64,N=32: %.3f ns/MMA -> %.2f us per 2 samples, %.2f us per SM at B = 2048\n", ns, per2, per2 * 13.84 / 2);
        }
    }
    printf("]}\n");
    return 0;
}
