// Resolution of %globaltimer on this GPU (is it fine enough to stamp 10-20 us kernels from the device?)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned long long* out, int n) {
    unsigned long long prev;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prev));
    int i = 0;
    while (i < n) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t != prev) { out[i++] = t - prev; prev = t; }
    }
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 64 * 8);
    k<<<1, 1>>>(d, 64);
    unsigned long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    unsigned long long mn = ~0ull, mx = 0; for (int i = 0; i < 64; ++i) { if (h[i] < mn) mn = h[i]; if (h[i] > mx) mx = h[i]; }
    printf("{\"globaltimer_step_ns_min\": %llu, \"globaltimer_step_ns_max\": %llu}\n", mn, mx);
    return 0;
}
