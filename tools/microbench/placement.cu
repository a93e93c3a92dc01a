// Developer microbenchmark: which SM does cluster c of a one-wave launch land on?  Same launch shape as the
// C2 launch of ctc_lin_kernel (2-CTA clusters, 128 threads, 56704 bytes of dynamic shared memory, 128
// registers worth of occupancy is emulated by the shared-memory footprint alone).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o placement placement.cu ; run: ./placement 512
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 4) probe(int* smid, long long* t0, int spin) {
    extern __shared__ unsigned char sm[];
    unsigned id;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
    if (threadIdx.x == 0) { smid[blockIdx.x] = (int)id; t0[blockIdx.x] = clock64(); }
    // stay resident long enough for the whole grid to be placed
    long long t = clock64();
    while (clock64() - t < spin) { sm[threadIdx.x] = (unsigned char)t; }
}
int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 512;
    int* d; long long* dt;
    cudaMalloc(&d, n * sizeof(int)); cudaMalloc(&dt, n * sizeof(long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 56704);
    for (int rep = 0; rep < 2; ++rep) {
        probe<<<n, 128, 56704>>>(d, dt, 200000);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    }
    std::vector<int> h(n); cudaMemcpy(h.data(), d, n * sizeof(int), cudaMemcpyDeviceToHost);
    printf("cluster -> smid of its two CTAs\n");
    for (int c = 0; c < n / 2; ++c) printf("%d:%d,%d%s", c, h[2 * c], h[2 * c + 1], (c % 8 == 7) ? "\n" : "  ");
    printf("\n");
    std::vector<int> cnt(256, 0);
    for (int i = 0; i < n; ++i) cnt[h[i]]++;
    printf("CTAs per SM: ");
    for (int s = 0; s < 160; ++s) if (cnt[s]) printf("%d:%d ", s, cnt[s]);
    printf("\n");
    return 0;
}
