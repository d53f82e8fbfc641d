// Empirical co-residency probe: do 2 CTAs of S bytes of dynamic shared memory share an SM on this part?
#include <cstdio>
#include <cuda_runtime.h>
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 2) probe(unsigned* ctr, unsigned* ok, unsigned want) {
    extern __shared__ unsigned char sm[];
    sm[threadIdx.x] = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(ctr, 1u);
        long long t0 = clock64();
        bool good = false;
        while (clock64() - t0 < 200000000ll) {
            if (*((volatile unsigned*)ctr) >= want) { good = true; break; }
        }
        if (good) atomicAdd(ok, 1u);
    }
}
template <int THREADS>
void run(size_t smem, int carve) {
    auto k = probe<THREADS>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (carve >= 0) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, THREADS, smem);
    unsigned *ctr, *ok;
    cudaMalloc(&ctr, 8); ok = ctr + 1;
    cudaMemset(ctr, 0, 8);
    k<<<296, THREADS, smem>>>(ctr, ok, 296);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned h[2];
    cudaMemcpy(h, ctr, 8, cudaMemcpyDeviceToHost);
    printf("threads=%d smem=%zu carve=%d: occupancy API=%d, co-resident CTAs that saw all 296: %u (%s)\n", THREADS, smem, carve, occ,
           h[1], cudaGetErrorString(e));
    cudaFree(ctr);
}
int main() {
    for (int carve : {-1, 100}) {
        run<256>(48 * 1024, carve);
        run<256>(80 * 1024, carve);
        run<256>(100 * 1024, carve);
        run<256>(105136, carve);
        run<256>(112 * 1024, carve);
        run<192>(104624, carve);
    }
    return 0;
}
