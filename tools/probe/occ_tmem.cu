// Empirical co-residency probe: do 2 CTAs of S bytes of dynamic shared memory share an SM on this part?
#include <cstdio>
#include <cuda_runtime.h>
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) probe(unsigned* ctr, unsigned* ok, unsigned want) {
    extern __shared__ unsigned char sm[];
    sm[threadIdx.x] = 1;
    __shared__ unsigned tslot;
    if (threadIdx.x < 32) {
        unsigned dst = (unsigned)__cvta_generic_to_shared(&tslot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
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
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tslot), "r"(64u) : "memory");
}
template <int THREADS, int MINB>
void run(size_t smem, int carve) {
    auto k = probe<THREADS, MINB>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (carve >= 0) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, THREADS, smem);
    unsigned *ctr, *ok;
    cudaMalloc(&ctr, 8); ok = ctr + 1;
    cudaMemset(ctr, 0, 8);
    k<<<296, THREADS, smem>>>(ctr, ok, 296);
    cudaError_t le = cudaGetLastError();
    cudaError_t e = cudaDeviceSynchronize();
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
    printf("  minb=%d regs=%d static=%zu launch=%s | ", MINB, fa.numRegs, fa.sharedSizeBytes, cudaGetErrorString(le));
    unsigned h[2];
    cudaMemcpy(h, ctr, 8, cudaMemcpyDeviceToHost);
    printf("threads=%d smem=%zu carve=%d: occupancy API=%d, co-resident CTAs that saw all 296: %u (%s)\n", THREADS, smem, carve, occ,
           h[1], cudaGetErrorString(e));
    cudaFree(ctr);
}
int main() {
    for (int carve : {-1, 100}) {
        run<256, 1>(48 * 1024, carve);
        run<256, 2>(48 * 1024, carve);
        run<256, 1>(105136, carve);
        run<256, 2>(105136, carve);
        run<192, 2>(104624, carve);
    }
    return 0;
}
