// Probe: does cp.async.bulk.prefetch.tensor.L2 make a later TMA stream of the same tiles faster than HBM?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../fish_speech_b200/csrc/gemm_tc.cuh"
using namespace fsb;

// mode 0: TMA loads into a ring (the consumer); mode 1: L2 prefetch only (TMA), mode 2: L2 prefetch with prefetch.global.L2
__global__ void __launch_bounds__(64, 2) k(const __grid_constant__ CUtensorMap tm, const char* base, int tiles, int kblocks, int stages,
                                           int mode, unsigned long long hint, unsigned* sink) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t b0 = (raw + 1023u) & ~1023u;
    const uint32_t full0 = b0 + stages * 16384u, empty0 = full0 + 8u * stages;
    const long long U = static_cast<long long>(tiles) * kblocks;
    const long long u0 = U * blockIdx.x / gridDim.x, u1 = U * (blockIdx.x + 1) / gridDim.x;
    if (mode == 1) {
        if (threadIdx.x == 0)
            for (long long u = u0; u < u1; ++u) tma_prefetch_l2_3d(&tm, static_cast<int>(u % kblocks) * 64, static_cast<int>(u / kblocks) * 128, 0);
        return;
    }
    if (mode == 2) {
        // each (tile, kblock) = 128 rows x 128 B; row pitch = kblocks*128 B
        for (long long u = u0; u < u1; ++u) {
            const int t = static_cast<int>(u / kblocks), kb = static_cast<int>(u % kblocks);
            for (int r = threadIdx.x; r < 128; r += blockDim.x) {
                const char* p = base + (static_cast<size_t>(t) * 128 + r) * (static_cast<size_t>(kblocks) * 128) + kb * 128;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            }
        }
        return;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(full0 + 8u * s, 1); mbar_init(empty0 + 8u * s, 1); }
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int it = 0;
        for (long long u = u0; u < u1; ++u, ++it) {
            const int s = it % stages; const uint32_t ph = (it / stages) & 1u;
            mbar_wait(empty0 + 8u * s, ph ^ 1u);
            mbar_expect_tx(full0 + 8u * s, 16384);
            tma_load_3d(b0 + s * 16384u, &tm, full0 + 8u * s, static_cast<int>(u % kblocks) * 64, static_cast<int>(u / kblocks) * 128, 0, hint);
        }
    } else if (threadIdx.x == 32) {
        int it = 0; unsigned acc = 0;
        for (long long u = u0; u < u1; ++u, ++it) {
            const int s = it % stages; const uint32_t ph = (it / stages) & 1u;
            mbar_wait(full0 + 8u * s, ph);
            acc += *reinterpret_cast<volatile unsigned*>(smem_raw + (b0 - raw) + s * 16384);
            mbar_arrive(empty0 + 8u * s);
        }
        if (acc == 0x12345678u) *sink = acc;
    }
}

int main() {
    const int K = 2560, kblocks = K / 64;
    __nv_bfloat16* w;
    const size_t cap = 200ull << 20;
    cudaMalloc(&w, cap * 4);
    cudaMemset(w, 1, cap * 4);
    unsigned* sink; cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int stages = 5;
    const size_t smem = 1024 + stages * 16384 + 16 * stages + 64;
    for (int mb : {25, 50, 100}) {
        const int N = (mb << 20) / (K * 2) / 128 * 128, tiles = N / 128;
        const double bytes = double(N) * K * 2;
        for (int mode : {0, 1, 2}) {
            for (unsigned long long hint : {kEvictNormal, kEvictFirst}) {
                float tsum = 0, psum = 0;
                const int reps = 6;
                for (int r = 0; r < reps; ++r) {
                    // rotate over 4 distinct matrices so that nothing is in L2 by accident
                    __nv_bfloat16* base = w + (cap / 2) * (r % 4);
                    GemmOperand A{base, K, N, 1, K, (long long)N * K};
                    CUtensorMap tm; gemm_make_tmap(&tm, A, 128);
                    float ms = 0;
                    if (mode != 0) {
                        cudaEventRecord(e0);
                        k<<<296, 64, smem>>>(tm, reinterpret_cast<const char*>(base), tiles, kblocks, stages, mode, hint, sink);
                        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); psum += ms;
                    }
                    cudaEventRecord(e0);
                    k<<<296, 64, smem>>>(tm, reinterpret_cast<const char*>(base), tiles, kblocks, stages, 0, hint, sink);
                    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); tsum += ms;
                }
                printf("%3d MB  prefetch=%s hint=%s: prefetch kernel %.1f us, stream %.1f us = %.0f GB/s\n", mb,
                       mode == 0 ? "none" : (mode == 1 ? "tma.L2" : "prefetch.global.L2"), hint == kEvictFirst ? "evict_first" : "normal",
                       psum * 1e3 / reps, tsum * 1e3 / reps, bytes * reps / (tsum * 1e-3) / 1e9);
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
