// UMMA (tcgen05.mma) descriptor helpers shared by the GEMM kernels.
#pragma once
#include "common.cuh"

namespace fsb {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int kATileBytes = kBlockM * kBlockK * 2;

// UMMA shared-memory descriptor for a K-major, 128B-swizzled tile whose rows are 128 B apart and
// whose 8-row groups are 1024 B apart (exactly what TMA SWIZZLE_128B writes for a {64, rows} box).
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address, 16 B units
    d |= static_cast<uint64_t>(1) << 16;                      // LBO (unused for swizzled K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO: 8 rows * 128 B
    d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BN.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(kBlockM >> 4) << 24);
}


}  // namespace fsb
