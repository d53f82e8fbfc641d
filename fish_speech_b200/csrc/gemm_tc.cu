// tcgen05 / TMEM / TMA multi-tap GEMM — see gemm_tc.cuh for the contract.
//
// CTA = 192 threads: warps 0-3 epilogue (TMEM lane quadrant = warp id), warp 4 TMA producer,
// warp 5 TMEM allocator + single-thread MMA issuer.  smem ring of `stages` x {A 128x64 bf16 (16 KB),
// B BNx64 bf16}, both written by TMA with SWIZZLE_128B and consumed through UMMA shared-memory
// descriptors (K-major, SBO = 1024 B).  A CTA runs one output tile, or a stream-K range of segments
// with two TMEM accumulators so the epilogue of segment n overlaps the MMAs of segment n+1.
#include "gemm_tc.cuh"

#include <vector>

namespace fsb {

int gemm_init();

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle atom row
constexpr int kThreads = 192;
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

__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}

// Direct epilogue for one accumulator: bias -> GELU -> gamma -> +residual -> tanh -> store raw (out0)
// and/or snake-activated (out1) in bf16 (or fp32).
template <int BN>
__device__ __forceinline__ void epilogue_direct(const GemmParams& p, uint32_t taddr, int i, bool i_ok,
                                                int j0, int z) {
    const size_t zoff = static_cast<size_t>(z) * p.o_zs;
    const bool vec = (p.o_js == 1) && !p.out_f32 && ((p.o_is & 7) == 0) && ((p.o_zs & 7) == 0);
    float bi = 0.f, gi = 1.f, sa = 0.f, sia = 0.f;
    if (p.chan_on_i && i_ok) {
        if (p.bias) bi = p.bias[i];
        if (p.gamma) gi = p.gamma[i];
        if (p.out1) {
            sa = p.snake_alpha[i];
            sia = p.snake_inv_alpha[i];
        }
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        if (!i_ok) continue;
        const int jbase = j0 + c0;
        if (jbase >= p.rows_j) continue;
        const size_t rowoff = zoff + static_cast<size_t>(i) * p.o_is;
        float v[32], w[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
            const int j = jbase + jj;
            float x = __uint_as_float(r[jj]);
            float b = bi, g = gi, a = sa, ia = sia;
            if (!p.chan_on_i && j < p.rows_j) {
                if (p.bias) b = p.bias[j];
                if (p.gamma) g = p.gamma[j];
                if (p.out1) {
                    a = p.snake_alpha[j];
                    ia = p.snake_inv_alpha[j];
                }
            }
            x += b;
            if (p.act == ACT_GELU) x = gelu_erf(x);
            x *= g;
            v[jj] = x;
            w[jj] = a;  // stash alpha / inv_alpha for the snake pass
            r[jj] = __float_as_uint(ia);
        }
        const bool full32 = (jbase + 32 <= p.rows_j);
        if (p.resid) {
            if (vec && full32) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.resid + rowoff + jbase);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 u = rp[q];
                    v[q * 8 + 0] += bf_lo(u.x); v[q * 8 + 1] += bf_hi(u.x);
                    v[q * 8 + 2] += bf_lo(u.y); v[q * 8 + 3] += bf_hi(u.y);
                    v[q * 8 + 4] += bf_lo(u.z); v[q * 8 + 5] += bf_hi(u.z);
                    v[q * 8 + 6] += bf_lo(u.w); v[q * 8 + 7] += bf_hi(u.w);
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    const int j = jbase + jj;
                    if (j < p.rows_j)
                        v[jj] += bf2f(p.resid[rowoff + static_cast<size_t>(j) * p.o_js]);
                }
            }
        }
        if (p.act == ACT_TANH) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] = tanhf(v[jj]);
        }
        if (p.out1) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
                const float s = sinf(w[jj] * v[jj]);
                w[jj] = v[jj] + __uint_as_float(r[jj]) * s * s;
            }
        }
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            void* outp = o == 0 ? p.out0 : p.out1;
            if (outp == nullptr) continue;
            const float* src = o == 0 ? v : w;
            if (vec && full32) {
                uint4* op =
                    reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(outp) + rowoff + jbase);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 u;
                    u.x = pack_bf2(src[q * 8 + 0], src[q * 8 + 1]);
                    u.y = pack_bf2(src[q * 8 + 2], src[q * 8 + 3]);
                    u.z = pack_bf2(src[q * 8 + 4], src[q * 8 + 5]);
                    u.w = pack_bf2(src[q * 8 + 6], src[q * 8 + 7]);
                    op[q] = u;
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    const int j = jbase + jj;
                    if (j < p.rows_j) {
                        const size_t idx = rowoff + static_cast<size_t>(j) * p.o_js;
                        if (p.out_f32)
                            reinterpret_cast<float*>(outp)[idx] = src[jj];
                        else
                            reinterpret_cast<__nv_bfloat16*>(outp)[idx] = f2bf(src[jj]);
                    }
                }
            }
        }
    }
}

// MODE 0: fp32 partial sums to the workspace (decode / prefill of the LM). Lean in registers and shared
//         memory so that TWO CTAs fit on an SM: the CTA of the next GEMM in the stream becomes resident
//         (programmatic dependent launch) and prefetches its weight tiles while this one still runs.
// MODE 1: direct fused epilogue (codec convolutions / linears).
template <int BN, int MODE>
__global__ void __launch_bounds__(kThreads, MODE == 0 ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int kBTileBytes = BN * kBlockK * 2;
    constexpr int kStageBytes = kATileBytes + kBTileBytes;
    constexpr int kTmemCols = 2 * BN;  // two accumulators: epilogue of item n overlaps MMA of n+1

    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
    const int stages = p.stages;
    const uint32_t bars = tiles + static_cast<uint32_t>(stages) * kStageBytes;
    // barrier block: full[stages], empty[stages], tmem_full[2], tmem_empty[2], TMEM base word
    const uint32_t full0 = bars, empty0 = bars + 8u * stages, tfull0 = bars + 16u * stages;
    const uint32_t tempty0 = tfull0 + 16u;
    const uint32_t tmem_slot = tempty0 + 16u;
    uint32_t* tmem_slot_ptr =
        reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long* trace =
        p.trace ? p.trace + (static_cast<size_t>(p.trace_id) * gridDim.x + blockIdx.x) * 3 : nullptr;
    if (trace && threadIdx.x == 0) trace[0] = globaltimer_ns();

    // Work items of this CTA: a stream-K range (several segments, each of one output tile), a range of
    // the persistent tiled schedule, or the single tile named by blockIdx.
    int item_begin = 0, item_end = 1;
    if (p.sched != nullptr) {
        item_begin = p.cta_items[blockIdx.x];
        item_end = p.cta_items[blockIdx.x + 1];
    } else if (p.tiled_total > 0) {
        item_begin = static_cast<int>(static_cast<long long>(p.tiled_total) * blockIdx.x / gridDim.x);
        item_end = static_cast<int>(static_cast<long long>(p.tiled_total) * (blockIdx.x + 1) / gridDim.x);
    }
    int z = blockIdx.z;
    auto get_item = [&](int n, int& i0, int& j0, int& kb0, int& kb1, int& slot) {
        if (p.sched != nullptr) {
            const int4 w = p.sched[n];
            i0 = (w.x & 0xffff) * kBlockM;
            j0 = ((w.x >> 16) & 0xffff) * BN;
            kb0 = w.y;
            kb1 = w.z;
            slot = w.w;
        } else if (p.tiled_total > 0) {
            const int per_z = p.tiled_ti * p.tiled_tj;
            z = n / per_z;
            const int rem = n - z * per_z;
            const int ti = rem / p.tiled_tj;
            i0 = ti * kBlockM;
            j0 = (rem - ti * p.tiled_tj) * BN;
            kb0 = 0;
            kb1 = p.kb_per_tap * p.num_taps;
            slot = 0;
        } else {
            i0 = blockIdx.x * kBlockM;
            j0 = blockIdx.y * BN;
            kb0 = 0;
            kb1 = p.kb_per_tap * p.num_taps;
            slot = 0;
        }
    };

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < stages; ++s) {
            mbar_init(full0 + 8u * s, 1);
            mbar_init(empty0 + 8u * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8u * a, 1);
            mbar_init(tempty0 + 8u * a, 4);  // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    pdl_launch_dependents();  // let the consumer kernel get resident; it blocks in griddepcontrol.wait
    if (warp == 4) {
        // ===== TMA producer: the ring keeps flowing across item boundaries =====
        // Operand A (weights when a_static) does not depend on the upstream kernel: its tiles for the
        // whole ring are requested BEFORE griddepcontrol.wait, so the HBM stream of this GEMM starts
        // while the previous kernel (a small consumer kernel) is still running. Operand B (the
        // activations that kernel produces) is only fetched after the wait.
        if (lane == 0) {
            int it = 0;
            int pre = 0;  // ring slots whose A tile was requested early
            bool waited = false;
            if (p.a_static) {
                int n = item_begin, kb = 0, i0 = 0, j0 = 0, kb0 = 0, kb1 = 0, slot = 0;
                if (n < item_end) {
                    get_item(n, i0, j0, kb0, kb1, slot);
                    kb = kb0;
                }
                while (n < item_end && pre < stages) {
                    const int s = pre;
                    mbar_expect_tx(full0 + 8u * s, kStageBytes);
                    const int tap = kb / p.kb_per_tap;
                    const int kc = (kb - tap * p.kb_per_tap) * kBlockK;
                    tma_load_3d(tiles + static_cast<uint32_t>(s) * kStageBytes, &tmA, full0 + 8u * s,
                                p.a_tapk * tap + kc, i0 + p.a_shift[tap], p.a_batched ? z : 0, p.a_hint);
                    ++pre;
                    if (++kb >= kb1) {
                        if (++n < item_end) {
                            get_item(n, i0, j0, kb0, kb1, slot);
                            kb = kb0;
                        }
                    }
                }
                // ... and the next `l2_prefetch` k-blocks go to L2: the HBM stream of this GEMM keeps
                // running through the tail of the previous kernel and the small consumer kernel between.
                for (int q = 0; q < p.l2_prefetch && n < item_end; ++q) {
                    const int tap = kb / p.kb_per_tap;
                    const int kc = (kb - tap * p.kb_per_tap) * kBlockK;
                    tma_prefetch_l2_3d(&tmA, p.a_tapk * tap + kc, i0 + p.a_shift[tap], p.a_batched ? z : 0);
                    if (++kb >= kb1) {
                        if (++n < item_end) {
                            get_item(n, i0, j0, kb0, kb1, slot);
                            kb = kb0;
                        }
                    }
                }
            }
            pdl_wait();
            if (trace) trace[1] = globaltimer_ns();
            waited = true;
            (void)waited;
            for (int n = item_begin; n < item_end; ++n) {
                int i0, j0, kb0, kb1, slot;
                get_item(n, i0, j0, kb0, kb1, slot);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = static_cast<uint32_t>(it / stages) & 1u;
                    const int tap = kb / p.kb_per_tap;
                    const int kc = (kb - tap * p.kb_per_tap) * kBlockK;
                    const uint32_t a_dst = tiles + static_cast<uint32_t>(s) * kStageBytes;
                    if (it >= pre) {
                        mbar_wait(empty0 + 8u * s, ph ^ 1u);
                        mbar_expect_tx(full0 + 8u * s, kStageBytes);
                        tma_load_3d(a_dst, &tmA, full0 + 8u * s, p.a_tapk * tap + kc,
                                    i0 + p.a_shift[tap], p.a_batched ? z : 0, p.a_hint);
                    }
                    tma_load_3d(a_dst + kATileBytes, &tmB, full0 + 8u * s, p.b_tapk * tap + kc,
                                j0 + p.b_shift[tap], p.b_batched ? z : 0, p.b_hint);
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            int it = 0;
            for (int n = item_begin; n < item_end; ++n) {
                int i0, j0, kb0, kb1, slot;
                get_item(n, i0, j0, kb0, kb1, slot);
                const int a = (n - item_begin) & 1;
                const uint32_t aph = static_cast<uint32_t>((n - item_begin) >> 1) & 1u;
                mbar_wait(tempty0 + 8u * a, aph ^ 1u);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(a * BN);
                uint32_t acc = 0;
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = static_cast<uint32_t>(it / stages) & 1u;
                    mbar_wait(full0 + 8u * s, ph);
                    tc_fence_after();
                    const uint32_t a_src = tiles + static_cast<uint32_t>(s) * kStageBytes;
                    const uint64_t ad = make_sdesc(a_src), bd = make_sdesc(a_src + kATileBytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // advance 16 elements (32 B) along K inside the swizzle atom: +2 (16 B units)
                        umma_bf16(d_tmem, ad + 2u * k, bd + 2u * k, idesc, acc);
                        acc = 1;
                    }
                    umma_commit(empty0 + 8u * s);  // frees the smem slot once these MMAs retire
                }
                umma_commit(tfull0 + 8u * a);  // accumulator complete
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        for (int n = item_begin; n < item_end; ++n) {
            int i0, j0, kb0, kb1, slot;
            get_item(n, i0, j0, kb0, kb1, slot);
            const int a = (n - item_begin) & 1;
            const uint32_t aph = static_cast<uint32_t>((n - item_begin) >> 1) & 1u;
            mbar_wait(tfull0 + 8u * a, aph);
            tc_fence_after();
            const int i = i0 + warp * 32 + lane;
            const bool i_ok = i < p.rows_i;
            const uint32_t taddr =
                tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(a * BN);
            if (MODE == 0) {
                float* base = p.ws + static_cast<size_t>(slot) * p.ws_slot_stride;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
                    if (i_ok) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) {
                            const int j = j0 + c0 + jj;
                            if (j < p.rows_j)
                                base[static_cast<size_t>(j) * p.ws_ld + i] = __uint_as_float(r[jj]);
                        }
                    }
                }
            } else {
                epilogue_direct<BN>(p, taddr, i, i_ok, j0, z);
            }
            // release the accumulator to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8u * a);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (trace && threadIdx.x == 0) trace[2] = globaltimer_ns();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
        set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s",
                  cudaGetErrorString(e));
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

int make_tmap(CUtensorMap* tm, const GemmOperand& op, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return 1;
    FSB_CHECK((reinterpret_cast<uintptr_t>(op.ptr) & 15) == 0, "TMA operand not 16-byte aligned");
    FSB_CHECK((op.row_stride * 2) % 16 == 0 && (op.batch_stride * 2) % 16 == 0,
              "TMA strides must be multiples of 16 bytes (row_stride=%lld batch_stride=%lld)",
              op.row_stride, op.batch_stride);
    FSB_CHECK(op.k > 0 && op.rows > 0 && op.batch > 0, "empty TMA operand");
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(op.k), static_cast<cuuint64_t>(op.rows),
                          static_cast<cuuint64_t>(op.batch)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(op.row_stride) * 2,
                             static_cast<cuuint64_t>(op.batch_stride > 0 ? op.batch_stride
                                                                         : op.row_stride * op.rows) *
                                 2};
    cuuint32_t box[3] = {kBlockK, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<__nv_bfloat16*>(op.ptr),
                    dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FSB_CHECK(r == CUDA_SUCCESS,
              "cuTensorMapEncodeTiled failed (%d): k=%lld rows=%lld batch=%lld rs=%lld bs=%lld",
              static_cast<int>(r), op.k, op.rows, op.batch, op.row_stride, op.batch_stride);
    return 0;
}

template <int BN, int MODE>
int launch_bn_mode(const GemmPlan& plan, cudaStream_t stream) {
    auto k = gemm_tc_kernel<BN, MODE>;
    FSB_LAUNCH(k, plan.grid, dim3(kThreads), plan.smem, stream, plan.tmA, plan.tmB, plan.p);
    return 0;
}

template <int BN>
int launch_bn(const GemmPlan& plan, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        FSB_TRY(gemm_init());
        attr_set = true;
    }
    return plan.p.mode == 0 ? launch_bn_mode<BN, 0>(plan, stream) : launch_bn_mode<BN, 1>(plan, stream);
}

}  // namespace

int gemm_plan_init(GemmPlan* plan, const GemmOperand& A, const GemmOperand& B, int bn, int stages,
                   int tiles_i, int tiles_j, int batch) {
    FSB_CHECK(bn == 32 || bn == 64 || bn == 128 || bn == 256, "unsupported BN %d", bn);
    FSB_TRY(make_tmap(&plan->tmA, A, kBlockM));
    FSB_TRY(make_tmap(&plan->tmB, B, bn));
    const int stage_bytes = kATileBytes + bn * kBlockK * 2;
    const int max_stages = (227 * 1024 - 1024 - 256) / stage_bytes;
    if (stages > max_stages) stages = max_stages;
    if (stages > kMaxStages) stages = kMaxStages;
    FSB_CHECK(stages >= 2, "not enough shared memory for 2 stages");
    plan->bn = bn;
    plan->p.stages = stages;
    plan->smem = static_cast<size_t>(stages) * stage_bytes + 1024 + 16 * stages + 64;
    plan->grid = dim3(tiles_i, tiles_j, batch);
    plan->p.sched = nullptr;
    plan->p.cta_items = nullptr;
    plan->cta_items_dev = nullptr;
    plan->sched_dev = nullptr;
    plan->nparts_dev = nullptr;
    plan->max_parts = 1;
    if (plan->p.a_hint == 0) plan->p.a_hint = kEvictNormal;
    if (plan->p.b_hint == 0) plan->p.b_hint = kEvictNormal;
    return 0;
}

int gemm_plan_streamk(GemmPlan* plan, int tiles_i, int kblocks, int num_ctas) {
    // Units are (tile, k-block) pairs in tile-major order; CTA c streams units [c*U/n, (c+1)*U/n).
    // Where that range crosses a tile boundary it becomes several items (one per tile touched); the
    // CTA runs them back to back with the TMA ring never draining, so every SM pulls the same number
    // of weight bytes. Partials of tile t land in workspace slots [0, nparts[t]).
    const long long U = static_cast<long long>(tiles_i) * kblocks;
    if (num_ctas > U) num_ctas = static_cast<int>(U);
    std::vector<int4> items;
    std::vector<int> cta_items(num_ctas + 1, 0);
    std::vector<int> nparts(tiles_i, 0);
    for (int c = 0; c < num_ctas; ++c) {
        long long u0 = U * c / num_ctas, u1 = U * (c + 1) / num_ctas;
        cta_items[c] = static_cast<int>(items.size());
        while (u0 < u1) {
            const int t = static_cast<int>(u0 / kblocks);
            const long long tend = static_cast<long long>(t + 1) * kblocks;
            const long long e = u1 < tend ? u1 : tend;
            int4 w;
            w.x = t;  // tile_j = 0
            w.y = static_cast<int>(u0 - static_cast<long long>(t) * kblocks);
            w.z = static_cast<int>(e - static_cast<long long>(t) * kblocks);
            w.w = nparts[t]++;
            items.push_back(w);
            u0 = e;
        }
    }
    cta_items[num_ctas] = static_cast<int>(items.size());
    int maxp = 0;
    for (int t = 0; t < tiles_i; ++t) maxp = nparts[t] > maxp ? nparts[t] : maxp;
    plan->max_parts = maxp;
    FSB_CUDA(cudaMalloc(&plan->sched_dev, items.size() * sizeof(int4)));
    FSB_CUDA(cudaMemcpy(plan->sched_dev, items.data(), items.size() * sizeof(int4),
                        cudaMemcpyHostToDevice));
    FSB_CUDA(cudaMalloc(&plan->cta_items_dev, cta_items.size() * sizeof(int)));
    FSB_CUDA(cudaMemcpy(plan->cta_items_dev, cta_items.data(), cta_items.size() * sizeof(int),
                        cudaMemcpyHostToDevice));
    FSB_CUDA(cudaMalloc(&plan->nparts_dev, tiles_i * sizeof(int)));
    FSB_CUDA(cudaMemcpy(plan->nparts_dev, nparts.data(), tiles_i * sizeof(int),
                        cudaMemcpyHostToDevice));
    plan->p.sched = reinterpret_cast<const int4*>(plan->sched_dev);
    plan->p.cta_items = plan->cta_items_dev;
    plan->grid = dim3(static_cast<unsigned>(num_ctas), 1, 1);
    return 0;
}

int gemm_plan_tiled(GemmPlan* plan, int tiles_i, int tiles_j, int batch) {
    const long long total = static_cast<long long>(tiles_i) * tiles_j * batch;
    FSB_CHECK(total > 0 && total < (1ll << 31), "gemm_plan_tiled: bad tile count");
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long ctas = total < sms ? total : sms;  // one persistent CTA per SM (smem-limited)
    plan->p.tiled_total = static_cast<int>(total);
    plan->p.tiled_ti = tiles_i;
    plan->p.tiled_tj = tiles_j;
    plan->grid = dim3(static_cast<unsigned>(ctas), 1, 1);
    return 0;
}

void gemm_plan_free(GemmPlan* plan) {
    if (plan->sched_dev) cudaFree(plan->sched_dev);
    if (plan->nparts_dev) cudaFree(plan->nparts_dev);
    if (plan->cta_items_dev) cudaFree(plan->cta_items_dev);
    plan->cta_items_dev = nullptr;
    plan->sched_dev = nullptr;
    plan->nparts_dev = nullptr;
}

int gemm_init() {
    // opt every instantiation into the large dynamic shared-memory carve-out up front, so that no
    // attribute call is needed later (e.g. while a stream is being captured into a CUDA graph)
#define FSB_GEMM_ATTR(BN_, M_)                                                                       \
    FSB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN_, M_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  227 * 1024));
    FSB_GEMM_ATTR(32, 0) FSB_GEMM_ATTR(64, 0) FSB_GEMM_ATTR(128, 0) FSB_GEMM_ATTR(256, 0)
    FSB_GEMM_ATTR(32, 1) FSB_GEMM_ATTR(64, 1) FSB_GEMM_ATTR(128, 1) FSB_GEMM_ATTR(256, 1)
#undef FSB_GEMM_ATTR
    return 0;
}

int gemm_launch(const GemmPlan& plan, cudaStream_t stream) {
    switch (plan.bn) {
        case 32: return launch_bn<32>(plan, stream);
        case 64: return launch_bn<64>(plan, stream);
        case 128: return launch_bn<128>(plan, stream);
        case 256: return launch_bn<256>(plan, stream);
    }
    set_error("gemm_launch: bad BN %d", plan.bn);
    return 1;
}

}  // namespace fsb
