// See lm_gemm.cuh.  CTA = 8 warps (6 when operand X needs no normalisation):
//   warps 0-3  epilogue: TMEM lane quadrant = warp id; stream-K partial store / arrival / slot-ordered fix-up of this
//              CTA's slice of the batch rows / fused epilogue
//   warp  4    TMA producer (weights and operand X)
//   warp  5    TMEM allocator + single-thread tcgen05.mma issuer
//   warps 6-7  operand-X normalisers (NORM == 1): RMSNorm of the TMA-delivered residual rows, in place in the ring
#include "lm_gemm.cuh"
#include "umma.cuh"

#include <stdlib.h>

#include <vector>

namespace fsb {

namespace {

constexpr int kBN = kStepRows;
constexpr int kBTileBytes = kBN * kBlockK * 2;  // 4 KB
constexpr int kStageBytes = kATileBytes + kBTileBytes;
constexpr int kTmemCols = 2 * kBN;  // two accumulators: the epilogue of item n overlaps the MMAs of item n+1
constexpr int kEpiThreads = 128;
constexpr int kLoaderThreads = 64;
constexpr int kScratchBytes = 1024;  // r_s[32] | red[4][32]

__device__ __forceinline__ void bar_sync(int id, int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- fused epilogues on a slice of R batch rows [j0, j0 + R): acc[r] = complete fp32 dot product of feature
// (tile*128 + tid) with batch row j0 + r.  `red` = shared float[4][32]. ----

// Per-row sum over the 128 features of the tile: result for row r in red[q*32 + r], q = 0..3 (one per warp).
template <int R>
__device__ __forceinline__ void rows_warp_sums(const float (&sq)[R], float* red, int quad, int lane) {
    float keep = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float s = warp_sum(sq[r]);
        if (lane == r) keep = s;
    }
    red[quad * 32 + lane] = keep;
}

// Every epilogue is split in two: `*_pre` issues the loads that do not depend on the GEMM result (residual rows, RoPE
// table entries, bias, norm weights) -- it runs BEFORE the CTA waits for the other contributors' partials, so those
// round trips overlap the wait -- and `*_apply` consumes them once the complete dot products are known.

// llama.py:842-845 / 944-946: y = Linear(...) (bf16), x = x + y (bf16); plus sum_i x^2 of this tile per row
template <int R>
struct ResidPre {
    float xin[R];
    float b;
};
template <int R>
__device__ __forceinline__ void epi_resid_pre(const StepGemmParams& p, int tile, int tid, int j0, ResidPre<R>& q) {
    const int i = tile * 128 + tid;
    const bool ok = i < p.n_out;
#pragma unroll
    for (int r = 0; r < R; ++r)
        q.xin[r] = (p.resid != nullptr && ok && j0 + r < p.rows) ? bf2f(p.resid[static_cast<size_t>(j0 + r) * p.n_out + i]) : 0.f;
    q.b = (p.bias != nullptr && ok) ? bf2f(p.bias[i]) : 0.f;
}
template <int R>
__device__ __forceinline__ void epi_resid_apply(const StepGemmParams& p, const float (&acc)[R], const ResidPre<R>& q,
                                                int tile, int tid, int j0, float* red) {
    const int quad = tid >> 5, lane = tid & 31;
    const int i = tile * 128 + tid;
    const bool ok = i < p.n_out;
    float sq[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float y = p.bias != nullptr ? rbf(acc[r] + q.b) : rbf(acc[r]);
        const float x = p.resid != nullptr ? rbf(q.xin[r] + y) : y;
        const bool live = ok && j0 + r < p.rows;
        if (live) p.x_out[static_cast<size_t>(j0 + r) * p.n_out + i] = f2bf(x);
        sq[r] = live ? x * x : 0.f;
    }
    rows_warp_sums<R>(sq, red, quad, lane);
    bar_sync(1, kEpiThreads);
    if (quad == 0 && lane < R && j0 + lane < p.rows)
        p.ssq_out[(j0 + lane) * kSsqStride + tile] = ((red[lane] + red[32 + lane]) + red[64 + lane]) + red[96 + lane];
    bar_sync(1, kEpiThreads);
}

// llama.py:979-987: h = silu(w1 x) * w3 x, every intermediate a bf16 tensor
template <int R>
__device__ __forceinline__ void epi_swiglu(const StepGemmParams& p, const float (&acc)[R], int tile, int tid, int j0) {
    const int quad = tid >> 5, lane = tid & 31;
    const bool hi = (lane & 16) != 0;  // lanes 16..31 hold the w3 ("up") rows of the features lanes 0..15 gate
    const int f = tile * 64 + quad * 16 + (lane & 15);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float up = __shfl_xor_sync(0xffffffffu, acc[r], 16);
        if (!hi && j0 + r < p.rows && f < p.I) {
            const float g = rbf(acc[r]), u = rbf(up);
            const float s = rbf(g / (1.f + expf(-g)));
            p.h[static_cast<size_t>(j0 + r) * p.I + f] = f2bf(s * u);
        }
    }
}

template <int R>
__device__ __forceinline__ void epi_logits(const StepGemmParams& p, const float (&acc)[R], int tile, int tid, int j0) {
    const int i = tile * 128 + tid;
    if (i >= p.n_out) return;
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (j0 + r < p.rows) p.logits[static_cast<size_t>(j0 + r) * p.logits_ld + i] = rbf(acc[r]);
}

// llama.py:891-911: q/k/v = wqkv(x) (bf16), per-head nn.RMSNorm on q and k (fp32 math, one rounding),
// interleaved-pair RoPE in fp32 with the bf16 table, KVCache.update (llama.py:196-214).
template <int R>
struct QkvPre {
    uint32_t cs[R];  // (cos, sin) bf16 pair of this lane's rotary pair at the row's position
    float b, wn;
};
template <int R>
__device__ __forceinline__ void epi_qkv_pre(const StepGemmParams& p, int tile, int tid, int j0, QkvPre<R>& q) {
    const int quad = tid >> 5, lane = tid & 31;
    const int fb = tile * 128 + quad * 32;  // a warp's 32 features never straddle a head (Dh % 32 == 0)
    const bool wok = fb < p.n_out;
    const int head = fb / p.Dh;
    const int d = fb - head * p.Dh + lane;
    const int kind = head < p.H ? 0 : (head < p.H + p.Hkv ? 1 : 2);
    const __nv_bfloat16* nw = kind == 0 ? p.q_norm : (kind == 1 ? p.k_norm : nullptr);
    q.b = (p.bias != nullptr && wok) ? bf2f(p.bias[fb + lane]) : 0.f;
    q.wn = (nw != nullptr && wok) ? bf2f(nw[d]) : 1.f;
    const int half = p.Dh >> 1;
    // (row positions are re-read in the apply step: by then they sit in L1; keeping them would cost R registers)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool live = wok && kind != 2 && j0 + r < p.rows;
        const int pos = live ? p.row_pos[j0 + r] : 0;
        q.cs[r] = live ? *reinterpret_cast<const uint32_t*>(
                             p.freqs + (static_cast<size_t>(max(0, min(pos, p.S - 1))) * half + (d >> 1)) * 2)
                       : 0u;
    }
}
template <int R>
__device__ __forceinline__ void epi_qkv_apply(const StepGemmParams& p, const float (&acc)[R], const QkvPre<R>& q, int tile,
                                              int tid, int j0, float* red) {
    const int quad = tid >> 5, lane = tid & 31;
    const int fb = tile * 128 + quad * 32;
    const bool wok = fb < p.n_out;
    const int head = fb / p.Dh;
    const int d = fb - head * p.Dh + lane;
    const int kind = head < p.H ? 0 : (head < p.H + p.Hkv ? 1 : 2);
    float v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = p.bias != nullptr ? rbf(acc[r] + q.b) : rbf(acc[r]);
    if (p.q_norm != nullptr || p.k_norm != nullptr) {
        const bool normed = wok && (kind == 0 ? p.q_norm != nullptr : (kind == 1 && p.k_norm != nullptr));
        float sq[R];
#pragma unroll
        for (int r = 0; r < R; ++r) sq[r] = normed ? v[r] * v[r] : 0.f;
        rows_warp_sums<R>(sq, red, quad, lane);
        bar_sync(1, kEpiThreads);
        if (normed) {
            const int wph = p.Dh >> 5;  // warps per head
            const int q0 = (quad / wph) * wph;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float tot = 0.f;
                for (int u = 0; u < wph; ++u) tot += red[(q0 + u) * 32 + r];
                const float rl = rsqrtf(tot / static_cast<float>(p.Dh) + p.qk_eps);
                v[r] = rbf(v[r] * rl * q.wn);
            }
        }
        bar_sync(1, kEpiThreads);
    }
    if (!wok) return;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = j0 + r;
        if (j >= p.rows) break;
        float o = v[r];
        if (kind != 2) {
            const float c = bf_lo(q.cs[r]), s = bf_hi(q.cs[r]);
            const float partner = __shfl_xor_sync(0xffffffffu, o, 1);
            o = (lane & 1) ? __fadd_rn(__fmul_rn(o, c), __fmul_rn(partner, s))
                           : __fsub_rn(__fmul_rn(o, c), __fmul_rn(partner, s));
            o = rbf(o);
        }
        if (kind == 0) {
            p.q[(static_cast<size_t>(j) * p.H + head) * p.Dh + d] = f2bf(o);
        } else {
            const int pos = p.row_pos[j];
            if (pos >= 0 && pos < p.S) {  // a row parked at position -1 (idle slot) leaves the cache alone
                const int g = kind == 1 ? head - p.H : head - p.H - p.Hkv;
                __nv_bfloat16* cache = kind == 1 ? p.kcache : p.vcache;
                cache[((static_cast<size_t>(p.row_seq[j]) * p.Hkv + g) * p.S + pos) * p.Dh + d] = f2bf(o);
            }
        }
    }
}

template <int EPI, int R>
struct EpiPre {
    ResidPre<EPI == EPI_RESID ? R : 1> resid;
    QkvPre<EPI == EPI_QKV ? R : 1> qkv;
};
template <int EPI, int R>
__device__ __forceinline__ void epi_pre(const StepGemmParams& p, int tile, int tid, int j0, EpiPre<EPI, R>& q) {
    if constexpr (EPI == EPI_QKV) epi_qkv_pre<R>(p, tile, tid, j0, q.qkv);
    if constexpr (EPI == EPI_RESID) epi_resid_pre<R>(p, tile, tid, j0, q.resid);
}
template <int EPI, int R>
__device__ __forceinline__ void epi_apply(const StepGemmParams& p, const float (&acc)[R], const EpiPre<EPI, R>& q, int tile,
                                          int tid, int j0, float* red) {
    if constexpr (EPI == EPI_QKV) epi_qkv_apply<R>(p, acc, q.qkv, tile, tid, j0, red);
    else if constexpr (EPI == EPI_RESID) epi_resid_apply<R>(p, acc, q.resid, tile, tid, j0, red);
    else if constexpr (EPI == EPI_SWIGLU) epi_swiglu<R>(p, acc, tile, tid, j0);
    else epi_logits<R>(p, acc, tile, tid, j0);
}

// One shared tile, seen from one of its `np` contributors: wait until every partial of the tile has been published,
// then sum them IN SLOT ORDER for this CTA's slice of the batch rows [j0, j0 + R) and run the fused epilogue on it.
// Up to 32 independent loads are in flight per thread; the order of the additions never depends on arrival order.
template <int EPI, int R>
__device__ __forceinline__ void finish_shared_tile(const StepGemmParams& p, int tile, int tid, int j0, int np, float* red,
                                                   unsigned* arrive, unsigned* done, unsigned long long* trace) {
    EpiPre<EPI, R> pre;
    const bool mine = j0 < p.rows;  // with few live rows some contributors have no slice
    if (mine) epi_pre<EPI, R>(p, tile, tid, j0, pre);
    if (tid == 0) {
        // every contributor is resident (the grid fits the GPU in one wave) and has published before it waits:
        // bounded spin, a protocol bug becomes a trap instead of a hung GPU
        const long long t0 = clock64();
        while (ld_acquire_gpu(arrive + tile) < static_cast<unsigned>(np)) {
            if (clock64() - t0 > 4000000000ll) {
                printf("fsb: stream-K arrival timeout block=%d tile=%d have=%u want=%d\n", blockIdx.x, tile,
                       ld_acquire_gpu(arrive + tile), np);
                __trap();
            }
        }
        if (trace) trace[4] = globaltimer_ns();
        // the last contributor past the wait re-arms both counters for the next launch
        if (atomicAdd(done + tile, 1u) == static_cast<unsigned>(np - 1)) {
            arrive[tile] = 0;
            done[tile] = 0;
        }
    }
    bar_sync(1, kEpiThreads);
    if (!mine) return;
    constexpr int UQ = 32 / R;  // partials fetched per round
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    const float* src = p.ws + (static_cast<size_t>(tile) * 32 + j0) * 128 + tid;
    const size_t sstride = static_cast<size_t>(p.tiles) * 32 * 128;
    for (int q0 = 0; q0 < np; q0 += UQ) {
        float t[UQ][R];
#pragma unroll
        for (int u = 0; u < UQ; ++u)
#pragma unroll
            for (int r = 0; r < R; ++r)
                t[u][r] = (q0 + u < np && j0 + r < p.rows) ? __ldcg(src + static_cast<size_t>(q0 + u) * sstride + r * 128) : 0.f;
#pragma unroll
        for (int u = 0; u < UQ; ++u) {
            if (q0 + u < np) {
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] += t[u][r];
            }
        }
    }
    epi_apply<EPI, R>(p, acc, pre, tile, tid, j0, red);
}

template <int EPI, int NORM>
__global__ void __launch_bounds__(NORM ? 256 : 192, 2)
step_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ StepGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
    const int stages = p.stages;
    const uint32_t bars = tiles + static_cast<uint32_t>(stages) * kStageBytes;
    // barrier block: full[stages] (tile ready for the MMA), empty[stages], rawx[stages] (un-normalised X landed),
    // tmem_full[2], tmem_empty[2], TMEM base word
    const uint32_t full0 = bars, empty0 = bars + 8u * stages, rawx0 = bars + 16u * stages;
    const uint32_t tfull0 = bars + 24u * stages, tempty0 = tfull0 + 16u;
    const int scratch_off = ((24 * stages + 48 + 15) / 16) * 16;
    uint8_t* gen = smem_raw + (tiles - raw);
    uint8_t* bar_gen = gen + static_cast<size_t>(stages) * kStageBytes;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(bar_gen + 24 * stages + 32);
    const uint32_t tmem_slot = bars + 24u * stages + 32u;
    float* scratch = reinterpret_cast<float*>(bar_gen + scratch_off);
    float* r_s = scratch;       // [32] per-row rsqrt (normalisers)
    float* red = scratch + 32;  // [4][32] cross-warp reductions (epilogue)
    uint4* normw_s = reinterpret_cast<uint4*>(bar_gen + scratch_off + kScratchBytes);  // [kblocks*8] norm weights

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item_begin = p.cta_items[blockIdx.x], item_end = p.cta_items[blockIdx.x + 1];
    unsigned long long* trace = p.trace ? p.trace + static_cast<size_t>(blockIdx.x) * 8 : nullptr;
    if (trace && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        trace[0] = globaltimer_ns();
        trace[6] = smid;
        trace[7] = static_cast<unsigned long long>(item_end - item_begin);
    }

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < stages; ++s) {
            mbar_init(full0 + 8u * s, 1 + (NORM ? kLoaderThreads / 32 : 0));
            mbar_init(empty0 + 8u * s, 1);
            mbar_init(rawx0 + 8u * s, 1);
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

    pdl_launch_dependents();
    if (warp == 4) {
        // ===== TMA producer. Weights do not depend on the upstream kernel: the first ring-full of weight
        // tiles is requested BEFORE griddepcontrol.wait, so this GEMM's HBM stream starts while the previous
        // kernel is still in its tail. Operand X (what that kernel produces) follows after the wait. =====
        if (lane == 0) {
            int pre = 0;
            {
                int n = item_begin, kb = 0;
                int4 w = make_int4(0, 0, 0, 0);
                if (n < item_end) {
                    w = p.sched[n];
                    kb = w.y;
                }
                while (n < item_end && pre < stages) {
                    const int s = pre;
                    mbar_expect_tx(full0 + 8u * s, NORM ? kATileBytes : kStageBytes);
                    tma_load_3d(tiles + static_cast<uint32_t>(s) * kStageBytes, &tmA, full0 + 8u * s, kb * kBlockK,
                                w.x * kBlockM, 0, p.a_hint);
                    ++pre;
                    if (++kb >= w.z) {
                        if (++n < item_end) {
                            w = p.sched[n];
                            kb = w.y;
                        }
                    }
                }
            }
            pdl_wait();
            if (trace) trace[1] = globaltimer_ns();
            int it = 0;
            for (int n = item_begin; n < item_end; ++n) {
                const int4 w = p.sched[n];
                for (int kb = w.y; kb < w.z; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = static_cast<uint32_t>(it / stages) & 1u;
                    const uint32_t a_dst = tiles + static_cast<uint32_t>(s) * kStageBytes;
                    if (it >= pre) {
                        mbar_wait(empty0 + 8u * s, ph ^ 1u);
                        mbar_expect_tx(full0 + 8u * s, NORM ? kATileBytes : kStageBytes);
                        tma_load_3d(a_dst, &tmA, full0 + 8u * s, kb * kBlockK, w.x * kBlockM, 0, p.a_hint);
                    }
                    if (NORM) {
                        mbar_expect_tx(rawx0 + 8u * s, kBTileBytes);
                        tma_load_3d(a_dst + kATileBytes, &tmB, rawx0 + 8u * s, kb * kBlockK, 0, 0, p.b_hint);
                    } else {
                        tma_load_3d(a_dst + kATileBytes, &tmB, full0 + 8u * s, kb * kBlockK, 0, 0, p.b_hint);
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(kBN);
            int it = 0;
            for (int n = item_begin; n < item_end; ++n) {
                const int4 w = p.sched[n];
                const int a = (n - item_begin) & 1;
                const uint32_t aph = static_cast<uint32_t>((n - item_begin) >> 1) & 1u;
                mbar_wait(tempty0 + 8u * a, aph ^ 1u);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(a * kBN);
                uint32_t acc = 0;
                for (int kb = w.y; kb < w.z; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = static_cast<uint32_t>(it / stages) & 1u;
                    mbar_wait(full0 + 8u * s, ph);
                    tc_fence_after();
                    const uint32_t a_src = tiles + static_cast<uint32_t>(s) * kStageBytes;
                    const uint64_t ad = make_sdesc(a_src), bd = make_sdesc(a_src + kATileBytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        umma_bf16(d_tmem, ad + 2u * k, bd + 2u * k, idesc, acc);
                        acc = 1;
                    }
                    umma_commit(empty0 + 8u * s);  // frees the ring slot once these MMAs retire
                }
                umma_commit(tfull0 + 8u * a);  // accumulator complete
            }
        }
    } else if (warp >= 6) {
        // ===== operand-X normalisers: rbf(rbf(x * r_row) * w) in place on the swizzled tile TMA delivered =====
        if (NORM) {
            const int t = threadIdx.x - 192;  // 0..63
            // the norm weights are constants: stage them before waiting for the upstream kernel
            const int nchunks = ((p.K + kBlockK - 1) / kBlockK) * 8;
            for (int ch = t; ch < nchunks; ch += kLoaderThreads)
                normw_s[ch] = ch * 8 < p.K ? __ldg(reinterpret_cast<const uint4*>(p.norm_w + ch * 8)) : make_uint4(0, 0, 0, 0);
            pdl_wait();
            if (t < 32) {
                // rsqrt(mean(x^2) + eps) of row t: the producer's per-tile sums, added in tile order (all loads in flight)
                const float* q = p.x_ssq + t * kSsqStride;
                float a[kSsqStride];
#pragma unroll
                for (int u = 0; u < kSsqStride; ++u) a[u] = u < p.x_nt ? __ldcg(q + u) : 0.f;
                float tot = 0.f;
#pragma unroll
                for (int u = 0; u < kSsqStride; ++u)
                    if (u < p.x_nt) tot += a[u];
                r_s[t] = rsqrtf(tot / static_cast<float>(p.K) + p.eps);
            }
            bar_sync(2, kLoaderThreads);
            // 32 rows x 8 sixteen-byte chunks per k-block; thread t owns chunk (t & 7) of rows (t >> 3) + 8e
            const int c = t & 7, rb = t >> 3;
            float rr[4];
            uint32_t off[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int row = rb + 8 * e;
                rr[e] = r_s[row];
                off[e] = static_cast<uint32_t>(row * 128 + ((c ^ (row & 7)) << 4));  // SWIZZLE_128B
            }
            int it = 0;
            for (int n = item_begin; n < item_end; ++n) {
                const int4 w = p.sched[n];
                for (int kb = w.y; kb < w.z; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = static_cast<uint32_t>(it / stages) & 1u;
                    const uint4 wv = normw_s[kb * 8 + c];
                    const float wf[8] = {bf_lo(wv.x), bf_hi(wv.x), bf_lo(wv.y), bf_hi(wv.y),
                                         bf_lo(wv.z), bf_hi(wv.z), bf_lo(wv.w), bf_hi(wv.w)};
                    uint8_t* btile = gen + static_cast<size_t>(s) * kStageBytes + kATileBytes;
                    mbar_wait(rawx0 + 8u * s, ph);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        uint4* cell = reinterpret_cast<uint4*>(btile + off[e]);
                        const uint4 xv = *cell;
                        const float xf[8] = {bf_lo(xv.x), bf_hi(xv.x), bf_lo(xv.y), bf_hi(xv.y),
                                             bf_lo(xv.z), bf_hi(xv.z), bf_lo(xv.w), bf_hi(xv.w)};
                        float nf[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) nf[q] = rbf(rbf(xf[q] * rr[e]) * wf[q]);
                        uint4 o;
                        o.x = pack_bf2(nf[0], nf[1]);
                        o.y = pack_bf2(nf[2], nf[3]);
                        o.z = pack_bf2(nf[4], nf[5]);
                        o.w = pack_bf2(nf[6], nf[7]);
                        *cell = o;
                    }
                    fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(full0 + 8u * s);
                }
            }
        }
    } else {
        // ===== epilogue warps =====
        const int tid = threadIdx.x;  // 0..127 = TMEM lane = feature inside the tile
        unsigned* arrive = p.tile_ctr;
        unsigned* done = p.tile_ctr + p.tile_ctr_len;
        pdl_wait();
        // Phase 1: as each accumulator completes, publish its fp32 partial (or, when this CTA ran the whole
        // reduction of the tile, finish it). Nothing here waits for another CTA, so the partials of a tile's
        // contributors appear as soon as each of them has streamed its share of the weights.
        for (int n = item_begin; n < item_end; ++n) {
            const int4 w = p.sched[n];
            const int tile = w.x, slot = w.w;
            const int a = (n - item_begin) & 1;
            const uint32_t aph = static_cast<uint32_t>((n - item_begin) >> 1) & 1u;
            mbar_wait(tfull0 + 8u * a, aph);
            tc_fence_after();
            if (trace && tid == 0 && n == item_begin) trace[2] = globaltimer_ns();
            float v[32];
            {
                uint32_t r[32];
                tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(a * kBN), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            }
            // the accumulator is in registers: hand it back to the MMA warp right away
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8u * a);

            if (p.nparts[tile] == 1) {  // the whole reduction ran in this CTA
                EpiPre<EPI, 32> pre;
                epi_pre<EPI, 32>(p, tile, tid, 0, pre);
                epi_apply<EPI, 32>(p, v, pre, tile, tid, 0, red);
                continue;
            }
            float* dst = p.ws + ((static_cast<size_t>(slot) * p.tiles + tile) * 32) * 128 + tid;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < p.rows) __stcg(dst + j * 128, v[j]);
            bar_sync(1, kEpiThreads);
            if (tid == 0) {
                __threadfence();  // cumulative: orders the stores of all 128 threads (joined by the barrier) before the arrival
                atomicAdd(arrive + tile, 1u);
            }
        }
        if (trace && tid == 0) trace[3] = globaltimer_ns();
        // Phase 2: for every shared tile, wait until all of its partials are there, then sum them IN SLOT ORDER
        // for this CTA's slice of the batch rows (contributor `slot` takes rows [slot*R, slot*R + R)) and run the
        // fused epilogue on that slice.
        for (int n = item_begin; n < item_end; ++n) {
            const int4 w = p.sched[n];
            const int tile = w.x, slot = w.w;
            const int np = p.nparts[tile];
            if (np == 1) continue;
            int R = 1;
            while (R * np < p.rows) R <<= 1;
            const int j0 = slot * R;
            switch (R) {
                case 1: finish_shared_tile<EPI, 1>(p, tile, tid, j0, np, red, arrive, done, trace); break;
                case 2: finish_shared_tile<EPI, 2>(p, tile, tid, j0, np, red, arrive, done, trace); break;
                case 4: finish_shared_tile<EPI, 4>(p, tile, tid, j0, np, red, arrive, done, trace); break;
                case 8: finish_shared_tile<EPI, 8>(p, tile, tid, j0, np, red, arrive, done, trace); break;
                default: finish_shared_tile<EPI, 16>(p, tile, tid, j0, np, red, arrive, done, trace); break;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (trace && threadIdx.x == 0) trace[5] = globaltimer_ns();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int EPI, int NORM>
int launch_t(const StepGemmPlan& plan, cudaStream_t st) {
    auto k = step_gemm_kernel<EPI, NORM>;
    FSB_LAUNCH(k, plan.grid, dim3(NORM ? 256 : 192), plan.smem, st, plan.tmA, plan.tmB, plan.p);
    return 0;
}

const void* kernel_of(int epi, int norm) {
    if (epi == EPI_QKV && norm) return reinterpret_cast<const void*>(step_gemm_kernel<EPI_QKV, 1>);
    if (epi == EPI_RESID && !norm) return reinterpret_cast<const void*>(step_gemm_kernel<EPI_RESID, 0>);
    if (epi == EPI_SWIGLU && norm) return reinterpret_cast<const void*>(step_gemm_kernel<EPI_SWIGLU, 1>);
    if (epi == EPI_LOGITS && norm) return reinterpret_cast<const void*>(step_gemm_kernel<EPI_LOGITS, 1>);
    return nullptr;
}

}  // namespace

int step_gemm_init() {
    static bool done = false;
    if (done) return 0;
#define FSB_STEP_ATTR(E_, B_)                                                                                  \
    FSB_CUDA(cudaFuncSetAttribute(step_gemm_kernel<E_, B_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
    FSB_CUDA(cudaFuncSetAttribute(step_gemm_kernel<E_, B_>, cudaFuncAttributePreferredSharedMemoryCarveout,            \
                                  cudaSharedmemCarveoutMaxShared));
    FSB_STEP_ATTR(EPI_QKV, 1) FSB_STEP_ATTR(EPI_RESID, 0) FSB_STEP_ATTR(EPI_SWIGLU, 1) FSB_STEP_ATTR(EPI_LOGITS, 1)
#undef FSB_STEP_ATTR
    done = true;
    return 0;
}

int step_plan_init(StepGemmPlan* plan, int epi, const __nv_bfloat16* w, int n_out, int K, const __nv_bfloat16* act,
                   bool norm_on_load, int num_ctas, int stages, float* ws, size_t ws_floats, unsigned* tile_ctr,
                   int tile_ctr_len) {
    memset(plan, 0, sizeof(*plan));
    FSB_CHECK(K % 8 == 0, "step GEMM: K=%d must be a multiple of 8", K);
    FSB_CHECK(act != nullptr, "step GEMM: operand X missing");
    const int norm = norm_on_load ? 1 : 0;
    const void* kernel = kernel_of(epi, norm);
    FSB_CHECK(kernel != nullptr, "step GEMM: EPI_RESID takes a ready operand, the other epilogues normalise on load");
    FSB_TRY(step_gemm_init());
    GemmOperand A{w, K, n_out, 1, K, static_cast<long long>(n_out) * K};
    FSB_TRY(gemm_make_tmap(&plan->tmA, A, kBlockM));
    GemmOperand B{act, K, kStepRows, 1, K, static_cast<long long>(kStepRows) * K};
    FSB_TRY(gemm_make_tmap(&plan->tmB, B, kBN));
    const int tiles = cdiv(n_out, kBlockM), kblocks = cdiv(K, kBlockK);
    FSB_CHECK(tiles <= tile_ctr_len, "step GEMM: %d tiles exceed the ticket array (%d)", tiles, tile_ctr_len);
    const int normw_bytes = norm ? kblocks * 128 : 0;
    auto smem_of = [&](int st) {
        return static_cast<size_t>(1024) + static_cast<size_t>(st) * kStageBytes + ((24 * st + 48 + 15) / 16) * 16 +
               kScratchBytes + normw_bytes;
    };
    while (stages > 2 && 2 * (smem_of(stages) + 1024) > 228 * 1024) --stages;  // two CTAs per SM
    FSB_CHECK(stages >= 2, "step GEMM: ring too shallow");
    // The fix-up waits for the other contributors of a tile: the whole grid must be resident at once.
    // cudaOccupancyMaxActiveBlocksPerMultiprocessor reports 1 for every kernel that allocates tensor memory, although
    // two such CTAs do share an SM (verified with %smid stamps, tools/trace_step_gemms.py): count registers and
    // shared memory ourselves.
    int dev = 0, sms = 0, smem_sm = 0, regs_sm = 0, resv = 0;
    FSB_CUDA(cudaGetDevice(&dev));
    FSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    FSB_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    FSB_CUDA(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev));
    FSB_CUDA(cudaDeviceGetAttribute(&resv, cudaDevAttrReservedSharedMemoryPerBlock, dev));
    cudaFuncAttributes fa;
    FSB_CUDA(cudaFuncGetAttributes(&fa, kernel));
    const int threads = norm ? 256 : 192;
    const int regs_cta = ((fa.numRegs + 7) / 8) * 8 * threads;
    const int per_sm = std::min<int>(regs_sm / regs_cta, smem_sm / static_cast<int>(smem_of(stages) + resv + fa.sharedSizeBytes));
    FSB_CHECK(per_sm >= 1, "step GEMM: kernel does not fit an SM (smem %zu, %d registers)", smem_of(stages), fa.numRegs);
    if (num_ctas > per_sm * sms) num_ctas = per_sm * sms;
    // stream-K: units are (tile, k-block) pairs in tile-major order; CTA c streams units [c*U/n, (c+1)*U/n)
    const long long U = static_cast<long long>(tiles) * kblocks;
    if (num_ctas > U) num_ctas = static_cast<int>(U);
    std::vector<int4> items;
    std::vector<int> cta_items(num_ctas + 1, 0), nparts(tiles, 0);
    for (int c = 0; c < num_ctas; ++c) {
        long long u0 = U * c / num_ctas;
        const long long u1 = U * (c + 1) / num_ctas;
        cta_items[c] = static_cast<int>(items.size());
        while (u0 < u1) {
            const int t = static_cast<int>(u0 / kblocks);
            const long long tend = static_cast<long long>(t + 1) * kblocks;
            const long long e = u1 < tend ? u1 : tend;
            items.push_back(make_int4(t, static_cast<int>(u0 - static_cast<long long>(t) * kblocks),
                                      static_cast<int>(e - static_cast<long long>(t) * kblocks), nparts[t]++));
            u0 = e;
        }
    }
    cta_items[num_ctas] = static_cast<int>(items.size());
    int maxp = 0;
    for (int t = 0; t < tiles; ++t) maxp = std::max(maxp, nparts[t]);
    plan->max_parts = maxp;
    FSB_CHECK(static_cast<size_t>(maxp) * tiles * 128 * 32 <= ws_floats, "step GEMM: partial workspace too small");
    FSB_CUDA(cudaMalloc(&plan->sched_dev, items.size() * sizeof(int4)));
    FSB_CUDA(cudaMemcpy(plan->sched_dev, items.data(), items.size() * sizeof(int4), cudaMemcpyHostToDevice));
    FSB_CUDA(cudaMalloc(&plan->cta_items_dev, cta_items.size() * sizeof(int)));
    FSB_CUDA(cudaMemcpy(plan->cta_items_dev, cta_items.data(), cta_items.size() * sizeof(int), cudaMemcpyHostToDevice));
    FSB_CUDA(cudaMalloc(&plan->nparts_dev, tiles * sizeof(int)));
    FSB_CUDA(cudaMemcpy(plan->nparts_dev, nparts.data(), tiles * sizeof(int), cudaMemcpyHostToDevice));
    StepGemmParams& p = plan->p;
    p.sched = reinterpret_cast<const int4*>(plan->sched_dev);
    p.cta_items = plan->cta_items_dev;
    p.nparts = plan->nparts_dev;
    p.tiles = tiles;
    p.stages = stages;
    p.n_out = n_out;
    p.K = K;
    p.rows = kStepRows;
    p.a_hint = kEvictFirst;  // weights are streamed once per step (>> L2)
    p.b_hint = kEvictLast;   // the activation tile is re-read by every CTA
    p.ws = ws;
    p.tile_ctr = tile_ctr;
    p.tile_ctr_len = tile_ctr_len;
    plan->grid = dim3(static_cast<unsigned>(num_ctas), 1, 1);
    plan->smem = smem_of(stages);
    plan->epi = epi;
    plan->bload = norm;
    plan->weight_bytes = static_cast<double>(n_out) * K * 2;
    return 0;
}

void step_plan_free(StepGemmPlan* plan) {
    if (plan->sched_dev) cudaFree(plan->sched_dev);
    if (plan->cta_items_dev) cudaFree(plan->cta_items_dev);
    if (plan->nparts_dev) cudaFree(plan->nparts_dev);
    plan->sched_dev = nullptr;
    plan->cta_items_dev = nullptr;
    plan->nparts_dev = nullptr;
}

int step_gemm_launch(const StepGemmPlan& plan, cudaStream_t st) {
    switch (plan.epi) {
        case EPI_QKV: return launch_t<EPI_QKV, 1>(plan, st);
        case EPI_RESID: return launch_t<EPI_RESID, 0>(plan, st);
        case EPI_SWIGLU: return launch_t<EPI_SWIGLU, 1>(plan, st);
        case EPI_LOGITS: return launch_t<EPI_LOGITS, 1>(plan, st);
    }
    set_error("step_gemm_launch: bad epilogue %d", plan.epi);
    return 1;
}

}  // namespace fsb
