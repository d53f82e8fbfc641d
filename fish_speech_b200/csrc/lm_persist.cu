// Persistent decode kernel: ONE launch runs every layer of a transformer stack for the <=32 decode rows.
//
// Why: as separate kernels each weight-streaming GEMM pays a ~5 us ramp/drain bubble and every consumer
// kernel between two GEMMs leaves HBM idle (profiles/r01_gemm_chain_trace_pdl.txt). Here the TMA producer
// warp is decoupled from the phase structure: as soon as ring slots free up it requests the NEXT GEMM's
// weight tiles, while the epilogue of the current GEMM, the grid-wide barriers and the consumer phases
// (qkv post-processing + attention, residual + RMSNorm, SwiGLU — the same math as lm_kernels.cu) run.
// Only the activation tile of a stage waits for the barrier that publishes the consumer's output.
//
// CTA (one per SM, all co-resident): warps 0-3 GEMM epilogue + workers, warps 4-7 workers, warp 8 TMA
// producer, warp 9 MMA issuer. Phases per layer: G(qkv) | B | prep+attn | B | G(wo) | B | resid+norm | B |
// G(w1|w3) | B | swiglu | B | G(w2) | B | resid+norm | B      (B = grid barrier on a global counter).
#include "lm_kernels.cuh"
#include "lm_persist.cuh"
#include "umma.cuh"

namespace fsb {

namespace {

constexpr int kPkWorkers = 256;            // warps 0-7
constexpr int kPkProducerWarp = 8, kPkMmaWarp = 9;
constexpr int kPkThreads = 320;
constexpr int kPkBN = 32;
constexpr int kPkStageBytes = kATileBytes + kPkBN * kBlockK * 2;  // 20 KB
constexpr int kPkTmemCols = 2 * kPkBN;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void spin_until(const unsigned* ctr, unsigned target) {
    if (ld_acquire_u32(ctr) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire_u32(ctr) < target) {
        if (clock64() - t0 > 4000000000ll) {
            printf("fsb: grid barrier timeout block=%d thread=%d have=%u want=%u\n", blockIdx.x, threadIdx.x,
                   ld_acquire_u32(ctr), target);
            __trap();
        }
    }
}

// grid-wide barrier for the 256 worker threads of every CTA
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
    __threadfence();
    named_bar_sync(1, kPkWorkers);
    if (threadIdx.x == 0) {
        atomicAdd(ctr, 1u);
        spin_until(ctr, target);
    }
    named_bar_sync(1, kPkWorkers);
}

struct Ring {
    uint32_t tiles, full0, empty0, tfull0, tempty0;
    int stages;
};

__device__ __forceinline__ void get_item(const PkGemm& G, int n, int& i0, int& kb0, int& kb1, int& slot) {
    const int4 w = G.sched[n];
    i0 = (w.x & 0xffff) * kBlockM;
    kb0 = w.y;
    kb1 = w.z;
    slot = w.w;
}

// ---- producer: stream one GEMM; `it` is the global k-block counter of the ring ----
__device__ void produce_gemm(const PkGemm& G, const Ring& R, int& it, const unsigned* bar, unsigned need,
                             bool first, int l2_prefetch) {
    const int item_begin = G.cta_items[blockIdx.x], item_end = G.cta_items[blockIdx.x + 1];
    // pass 1: weight tiles for as many k-blocks as the ring holds
    int pre = 0;
    {
        int n = item_begin, kb = 0, i0 = 0, kb0 = 0, kb1 = 0, slot = 0;
        if (n < item_end) {
            get_item(G, n, i0, kb0, kb1, slot);
            kb = kb0;
        }
        while (n < item_end && pre < R.stages) {
            const int j = it + pre;
            const int s = j % R.stages;
            const uint32_t ph = static_cast<uint32_t>(j / R.stages) & 1u;
            mbar_wait(R.empty0 + 8u * s, ph ^ 1u);
            mbar_expect_tx(R.full0 + 8u * s, kPkStageBytes);
            tma_load_3d(R.tiles + static_cast<uint32_t>(s) * kPkStageBytes, &G.tmA, R.full0 + 8u * s, kb * kBlockK, i0,
                        0, kEvictFirst);
            ++pre;
            if (++kb >= kb1) {
                if (++n < item_end) {
                    get_item(G, n, i0, kb0, kb1, slot);
                    kb = kb0;
                }
            }
        }
    }
    // the activation operand is published by the barrier `need` (or by the upstream kernel for the first GEMM)
    if (first) {
        pdl_wait();
    } else if (ld_acquire_u32(bar) < need) {
        // While the consumers of the previous GEMM are still running HBM would idle: use the wait to pull
        // the NEXT weight tiles of this CTA's range (beyond what the ring already holds) into L2, paced so
        // the requests do not pile up in front of other CTAs' demand loads.
        int n = item_begin, kb = 0, i0 = 0, kb0 = 0, kb1 = 0, slot = 0, skip = pre, issued = 0;
        if (n < item_end) {
            get_item(G, n, i0, kb0, kb1, slot);
            kb = kb0;
        }
        const long long t0 = clock64();
        long long next_t = t0;
        while (ld_acquire_u32(bar) < need) {
            const long long now = clock64();
            if (n < item_end && issued < l2_prefetch && now >= next_t) {
                if (skip > 0) {
                    --skip;
                } else {
                    tma_prefetch_l2_3d(&G.tmA, kb * kBlockK, i0, 0);
                    ++issued;
                    next_t = now + 400;
                }
                if (++kb >= kb1) {
                    if (++n < item_end) {
                        get_item(G, n, i0, kb0, kb1, slot);
                        kb = kb0;
                    }
                }
            }
            if (now - t0 > 4000000000ll) {
                printf("fsb: producer barrier timeout block=%d have=%u want=%u\n", blockIdx.x, ld_acquire_u32(bar), need);
                __trap();
            }
        }
    }
    fence_proxy_async_all();
    int local = 0;
    for (int n = item_begin; n < item_end; ++n) {
        int i0, kb0, kb1, slot;
        get_item(G, n, i0, kb0, kb1, slot);
        for (int kb = kb0; kb < kb1; ++kb, ++local) {
            const int j = it + local;
            const int s = j % R.stages;
            const uint32_t ph = static_cast<uint32_t>(j / R.stages) & 1u;
            const uint32_t dst = R.tiles + static_cast<uint32_t>(s) * kPkStageBytes;
            if (local >= pre) {
                mbar_wait(R.empty0 + 8u * s, ph ^ 1u);
                mbar_expect_tx(R.full0 + 8u * s, kPkStageBytes);
                tma_load_3d(dst, &G.tmA, R.full0 + 8u * s, kb * kBlockK, i0, 0, kEvictFirst);
            }
            tma_load_3d(dst + kATileBytes, &G.tmB, R.full0 + 8u * s, kb * kBlockK, 0, 0, kEvictLast);
        }
    }
    it += local;
}

// ---- MMA issuer ----
__device__ void mma_gemm(const PkGemm& G, const Ring& R, int& it, int& acc_it, uint32_t tmem_base) {
    constexpr uint32_t idesc = make_idesc(kPkBN);
    const int item_begin = G.cta_items[blockIdx.x], item_end = G.cta_items[blockIdx.x + 1];
    for (int n = item_begin; n < item_end; ++n, ++acc_it) {
        int i0, kb0, kb1, slot;
        get_item(G, n, i0, kb0, kb1, slot);
        const int a = acc_it & 1;
        const uint32_t aph = static_cast<uint32_t>(acc_it >> 1) & 1u;
        mbar_wait(R.tempty0 + 8u * a, aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(a * kPkBN);
        uint32_t acc = 0;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const int s = it % R.stages;
            const uint32_t ph = static_cast<uint32_t>(it / R.stages) & 1u;
            mbar_wait(R.full0 + 8u * s, ph);
            tc_fence_after();
            const uint32_t src = R.tiles + static_cast<uint32_t>(s) * kPkStageBytes;
            const uint64_t ad = make_sdesc(src), bd = make_sdesc(src + kATileBytes);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
                umma_bf16(d_tmem, ad + 2u * k, bd + 2u * k, idesc, acc);
                acc = 1;
            }
            umma_commit(R.empty0 + 8u * s);
        }
        umma_commit(R.tfull0 + 8u * a);
    }
}

// ---- epilogue (warps 0-3): TMEM -> fp32 partials ----
__device__ void epilogue_gemm(const PkGemm& G, const Ring& R, int& acc_it, uint32_t tmem_base, float* ws, int rows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item_begin = G.cta_items[blockIdx.x], item_end = G.cta_items[blockIdx.x + 1];
    for (int n = item_begin; n < item_end; ++n, ++acc_it) {
        int i0, kb0, kb1, slot;
        get_item(G, n, i0, kb0, kb1, slot);
        const int a = acc_it & 1;
        const uint32_t aph = static_cast<uint32_t>(acc_it >> 1) & 1u;
        mbar_wait(R.tfull0 + 8u * a, aph);
        tc_fence_after();
        const int i = i0 + warp * 32 + lane;
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(a * kPkBN), r);
        tmem_ld_wait();
        if (i < G.n_out) {
            float* base = ws + static_cast<size_t>(slot) * G.slot_stride + i;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < rows) base[static_cast<size_t>(j) * G.n_out] = __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(R.tempty0 + 8u * a);
    }
}

__device__ __forceinline__ Partials parts_of(const PkGemm& G, const float* ws) {
    Partials P;
    P.ws = ws;
    P.slot_stride = G.slot_stride;
    P.ld = G.n_out;
    P.nparts = G.nparts;
    P.max_parts = G.max_parts;
    return P;
}

// ---- consumer phases (256 worker threads; same rounding points as lm_kernels.cu) ----
__device__ float worker_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    named_bar_sync(2, kPkWorkers);
    if (lane == 0) red[w] = v;
    named_bar_sync(2, kPkWorkers);
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < kPkWorkers / 32; ++q) t += red[q];
    return t;
}

// x = rbf(x + rbf(sum parts + bias)); xn = rbf(rbf(x * r) * w)      (llama.py:842-845, 990-1001)
__device__ void phase_resid_norm(const PkArgs& A, const PkGemm& G, const __nv_bfloat16* bias,
                                 const __nv_bfloat16* norm_w, float* scratch) {
    const Partials P = parts_of(G, A.ws);
    const int D = A.D;
    constexpr int NE = 16;  // D <= 4096
    for (int row = blockIdx.x; row < A.rows; row += gridDim.x) {
        int rws[NE], fts[NE];
        bool ok[NE];
        float y[NE], v[NE], xin[NE], nw[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            fts[e] = threadIdx.x + e * kPkWorkers;
            rws[e] = row;
            ok[e] = fts[e] < D;
            xin[e] = ok[e] ? bf2f(A.xres[static_cast<size_t>(row) * D + fts[e]]) : 0.f;
            nw[e] = ok[e] ? bf2f(norm_w[fts[e]]) : 0.f;
        }
        sum_parts_n<NE>(P, rws, fts, ok, y, G.max_parts);
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            v[e] = 0.f;
            if (ok[e]) {
                float yy = y[e];
                if (bias) yy += bf2f(bias[fts[e]]);
                const float x = rbf(xin[e] + rbf(yy));
                v[e] = x;
                ss += x * x;
                A.xres[static_cast<size_t>(row) * D + fts[e]] = f2bf(x);
            }
        }
        const float tot = worker_sum(ss, scratch);
        const float r = rsqrtf(tot / static_cast<float>(D) + A.eps);
#pragma unroll
        for (int e = 0; e < NE; ++e)
            if (ok[e]) A.xn[static_cast<size_t>(row) * D + fts[e]] = f2bf(rbf(rbf(v[e] * r) * nw[e]));
        named_bar_sync(2, kPkWorkers);
    }
}

__device__ void phase_swiglu(const PkArgs& A, const PkGemm& G) {
    const Partials P = parts_of(G, A.ws);
    const long long total = static_cast<long long>(A.rows) * A.I;
    const long long stride = static_cast<long long>(gridDim.x) * kPkWorkers;
    constexpr int NE = 4;
    for (long long e0 = static_cast<long long>(blockIdx.x) * kPkWorkers + threadIdx.x; e0 < total; e0 += NE * stride) {
        int rws[2 * NE], fts[2 * NE];
        bool ok[2 * NE];
        float y[2 * NE];
#pragma unroll
        for (int k = 0; k < NE; ++k) {
            const long long e = e0 + k * stride;
            const bool o = e < total;
            const int row = o ? static_cast<int>(e / A.I) : 0;
            const int i = o ? static_cast<int>(e - static_cast<long long>(row) * A.I) : 0;
            rws[2 * k] = rws[2 * k + 1] = row;
            fts[2 * k] = i;
            fts[2 * k + 1] = A.I + i;
            ok[2 * k] = ok[2 * k + 1] = o;
        }
        sum_parts_n<2 * NE>(P, rws, fts, ok, y, G.max_parts);
#pragma unroll
        for (int k = 0; k < NE; ++k) {
            if (ok[2 * k]) {
                const float g = rbf(y[2 * k]), c = rbf(y[2 * k + 1]);
                const float sg = rbf(g / (1.f + expf(-g)));
                A.hbuf[static_cast<size_t>(rws[2 * k]) * A.I + fts[2 * k]] = f2bf(sg * c);
            }
        }
    }
}

// q/k/v post-processing + KV append + attention for one (row, kv-head) item   (llama.py:891-934 / 948-976).
// Each HALF of the worker threads (128 threads, 4 warps) takes its own item, so the 32 x Hkv items of a
// decode step fit in one round over 2 x #SM half-CTAs.
constexpr int kHalf = 128, kHalfWarps = 4;

template <int DH, int G>
__device__ void phase_prep_attn(const PkArgs& A, const PkLayer& L, float* sm_all, bool kv_only) {
    const Partials P = parts_of(L.qkv, A.ws);
    const int half = threadIdx.x / kHalf, ht = threadIdx.x % kHalf;
    const int hw = ht >> 5, lane = ht & 31;
    const int bar_id = 2 + half;
    const int lcap = A.S;
    const size_t per_half = static_cast<size_t>(G + 2) * DH + 8 + static_cast<size_t>(G) * lcap +
                            static_cast<size_t>(kHalfWarps) * G * DH;
    float* sm = sm_all + half * per_half;
    float* vals = sm;                    // [(G+2)][DH]  q heads, k, v (post norm / rope, bf16-rounded)
    float* hss = vals + (G + 2) * DH;    // [(G+2)] per-head sum of squares
    float* sc = hss + 8;                 // [G][lcap]
    float* red = sc + G * lcap;          // [4][G][DH]
    const float scale = rsqrtf(static_cast<float>(DH));
    const int items = A.rows * A.Hkv;
    constexpr int NV = ((G + 2) * DH + kHalf - 1) / kHalf;
    for (int item = blockIdx.x * 2 + half; item < items; item += gridDim.x * 2) {
        const int row = item / A.Hkv, g = item - row * A.Hkv;
        const int b = A.row_seq[row], pos = A.row_pos[row];
        // 1. raw values: q heads g*G..g*G+G-1, then k head g, then v head g
        {
            int rws[NV], fts[NV];
            bool ok[NV];
            float y[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const int e = ht + k * kHalf;
                ok[k] = e < (G + 2) * DH;
                const int hh = ok[k] ? e / DH : 0, d = ok[k] ? e - hh * DH : 0;
                fts[k] = hh < G ? (g * G + hh) * DH + d : (hh == G ? (A.H + g) * DH + d : (A.H + A.Hkv + g) * DH + d);
                rws[k] = row;
            }
            sum_parts_n<NV>(P, rws, fts, ok, y, L.qkv.max_parts);
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                if (ok[k]) {
                    float v = y[k];
                    if (L.bqkv) v += bf2f(L.bqkv[fts[k]]);
                    vals[ht + k * kHalf] = rbf(v);
                }
            }
        }
        named_bar_sync(bar_id, kHalf);
        if (A.qk_norm) {
            for (int hh = hw; hh < G + 1; hh += kHalfWarps) {  // q heads and the k head
                float s = 0.f;
                for (int d = lane; d < DH; d += 32) s += vals[hh * DH + d] * vals[hh * DH + d];
                s = warp_sum(s);
                if (lane == 0) hss[hh] = s;
            }
            named_bar_sync(bar_id, kHalf);
        }
        // 2. norm + RoPE on pairs, write K/V to the cache
        for (int e = ht; e < (G + 2) * DH / 2; e += kHalf) {
            const int hh = e / (DH / 2), t = e - hh * (DH / 2);
            float v0 = vals[hh * DH + 2 * t], v1 = vals[hh * DH + 2 * t + 1];
            if (hh <= G) {
                if (A.qk_norm) {
                    const __nv_bfloat16* nw = hh < G ? L.q_norm : L.k_norm;
                    const float r = rsqrtf(hss[hh] / static_cast<float>(DH) + A.eps);
                    v0 = rbf(v0 * r * bf2f(nw[2 * t]));
                    v1 = rbf(v1 * r * bf2f(nw[2 * t + 1]));
                }
                const __nv_bfloat16* f = A.freqs + (static_cast<size_t>(pos) * (DH / 2) + t) * 2;
                const float c = bf2f(f[0]), sn = bf2f(f[1]);
                const float o0 = __fsub_rn(__fmul_rn(v0, c), __fmul_rn(v1, sn));
                const float o1 = __fadd_rn(__fmul_rn(v1, c), __fmul_rn(v0, sn));
                v0 = rbf(o0);
                v1 = rbf(o1);
            }
            if (hh >= G) {
                __nv_bfloat16* cache = hh == G ? L.kcache : L.vcache;
                uint32_t* dst = reinterpret_cast<uint32_t*>(
                    cache + ((static_cast<size_t>(b) * A.Hkv + g) * A.S + pos) * DH);
                dst[t] = pack_bf2(v0, v1);
            }
            vals[hh * DH + 2 * t] = v0;
            vals[hh * DH + 2 * t + 1] = v1;
        }
        __threadfence_block();
        named_bar_sync(bar_id, kHalf);
        if (kv_only) continue;
        // 3. attention over cache positions [0, pos] (the row for `pos` was just written by this half-CTA)
        const int Lq = pos + 1;
        const size_t cache_base = (static_cast<size_t>(b) * A.Hkv + g) * A.S * DH;
        const __nv_bfloat16* kc = L.kcache + cache_base;
        const __nv_bfloat16* vc = L.vcache + cache_base;
        constexpr int LPR = DH / 8, RPW = 32 / LPR, UNR = 4;
        const int sub = lane / LPR, li = lane % LPR;
        float qr[G][8];
#pragma unroll
        for (int gg = 0; gg < G; ++gg)
#pragma unroll
            for (int e = 0; e < 8; ++e) qr[gg][e] = vals[gg * DH + li * 8 + e];
        for (int pb = hw * RPW * UNR; pb < Lq; pb += kHalfWarps * RPW * UNR) {
            uint4 u[UNR];
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int p = pb + j * RPW + sub;
                u[j] = make_uint4(0, 0, 0, 0);
                if (p < Lq) u[j] = *reinterpret_cast<const uint4*>(kc + static_cast<size_t>(p) * DH + li * 8);
            }
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int p = pb + j * RPW + sub;
                const bool okp = p < Lq;
                const float kf[8] = {bf_lo(u[j].x), bf_hi(u[j].x), bf_lo(u[j].y), bf_hi(u[j].y),
                                     bf_lo(u[j].z), bf_hi(u[j].z), bf_lo(u[j].w), bf_hi(u[j].w)};
#pragma unroll
                for (int gg = 0; gg < G; ++gg) {
                    float d = 0.f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) d += qr[gg][e] * kf[e];
#pragma unroll
                    for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                    if (okp && li == 0) sc[gg * lcap + p] = A.bf16_math ? rbf(rbf(d) * scale) : d * scale;
                }
            }
        }
        named_bar_sync(bar_id, kHalf);
        for (int gg = hw; gg < G; gg += kHalfWarps) {
            float* s = sc + gg * lcap;
            float m = -INFINITY;
            for (int p = lane; p < Lq; p += 32) m = fmaxf(m, s[p]);
            m = warp_max(m);
            float z = 0.f;
            for (int p = lane; p < Lq; p += 32) {
                const float e = expf(s[p] - m);
                s[p] = e;
                z += e;
            }
            z = warp_sum(z);
            for (int p = lane; p < Lq; p += 32) {
                const float pr = s[p] / z;
                s[p] = A.bf16_math ? rbf(pr) : pr;
            }
        }
        named_bar_sync(bar_id, kHalf);
        constexpr int DPL = DH / 32;
        float acc[G][DPL];
#pragma unroll
        for (int gg = 0; gg < G; ++gg)
#pragma unroll
            for (int e = 0; e < DPL; ++e) acc[gg][e] = 0.f;
        for (int pb = hw; pb < Lq; pb += kHalfWarps * UNR) {
            float vf[UNR][DPL];
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int p = pb + j * kHalfWarps;
#pragma unroll
                for (int e = 0; e < DPL; ++e) vf[j][e] = 0.f;
                if (p < Lq) {
                    if (DPL == 4) {
                        const uint2 u2 = *reinterpret_cast<const uint2*>(vc + static_cast<size_t>(p) * DH + lane * 4);
                        vf[j][0] = bf_lo(u2.x); vf[j][1] = bf_hi(u2.x); vf[j][2 % DPL] = bf_lo(u2.y); vf[j][3 % DPL] = bf_hi(u2.y);
                    } else {
                        const uint32_t u1 = *reinterpret_cast<const uint32_t*>(vc + static_cast<size_t>(p) * DH + lane * 2);
                        vf[j][0] = bf_lo(u1); vf[j][1] = bf_hi(u1);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < UNR; ++j) {
                const int p = pb + j * kHalfWarps;
                if (p < Lq) {
#pragma unroll
                    for (int gg = 0; gg < G; ++gg) {
                        const float w = sc[gg * lcap + p];
#pragma unroll
                        for (int e = 0; e < DPL; ++e) acc[gg][e] += w * vf[j][e];
                    }
                }
            }
        }
#pragma unroll
        for (int gg = 0; gg < G; ++gg)
#pragma unroll
            for (int e = 0; e < DPL; ++e) red[(hw * G + gg) * DH + lane * DPL + e] = acc[gg][e];
        named_bar_sync(bar_id, kHalf);
        for (int e = ht; e < G * DH; e += kHalf) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kHalfWarps; ++w) s += red[w * G * DH + e];
            const int gg = e / DH, d = e - gg * DH;
            A.attn[(static_cast<size_t>(row) * A.H + g * G + gg) * DH + d] = f2bf(s);
        }
        named_bar_sync(bar_id, kHalf);
    }
}

template <int DH, int G>
__global__ void __launch_bounds__(kPkThreads, 1) stack_kernel(const __grid_constant__ PkArgs A) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    Ring R;
    R.stages = A.stages;
    R.tiles = (raw + 1023u) & ~1023u;
    const uint32_t bars = R.tiles + static_cast<uint32_t>(R.stages) * kPkStageBytes;
    R.full0 = bars;
    R.empty0 = bars + 8u * R.stages;
    R.tfull0 = bars + 16u * R.stages;
    R.tempty0 = R.tfull0 + 16u;
    const uint32_t tmem_slot = R.tempty0 + 16u;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
    float* scratch = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16u - raw));  // consumer-phase scratch

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == kPkProducerWarp && lane == 0) {
        for (int s = 0; s < R.stages; ++s) {
            mbar_init(R.full0 + 8u * s, 1);
            mbar_init(R.empty0 + 8u * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(R.tfull0 + 8u * a, 1);
            mbar_init(R.tempty0 + 8u * a, 4);
        }
        fence_mbar_init();
    }
    if (warp == kPkMmaWarp) tmem_alloc(tmem_slot, kPkTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_launch_dependents();

    const unsigned grid = gridDim.x;
    const int ngemm_last = A.kv_only_last ? 1 : 4;
    if (warp == kPkProducerWarp) {
        if (lane == 0) {
            int it = 0;
            unsigned n = 0;  // GEMM index inside the kernel
            for (int l = 0; l < A.nl; ++l) {
                const PkLayer& L = A.layers[l];
                const int ng = (l == A.nl - 1) ? ngemm_last : 4;
                for (int g = 0; g < ng; ++g, ++n) {
                    const PkGemm& G2 = g == 0 ? L.qkv : (g == 1 ? L.wo : (g == 2 ? L.w13 : L.w2));
                    produce_gemm(G2, R, it, A.bar, 2u * n * grid, n == 0, A.l2_prefetch);
                }
            }
        }
    } else if (warp == kPkMmaWarp) {
        if (lane == 0) {
            int it = 0, acc_it = 0;
            for (int l = 0; l < A.nl; ++l) {
                const PkLayer& L = A.layers[l];
                const int ng = (l == A.nl - 1) ? ngemm_last : 4;
                for (int g = 0; g < ng; ++g) {
                    const PkGemm& G2 = g == 0 ? L.qkv : (g == 1 ? L.wo : (g == 2 ? L.w13 : L.w2));
                    mma_gemm(G2, R, it, acc_it, tmem_base);
                }
            }
        }
    } else {
        // ===== workers (warps 0-7); warps 0-3 also run the GEMM epilogues =====
        // The barrier counter is shared with the previous launch (which resets it on exit): under
        // programmatic dependent launch this CTA may be resident before that launch has finished, so the
        // thread that touches the counter first waits for the upstream grids.
        if (threadIdx.x == 0) pdl_wait();
        int acc_it = 0;
        unsigned bidx = 0;
        int tix = 0;
        auto stamp = [&]() {
            if (A.trace && blockIdx.x == 0 && threadIdx.x == 0 && tix < A.trace_max) A.trace[tix] = globaltimer_ns();
            ++tix;
        };
        stamp();
        for (int l = 0; l < A.nl; ++l) {
            const PkLayer& L = A.layers[l];
            const bool last = l == A.nl - 1;
            const int ng = last ? ngemm_last : 4;
            for (int g = 0; g < ng; ++g) {
                const PkGemm& G2 = g == 0 ? L.qkv : (g == 1 ? L.wo : (g == 2 ? L.w13 : L.w2));
                if (warp < 4) epilogue_gemm(G2, R, acc_it, tmem_base, A.ws, A.rows);
                stamp();
                grid_barrier(A.bar, ++bidx * grid);
                stamp();
                if (g == 0) {
                    phase_prep_attn<DH, G>(A, L, scratch, last && A.kv_only_last);
                } else if (g == 1) {
                    phase_resid_norm(A, L.wo, L.bo, L.ffn_norm, scratch);
                } else if (g == 2) {
                    phase_swiglu(A, L.w13);
                } else {
                    phase_resid_norm(A, L.w2, nullptr, L.next_norm, scratch);
                }
                stamp();
                if (!(last && g == ng - 1)) grid_barrier(A.bar, ++bidx * grid);
                stamp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kPkMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kPkTmemCols);
    }
    // the last CTA out resets the barrier counters for the next launch
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned old = atomicAdd(A.bar + 1, 1u);
        if (old == grid - 1) {
            A.bar[0] = 0;
            A.bar[1] = 0;
            __threadfence();
        }
    }
}

// Stand-alone use of the fused q/k/v post-processing + KV append + attention phase (default decode path):
// one launch per layer instead of qkv_prep + attn.
template <int DH, int G>
__global__ void __launch_bounds__(kPkWorkers) prep_attn_kernel(const __grid_constant__ PkArgs A,
                                                               const __grid_constant__ PkLayer L, int kv_only) {
    extern __shared__ uint8_t pa_smem[];
    pdl_launch_dependents();
    pdl_wait();
    phase_prep_attn<DH, G>(A, L, reinterpret_cast<float*>(pa_smem), kv_only != 0);
}

template <int DH, int G>
int launch_pa_t(const PkArgs& A, const PkLayer& L, int kv_only, cudaStream_t st) {
    const size_t smem = pk_scratch_bytes(A.H, A.Hkv, A.Dh, A.S);
    FSB_CHECK(smem <= 200 * 1024, "prep_attn: context %d too long for the shared-memory score buffer", A.S);
    const int items = A.rows * A.Hkv;
    auto k = prep_attn_kernel<DH, G>;
    FSB_LAUNCH(k, dim3((items + 1) / 2), dim3(kPkWorkers), smem, st, A, L, kv_only);
    return 0;
}

template <int DH, int G>
int launch_t(const PkArgs& A, int grid, size_t smem, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        FSB_CUDA(cudaFuncSetAttribute(stack_kernel<DH, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr = true;
    }
    auto k = stack_kernel<DH, G>;
    FSB_LAUNCH(k, dim3(grid), dim3(kPkThreads), smem, st, A);
    return 0;
}

}  // namespace

size_t pk_scratch_bytes(int H, int Hkv, int Dh, int S) {
    const int G = H / Hkv;
    // two half-CTA work areas of the attention phase
    return 2 * (static_cast<size_t>(G + 2) * Dh + 8 + static_cast<size_t>(G) * S + static_cast<size_t>(4) * G * Dh) * sizeof(float) + 64;
}

int pk_init() {
#define FSB_PK_ATTR(DH_, G_)                                                                                       \
    FSB_CUDA(cudaFuncSetAttribute(prep_attn_kernel<DH_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
    FSB_CUDA(cudaFuncSetAttribute(stack_kernel<DH_, G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    FSB_PK_ATTR(128, 4) FSB_PK_ATTR(128, 1) FSB_PK_ATTR(128, 2) FSB_PK_ATTR(64, 4) FSB_PK_ATTR(64, 1) FSB_PK_ATTR(64, 2)
#undef FSB_PK_ATTR
    return 0;
}

int launch_prep_attn(const PkArgs& A, const PkLayer& L, int kv_only, cudaStream_t st) {
    if (A.rows <= 0) return 0;
    const int G = A.H / A.Hkv;
#define FSB_PA_CASE(DH_, G_) \
    if (A.Dh == DH_ && G == G_) return launch_pa_t<DH_, G_>(A, L, kv_only, st);
    FSB_PA_CASE(128, 4) FSB_PA_CASE(128, 1) FSB_PA_CASE(128, 2) FSB_PA_CASE(64, 4) FSB_PA_CASE(64, 1) FSB_PA_CASE(64, 2)
#undef FSB_PA_CASE
    return -1;  // unsupported shape: caller falls back to the two-kernel sequence
}

int launch_stack_persistent(const PkArgs& A, int grid, cudaStream_t st) {
    const size_t scratch = pk_scratch_bytes(A.H, A.Hkv, A.Dh, A.S);
    const size_t smem = static_cast<size_t>(A.stages) * kPkStageBytes + 1024 + 16 * A.stages + 64 + scratch;
    FSB_CHECK(smem <= 227 * 1024, "persistent stack kernel: %zu bytes of shared memory needed (kv_len too large)", smem);
    const int G = A.H / A.Hkv;
#define FSB_PK_CASE(DH_, G_) \
    if (A.Dh == DH_ && G == G_) return launch_t<DH_, G_>(A, grid, smem, st);
    FSB_PK_CASE(128, 4) FSB_PK_CASE(128, 1) FSB_PK_CASE(128, 2) FSB_PK_CASE(64, 4) FSB_PK_CASE(64, 1) FSB_PK_CASE(64, 2)
#undef FSB_PK_CASE
    set_error("persistent stack kernel: unsupported head_dim=%d group=%d", A.Dh, G);
    return 1;
}

}  // namespace fsb
