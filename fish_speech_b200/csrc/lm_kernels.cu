// See lm_kernels.cuh.  All kernels here are HBM/L2-bound glue between the GEMMs: vectorised,
// coalesced, warp-shuffle reductions, fp32 math with bf16 rounding at the reference's rounding points.
#include "lm_kernels.cuh"

namespace fsb {

namespace {

// ------------------------------------------------------------------------------------------------
// embed: llama.py:399-420
// ------------------------------------------------------------------------------------------------
__global__ void embed_kernel(EmbedArgs a, float inv_div) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    const int* t = a.tokens + static_cast<size_t>(row) * (a.C + 1);
    int tok = t[0];
    const bool sem = tok >= a.sem_begin && tok <= a.sem_end;
    tok = min(max(tok, 0), a.vocab - 1);
    for (int d = threadIdx.x; d < a.D; d += blockDim.x) {
        float s = 0.f;
        if (sem) {
            for (int c = 0; c < a.C; ++c) {
                int code = min(max(t[c + 1], 0), a.cs - 1);
                s += bf2f(a.cb_emb[(static_cast<size_t>(c) * a.cs + code) * a.D + d]);
            }
            s = rbf(s);  // torch.stack(...).sum(dim=1) -> one bf16 rounding
        }
        float x = rbf(bf2f(a.emb[static_cast<size_t>(tok) * a.D + d]) + s);
        if (sem && a.scale) x = rbf(x / inv_div);  // x / sqrt(C+1)
        a.x[static_cast<size_t>(row) * a.D + d] = f2bf(x);
    }
}

// ------------------------------------------------------------------------------------------------
// residual add + fish RMSNorm
// ------------------------------------------------------------------------------------------------
constexpr int kRnThreads = 512;
constexpr int kRnMaxPer = 8;  // D <= 4096

// kFlags = device-side dependency flags (opt-in experiment). The default instantiation keeps exactly the
// grid-dependency code: the flag plumbing changed the register allocation of this kernel (128 -> 99, fewer
// partial-sum loads in flight) and cost 0.7 ms per frame even when unused.
template <bool kFlags>
__global__ void __launch_bounds__(kRnThreads) resid_norm_kernel(ResidNormCore a, int gstride, DepFlag wait,
                                                                unsigned* done_ctr) {
    pdl_launch_dependents();
    if constexpr (kFlags) dep_wait_cta(wait);
    else pdl_wait();
    __shared__ float red[33];
    const int row = blockIdx.x;
    const int grow = a.gather_map ? a.gather_map[row] : row;
    const int src = a.gather ? a.gather[static_cast<size_t>(grow) * gstride] : row;
    float v[kRnMaxPer], y[kRnMaxPer];
    int rws[kRnMaxPer], fts[kRnMaxPer];
    bool ok[kRnMaxPer];
#pragma unroll
    for (int e = 0; e < kRnMaxPer; ++e) {
        fts[e] = threadIdx.x + e * kRnThreads;
        rws[e] = row;
        ok[e] = fts[e] < a.D;
        v[e] = (ok[e] && a.x_in) ? bf2f(a.x_in[static_cast<size_t>(src) * a.D + fts[e]]) : 0.f;
        y[e] = 0.f;
    }
    if (a.parts.ws) sum_parts_n<kRnMaxPer>(a.parts, rws, fts, ok, y, a.parts.max_parts > 0 ? a.parts.max_parts : 1);
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < kRnMaxPer; ++e) {
        if (ok[e]) {
            float x = v[e];
            if (a.parts.ws) {
                float yy = y[e];
                if (a.bias) yy += bf2f(a.bias[fts[e]]);
                yy = rbf(yy);
                if (a.scale) yy *= bf2f(a.scale[fts[e]]);
                x = rbf(x + yy);
            }
            v[e] = x;
            ss += x * x;
            if (a.x_out) a.x_out[static_cast<size_t>(row) * a.D + fts[e]] = f2bf(x);
        }
    }
    if (a.norm_w != nullptr) {
        const float tot = block_sum(ss, red);
        const float r = rsqrtf(tot / static_cast<float>(a.D) + a.eps);
#pragma unroll
        for (int e = 0; e < kRnMaxPer; ++e) {
            if (ok[e]) {
                const float n = rbf(rbf(v[e] * r) * bf2f(a.norm_w[fts[e]]));
                a.n_out[static_cast<size_t>(row) * a.D + fts[e]] = f2bf(n);
            }
        }
    }
    if constexpr (kFlags) dep_signal_cta(done_ctr);
}

// Vectorised form (D % 4 == 0): a thread owns groups of 4 consecutive features, so one 16-byte load fetches a
// slot's partials for all four and the slot loop issues 4x fewer instructions (the scalar kernel executes ~1400
// instructions per warp for 80 useful loads and is issue/latency bound at 16 warps per SM, ncu). The per-feature
// additions are in the same slot order as sum_parts(): identical partial sums.
constexpr int kRn4MaxThreads = 640;
template <bool kFlags, int GP, int U>
__global__ void __launch_bounds__(kRn4MaxThreads) resid_norm4_kernel(ResidNormCore a, int gstride, DepFlag wait,
                                                                      unsigned* done_ctr) {
    pdl_launch_dependents();
    if constexpr (kFlags) dep_wait_cta(wait);
    else pdl_wait();
    __shared__ float red[33];
    const int row = blockIdx.x;
    const int grow = a.gather_map ? a.gather_map[row] : row;
    const int src = a.gather ? a.gather[static_cast<size_t>(grow) * gstride] : row;
    const int G = a.D >> 2;
    int f[GP];
    bool ok[GP];
    float4 y[GP];
#pragma unroll
    for (int e = 0; e < GP; ++e) {
        const int g = threadIdx.x + e * blockDim.x;
        ok[e] = g < G;
        f[e] = ok[e] ? g * 4 : 0;
        y[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (a.parts.ws) {
        int np[GP];
        const float4* p[GP];
        const size_t ss4 = static_cast<size_t>(a.parts.slot_stride) >> 2;
        const int maxp = a.parts.max_parts > 0 ? a.parts.max_parts : 1;
#pragma unroll
        for (int e = 0; e < GP; ++e) {
            np[e] = ok[e] ? (a.parts.nparts ? __ldg(a.parts.nparts + (f[e] >> 7)) : 1) : 0;
            p[e] = reinterpret_cast<const float4*>(a.parts.ws + static_cast<size_t>(row) * a.parts.ld + f[e]);
        }
        for (int q = 0; q < maxp; q += U) {
            float4 t[GP][U];
#pragma unroll
            for (int e = 0; e < GP; ++e)
#pragma unroll
                for (int u = 0; u < U; ++u)
                    t[e][u] = (q + u < np[e]) ? p[e][static_cast<size_t>(q + u) * ss4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int e = 0; e < GP; ++e) {
                if (q < np[e]) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        y[e].x += t[e][u].x;
                        y[e].y += t[e][u].y;
                        y[e].z += t[e][u].z;
                        y[e].w += t[e][u].w;
                    }
                }
            }
        }
    }
    float v[GP][4];
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < GP; ++e) {
        if (!ok[e]) continue;
        float x[4] = {0.f, 0.f, 0.f, 0.f};
        if (a.x_in) {
            const uint2 u = *reinterpret_cast<const uint2*>(a.x_in + static_cast<size_t>(src) * a.D + f[e]);
            x[0] = bf_lo(u.x); x[1] = bf_hi(u.x); x[2] = bf_lo(u.y); x[3] = bf_hi(u.y);
        }
        if (a.parts.ws) {
            float yy[4] = {y[e].x, y[e].y, y[e].z, y[e].w};
            if (a.bias) {
                const uint2 b = *reinterpret_cast<const uint2*>(a.bias + f[e]);
                yy[0] += bf_lo(b.x); yy[1] += bf_hi(b.x); yy[2] += bf_lo(b.y); yy[3] += bf_hi(b.y);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) yy[c] = rbf(yy[c]);
            if (a.scale) {
                const uint2 sc = *reinterpret_cast<const uint2*>(a.scale + f[e]);
                yy[0] *= bf_lo(sc.x); yy[1] *= bf_hi(sc.x); yy[2] *= bf_lo(sc.y); yy[3] *= bf_hi(sc.y);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) x[c] = rbf(x[c] + yy[c]);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            v[e][c] = x[c];
            ss += x[c] * x[c];
        }
        if (a.x_out) {
            uint2 o;
            o.x = pack_bf2(x[0], x[1]);
            o.y = pack_bf2(x[2], x[3]);
            *reinterpret_cast<uint2*>(a.x_out + static_cast<size_t>(row) * a.D + f[e]) = o;
        }
    }
    if (a.norm_w != nullptr) {
        const float tot = block_sum(ss, red);
        const float r = rsqrtf(tot / static_cast<float>(a.D) + a.eps);
#pragma unroll
        for (int e = 0; e < GP; ++e) {
            if (!ok[e]) continue;
            const uint2 w = *reinterpret_cast<const uint2*>(a.norm_w + f[e]);
            const float wf[4] = {bf_lo(w.x), bf_hi(w.x), bf_lo(w.y), bf_hi(w.y)};
            float n[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) n[c] = rbf(rbf(v[e][c] * r) * wf[c]);
            uint2 o;
            o.x = pack_bf2(n[0], n[1]);
            o.y = pack_bf2(n[2], n[3]);
            *reinterpret_cast<uint2*>(a.n_out + static_cast<size_t>(row) * a.D + f[e]) = o;
        }
    }
    if constexpr (kFlags) dep_signal_cta(done_ctr);
}

__global__ void linear_out_kernel(LinearOutArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.N) return;
    float y = sum_parts(a.parts, row, i);
    if (a.bias) y += bf2f(a.bias[i]);
    a.y[static_cast<size_t>(row) * a.N + i] = f2bf(y);
}

// ------------------------------------------------------------------------------------------------
// q/k/v post-processing: llama.py:891-911
// ------------------------------------------------------------------------------------------------
template <bool kFlags>
__global__ void qkv_prep_kernel(QkvPrepCore a, DepFlag wait) {
    pdl_launch_dependents();
    if constexpr (kFlags) dep_wait_cta(wait);
    else pdl_wait();
    __shared__ float red[33];
    const int row = blockIdx.x, head = blockIdx.y;
    const int t = threadIdx.x;  // pair index, Dh/2 threads
    const int Dh = a.Dh;
    const int kind = head < a.H ? 0 : (head < a.H + a.Hkv ? 1 : 2);  // q, k, v
    const int f0 = head * Dh + 2 * t;
    float v0, v1;
    {
        const int rws[2] = {row, row}, fts[2] = {f0, f0 + 1};
        const bool ok[2] = {true, true};
        float y[2];
        sum_parts_n<2>(a.parts, rws, fts, ok, y, a.parts.max_parts > 0 ? a.parts.max_parts : 1);
        v0 = y[0];
        v1 = y[1];
    }
    if (a.bias) {
        v0 += bf2f(a.bias[f0]);
        v1 += bf2f(a.bias[f0 + 1]);
    }
    v0 = rbf(v0);
    v1 = rbf(v1);
    const __nv_bfloat16* nw = kind == 0 ? a.q_norm : (kind == 1 ? a.k_norm : nullptr);
    if (nw != nullptr) {
        // nn.RMSNorm(head_dim): fp32 math, weight multiply included, ONE rounding
        const float tot = block_sum(v0 * v0 + v1 * v1, red);
        const float r = rsqrtf(tot / static_cast<float>(Dh) + a.eps);
        v0 = rbf(v0 * r * bf2f(nw[2 * t]));
        v1 = rbf(v1 * r * bf2f(nw[2 * t + 1]));
    }
    const int pos = a.row_pos[row];
    if (kind != 2) {
        const __nv_bfloat16* f = a.freqs + (static_cast<size_t>(pos) * (Dh / 2) + t) * 2;
        const float c = bf2f(f[0]), s = bf2f(f[1]);
        const float o0 = __fsub_rn(__fmul_rn(v0, c), __fmul_rn(v1, s));
        const float o1 = __fadd_rn(__fmul_rn(v1, c), __fmul_rn(v0, s));
        v0 = rbf(o0);
        v1 = rbf(o1);
    }
    const uint32_t packed = pack_bf2(v0, v1);
    if (kind == 0) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(a.q + (static_cast<size_t>(row) * a.H + head) * Dh);
        dst[t] = packed;
    } else {
        const int g = kind == 1 ? head - a.H : head - a.H - a.Hkv;
        __nv_bfloat16* cache = kind == 1 ? a.kcache : a.vcache;
        const int b = a.row_seq[row];
        uint32_t* dst = reinterpret_cast<uint32_t*>(
            cache + ((static_cast<size_t>(b) * a.Hkv + g) * a.S + pos) * Dh);
        dst[t] = packed;
    }
}

// ------------------------------------------------------------------------------------------------
// length-aware GQA attention over the KV cache (one query token per CTA, all G heads of a KV group)
// ------------------------------------------------------------------------------------------------
constexpr int kAttnThreads = 256;
constexpr int kAttnWarps = kAttnThreads / 32;

template <int DH, int G, bool kFlags>
__global__ void __launch_bounds__(kAttnThreads) attn_kernel(AttnCore a, float scale, int lcap, unsigned* done_ctr) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float sm[];
    float* qs = sm;                      // [G][DH]
    float* sc = qs + G * DH;             // [G][lcap]
    float* red = sc + G * lcap;          // [kAttnWarps][G][DH]
    __shared__ float red2[33];
    const int row = blockIdx.y, g = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = a.row_seq[row], pos = a.row_pos[row];
    const int lo = (a.window > 0 && pos - a.window + 1 > 0) ? pos - a.window + 1 : 0;
    const int L = pos - lo + 1;
    const size_t cache_base = ((static_cast<size_t>(b) * a.Hkv + g) * a.S + lo) * DH;
    const __nv_bfloat16* kc = a.kcache + cache_base;
    const __nv_bfloat16* vc = a.vcache + cache_base;

    for (int e = threadIdx.x; e < G * DH; e += kAttnThreads) {
        const int gg = e / DH, d = e - gg * DH;
        qs[e] = bf2f(a.q[(static_cast<size_t>(row) * a.H + g * G + gg) * DH + d]);
    }
    __syncthreads();

    // ---- scores ----
    constexpr int LPR = DH / 8;    // lanes per cache row (16-byte loads)
    constexpr int RPW = 32 / LPR;  // rows per warp iteration
    const int sub = lane / LPR, li = lane % LPR;
    float qr[G][8];
#pragma unroll
    for (int gg = 0; gg < G; ++gg)
#pragma unroll
        for (int e = 0; e < 8; ++e) qr[gg][e] = qs[gg * DH + li * 8 + e];
    constexpr int UNR = 4;  // independent 16-byte loads in flight per lane
    for (int pb = warp * RPW * UNR; pb < L; pb += kAttnWarps * RPW * UNR) {
        uint4 u[UNR];
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            const int p = pb + j * RPW + sub;
            u[j] = make_uint4(0, 0, 0, 0);
            if (p < L) u[j] = *reinterpret_cast<const uint4*>(kc + static_cast<size_t>(p) * DH + li * 8);
        }
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            const int p = pb + j * RPW + sub;
            const bool ok = p < L;
            const float kf[8] = {bf_lo(u[j].x), bf_hi(u[j].x), bf_lo(u[j].y), bf_hi(u[j].y),
                                 bf_lo(u[j].z), bf_hi(u[j].z), bf_lo(u[j].w), bf_hi(u[j].w)};
#pragma unroll
            for (int gg = 0; gg < G; ++gg) {
                float d = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) d += qr[gg][e] * kf[e];
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                if (ok && li == 0) sc[gg * lcap + p] = a.bf16_math ? rbf(rbf(d) * scale) : d * scale;
            }
        }
    }
    __syncthreads();

    // ---- softmax (warp per head) ----
    for (int gg = warp; gg < G; gg += kAttnWarps) {
        float* s = sc + gg * lcap;
        float m = -INFINITY;
        for (int p = lane; p < L; p += 32) m = fmaxf(m, s[p]);
        m = warp_max(m);
        float z = 0.f;
        for (int p = lane; p < L; p += 32) {
            const float e = expf(s[p] - m);
            s[p] = e;
            z += e;
        }
        z = warp_sum(z);
        for (int p = lane; p < L; p += 32) {
            const float pr = s[p] / z;
            s[p] = a.bf16_math ? rbf(pr) : pr;
        }
    }
    __syncthreads();
    (void)red2;

    // ---- P.V ----
    constexpr int DPL = DH / 32;  // dims per lane
    float acc[G][DPL];
#pragma unroll
    for (int gg = 0; gg < G; ++gg)
#pragma unroll
        for (int e = 0; e < DPL; ++e) acc[gg][e] = 0.f;
    for (int pb = warp; pb < L; pb += kAttnWarps * UNR) {
        float vf[UNR][DPL];
#pragma unroll
        for (int j = 0; j < UNR; ++j) {
            const int p = pb + j * kAttnWarps;
#pragma unroll
            for (int e = 0; e < DPL; ++e) vf[j][e] = 0.f;
            if (p < L) {
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
            const int p = pb + j * kAttnWarps;
            if (p < L) {
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
        for (int e = 0; e < DPL; ++e) red[(warp * G + gg) * DH + lane * DPL + e] = acc[gg][e];
    __syncthreads();
    for (int e = threadIdx.x; e < G * DH; e += kAttnThreads) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kAttnWarps; ++w) s += red[w * G * DH + e];
        const int gg = e / DH, d = e - gg * DH;
        a.out[(static_cast<size_t>(row) * a.H + g * G + gg) * DH + d] = f2bf(s);
    }
    if constexpr (kFlags) dep_signal_cta(done_ctr);
}

template <bool kFlags>
__global__ void swiglu_kernel(SwigluCore a, DepFlag wait, unsigned* done_ctr) {
    pdl_launch_dependents();
    if constexpr (kFlags) dep_wait_cta(wait);
    else pdl_wait();
    const int row = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.I) {
        const int rws[2] = {row, row}, fts[2] = {i, a.I + i};
        const bool ok[2] = {true, true};
        float y[2];
        sum_parts_n<2>(a.parts, rws, fts, ok, y, a.parts.max_parts > 0 ? a.parts.max_parts : 1);
        const float g = rbf(y[0]), c = rbf(y[1]);
        const float s = rbf(g / (1.f + expf(-g)));
        a.h[static_cast<size_t>(row) * a.I + i] = f2bf(s * c);
    }
    if constexpr (kFlags) dep_signal_cta(done_ctr);
}

// Vectorised SwiGLU (I % 4 == 0): 4 consecutive features per thread, 16-byte partial loads.
template <bool kFlags>
__global__ void swiglu4_kernel(SwigluCore a, DepFlag wait, unsigned* done_ctr) {
    pdl_launch_dependents();
    if constexpr (kFlags) dep_wait_cta(wait);
    else pdl_wait();
    const int row = blockIdx.y;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i < a.I) {
        const int maxp = a.parts.max_parts > 0 ? a.parts.max_parts : 1;
        const int npg = a.parts.nparts ? __ldg(a.parts.nparts + (i >> 7)) : 1;
        const int npu = a.parts.nparts ? __ldg(a.parts.nparts + ((a.I + i) >> 7)) : 1;
        const float4* pg = reinterpret_cast<const float4*>(a.parts.ws + static_cast<size_t>(row) * a.parts.ld + i);
        const float4* pu = reinterpret_cast<const float4*>(a.parts.ws + static_cast<size_t>(row) * a.parts.ld + a.I + i);
        const size_t ss4 = static_cast<size_t>(a.parts.slot_stride) >> 2;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f), u = g;
        const float4 z = g;
        for (int q = 0; q < maxp; q += 4) {
            float4 tg[4], tu[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                tg[k] = (q + k < npg) ? pg[static_cast<size_t>(q + k) * ss4] : z;
                tu[k] = (q + k < npu) ? pu[static_cast<size_t>(q + k) * ss4] : z;
            }
            if (q < npg) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { g.x += tg[k].x; g.y += tg[k].y; g.z += tg[k].z; g.w += tg[k].w; }
            }
            if (q < npu) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { u.x += tu[k].x; u.y += tu[k].y; u.z += tu[k].z; u.w += tu[k].w; }
            }
        }
        const float gv[4] = {g.x, g.y, g.z, g.w}, uv[4] = {u.x, u.y, u.z, u.w};
        float h[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float gg = rbf(gv[c]), cc = rbf(uv[c]);
            const float sl = rbf(gg / (1.f + expf(-gg)));
            h[c] = sl * cc;
        }
        uint2 o;
        o.x = pack_bf2(h[0], h[1]);
        o.y = pack_bf2(h[2], h[3]);
        *reinterpret_cast<uint2*>(a.h + static_cast<size_t>(row) * a.I + i) = o;
    }
    if constexpr (kFlags) dep_signal_cta(done_ctr);
}

// ------------------------------------------------------------------------------------------------
// sampling: inference.py:43-93 (logits_to_probs / multinomial_sample_one_no_sync / sample) and the
// slow-token RAS rule inference.py:114-144.
// ------------------------------------------------------------------------------------------------
constexpr int kSampleThreads = 1024;
constexpr int kSampleMaxN = 8192;
constexpr int kSelCap = 256;  // ranks materialised for top-k / top-p

__device__ __forceinline__ void philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                           uint32_t c2, uint32_t c3, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct ArgMax {
    float v;
    int i;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
    // larger value wins; ties -> smaller index
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax x, ArgMax* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax y;
        y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
        y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
        x = better(x, y);
    }
    __syncthreads();
    if (lane == 0) red[w] = x;
    __syncthreads();
    if (w == 0) {
        ArgMax t = red[lane];  // kSampleThreads/32 == 32 warps
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ArgMax y;
            y.v = __shfl_xor_sync(0xffffffffu, t.v, o);
            y.i = __shfl_xor_sync(0xffffffffu, t.i, o);
            t = better(t, y);
        }
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

template <bool kFlags>
__global__ void __launch_bounds__(kSampleThreads) sample_kernel(SampleCore a, DepFlag wait, SlotCtl ctl) {
    pdl_launch_dependents();
    if constexpr (kFlags) dep_wait_cta(wait);
    else pdl_wait();
    __shared__ float lg[kSampleMaxN];
    __shared__ ArgMax red[33];
    __shared__ float fred[33];
    __shared__ float sel_v[kSelCap];
    __shared__ int sel_i[kSelCap];
    __shared__ float sel_cum[kSelCap];
    __shared__ int s_nsel;
    __shared__ int s_choice[2];
    const int row = blockIdx.x;
    const int slot = a.row_slot ? a.row_slot[row] : row;
    const int n = a.n;
    if (!slot_live(ctl, slot)) return;  // idle / frozen slot: nothing is sampled, its state stays as it is
    const bool per_slot = ctl.state != nullptr;
    const float temperature = per_slot ? ctl.temperature[slot] : a.temperature;
    const float top_p = per_slot ? ctl.top_p[slot] : a.top_p;
    const int top_k = per_slot ? ctl.top_k[slot] : a.top_k;
    const unsigned long long seed = per_slot ? ctl.seed[slot] : a.seed;
    for (int e = threadIdx.x; e < n; e += kSampleThreads) {
        const float v = rbf(sum_parts(a.parts, row, e));
        lg[e] = v;
        if (a.logits_out) a.logits_out[static_cast<size_t>(slot) * n + e] = v;
    }
    __syncthreads();

    const bool two = a.slow && a.use_ras && top_k != 1;
    if (top_k == 1) {
        ArgMax x{-INFINITY, 0x7fffffff};
        for (int e = threadIdx.x; e < n; e += kSampleThreads) x = better(x, ArgMax{lg[e], e});
        x = block_argmax(x, red);
        if (threadIdx.x == 0) s_choice[0] = s_choice[1] = x.i;
        __syncthreads();
    } else {
        // softmax denominator over every candidate (the -inf-biased vocabulary contributes 0)
        float m = -INFINITY;
        for (int e = threadIdx.x; e < n; e += kSampleThreads) m = fmaxf(m, lg[e]);
        m = block_max(m, fred);
        float z = 0.f;
        for (int e = threadIdx.x; e < n; e += kSampleThreads) z += expf(lg[e] - m);
        z = block_sum(z, fred);
        // descending ranks until neither criterion can keep anything further
        const float p_lim = two ? fmaxf(top_p, 0.9f) : top_p;
        int kcap = top_k < n ? top_k : n;
        if (kcap > kSelCap) kcap = kSelCap;
        float cum = 0.f;
        int nsel = 0;
        for (int r = 0; r < kcap; ++r) {
            ArgMax x{-INFINITY, 0x7fffffff};
            for (int e = threadIdx.x; e < n; e += kSampleThreads) x = better(x, ArgMax{lg[e], e});
            x = block_argmax(x, red);
            const float pr = rbf(expf(x.v - m) / z);
            cum = rbf(cum + pr);
            if (threadIdx.x == 0) {
                sel_v[r] = x.v;
                sel_i[r] = x.i;
                sel_cum[r] = cum;
                lg[x.i] = -INFINITY;
            }
            nsel = r + 1;
            __syncthreads();
            if (cum > p_lim) break;
        }
        if (threadIdx.x == 0) s_nsel = nsel;
        __syncthreads();
        // the two draws (normal, RAS high-temperature); warp 0 / warp 1
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (w < (two ? 2 : 1)) {
            const float T = w == 0 ? temperature : 1.0f;
            const float tp = w == 0 ? top_p : 0.9f;
            const float Tc = fmaxf(T, 1e-5f);
            // survivors: rank 0 always; rank r kept iff cum[r] <= top_p (and r < top_k, implied)
            int ns = 1;
            while (ns < nsel && !(sel_cum[ns] > tp)) ++ns;
            float mx = -INFINITY;
            for (int r = lane; r < ns; r += 32) mx = fmaxf(mx, rbf(sel_v[r] / Tc));
            mx = warp_max(mx);
            float zz = 0.f;
            for (int r = lane; r < ns; r += 32) zz += expf(rbf(sel_v[r] / Tc) - mx);
            zz = warp_sum(zz);
            // RNG stream: per call (seed, global frame counter, slot) or, with slot control, per request
            // (its own seed and frame index, no slot) so that the draw does not depend on the schedule
            const unsigned long long off =
                per_slot ? static_cast<unsigned long long>(ctl.n_out[slot]) : (a.rng_offset ? *a.rng_offset : 0ull);
            const uint32_t lane_id = per_slot ? 0u : static_cast<uint32_t>(slot);
            ArgMax best{-INFINITY, 0x7fffffff};
            for (int r = lane; r < ns; r += 32) {
                const float pr = rbf(expf(rbf(sel_v[r] / Tc) - mx) / zz);
                uint32_t rnd[4];
                philox4x32(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32),
                           static_cast<uint32_t>(off), static_cast<uint32_t>(off >> 32), lane_id,
                           static_cast<uint32_t>((a.draw_id * 2 + w) * kSelCap + r), rnd);
                const float u = (static_cast<float>(rnd[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
                const float q = -logf(u);
                best = better(best, ArgMax{pr / q, r});
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ArgMax y;
                y.v = __shfl_xor_sync(0xffffffffu, best.v, o);
                y.i = __shfl_xor_sync(0xffffffffu, best.i, o);
                best = better(best, y);
            }
            if (lane == 0) s_choice[w] = sel_i[best.i];
        }
        __syncthreads();
        if (!two && threadIdx.x == 0) s_choice[1] = s_choice[0];
        __syncthreads();
    }

    if (threadIdx.x == 0) {
        int* ct = a.cur_tok + static_cast<size_t>(slot) * (a.num_cb + 1);
        if (a.slow) {
            int e = s_choice[0];
            int tok = e < a.n_sem ? a.sem_begin + e : a.im_end_id;
            if (a.use_ras) {
                const int eh = s_choice[1];
                const int tok_h = eh < a.n_sem ? a.sem_begin + eh : a.im_end_id;
                bool in_win = false;
                for (int k = 0; k < 10; ++k) in_win |= (a.ras_window[slot * 10 + k] == tok);
                const bool is_sem = tok >= a.sem_begin && tok <= a.sem_begin + a.n_sem - 1;
                if (in_win && is_sem) tok = tok_h;
            }
            int c0 = tok - a.sem_begin;
            c0 = c0 < 0 ? 0 : (c0 > a.codebook_size - 1 ? a.codebook_size - 1 : c0);
            ct[0] = tok;
            ct[1] = c0;
            if (a.ras_update) {
                for (int k = 0; k < 9; ++k) a.ras_window[slot * 10 + k] = a.ras_window[slot * 10 + k + 1];
                a.ras_window[slot * 10 + 9] = tok;
            }
            if (a.finished && tok == a.im_end_id) a.finished[slot] = 1;
            // the reference's loop tests <|im_end|> from the second frame on (the prefill's token is not
            // tested, inference.py:336-352 then :233)
            if (per_slot && tok == a.im_end_id && ctl.n_out[slot] >= 1) ctl.state[slot] = 3;
        } else {
            ct[a.cb_index + 1] = s_choice[0];
        }
    }
}

__global__ void frame_end_kernel(FrameEndCore a, SlotCtl ctl) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    const int slot = a.row_slot ? a.row_slot[row] : row;
    if (!slot_live(ctl, slot)) return;
    const int f = a.n_out[slot];
    if (threadIdx.x < a.ncols && f < a.T_cap)
        a.out_tokens[(static_cast<size_t>(slot) * a.ncols + threadIdx.x) * a.T_cap + f] =
            a.cur_tok[slot * a.ncols + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        a.n_out[slot] = f + 1;
        bool advance = true;
        if (ctl.state != nullptr && (ctl.state[slot] == 3 || f + 1 >= ctl.limit[slot])) {
            ctl.state[slot] = 2;  // frozen: position and counters stay at the last frame
            advance = false;
        }
        if (a.set_pos_rows)
            a.pos[slot] = a.row_pos_src[a.set_pos_rows[row]] + (advance ? 1 : 0);
        else if (advance)
            a.pos[slot] = a.pos[slot] + 1;
    }
}
__global__ void step_inc_kernel(unsigned long long* step) {     pdl_launch_dependents();
    pdl_wait();
*step += 1; }

__global__ void gather_rows_kernel(const __nv_bfloat16* src, const int* idx, __nv_bfloat16* dst, int D) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x;
    const size_t s = static_cast<size_t>(idx[row]) * D, d = static_cast<size_t>(row) * D;
    for (int e = threadIdx.x; e < D; e += blockDim.x) dst[d + e] = src[s + e];
}

}  // namespace

int launch_embed(const EmbedArgs& a, cudaStream_t st) {
    if (a.rows <= 0) return 0;
    FSB_LAUNCH(embed_kernel, dim3(a.rows), dim3(256), 0, st, a, sqrtf(static_cast<float>(a.C + 1)));
    return 0;
}

int launch_resid_norm(const ResidNormArgs& a, cudaStream_t st) { return launch_resid_norm_g(a, 1, st); }

static int rn_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FSB_RN_VARIANT");
        v = e ? atoi(e) : 2;  // 0: scalar kernel, 1/2/3: vectorised with 8/16/4 slots per round (6.21 / 5.80 / 5.76 / 6.11 ms per frame)
    }
    return v;
}

int launch_resid_norm_g(const ResidNormArgs& a, int gather_stride, cudaStream_t st) {
    if (a.rows <= 0) return 0;
    FSB_CHECK(a.D <= kRnThreads * kRnMaxPer, "resid_norm: D=%d too large", a.D);
    const bool fl = a.wait.ctr != nullptr || a.done_ctr != nullptr;
    const ResidNormCore core = static_cast<ResidNormCore>(a);
    const int var = rn_variant();
    const bool vec_ok = var > 0 && (a.D & 3) == 0 && (!a.parts.ws || ((a.parts.ld & 3) == 0 && (a.parts.slot_stride & 3) == 0));
    if (vec_ok) {
        const int G = a.D >> 2;
#define FSB_RN4(GP_, U_, T_)                                                                                        \
    do {                                                                                                            \
        if (fl)                                                                                                     \
            FSB_LAUNCH((resid_norm4_kernel<true, GP_, U_>), dim3(a.rows), dim3(T_), 0, st, core, gather_stride,     \
                       a.wait, a.done_ctr);                                                                         \
        else                                                                                                        \
            FSB_LAUNCH((resid_norm4_kernel<false, GP_, U_>), dim3(a.rows), dim3(T_), 0, st, core, gather_stride,    \
                       a.wait, a.done_ctr);                                                                         \
    } while (0)
        if (G <= kRn4MaxThreads) {
            const int T = ((G + 31) / 32) * 32;
            if (var == 2) FSB_RN4(1, 16, T);
            else if (var == 3) FSB_RN4(1, 4, T);
            else FSB_RN4(1, 8, T);
        } else {
            FSB_RN4(2, 4, 512);
        }
#undef FSB_RN4
        return 0;
    }
    if (fl)
        FSB_LAUNCH(resid_norm_kernel<true>, dim3(a.rows), dim3(kRnThreads), 0, st, core, gather_stride, a.wait, a.done_ctr);
    else
        FSB_LAUNCH(resid_norm_kernel<false>, dim3(a.rows), dim3(kRnThreads), 0, st, core, gather_stride, a.wait, a.done_ctr);
    return 0;
}

int launch_linear_out(const LinearOutArgs& a, cudaStream_t st) {
    if (a.rows <= 0) return 0;
    FSB_LAUNCH(linear_out_kernel, dim3(cdiv(a.N, 256), a.rows), dim3(256), 0, st, a);
    return 0;
}

int launch_qkv_prep(const QkvPrepArgs& a, cudaStream_t st) {
    if (a.rows <= 0) return 0;
    FSB_CHECK(a.Dh % 64 == 0 && a.Dh <= 256, "qkv_prep: head_dim %d unsupported", a.Dh);
    if (a.wait.ctr != nullptr)
        FSB_LAUNCH(qkv_prep_kernel<true>, dim3(a.rows, a.H + 2 * a.Hkv), dim3(a.Dh / 2), 0, st,
                   static_cast<QkvPrepCore>(a), a.wait);
    else
        FSB_LAUNCH(qkv_prep_kernel<false>, dim3(a.rows, a.H + 2 * a.Hkv), dim3(a.Dh / 2), 0, st,
                   static_cast<QkvPrepCore>(a), a.wait);
    return 0;
}

template <int DH, int G>
static int launch_attn_t(const AttnArgs& a, cudaStream_t st) {
    int lcap = a.window > 0 && a.window < a.S ? a.window : a.S;
    if (a.lcap > 0 && a.lcap < lcap) lcap = a.lcap;
    const size_t smem = (static_cast<size_t>(G) * DH + static_cast<size_t>(G) * lcap +
                         static_cast<size_t>(kAttnWarps) * G * DH) * sizeof(float);
    FSB_CHECK(smem <= 200 * 1024, "attention: context %d too long for the shared-memory score buffer", lcap);
    const float scale = 1.0f / sqrtf(static_cast<float>(DH));
    if (a.done_ctr != nullptr)
        FSB_LAUNCH((attn_kernel<DH, G, true>), dim3(a.Hkv, a.rows), dim3(kAttnThreads), smem, st,
                   static_cast<AttnCore>(a), scale, lcap, a.done_ctr);
    else
        FSB_LAUNCH((attn_kernel<DH, G, false>), dim3(a.Hkv, a.rows), dim3(kAttnThreads), smem, st,
                   static_cast<AttnCore>(a), scale, lcap, a.done_ctr);
    return 0;
}

int attn_init() {
#define FSB_ATTN_ATTR(DH_, G_)                                                                                       \
    FSB_CUDA(cudaFuncSetAttribute(attn_kernel<DH_, G_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
    FSB_CUDA(cudaFuncSetAttribute(attn_kernel<DH_, G_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    FSB_ATTN_ATTR(128, 1) FSB_ATTN_ATTR(128, 2) FSB_ATTN_ATTR(128, 4) FSB_ATTN_ATTR(128, 8)
    FSB_ATTN_ATTR(64, 1) FSB_ATTN_ATTR(64, 2) FSB_ATTN_ATTR(64, 4) FSB_ATTN_ATTR(64, 8)
#undef FSB_ATTN_ATTR
    return 0;
}

int launch_attn(const AttnArgs& a, cudaStream_t st) {
    if (a.rows <= 0) return 0;
    const int G = a.H / a.Hkv;
    FSB_CHECK(a.H % a.Hkv == 0, "attention: H %% Hkv != 0");
#define FSB_ATTN_CASE(DH_, G_) \
    if (a.Dh == DH_ && G == G_) return launch_attn_t<DH_, G_>(a, st);
    FSB_ATTN_CASE(128, 1) FSB_ATTN_CASE(128, 2) FSB_ATTN_CASE(128, 4) FSB_ATTN_CASE(128, 8)
    FSB_ATTN_CASE(64, 1) FSB_ATTN_CASE(64, 2) FSB_ATTN_CASE(64, 4) FSB_ATTN_CASE(64, 8)
#undef FSB_ATTN_CASE
    set_error("attention: unsupported head_dim=%d group=%d", a.Dh, G);
    return 1;
}

int swiglu_ctas(int rows, int I) {
    const bool vec = (I & 3) == 0 && rn_variant() > 0;  // partial workspaces are always 16-byte aligned in the engine
    return rows * (vec ? cdiv(I, 1024) : cdiv(I, 256));
}

int launch_swiglu(const SwigluArgs& a, cudaStream_t st) {
    if (a.rows <= 0) return 0;
    const bool fl = a.wait.ctr != nullptr || a.done_ctr != nullptr;
    const SwigluCore core = static_cast<SwigluCore>(a);
    const bool vec = (a.I & 3) == 0 && (a.parts.ld & 3) == 0 && (a.parts.slot_stride & 3) == 0 && rn_variant() > 0;
    if (vec) {
        if (fl) FSB_LAUNCH(swiglu4_kernel<true>, dim3(cdiv(a.I, 1024), a.rows), dim3(256), 0, st, core, a.wait, a.done_ctr);
        else FSB_LAUNCH(swiglu4_kernel<false>, dim3(cdiv(a.I, 1024), a.rows), dim3(256), 0, st, core, a.wait, a.done_ctr);
        return 0;
    }
    if (fl) FSB_LAUNCH(swiglu_kernel<true>, dim3(cdiv(a.I, 256), a.rows), dim3(256), 0, st, core, a.wait, a.done_ctr);
    else FSB_LAUNCH(swiglu_kernel<false>, dim3(cdiv(a.I, 256), a.rows), dim3(256), 0, st, core, a.wait, a.done_ctr);
    return 0;
}

int launch_sample(const SampleArgs& a, cudaStream_t st) {
    if (a.rows <= 0) return 0;
    FSB_CHECK(a.n > 0 && a.n <= kSampleMaxN, "sample: n=%d out of range", a.n);
    FSB_CHECK(a.ctl.state != nullptr || a.top_k >= 1, "sample: top_k must be >= 1");
    if (a.wait.ctr != nullptr)
        FSB_LAUNCH(sample_kernel<true>, dim3(a.rows), dim3(kSampleThreads), 0, st, static_cast<SampleCore>(a), a.wait,
                   a.ctl);
    else
        FSB_LAUNCH(sample_kernel<false>, dim3(a.rows), dim3(kSampleThreads), 0, st, static_cast<SampleCore>(a), a.wait,
                   a.ctl);
    return 0;
}

int launch_frame_end(const FrameEndArgs& a, cudaStream_t st) {
    FSB_LAUNCH(frame_end_kernel, dim3(a.rows), dim3(32), 0, st, static_cast<FrameEndCore>(a), a.ctl);
    FSB_LAUNCH(step_inc_kernel, dim3(1), dim3(1), 0, st, a.step);
    return 0;
}

int launch_gather_rows(const __nv_bfloat16* src, const int* idx, __nv_bfloat16* dst, int rows, int D,
                       cudaStream_t st) {
    if (rows <= 0) return 0;
    FSB_LAUNCH(gather_rows_kernel, dim3(rows), dim3(256), 0, st, src, idx, dst, D);
    return 0;
}

}  // namespace fsb
