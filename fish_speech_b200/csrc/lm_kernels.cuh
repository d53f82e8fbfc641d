// Memory-bound kernels of the Dual-AR step that sit between the tensor-core GEMMs.  Each one is the
// consumer of a GEMM's fp32 stream-K partials: it sums the partials in a fixed order, rounds to bf16
// at exactly the points where the reference's bf16 tensors round (llama.py:990-1001 RMSNorm,
// :891-908 qk-norm/RoPE, :979-987 SwiGLU, :842-845 residual adds) and produces the next GEMM's
// operand.
#pragma once
#include "common.cuh"

namespace fsb {

// fp32 partials of one GEMM: value(row j, feature i) = sum_{s < nparts[i/128]} ws[s*slot_stride + j*ld + i]
struct Partials {
    const float* ws;
    long long slot_stride;
    int ld;
    const int* nparts;  // per 128-feature tile; null => 1
    int max_parts;      // upper bound of nparts (0 is read as 1)
};

#ifdef __CUDACC__
// Sum of the stream-K partials of one output element, always in slot order (deterministic and
// independent of the batch). Loads are issued four at a time so the L2 latencies overlap.
__device__ __forceinline__ float sum_parts(const Partials& P, int j, int i) {
    const int np = P.nparts ? __ldg(P.nparts + (i >> 7)) : 1;
    const float* p = P.ws + static_cast<size_t>(j) * P.ld + i;
    const size_t ss = static_cast<size_t>(P.slot_stride);
    float s = 0.f;
    for (int q = 0; q < np; q += 4) {
        const float a0 = p[static_cast<size_t>(q) * ss];
        const float a1 = q + 1 < np ? p[static_cast<size_t>(q + 1) * ss] : 0.f;
        const float a2 = q + 2 < np ? p[static_cast<size_t>(q + 2) * ss] : 0.f;
        const float a3 = q + 3 < np ? p[static_cast<size_t>(q + 3) * ss] : 0.f;
        s += a0;
        s += a1;
        s += a2;
        s += a3;
    }
    return s;
}

// Slot-ordered partial sums of N output elements at once: all loads of a 4-slot round are issued before
// any add, so a thread pays ~max_parts/4 L2 round trips for N elements instead of N * max_parts/4.
// The additions are in the same order as sum_parts() (bitwise identical results).
template <int N>
__device__ __forceinline__ void sum_parts_n(const Partials& P, const int (&row)[N], const int (&feat)[N],
                                            const bool (&ok)[N], float (&out)[N], int maxp) {
    int np[N];
    const float* p[N];
    const size_t ss = static_cast<size_t>(P.slot_stride);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        np[k] = ok[k] ? (P.nparts ? __ldg(P.nparts + (feat[k] >> 7)) : 1) : 0;
        p[k] = P.ws + static_cast<size_t>(ok[k] ? row[k] : 0) * P.ld + (ok[k] ? feat[k] : 0);
        out[k] = 0.f;
    }
    for (int q = 0; q < maxp; q += 4) {
        float a[N][4];
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
            for (int u = 0; u < 4; ++u) a[k][u] = (q + u < np[k]) ? p[k][static_cast<size_t>(q + u) * ss] : 0.f;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            if (q < np[k]) {
                out[k] += a[k][0];
                out[k] += a[k][1];
                out[k] += a[k][2];
                out[k] += a[k][3];
            }
        }
    }
}

#endif

struct EmbedArgs {
    const int* tokens;  // [rows][C+1] (row-major per token row)
    const __nv_bfloat16* emb;     // [V, D]
    const __nv_bfloat16* cb_emb;  // [C*cs, D]
    __nv_bfloat16* x;             // [rows, D]
    int rows, D, C, cs, vocab;
    int sem_begin, sem_end;
    int scale;  // scale_codebook_embeddings
};
int launch_embed(const EmbedArgs& a, cudaStream_t st);

// x_out = rbf(x_in + rbf(sum partials + bias))   (skip the add when parts.ws == null)
// n_out = rbf(rbf(x_out * rsqrt(mean(x_out^2)+eps)) * w)     [fish RMSNorm: round, then * weight]
// NOTE on the *Core / *Args split below: the kernels take the Core struct by value and the optional
// dependency-flag / slot-control fields as SEPARATE kernel parameters. Growing a by-value argument struct past
// 128 bytes changed the register allocation of resid_norm (128 -> 99 registers, fewer partial-sum loads in
// flight) and cost 0.7 ms per decode frame (measured A/B on the same B200).
struct ResidNormCore {
    Partials parts;
    const __nv_bfloat16* bias;  // [D] or null (attention_o_bias)
    const __nv_bfloat16* scale;  // [D] or null: y *= scale (codec LayerScale, modded_dac.py:329-341)
    const __nv_bfloat16* x_in;  // [rows, D] residual input (null => 0)
    __nv_bfloat16* x_out;       // [rows, D] (may alias x_in; null => don't store)
    const __nv_bfloat16* norm_w;  // [D] (null => no norm output)
    __nv_bfloat16* n_out;         // [rows, D]
    const int* gather;      // optional: output row r takes input row gather[map(r) * gather_stride]
    const int* gather_map;  // optional row -> slot map applied before the gather lookup
    int rows, D;
    float eps;
};
struct ResidNormArgs : ResidNormCore {
    DepFlag wait;        // optional flag dependency (else grid dependency)
    unsigned* done_ctr;  // optional: signalled once per CTA at the end
};
int launch_resid_norm(const ResidNormArgs& a, cudaStream_t st);
// gather[r * gather_stride] names the x_in row of output row r (embedding lookups)
int launch_resid_norm_g(const ResidNormArgs& a, int gather_stride, cudaStream_t st);

// y = rbf(sum partials + bias)  -> bf16 rows (fast_project_in)
struct LinearOutArgs {
    Partials parts;
    const __nv_bfloat16* bias;
    __nv_bfloat16* y;
    int rows, N;
};
int launch_linear_out(const LinearOutArgs& a, cudaStream_t st);

// q,k,v = rbf(partials [+bias]); optional per-head nn.RMSNorm (single rounding); interleaved RoPE in
// fp32 with bf16 tables; q -> qbuf[row][H][Dh]; k,v -> cache[b][hkv][pos][Dh].
struct QkvPrepCore {
    Partials parts;
    const __nv_bfloat16* bias;  // [(H+2Hkv)*Dh] or null
    const __nv_bfloat16* q_norm;  // [Dh] or null
    const __nv_bfloat16* k_norm;
    const __nv_bfloat16* freqs;  // [S, Dh/2, 2] bf16 (cos, sin)
    const int* row_seq;          // [rows] cache slot of each row
    const int* row_pos;          // [rows] position of each row
    __nv_bfloat16* q;            // [rows, H, Dh]
    __nv_bfloat16* kcache;       // [Bslots, Hkv, S, Dh]
    __nv_bfloat16* vcache;
    int rows, H, Hkv, Dh, S;
    float eps;
};
struct QkvPrepArgs : QkvPrepCore {
    DepFlag wait;
};
int launch_qkv_prep(const QkvPrepArgs& a, cudaStream_t st);

// out[row][h][:] = softmax(q.k^T * scale over cache positions [max(0,pos-window+1), pos]) . v
// bf16_math = 1 reproduces the fast-AR hand-rolled attention (llama.py:948-976): scores, scaled
// scores, probabilities and the output are each rounded to bf16.
struct AttnCore {
    const __nv_bfloat16* q;  // [rows, H, Dh]
    const __nv_bfloat16* kcache;
    const __nv_bfloat16* vcache;
    const int* row_seq;
    const int* row_pos;
    __nv_bfloat16* out;  // [rows, H*Dh]
    int rows, H, Hkv, Dh, S;
    int window;  // <=0: unlimited
    int lcap;    // score-buffer length: an upper bound of (row_pos + 1); 0 = cache capacity S
    int bf16_math;
};
struct AttnArgs : AttnCore {
    unsigned* done_ctr;  // optional: signalled once per CTA at the end
};
int launch_attn(const AttnArgs& a, cudaStream_t st);
int attn_init();  // set kernel attributes (idempotent)

// h = rbf( rbf(silu(rbf(a))) * rbf(c) ), a = feature i, c = feature I+i of the fused w1|w3 GEMM
struct SwigluCore {
    Partials parts;
    __nv_bfloat16* h;  // [rows, I]
    int rows, I;
};
struct SwigluArgs : SwigluCore {
    DepFlag wait;
    unsigned* done_ctr;
};
int launch_swiglu(const SwigluArgs& a, cudaStream_t st);
int swiglu_ctas(int rows, int I);  // CTAs of one launch (what a dependency flag must count)

// Per-slot request control (continuous batching). When `state` is non-null the frame kernels
// take the sampling parameters, the RNG stream and the stop rule from these arrays instead of the per-call
// values, so that every request computes exactly what it would compute alone, whatever shares the batch.
//   state: 0 idle, 1 active, 2 finished (frozen until the host retires it), 3 finishing (set by the slow
//   sampler on <|im_end|>; this frame is still recorded, inference.py:233-234)
struct SlotCtl {
    int* state;
    const int* limit;  // frames this request may produce (the prefill frame included)
    const float* temperature;
    const float* top_p;
    const int* top_k;
    const unsigned long long* seed;
    const int* n_out;  // frames produced so far = RNG counter of the request
};
__device__ __forceinline__ bool slot_live(const SlotCtl& c, int slot) {
    if (c.state == nullptr) return true;
    const int s = c.state[slot];
    return s == 1 || s == 3;
}

struct SampleCore {
    Partials parts;  // logits of the (restricted) head: n entries per row
    int n;           // number of candidate entries (<= 8192)
    int rows;
    // sampling parameters
    float temperature, top_p;
    int top_k;
    // token mapping for the slow head: entry e -> token id (e < n_sem ? sem_begin + e : im_end_id)
    int slow;  // 1: slow head (RAS + token mapping), 0: fast head (codes)
    int n_sem, sem_begin, im_end_id, codebook_size;
    int use_ras;              // RAS only in decode frames (previous_tokens given)
    int* ras_window;          // [rows][10] ring of previous main tokens (slow only)
    int ras_update;           // push the chosen token into the window
    unsigned long long seed;  // Philox key
    const unsigned long long* rng_offset;  // device counter (frame index) mixed into the stream
    int draw_id;              // distinguishes the samples of one frame
    // outputs
    int* cur_tok;     // [rows][C+1]
    int cb_index;     // slow: writes column 0 (token) and 1 (code a0); fast: writes column cb_index+1
    int num_cb;
    float* logits_out;  // optional [rows][n] fp32 copy of the bf16-rounded logits (tests)
    int* finished;      // slow only: set when token == im_end
    const int* row_slot;  // optional: state (cur_tok / window / finished / logits_out) index of a row
};
struct SampleArgs : SampleCore {
    DepFlag wait;
    SlotCtl ctl;
};
int launch_sample(const SampleArgs& a, cudaStream_t st);

// bookkeeping at the end of a frame for each row's slot:
//   out_tokens[slot][c][n_out[slot]] = cur_tok[slot][c]; n_out[slot]++;
//   pos[slot] = set_pos_rows ? row_pos_src[set_pos_rows[row]] + 1 : pos[slot] + 1;   step++
struct FrameEndCore {
    const int* cur_tok;
    int* out_tokens;  // [slots][C+1][T_cap]
    int* n_out;
    int* pos;
    const int* row_slot;      // optional
    const int* set_pos_rows;  // optional (prefill): last token row of each sequence
    const int* row_pos_src;
    unsigned long long* step;
    int rows, ncols, T_cap;
};
struct FrameEndArgs : FrameEndCore {
    SlotCtl ctl;
};
int launch_frame_end(const FrameEndArgs& a, cudaStream_t st);

int launch_gather_rows(const __nv_bfloat16* src, const int* idx, __nv_bfloat16* dst, int rows, int D,
                       cudaStream_t st);

}  // namespace fsb
