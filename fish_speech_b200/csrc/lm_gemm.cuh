// Step GEMM: the weight-streaming matmul of the Dual-AR decode step with everything between two matmuls
// fused into it (llama.py:831-987 one TransformerBlock; inference.py:96-181 one frame).
//
//   D[i][j] = sum_k W[i][k] * X[j][k]        i = output feature (128 per tile, on the TMEM lanes)
//                                            j = sequence slot of the batch (<= 32, the UMMA N)
//
// * Weights stream HBM -> shared memory by TMA (SWIZZLE_128B, EVICT_FIRST) into an mbarrier ring and are
//   consumed by tcgen05.mma with the accumulator in TMEM.  Work is a host-built stream-K schedule: the
//   (tile, k-block) units are cut into equal contiguous ranges, one per CTA, two CTAs per SM.
// * Operand X is fetched by TMA into the same ring. Where the layer input is the residual stream, it is
//   *normalised on load*: TMA brings the un-normalised rows, two loader warps apply the reference's RMSNorm
//   (llama.py:990-1001: round(x * rsqrt(mean(x^2) + eps)) * w, two bf16 roundings) in place in shared
//   memory with the per-row sum of squares the producing kernel left behind, then hand the tile to the MMA.
// * Stream-K fix-up happens inside the kernel and is spread over the CTAs that share a tile: each stores its
//   fp32 partial, announces it on the tile's arrival counter and waits until all partials of the tile are
//   there; then contributor s sums ALL partials IN SLOT ORDER (deterministic, independent of arrival order and
//   of the batch) for its own slice of the batch rows and runs the fused epilogue on that slice:
//     EPI_QKV     bias, per-head nn.RMSNorm, interleaved RoPE, q -> q buffer, k/v -> KV cache (llama.py:891-911)
//     EPI_RESID   bias, residual add, per-tile sum of squares for the next norm          (llama.py:842-845)
//     EPI_SWIGLU  silu(w1 x) * w3 x with w1/w3 rows interleaved in the tile               (llama.py:979-987)
//     EPI_LOGITS  bf16-rounded logits for the sampler                                     (llama.py:447-457)
#pragma once
#include "gemm_tc.cuh"

namespace fsb {

enum StepEpi { EPI_QKV = 0, EPI_RESID = 1, EPI_SWIGLU = 2, EPI_LOGITS = 3 };

constexpr int kStepRows = 32;   // batch rows per step = UMMA N
constexpr int kSsqStride = 32;  // floats per row of a sum-of-squares array: one per 128-feature tile (D <= 4096)

struct StepGemmParams {
    // ---- schedule ----
    const int4* sched;      // items {tile, kb_begin, kb_end, slot}
    const int* cta_items;   // [grid + 1]
    const int* nparts;      // [tiles] partial count of each tile
    int tiles, stages;
    int n_out, K;           // output features, reduction length
    int rows;               // live batch rows (<= 32)
    unsigned long long a_hint, b_hint;
    // ---- stream-K fix-up ----
    float* ws;              // [slot][tile][32 rows][128 features] fp32 partials
    unsigned* tile_ctr;     // [2][tile_ctr_len]: arrivals, then completions; zero between launches
    int tile_ctr_len;
    // ---- operand X by normalise-on-load (BLOAD = 1) ----
    const float* x_ssq;           // [32][kSsqStride] per-tile sum of squares of x
    const __nv_bfloat16* norm_w;  // [K]
    int x_nt;                     // tiles per row in x_ssq
    float eps;
    // ---- epilogue ----
    const __nv_bfloat16* bias;    // [n_out] or null
    // EPI_RESID: x_out[j][i] = rbf(resid[j][i] + rbf(acc + bias)); ssq_out[j][tile] = sum_i x_out^2
    const __nv_bfloat16* resid;   // may alias x_out; null => no add
    __nv_bfloat16* x_out;
    float* ssq_out;
    // EPI_SWIGLU: tile t = h features [64t, 64t+64): lanes 32w+l (l < 16) = w1 row, lanes 32w+16+l = w3 row
    __nv_bfloat16* h;
    int I;
    // EPI_QKV
    const __nv_bfloat16 *q_norm, *k_norm, *freqs;
    const int *row_seq, *row_pos;
    __nv_bfloat16 *q, *kcache, *vcache;
    int H, Hkv, Dh, S;
    float qk_eps;
    // EPI_LOGITS
    float* logits;
    int logits_ld;
    // diagnostics: optional [grid][8] globaltimer stamps {start, X may be fetched, first accumulator done, partials
    // published, all partials of the last shared tile present, end, smid, items}
    unsigned long long* trace;
};

struct StepGemmPlan {
    CUtensorMap tmA, tmB;
    StepGemmParams p;
    dim3 grid;
    size_t smem;
    int epi, bload;
    void* sched_dev;
    int* cta_items_dev;
    int* nparts_dev;
    int max_parts;
    double weight_bytes;
};

// Operand X = act [32][K], fetched by TMA. norm_on_load: act is the un-normalised residual stream (the caller
// fills p.x_ssq / p.norm_w / p.x_nt / p.eps). The caller fills the epilogue fields of plan->p afterwards.
int step_plan_init(StepGemmPlan* plan, int epi, const __nv_bfloat16* w, int n_out, int K, const __nv_bfloat16* act,
                   bool norm_on_load, int num_ctas, int stages, float* ws, size_t ws_floats, unsigned* tile_ctr,
                   int tile_ctr_len);
void step_plan_free(StepGemmPlan* plan);
int step_gemm_launch(const StepGemmPlan& plan, cudaStream_t stream);
int step_gemm_init();  // kernel attributes (idempotent)

// Row index of the fused w1|w3 weight for SwiGLU-in-epilogue: h feature f -> (row of w1[f], row of w3[f]).
__host__ __device__ inline int w13_gate_row(int f) { return (f >> 6) * 128 + ((f >> 4) & 3) * 32 + (f & 15); }

}  // namespace fsb
