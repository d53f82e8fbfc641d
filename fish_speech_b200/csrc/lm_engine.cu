// Dual-AR engine: owns the KV caches, workspaces and GEMM plans of one model replica and runs
// prefill / per-frame decode as a fixed kernel sequence (captured into a CUDA graph for decode).
//
// HBM layout (all bf16 unless noted)
//   weights            caller-owned, row-major [out, in] as in the checkpoint (TMA-ready, K-major)
//   slow KV cache      [n_layer][max_batch][Hkv][kv_len][Dh]   x2 (K, V)
//   fast KV cache      [n_fast_layer][max_batch][fHkv][num_codebooks][fDh]  x2
//   decode workspaces  xres_d/xn_d [32][max(D,Df)], q_d/attn_d [32][max(H*Dh)], h_d [32][max(I)]
//   prefill workspaces same with max_rows rows
//   ws                 fp32 GEMM partial sums: [slot][row][feature]
//
// Reference: fish_speech/models/text2semantic/llama.py:390-466 (slow step), :799-817 (fast step),
// fish_speech/models/text2semantic/inference.py:96-181 (one frame), :184-238 (frame loop).
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/fishb200.h"
#include "gemm_tc.cuh"
#include "lm_kernels.cuh"
#include "lm_persist.cuh"

using namespace fsb;

namespace {

typedef __nv_bfloat16 bf16;
constexpr int kDecRows = 32;  // decode GEMM N tile: up to 32 sequences per step

struct LayerW {
    const bf16 *attn_norm, *wqkv, *bqkv, *q_norm, *k_norm, *wo, *bo, *ffn_norm, *w13, *w2;
};
struct LayerPlans {
    GemmPlan qkv, wo, w13, w2;
};

struct Stack {
    int D = 0, H = 0, Hkv = 0, Dh = 0, I = 0, nl = 0, S = 0;
    bool qk_norm = false;
    int bf16_math = 0;
    std::vector<LayerW> w;
    std::vector<LayerPlans> dec;  // BN=32, stream-K, B operand = decode workspaces
    std::vector<LayerPlans> pf;   // BN=128 tiles, B operand = prefill workspaces (slow only)
    std::vector<LayerPlans> pk;   // stream-K over one CTA per SM, for the persistent stack kernel
    PkLayer* pk_layers = nullptr;  // device array
    bf16 *kcache = nullptr, *vcache = nullptr;
    const bf16* freqs = nullptr;
    size_t cache_layer_stride = 0;
};

}  // namespace

struct fsb_lm {
    // device-side dependency flags for the decode chain (FSB_FLAGS, see common.cuh DepFlag)
    bool flags_on = false;
    unsigned* flag_base = nullptr;
    int flag_next = 0;
    static constexpr int kFlagCap = 4096;
    // per-slot request control (continuous batching): see SlotCtl in lm_kernels.cuh
    bool fused_fast_only = false;
    bool slot_control = false;
    int* slot_state = nullptr;
    int* slot_limit = nullptr;
    float* slot_temperature = nullptr;
    float* slot_top_p = nullptr;
    int* slot_top_k = nullptr;
    unsigned long long* slot_seed = nullptr;
    bool graph_slot_control = false;
    fsb_lm_config cfg;
    int num_sms = 148;
    Stack slow, fast;
    const bf16 *emb = nullptr, *cb_emb = nullptr, *norm_w = nullptr, *head_w = nullptr;
    const bf16 *fast_emb = nullptr, *fast_norm_w = nullptr, *fast_out_w = nullptr;
    const bf16 *fast_proj_w = nullptr, *fast_proj_b = nullptr;
    int head_rows = 0;
    GemmPlan head_plan, fast_out_plan, proj_plan;
    bool has_proj = false;
    // workspaces
    bf16 *xres_d = nullptr, *xn_d = nullptr, *q_d = nullptr, *attn_d = nullptr, *h_d = nullptr;
    bf16 *xres_p = nullptr, *xn_p = nullptr, *q_p = nullptr, *attn_p = nullptr, *h_p = nullptr;
    bf16* proj_d = nullptr;
    float* ws = nullptr;
    size_t ws_floats = 0;
    // state
    int *cur_tok = nullptr, *out_tokens = nullptr, *n_out = nullptr, *pos = nullptr, *finished = nullptr;
    int *ras_window = nullptr, *iota = nullptr, *fpos = nullptr;
    unsigned long long* step = nullptr;
    // debug
    float *slow_logits = nullptr, *fast_logits = nullptr;
    bf16* dbg_x = nullptr;
    int ctx_lcap = 0;  // score-buffer bound for the slow attention (0 = capacity)
    int graph_lcap = -1;
    bool persistent = false;
    bool fused_prep_attn = false;
    int pk_stages = 8;
    unsigned* pk_bar = nullptr;
    unsigned long long* pk_trace = nullptr;
    // decode graph cache
    cudaGraphExec_t graph_exec = nullptr;
    int graph_batch = -1;
    int graph_kernels = 0;  // kernels inside one captured frame
    fsb_sampling graph_sampling{};
    std::vector<void*> owned;
    std::map<std::string, std::pair<void*, size_t>> named;
};

namespace {

template <typename T>
int dalloc(fsb_lm* h, T** p, size_t count, const char* name = nullptr) {
    void* q = nullptr;
    const size_t bytes = std::max<size_t>(count * sizeof(T), 256);
    FSB_CUDA(cudaMalloc(&q, bytes));
    FSB_CUDA(cudaMemset(q, 0, bytes));
    h->owned.push_back(q);
    *p = reinterpret_cast<T*>(q);
    if (name) h->named[name] = {q, count * sizeof(T)};
    return 0;
}

// Decode-orientation plan: A = weight [n_out, k] on the TMEM lanes, B = activations [rows, k].
int make_plan(fsb_lm* h, GemmPlan* plan, const bf16* w, int n_out, int k, const bf16* act, int act_rows,
              bool decode, int force_per_sm = 0) {
    memset(plan, 0, sizeof(*plan));
    GemmOperand A{w, k, n_out, 1, k, static_cast<long long>(n_out) * k};
    GemmOperand B{act, k, act_rows, 1, k, static_cast<long long>(act_rows) * k};
    GemmParams& p = plan->p;
    const int kblocks = cdiv(k, 64);
    p.kb_per_tap = kblocks;
    p.num_taps = 1;
    p.a_hint = kEvictFirst;  // weights are streamed once per step (>> L2)
    p.a_static = 1;
    {
        const char* e = getenv("FSB_L2_PREFETCH");
        p.l2_prefetch = decode ? (e ? atoi(e) : 0) : 0;
    }
    p.b_hint = kEvictLast;   // the activation tile is re-read by every CTA
    p.rows_i = n_out;
    p.mode = 0;
    p.ws = h->ws;
    p.ws_ld = n_out;
    const int tiles_i = cdiv(n_out, 128);
    if (decode) {
        // ring depth / CTAs per SM are tunable for experiments (FSB_STAGES, FSB_CTAS_PER_SM)
        const char* es = getenv("FSB_STAGES");
        const char* ec = getenv("FSB_CTAS_PER_SM");
        const int stages = es ? atoi(es) : 4;   // measured best on B200: 4 stages x 2 CTAs per SM
        const int per_sm = force_per_sm ? force_per_sm : (ec ? atoi(ec) : 2);
        FSB_TRY(gemm_plan_init(plan, A, B, 32, stages, tiles_i, 1, 1));
        p.rows_j = kDecRows;
        p.ws_slot_stride = static_cast<long long>(kDecRows) * n_out;
        // the persistent kernel indexes the schedule by blockIdx.x of a full grid: keep empty CTA ranges
        FSB_TRY(gemm_plan_streamk(plan, tiles_i, kblocks, h->num_sms * per_sm, force_per_sm != 0));
        FSB_CHECK(static_cast<size_t>(plan->max_parts) * kDecRows * n_out <= h->ws_floats,
                  "partial workspace too small");
    } else {
        FSB_TRY(gemm_plan_init(plan, A, B, 128, 4, tiles_i, cdiv(act_rows, 128), 1));
        p.rows_j = act_rows;
        p.ws_slot_stride = 0;
        FSB_CHECK(static_cast<size_t>(act_rows) * n_out <= h->ws_floats, "partial workspace too small");
    }
    return 0;
}

Partials parts_of(const GemmPlan& p) {
    Partials P;
    P.ws = p.p.ws;
    P.slot_stride = p.p.ws_slot_stride;
    P.ld = p.p.ws_ld;
    P.nparts = p.nparts_dev;
    P.max_parts = p.max_parts;
    return P;
}

// Every frame starts from zeroed counters; each dependency edge of the frame gets its own counter.
int flags_begin(fsb_lm* h, cudaStream_t st) {
    if (!h->flags_on) return 0;
    h->flag_next = 0;
    FSB_CUDA(cudaMemsetAsync(h->flag_base, 0, static_cast<size_t>(fsb_lm::kFlagCap) * 4, st));
    return 0;
}
unsigned* flag_new(fsb_lm* h) {
    if (!h->flags_on || h->flag_next >= fsb_lm::kFlagCap) return nullptr;
    return h->flag_base + h->flag_next++;
}

// Decode GEMM with flag dependencies: operand B becomes valid when `ready` is satisfied; the returned
// flag is satisfied when every CTA has written its partial sums.
int launch_dec_gemm(fsb_lm* h, const GemmPlan& plan, const DepFlag& ready, DepFlag* done, cudaStream_t st) {
    *done = DepFlag{nullptr, 0};
    if (!h->flags_on) return gemm_launch(plan, st);
    GemmPlan q = plan;
    q.p.b_ready = ready;
    q.p.done_ctr = flag_new(h);
    if (q.p.done_ctr) *done = DepFlag{q.p.done_ctr, plan.grid.x * plan.grid.y * plan.grid.z};
    return gemm_launch(q, st);
}

int launch_rows(const GemmPlan& plan, int rows, cudaStream_t st) {
    // prefill plans: restrict the column tiles to the live rows
    if (plan.p.sched == nullptr) {
        GemmPlan q = plan;
        q.grid.y = cdiv(rows, plan.bn);
        q.p.rows_j = rows;
        return gemm_launch(q, st);
    }
    return gemm_launch(plan, st);
}

struct RowCtx {
    int rows;
    const int* row_seq;
    const int* row_pos;
    bf16 *xres, *xn, *q, *attn, *hbuf;
    bool decode;
};

// One transformer stack over `rows` token rows. On entry xn = attention_norm_0(xres).
// `final_norm`: weight of the norm applied after the last layer (-> xn). `stop_after_kv`: fast pass 0
// only needs the last layer's K/V (its logits are discarded, inference.py:147).
int run_stack(fsb_lm* h, Stack& s, const RowCtx& c, const bf16* final_norm, bool stop_after_kv,
              bf16* dbg, cudaStream_t st, DepFlag* ready_io = nullptr) {
    const float eps = h->cfg.norm_eps;
    // flag mode (decode only): `ready` = "xn holds this layer's normed input"
    const bool fl = h->flags_on && c.decode && ready_io != nullptr && !h->persistent && !h->fused_prep_attn;
    DepFlag ready = fl ? *ready_io : DepFlag{nullptr, 0};
    DepFlag gd{nullptr, 0};
    if (h->persistent && c.decode && s.pk_layers != nullptr) {
        PkArgs A{};
        A.layers = s.pk_layers;
        A.nl = s.nl; A.rows = c.rows; A.D = s.D; A.H = s.H; A.Hkv = s.Hkv; A.Dh = s.Dh; A.I = s.I; A.S = s.S;
        A.eps = eps;
        A.bf16_math = s.bf16_math;
        A.qk_norm = s.qk_norm ? 1 : 0;
        A.kv_only_last = stop_after_kv ? 1 : 0;
        A.stages = h->pk_stages;
        {
            const char* e = getenv("FSB_PK_L2PF");
            A.l2_prefetch = e ? atoi(e) : 24;
        }
        A.row_seq = c.row_seq;
        A.row_pos = c.row_pos;
        A.freqs = s.freqs;
        A.xres = c.xres; A.xn = c.xn; A.attn = c.attn; A.hbuf = c.hbuf;
        A.ws = h->ws;
        A.bar = h->pk_bar;
        A.trace = (s.bf16_math == 0) ? h->pk_trace : nullptr;  // slow stack only
        A.trace_max = 1024;
        (void)final_norm;  // baked into the last layer's next_norm
        return launch_stack_persistent(A, h->num_sms, st);
    }
    for (int l = 0; l < s.nl; ++l) {
        const LayerW& w = s.w[l];
        LayerPlans& P = c.decode ? s.dec[l] : s.pf[l];
        if (fl) FSB_TRY(launch_dec_gemm(h, P.qkv, ready, &gd, st));
        else FSB_TRY(launch_rows(P.qkv, c.rows, st));
        if (c.decode && h->fused_prep_attn && (!h->fused_fast_only || s.bf16_math)) {
            // decode rows: q/k/v post-processing, KV append and attention in one launch
            PkArgs A{};
            A.rows = c.rows; A.D = s.D; A.H = s.H; A.Hkv = s.Hkv; A.Dh = s.Dh; A.I = s.I; A.S = s.S;
            A.eps = eps;
            A.bf16_math = s.bf16_math;
            A.qk_norm = s.qk_norm ? 1 : 0;
            A.row_seq = c.row_seq;
            A.row_pos = c.row_pos;
            A.freqs = s.freqs;
            A.attn = c.attn;
            A.ws = h->ws;
            PkLayer L{};
            L.qkv.nparts = P.qkv.nparts_dev;
            L.qkv.n_out = P.qkv.p.ws_ld;
            L.qkv.max_parts = P.qkv.max_parts;
            L.qkv.slot_stride = P.qkv.p.ws_slot_stride;
            L.bqkv = w.bqkv;
            L.q_norm = s.qk_norm ? w.q_norm : nullptr;
            L.k_norm = s.qk_norm ? w.k_norm : nullptr;
            L.kcache = s.kcache + l * s.cache_layer_stride;
            L.vcache = s.vcache + l * s.cache_layer_stride;
            const bool kv_only = stop_after_kv && l == s.nl - 1;
            const int rc = launch_prep_attn(A, L, kv_only ? 1 : 0, st);
            if (rc > 0) return rc;
            if (rc == 0) {
                if (kv_only) return 0;
                goto after_attention;
            }
        }
        {
        QkvPrepArgs qa{};
        qa.parts = parts_of(P.qkv);
        qa.bias = w.bqkv;
        qa.q_norm = s.qk_norm ? w.q_norm : nullptr;
        qa.k_norm = s.qk_norm ? w.k_norm : nullptr;
        qa.freqs = s.freqs;
        qa.row_seq = c.row_seq;
        qa.row_pos = c.row_pos;
        qa.q = c.q;
        qa.kcache = s.kcache + l * s.cache_layer_stride;
        qa.vcache = s.vcache + l * s.cache_layer_stride;
        qa.rows = c.rows; qa.H = s.H; qa.Hkv = s.Hkv; qa.Dh = s.Dh; qa.S = s.S;
        qa.eps = eps;
        qa.wait = gd;
        FSB_TRY(launch_qkv_prep(qa, st));
        if (stop_after_kv && l == s.nl - 1) {
            if (ready_io) *ready_io = DepFlag{nullptr, 0};  // the next kernel depends on this grid classically
            return 0;
        }
        AttnArgs aa{};
        aa.q = c.q;
        aa.kcache = qa.kcache;
        aa.vcache = qa.vcache;
        aa.row_seq = c.row_seq;
        aa.row_pos = c.row_pos;
        aa.out = c.attn;
        aa.rows = c.rows; aa.H = s.H; aa.Hkv = s.Hkv; aa.Dh = s.Dh; aa.S = s.S;
        aa.window = 0;
        aa.lcap = (s.bf16_math == 0) ? h->ctx_lcap : 0;  // slow stack: bounded by the live context, not the capacity
        aa.bf16_math = s.bf16_math;
        aa.done_ctr = fl ? flag_new(h) : nullptr;
        FSB_TRY(launch_attn(aa, st));
        ready = DepFlag{aa.done_ctr, static_cast<unsigned>(s.Hkv * c.rows)};
        }
    after_attention:
        if (fl) FSB_TRY(launch_dec_gemm(h, P.wo, ready, &gd, st));
        else FSB_TRY(launch_rows(P.wo, c.rows, st));
        ResidNormArgs r1{};
        r1.parts = parts_of(P.wo);
        r1.bias = w.bo;
        r1.x_in = c.xres; r1.x_out = c.xres;
        r1.norm_w = w.ffn_norm; r1.n_out = c.xn;
        r1.rows = c.rows; r1.D = s.D; r1.eps = eps;
        r1.wait = gd;
        r1.done_ctr = fl ? flag_new(h) : nullptr;
        FSB_TRY(launch_resid_norm(r1, st));
        ready = DepFlag{r1.done_ctr, static_cast<unsigned>(c.rows)};
        if (fl) FSB_TRY(launch_dec_gemm(h, P.w13, ready, &gd, st));
        else FSB_TRY(launch_rows(P.w13, c.rows, st));
        SwigluArgs sa{};
        sa.parts = parts_of(P.w13);
        sa.h = c.hbuf; sa.rows = c.rows; sa.I = s.I;
        sa.wait = gd;
        sa.done_ctr = fl ? flag_new(h) : nullptr;
        FSB_TRY(launch_swiglu(sa, st));
        ready = DepFlag{sa.done_ctr, static_cast<unsigned>(swiglu_ctas(c.rows, s.I))};
        if (fl) FSB_TRY(launch_dec_gemm(h, P.w2, ready, &gd, st));
        else FSB_TRY(launch_rows(P.w2, c.rows, st));
        ResidNormArgs r2{};
        r2.parts = parts_of(P.w2);
        r2.x_in = c.xres; r2.x_out = c.xres;
        r2.norm_w = (l + 1 < s.nl) ? s.w[l + 1].attn_norm : final_norm;
        r2.n_out = c.xn;
        r2.rows = c.rows; r2.D = s.D; r2.eps = eps;
        r2.wait = gd;
        r2.done_ctr = fl ? flag_new(h) : nullptr;
        FSB_TRY(launch_resid_norm(r2, st));
        ready = DepFlag{r2.done_ctr, static_cast<unsigned>(c.rows)};
        if (dbg)
            FSB_CUDA(cudaMemcpyAsync(dbg + static_cast<size_t>(l + 1) * kDecRows * s.D, c.xres,
                                     static_cast<size_t>(std::min(c.rows, kDecRows)) * s.D * 2,
                                     cudaMemcpyDeviceToDevice, st));
    }
    if (ready_io) *ready_io = ready;
    return 0;
}

SlotCtl slot_ctl(const fsb_lm* h) {
    SlotCtl c{};
    if (!h->slot_control) return c;
    c.state = h->slot_state;
    c.limit = h->slot_limit;
    c.temperature = h->slot_temperature;
    c.top_p = h->slot_top_p;
    c.top_k = h->slot_top_k;
    c.seed = h->slot_seed;
    c.n_out = h->n_out;
    return c;
}

// Head + sampling + fast passes + bookkeeping for `rows` sequences whose final-normed last hidden
// state is in xn_d[0..rows) (and un-normed residual in xres_d). inference.py:114-181.
int run_frame_tail(fsb_lm* h, int rows, const int* row_slot, bool use_ras, const int* set_pos_rows,
                   const int* row_pos_src, const fsb_sampling& sp, cudaStream_t st,
                   DepFlag ready = DepFlag{nullptr, 0}) {
    const fsb_lm_config& c = h->cfg;
    const int C = c.num_codebooks;
    const int* slots = row_slot ? row_slot : h->iota;
    // ---- slow head over the selectable rows ----
    DepFlag gd{nullptr, 0};
    FSB_TRY(launch_dec_gemm(h, h->head_plan, ready, &gd, st));
    SampleArgs sa{};
    sa.wait = gd;
    sa.ctl = slot_ctl(h);
    sa.parts = parts_of(h->head_plan);
    sa.n = h->head_rows;
    sa.rows = rows;
    sa.temperature = sp.temperature; sa.top_p = sp.top_p; sa.top_k = sp.top_k;
    sa.slow = 1;
    sa.n_sem = h->head_rows - 1;
    sa.sem_begin = c.semantic_begin_id;
    sa.im_end_id = c.im_end_id;
    sa.codebook_size = c.codebook_size;
    sa.use_ras = use_ras ? 1 : 0;
    sa.ras_window = h->ras_window;
    sa.ras_update = use_ras ? 1 : 0;
    sa.seed = sp.seed;
    sa.rng_offset = h->step;
    sa.draw_id = 0;
    sa.cur_tok = h->cur_tok;
    sa.cb_index = 0;
    sa.num_cb = C;
    sa.logits_out = h->slow_logits;
    sa.finished = h->finished;
    sa.row_slot = row_slot;
    FSB_TRY(launch_sample(sa, st));

    // ---- fast passes ----
    Stack& f = h->fast;
    const bf16* hidden = c.norm_fastlayer_input ? h->xn_d : h->xres_d;
    RowCtx ctx{rows, slots, nullptr, h->xres_d, h->xn_d, h->q_d, h->attn_d, h->h_d, true};
    for (int p = 0; p < C; ++p) {
        ctx.row_pos = h->fpos + p * kDecRows;
        ResidNormArgs r{};
        r.rows = rows; r.D = f.D; r.eps = c.norm_eps;
        r.x_out = h->xres_d;
        r.norm_w = f.w[0].attn_norm;
        r.n_out = h->xn_d;
        if (p == 0) {
            if (h->has_proj) {
                // hidden must be the GEMM operand: it already lives in xn_d when norm_fastlayer_input,
                // otherwise stage it there.
                if (!c.norm_fastlayer_input)
                    FSB_CUDA(cudaMemcpyAsync(h->xn_d, h->xres_d, static_cast<size_t>(kDecRows) * h->slow.D * 2,
                                             cudaMemcpyDeviceToDevice, st));
                FSB_TRY(gemm_launch(h->proj_plan, st));
                LinearOutArgs lo{};
                lo.parts = parts_of(h->proj_plan);
                lo.bias = h->fast_proj_b;
                lo.y = h->proj_d; lo.rows = rows; lo.N = f.D;
                FSB_TRY(launch_linear_out(lo, st));
                r.x_in = h->proj_d;
            } else {
                r.x_in = hidden;
            }
            r.done_ctr = flag_new(h);
            FSB_TRY(launch_resid_norm(r, st));
        } else {
            // input = fast_embeddings[code_{p-1}] ; codes live in cur_tok[slot][p]
            r.x_in = h->fast_emb;
            r.gather = h->cur_tok + p;
            r.gather_map = row_slot;
            r.done_ctr = flag_new(h);
            FSB_TRY(launch_resid_norm_g(r, C + 1, st));
        }
        DepFlag fr{r.done_ctr, static_cast<unsigned>(rows)};
        FSB_TRY(run_stack(h, f, ctx, h->fast_norm_w, p == 0, nullptr, st, &fr));
        if (p == 0) continue;
        FSB_TRY(launch_dec_gemm(h, h->fast_out_plan, fr, &gd, st));
        SampleArgs fa{};
        fa.wait = gd;
        fa.ctl = slot_ctl(h);
        fa.parts = parts_of(h->fast_out_plan);
        fa.n = c.codebook_size;
        fa.rows = rows;
        fa.temperature = sp.temperature; fa.top_p = sp.top_p; fa.top_k = sp.top_k;
        fa.slow = 0;
        fa.seed = sp.seed;
        fa.rng_offset = h->step;
        fa.draw_id = p;
        fa.cur_tok = h->cur_tok;
        fa.cb_index = p;
        fa.num_cb = C;
        fa.logits_out = h->fast_logits ? h->fast_logits + static_cast<size_t>(p - 1) * c.max_batch * c.codebook_size
                                       : nullptr;
        fa.row_slot = row_slot;
        FSB_TRY(launch_sample(fa, st));
    }
    FrameEndArgs fe{};
    fe.ctl = slot_ctl(h);
    fe.cur_tok = h->cur_tok;
    fe.out_tokens = h->out_tokens;
    fe.n_out = h->n_out;
    fe.pos = h->pos;
    fe.row_slot = row_slot;
    fe.set_pos_rows = set_pos_rows;
    fe.row_pos_src = row_pos_src;
    fe.step = h->step;
    fe.rows = rows; fe.ncols = C + 1; fe.T_cap = c.max_frames;
    FSB_TRY(launch_frame_end(fe, st));
    return 0;
}

int decode_one_frame(fsb_lm* h, int batch, const fsb_sampling& sp, cudaStream_t st) {
    const fsb_lm_config& c = h->cfg;
    Stack& s = h->slow;
    EmbedArgs ea{};
    ea.tokens = h->cur_tok;
    ea.emb = h->emb; ea.cb_emb = h->cb_emb; ea.x = h->xres_d;
    ea.rows = batch; ea.D = s.D; ea.C = c.num_codebooks; ea.cs = c.codebook_size; ea.vocab = c.vocab_size;
    ea.sem_begin = c.semantic_begin_id; ea.sem_end = c.semantic_end_id;
    ea.scale = c.scale_codebook_embeddings;
    FSB_TRY(flags_begin(h, st));
    FSB_TRY(launch_embed(ea, st));
    ResidNormArgs r{};
    r.x_in = h->xres_d;
    r.norm_w = s.w[0].attn_norm; r.n_out = h->xn_d;
    r.rows = batch; r.D = s.D; r.eps = c.norm_eps;
    r.done_ctr = flag_new(h);
    FSB_TRY(launch_resid_norm(r, st));
    DepFlag ready{r.done_ctr, static_cast<unsigned>(batch)};
    bf16* dbg = (h->dbg_x && h->graph_exec == nullptr) ? h->dbg_x : nullptr;
    if (dbg)
        FSB_CUDA(cudaMemcpyAsync(dbg, h->xres_d, static_cast<size_t>(batch) * s.D * 2,
                                 cudaMemcpyDeviceToDevice, st));
    RowCtx ctx{batch, h->iota, h->pos, h->xres_d, h->xn_d, h->q_d, h->attn_d, h->h_d, true};
    FSB_TRY(run_stack(h, s, ctx, h->norm_w, false, dbg, st, &ready));
    return run_frame_tail(h, batch, nullptr, true, nullptr, nullptr, sp, st, ready);
}

}  // namespace

extern "C" {

int fsb_lm_create(const fsb_lm_config* cfg, const fsb_lm_weights* w, fsb_lm** out) {
    FSB_CHECK(cfg && w && out, "fsb_lm_create: null argument");
    FSB_CHECK(cfg->max_batch >= 1 && cfg->max_batch <= kDecRows, "max_batch must be in [1,32]");
    FSB_CHECK(cfg->dim % 8 == 0 && cfg->fast_dim % 8 == 0 && cfg->intermediate % 8 == 0 &&
                  cfg->fast_intermediate % 8 == 0,
              "dims must be multiples of 8 (TMA 16-byte strides)");
    FSB_CHECK(w->head_rows == cfg->semantic_end_id - cfg->semantic_begin_id + 2, "head_rows mismatch");
    fsb_lm* h = new fsb_lm();
    h->cfg = *cfg;
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        set_error("fsb_lm_create: no CUDA device");
        delete h;
        return 1;
    }
    FSB_CHECK(prop.major == 10, "fishb200 kernels are built for sm_100a only (device is sm_%d%d)", prop.major,
              prop.minor);
    h->num_sms = prop.multiProcessorCount;
    if (gemm_init() != 0 || attn_init() != 0) {
        delete h;
        return 1;
    }

    auto B16 = [](const void* p) { return reinterpret_cast<const bf16*>(p); };
    h->emb = B16(w->d_embeddings);
    h->cb_emb = B16(w->d_codebook_embeddings);
    h->norm_w = B16(w->d_norm);
    h->head_w = B16(w->d_head);
    h->head_rows = w->head_rows;
    h->fast_emb = B16(w->d_fast_embeddings);
    h->fast_norm_w = B16(w->d_fast_norm);
    h->fast_out_w = B16(w->d_fast_output);
    h->fast_proj_w = B16(w->d_fast_proj_w);
    h->fast_proj_b = B16(w->d_fast_proj_b);
    h->has_proj = w->d_fast_proj_w != nullptr;

    Stack& s = h->slow;
    s.D = cfg->dim; s.H = cfg->n_head; s.Hkv = cfg->n_kv_head; s.Dh = cfg->head_dim; s.I = cfg->intermediate;
    s.nl = cfg->n_layer; s.S = cfg->kv_len; s.qk_norm = cfg->qk_norm != 0; s.bf16_math = 0;
    s.freqs = B16(w->d_freqs);
    Stack& f = h->fast;
    f.D = cfg->fast_dim; f.H = cfg->fast_n_head; f.Hkv = cfg->fast_n_kv_head; f.Dh = cfg->fast_head_dim;
    f.I = cfg->fast_intermediate; f.nl = cfg->n_fast_layer; f.S = cfg->num_codebooks;
    f.qk_norm = cfg->fast_qk_norm != 0; f.bf16_math = 1;
    f.freqs = B16(w->d_fast_freqs);
    auto copy_layers = [&](Stack& st, const fsb_lm_layer* L) {
        st.w.resize(st.nl);
        for (int l = 0; l < st.nl; ++l) {
            st.w[l] = LayerW{B16(L[l].d_attn_norm), B16(L[l].d_wqkv), B16(L[l].d_bqkv), B16(L[l].d_q_norm),
                            B16(L[l].d_k_norm), B16(L[l].d_wo), B16(L[l].d_bo), B16(L[l].d_ffn_norm),
                            B16(L[l].d_w13), B16(L[l].d_w2)};
        }
    };
    copy_layers(s, w->layers);
    copy_layers(f, w->fast_layers);

    const int C = cfg->num_codebooks;
    const int Dm = std::max(s.D, f.D), Qm = std::max(s.H * s.Dh, f.H * f.Dh), Im = std::max(s.I, f.I);
    const int Nqkv_s = (s.H + 2 * s.Hkv) * s.Dh, Nqkv_f = (f.H + 2 * f.Hkv) * f.Dh;
    const int Nmax = std::max({Nqkv_s, Nqkv_f, 2 * s.I, 2 * f.I, s.D, f.D, h->head_rows, cfg->codebook_size});
    const int R = std::max(cfg->max_rows, 128);
    // partial workspace: decode needs max_parts(<=8) x 32 x N ; prefill needs rows x N
    h->ws_floats = std::max<size_t>(static_cast<size_t>(24) * kDecRows * Nmax, static_cast<size_t>(R) * Nmax);
#define TRYC(x)                  \
    do {                         \
        if ((x) != 0) {          \
            fsb_lm_destroy(h);   \
            return 1;            \
        }                        \
    } while (0)
    TRYC(dalloc(h, &h->ws, h->ws_floats));
    TRYC(dalloc(h, &h->xres_d, static_cast<size_t>(kDecRows) * Dm));
    TRYC(dalloc(h, &h->xn_d, static_cast<size_t>(kDecRows) * Dm, "hidden"));
    TRYC(dalloc(h, &h->q_d, static_cast<size_t>(kDecRows) * Qm));
    TRYC(dalloc(h, &h->attn_d, static_cast<size_t>(kDecRows) * Qm));
    TRYC(dalloc(h, &h->h_d, static_cast<size_t>(kDecRows) * Im));
    TRYC(dalloc(h, &h->proj_d, static_cast<size_t>(kDecRows) * Dm));
    TRYC(dalloc(h, &h->xres_p, static_cast<size_t>(R) * s.D));
    TRYC(dalloc(h, &h->xn_p, static_cast<size_t>(R) * s.D));
    TRYC(dalloc(h, &h->q_p, static_cast<size_t>(R) * s.H * s.Dh));
    TRYC(dalloc(h, &h->attn_p, static_cast<size_t>(R) * s.H * s.Dh));
    TRYC(dalloc(h, &h->h_p, static_cast<size_t>(R) * s.I));
    s.cache_layer_stride = static_cast<size_t>(cfg->max_batch) * s.Hkv * s.S * s.Dh;
    f.cache_layer_stride = static_cast<size_t>(cfg->max_batch) * f.Hkv * f.S * f.Dh;
    TRYC(dalloc(h, &s.kcache, s.cache_layer_stride * s.nl));
    TRYC(dalloc(h, &s.vcache, s.cache_layer_stride * s.nl));
    TRYC(dalloc(h, &f.kcache, f.cache_layer_stride * f.nl));
    TRYC(dalloc(h, &f.vcache, f.cache_layer_stride * f.nl));
    TRYC(dalloc(h, &h->cur_tok, static_cast<size_t>(cfg->max_batch) * (C + 1), "cur_tok"));
    TRYC(dalloc(h, &h->out_tokens, static_cast<size_t>(cfg->max_batch) * (C + 1) * cfg->max_frames, "out_tokens"));
    TRYC(dalloc(h, &h->n_out, cfg->max_batch, "n_out"));
    TRYC(dalloc(h, &h->pos, kDecRows, "pos"));
    TRYC(dalloc(h, &h->finished, cfg->max_batch, "finished"));
    TRYC(dalloc(h, &h->ras_window, static_cast<size_t>(cfg->max_batch) * 10, "ras_window"));
    TRYC(dalloc(h, &h->slot_state, cfg->max_batch, "slot_state"));
    TRYC(dalloc(h, &h->slot_limit, cfg->max_batch, "slot_limit"));
    TRYC(dalloc(h, &h->slot_temperature, cfg->max_batch, "slot_temperature"));
    TRYC(dalloc(h, &h->slot_top_p, cfg->max_batch, "slot_top_p"));
    TRYC(dalloc(h, &h->slot_top_k, cfg->max_batch, "slot_top_k"));
    TRYC(dalloc(h, &h->slot_seed, cfg->max_batch, "slot_seed"));
    TRYC(dalloc(h, &h->iota, kDecRows));
    TRYC(dalloc(h, &h->fpos, static_cast<size_t>(C) * kDecRows));
    TRYC(dalloc(h, &h->step, 1));
    {
        const char* ef = getenv("FSB_FLAGS");
        h->flags_on = ef && ef[0] == '1';
        if (h->flags_on) TRYC(dalloc(h, &h->flag_base, static_cast<size_t>(fsb_lm::kFlagCap)));
    }
    if (cfg->debug) {
        TRYC(dalloc(h, &h->slow_logits, static_cast<size_t>(cfg->max_batch) * h->head_rows, "slow_logits"));
        TRYC(dalloc(h, &h->fast_logits, static_cast<size_t>(C) * cfg->max_batch * cfg->codebook_size, "fast_logits"));
        TRYC(dalloc(h, &h->dbg_x, static_cast<size_t>(s.nl + 1) * kDecRows * s.D, "dbg_x"));
    }
    {
        std::vector<int> io(kDecRows), fp(static_cast<size_t>(C) * kDecRows);
        for (int i = 0; i < kDecRows; ++i) io[i] = i;
        for (int p = 0; p < C; ++p)
            for (int i = 0; i < kDecRows; ++i) fp[p * kDecRows + i] = p;
        if (cudaMemcpy(h->iota, io.data(), io.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(h->fpos, fp.data(), fp.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("fsb_lm_create: state upload failed");
            fsb_lm_destroy(h);
            return 1;
        }
    }
    // ---- GEMM plans ----
    auto build = [&](Stack& st, bool with_prefill) -> int {
        st.dec.resize(st.nl);
        if (with_prefill) st.pf.resize(st.nl);
        const int Nqkv = (st.H + 2 * st.Hkv) * st.Dh;
        for (int l = 0; l < st.nl; ++l) {
            const LayerW& lw = st.w[l];
            FSB_TRY(make_plan(h, &st.dec[l].qkv, lw.wqkv, Nqkv, st.D, h->xn_d, kDecRows, true));
            FSB_TRY(make_plan(h, &st.dec[l].wo, lw.wo, st.D, st.H * st.Dh, h->attn_d, kDecRows, true));
            FSB_TRY(make_plan(h, &st.dec[l].w13, lw.w13, 2 * st.I, st.D, h->xn_d, kDecRows, true));
            FSB_TRY(make_plan(h, &st.dec[l].w2, lw.w2, st.D, st.I, h->h_d, kDecRows, true));
            if (with_prefill) {
                FSB_TRY(make_plan(h, &st.pf[l].qkv, lw.wqkv, Nqkv, st.D, h->xn_p, R, false));
                FSB_TRY(make_plan(h, &st.pf[l].wo, lw.wo, st.D, st.H * st.Dh, h->attn_p, R, false));
                FSB_TRY(make_plan(h, &st.pf[l].w13, lw.w13, 2 * st.I, st.D, h->xn_p, R, false));
                FSB_TRY(make_plan(h, &st.pf[l].w2, lw.w2, st.D, st.I, h->h_p, R, false));
            }
        }
        return 0;
    };
    TRYC(build(s, true));
    TRYC(build(f, false));
    {
        const char* ep = getenv("FSB_PERSISTENT");
        h->persistent = ep && ep[0] == '1';
        const char* es = getenv("FSB_PK_STAGES");
        if (es) h->pk_stages = atoi(es);
    }
    {
        const char* ef = getenv("FSB_FUSED_ATTN");
        h->fused_prep_attn = ef && (ef[0] == '1' || ef[0] == '2');  // measured slower than qkv_prep + attn (6.45 vs 6.22 ms/frame): off
        h->fused_fast_only = ef && ef[0] == '2';  // 2: only the fast stack (10-position KV)
        TRYC(pk_init());
    }
    if (h->persistent) {
        TRYC(dalloc(h, &h->pk_bar, 2));
        if (getenv("FSB_PK_TRACE")) TRYC(dalloc(h, &h->pk_trace, 1024, "pk_trace"));
        auto build_pk = [&](Stack& st, const bf16* final_norm) -> int {
            st.pk.resize(st.nl);
            std::vector<PkLayer> host(st.nl);
            const int Nqkv = (st.H + 2 * st.Hkv) * st.Dh;
            auto fill = [&](PkGemm& g, GemmPlan& p) {
                g.tmA = p.tmA;
                g.tmB = p.tmB;
                g.sched = reinterpret_cast<const int4*>(p.sched_dev);
                g.cta_items = p.cta_items_dev;
                g.nparts = p.nparts_dev;
                g.n_out = p.p.ws_ld;
                g.max_parts = p.max_parts;
                g.kblocks = p.p.kb_per_tap;
                g.slot_stride = p.p.ws_slot_stride;
            };
            for (int l = 0; l < st.nl; ++l) {
                const LayerW& lw = st.w[l];
                FSB_TRY(make_plan(h, &st.pk[l].qkv, lw.wqkv, Nqkv, st.D, h->xn_d, kDecRows, true, 1));
                FSB_TRY(make_plan(h, &st.pk[l].wo, lw.wo, st.D, st.H * st.Dh, h->attn_d, kDecRows, true, 1));
                FSB_TRY(make_plan(h, &st.pk[l].w13, lw.w13, 2 * st.I, st.D, h->xn_d, kDecRows, true, 1));
                FSB_TRY(make_plan(h, &st.pk[l].w2, lw.w2, st.D, st.I, h->h_d, kDecRows, true, 1));
                PkLayer& L = host[l];
                memset(&L, 0, sizeof(L));
                fill(L.qkv, st.pk[l].qkv);
                fill(L.wo, st.pk[l].wo);
                fill(L.w13, st.pk[l].w13);
                fill(L.w2, st.pk[l].w2);
                L.bqkv = lw.bqkv;
                L.q_norm = st.qk_norm ? lw.q_norm : nullptr;
                L.k_norm = st.qk_norm ? lw.k_norm : nullptr;
                L.bo = lw.bo;
                L.ffn_norm = lw.ffn_norm;
                L.next_norm = (l + 1 < st.nl) ? st.w[l + 1].attn_norm : final_norm;
                L.kcache = st.kcache + l * st.cache_layer_stride;
                L.vcache = st.vcache + l * st.cache_layer_stride;
            }
            FSB_TRY(dalloc(h, &st.pk_layers, static_cast<size_t>(st.nl)));
            FSB_CUDA(cudaMemcpy(st.pk_layers, host.data(), host.size() * sizeof(PkLayer), cudaMemcpyHostToDevice));
            return 0;
        };
        TRYC(build_pk(s, h->norm_w));
        TRYC(build_pk(f, h->fast_norm_w));
    }
    TRYC(make_plan(h, &h->head_plan, h->head_w, h->head_rows, s.D, h->xn_d, kDecRows, true));
    TRYC(make_plan(h, &h->fast_out_plan, h->fast_out_w, cfg->codebook_size, f.D, h->xn_d, kDecRows, true));
    if (h->has_proj) TRYC(make_plan(h, &h->proj_plan, h->fast_proj_w, f.D, s.D, h->xn_d, kDecRows, true));
#undef TRYC
    h->named["ws"] = {h->ws, h->ws_floats * 4};
    *out = h;
    return 0;
}

void fsb_lm_destroy(fsb_lm* h) {
    if (!h) return;
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    auto free_plans = [](Stack& st) {
        for (auto& p : st.dec) { gemm_plan_free(&p.qkv); gemm_plan_free(&p.wo); gemm_plan_free(&p.w13); gemm_plan_free(&p.w2); }
        for (auto& p : st.pf) { gemm_plan_free(&p.qkv); gemm_plan_free(&p.wo); gemm_plan_free(&p.w13); gemm_plan_free(&p.w2); }
        for (auto& p : st.pk) { gemm_plan_free(&p.qkv); gemm_plan_free(&p.wo); gemm_plan_free(&p.w13); gemm_plan_free(&p.w2); }
    };
    free_plans(h->slow);
    free_plans(h->fast);
    gemm_plan_free(&h->head_plan);
    gemm_plan_free(&h->fast_out_plan);
    if (h->has_proj) gemm_plan_free(&h->proj_plan);
    for (void* p : h->owned) cudaFree(p);
    delete h;
}

int fsb_lm_reset(fsb_lm* h, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const fsb_lm_config& c = h->cfg;
    FSB_CUDA(cudaMemsetAsync(h->n_out, 0, c.max_batch * 4, st));
    FSB_CUDA(cudaMemsetAsync(h->finished, 0, c.max_batch * 4, st));
    FSB_CUDA(cudaMemsetAsync(h->slot_state, 0, c.max_batch * 4, st));
    FSB_CUDA(cudaMemsetAsync(h->ras_window, 0, static_cast<size_t>(c.max_batch) * 10 * 4, st));
    FSB_CUDA(cudaMemsetAsync(h->pos, 0, kDecRows * 4, st));
    FSB_CUDA(cudaMemsetAsync(h->step, 0, 8, st));
    return 0;
}

int fsb_lm_prefill(fsb_lm* h, const int32_t* d_tokens, const int32_t* d_row_slot, const int32_t* d_row_pos,
                   int rows, const int32_t* d_last_rows, const int32_t* d_slots, int nseq, int do_sample,
                   const fsb_sampling* sp, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const fsb_lm_config& c = h->cfg;
    FSB_CHECK(rows >= 1 && rows <= std::max(c.max_rows, 128), "prefill: rows=%d exceeds max_rows=%d", rows, c.max_rows);
    FSB_CHECK(nseq >= 1 && nseq <= c.max_batch, "prefill: nseq=%d out of range", nseq);
    Stack& s = h->slow;
    EmbedArgs ea{};
    ea.tokens = d_tokens;
    ea.emb = h->emb; ea.cb_emb = h->cb_emb; ea.x = h->xres_p;
    ea.rows = rows; ea.D = s.D; ea.C = c.num_codebooks; ea.cs = c.codebook_size; ea.vocab = c.vocab_size;
    ea.sem_begin = c.semantic_begin_id; ea.sem_end = c.semantic_end_id;
    ea.scale = c.scale_codebook_embeddings;
    FSB_TRY(launch_embed(ea, st));
    ResidNormArgs r{};
    r.x_in = h->xres_p;
    r.norm_w = s.w[0].attn_norm; r.n_out = h->xn_p;
    r.rows = rows; r.D = s.D; r.eps = c.norm_eps;
    FSB_TRY(launch_resid_norm(r, st));
    RowCtx ctx{rows, d_row_slot, d_row_pos, h->xres_p, h->xn_p, h->q_p, h->attn_p, h->h_p, false};
    FSB_TRY(run_stack(h, s, ctx, h->norm_w, false, nullptr, st));
    if (!do_sample) return 0;
    FSB_CHECK(sp != nullptr || h->slot_control, "prefill: sampling parameters required");
    const fsb_sampling sp_none{1.f, 1.f, 1, 0};
    if (sp == nullptr) sp = &sp_none;
    // last-token rows -> decode workspaces (llama.py:447-448 keeps only the last position)
    FSB_TRY(launch_gather_rows(h->xn_p, d_last_rows, h->xn_d, nseq, s.D, st));
    FSB_TRY(launch_gather_rows(h->xres_p, d_last_rows, h->xres_d, nseq, s.D, st));
    // the reference resets the RAS window per generate() call and prefill uses no RAS
    FSB_TRY(flags_begin(h, st));
    return run_frame_tail(h, nseq, d_slots, false, d_last_rows, d_row_pos, *sp, st);
}

int fsb_lm_set_slot_control(fsb_lm* h, int enable) {
    FSB_CHECK(h != nullptr, "set_slot_control: null handle");
    h->slot_control = enable != 0;
    return 0;
}

int fsb_lm_decode(fsb_lm* h, int batch, int nframes, const fsb_sampling* sp, int use_graph, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    FSB_CHECK(batch >= 1 && batch <= h->cfg.max_batch, "decode: batch=%d out of range", batch);
    FSB_CHECK(sp != nullptr || h->slot_control, "decode: sampling parameters required");
    const fsb_sampling sp_none{1.f, 1.f, 1, 0};
    if (sp == nullptr || h->slot_control) sp = &sp_none;  // per-slot parameters: one graph for every mix of requests
    if (!use_graph) {
        if (h->graph_exec) {  // debug copies are only recorded outside graphs
            cudaGraphExecDestroy(h->graph_exec);
            h->graph_exec = nullptr;
            h->graph_batch = -1;
        }
        for (int i = 0; i < nframes; ++i) FSB_TRY(decode_one_frame(h, batch, *sp, st));
        return 0;
    }
    const bool same = h->graph_exec && h->graph_batch == batch && h->graph_lcap == h->ctx_lcap &&
                      h->graph_slot_control == h->slot_control &&
                      memcmp(&h->graph_sampling, sp, sizeof(fsb_sampling)) == 0;
    if (!same) {
        if (h->graph_exec) {
            cudaGraphExecDestroy(h->graph_exec);
            h->graph_exec = nullptr;
        }
        // capture on a private stream so the caller's stream mode does not matter
        cudaStream_t cs;
        FSB_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        cudaGraphExec_t sentinel = reinterpret_cast<cudaGraphExec_t>(1);
        h->graph_exec = sentinel;  // suppress debug copies while capturing
        cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
        int rc = 1;
        const int launches_before = g_launch_count;
        if (e == cudaSuccess) {
            rc = decode_one_frame(h, batch, *sp, cs);
            e = cudaStreamEndCapture(cs, &g);
        }
        h->graph_kernels = g_launch_count - launches_before;
        g_launch_count = launches_before;  // captured, not launched
        h->graph_exec = nullptr;
        if (e != cudaSuccess || rc != 0 || g == nullptr) {
            if (rc == 0) set_error("decode: graph capture failed: %s", cudaGetErrorString(e));
            cudaStreamDestroy(cs);
            if (g) cudaGraphDestroy(g);
            return 1;
        }
        cudaGraphExec_t ge = nullptr;
        e = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        cudaStreamDestroy(cs);
        if (e != cudaSuccess) {
            set_error("decode: cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
            return 1;
        }
        h->graph_exec = ge;
        h->graph_batch = batch;
        h->graph_lcap = h->ctx_lcap;
        h->graph_slot_control = h->slot_control;
        h->graph_sampling = *sp;
    }
    for (int i = 0; i < nframes; ++i) {
        FSB_CUDA(cudaGraphLaunch(h->graph_exec, st));
        g_launch_count += h->graph_kernels;
    }
    return 0;
}

int fsb_lm_trace_gemms(fsb_lm* h, unsigned long long* d_trace, int max_launches, int* grid_out, void* stream) {
    // One slow layer's four GEMMs repeated, each CTA recording globaltimer at start / after the
    // programmatic-dependency wait / at exit: shows whether consecutive kernels really overlap.
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Stack& s = h->slow;
    int id = 0;
    for (int l = 0; l < s.nl && id + 4 <= max_launches; ++l) {
        GemmPlan* plans[4] = {&s.dec[l].qkv, &s.dec[l].wo, &s.dec[l].w13, &s.dec[l].w2};
        for (int k = 0; k < 4; ++k) {
            GemmPlan q = *plans[k];
            q.p.trace = d_trace;
            q.p.trace_id = id++;
            FSB_TRY(gemm_launch(q, st));
        }
    }
    if (grid_out) *grid_out = static_cast<int>(s.dec[0].qkv.grid.x);
    return id;
}

int fsb_lm_bench_gemms(fsb_lm* h, int reps, double* weight_bytes_per_rep, int* launches_per_rep, void* stream) {
    // Every weight-streaming GEMM of one decode frame (36 slow layers x 4, head, 10 fast passes x
    // (4 layers x 4 + head)) back to back, without the glue kernels: the measured stream of the
    // dominant kernel for the roofline line of bench.py. Operands are whatever the workspaces hold.
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const fsb_lm_config& c = h->cfg;
    double bytes = 0;
    int launches = 0;
    auto run = [&](const GemmPlan& p, double wbytes) -> int {
        bytes += wbytes;
        ++launches;
        return gemm_launch(p, st);
    };
    for (int r = 0; r < reps; ++r) {
        bytes = 0;
        launches = 0;
        auto stack = [&](Stack& s, bool skip_tail) -> int {
            const double nq = static_cast<double>((s.H + 2 * s.Hkv) * s.Dh) * s.D * 2;
            const double no = static_cast<double>(s.D) * s.H * s.Dh * 2;
            const double n13 = 2.0 * s.I * s.D * 2, n2 = static_cast<double>(s.D) * s.I * 2;
            for (int l = 0; l < s.nl; ++l) {
                FSB_TRY(run(s.dec[l].qkv, nq));
                if (skip_tail && l == s.nl - 1) break;
                FSB_TRY(run(s.dec[l].wo, no));
                FSB_TRY(run(s.dec[l].w13, n13));
                FSB_TRY(run(s.dec[l].w2, n2));
            }
            return 0;
        };
        FSB_TRY(stack(h->slow, false));
        FSB_TRY(run(h->head_plan, static_cast<double>(h->head_rows) * c.dim * 2));
        for (int p = 0; p < c.num_codebooks; ++p) {
            FSB_TRY(stack(h->fast, p == 0));
            if (p > 0) FSB_TRY(run(h->fast_out_plan, static_cast<double>(c.codebook_size) * c.fast_dim * 2));
        }
    }
    if (weight_bytes_per_rep) *weight_bytes_per_rep = bytes;
    if (launches_per_rep) *launches_per_rep = launches;
    return 0;
}

int fsb_lm_set_context_bound(fsb_lm* h, int max_positions) {
    // The attention kernel keeps one fp32 score per live position in shared memory; sizing that buffer by
    // the KV capacity would cap max_seq_len at ~12k. The host knows an upper bound of every slot's length
    // (prompt + frames so far): round it up to a 1024 bucket (the decode graph is re-captured per bucket).
    FSB_CHECK(max_positions >= 0, "set_context_bound: negative bound");
    int b = ((max_positions + 1023) / 1024) * 1024;
    if (b > h->cfg.kv_len) b = h->cfg.kv_len;
    h->ctx_lcap = b;
    return 0;
}

int fsb_lm_buffer(fsb_lm* h, const char* name, void** d_ptr, size_t* bytes) {
    auto it = h->named.find(name);
    FSB_CHECK(it != h->named.end(), "fsb_lm_buffer: unknown buffer '%s'", name);
    if (d_ptr) *d_ptr = it->second.first;
    if (bytes) *bytes = it->second.second;
    return 0;
}

}  // extern "C"
