/* fishb200 — C-ABI of the B200-native Fish-Speech inference hot path.
 *
 * The reference (fishaudio/fish-speech @ 3dd1f85c) is pure Python/PyTorch and has no FFI of its own:
 * its boundary for this path is the Python API listed in SURVEY.md §8(b).  This header is what a
 * maintainer's `ctypes` stub binds instead of the eager PyTorch modules; every entry point names the
 * reference symbol (file:line under the reference checkout) whose work it replaces.  See
 * INTEGRATION.md for the reference-side binding.
 *
 * Conventions: plain pointers and sizes only (no torch types).  All `d_` pointers are CUDA device
 * pointers owned by the caller and must stay valid for the lifetime of the handle they are given to;
 * `h_` pointers are host memory.  Every function returns 0 on success, non-zero on failure with the
 * message available from fsb_last_error() (per host thread).  `stream` is a cudaStream_t passed as
 * void*.  Handles are not thread-safe; distinct handles may be driven from different host threads
 * (the reference runs the LM worker thread and the codec caller concurrently:
 * fish_speech/models/text2semantic/inference.py:748-799).
 */
#ifndef FISHB200_H
#define FISHB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* fsb_last_error(void);
/* Number of kernels this host thread has launched through the library (bench.py: gpu_launches). */
long long fsb_launch_count(void);
int fsb_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* cudaMemcpy helpers so Python hosts need no CUDA bindings of their own. */
int fsb_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream);
int fsb_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dual-AR text2semantic transformer
 *   replaces DualARTransformer.forward_generate / forward_generate_fast
 *     (fish_speech/models/text2semantic/llama.py:390-466, 799-828) and
 *   decode_one_token_ar / decode_n_tokens / the prefill call in generate
 *     (fish_speech/models/text2semantic/inference.py:96-181, 184-238, 322-335).
 * ---------------------------------------------------------------------------------------------- */
typedef struct fsb_lm fsb_lm;

typedef struct {
    /* slow (time-axis) stack: BaseModelArgs, llama.py:28-73 */
    int dim, n_layer, n_head, n_kv_head, head_dim, intermediate;
    /* fast (codebook-axis) stack: DualARModelArgs, llama.py:156-193 */
    int fast_dim, n_fast_layer, fast_n_head, fast_n_kv_head, fast_head_dim, fast_intermediate;
    int vocab_size, codebook_size, num_codebooks;
    int semantic_begin_id, semantic_end_id, im_end_id;
    float norm_eps;
    int qk_norm, fast_qk_norm;           /* attention_qk_norm / fast_attention_qk_norm */
    int scale_codebook_embeddings;       /* llama.py:416-420 */
    int norm_fastlayer_input;            /* llama.py:459-461 */
    int max_batch;                       /* resident sequences ("slots"), <= 32 */
    int kv_len;                          /* KV-cache positions per sequence */
    int max_rows;                        /* token rows per prefill pass */
    int max_frames;                      /* generated frames kept per sequence */
    int debug;                           /* keep logits / per-layer activations for tests */
} fsb_lm_config;

/* One TransformerBlock (llama.py:831-987). bf16, row-major [out_features, in_features]. */
typedef struct {
    const void* d_attn_norm;  /* [dim] */
    const void* d_wqkv;       /* [(H+2Hkv)*Dh, dim] */
    const void* d_bqkv;       /* optional */
    const void* d_q_norm;     /* [Dh] optional */
    const void* d_k_norm;     /* [Dh] optional */
    const void* d_wo;         /* [dim, H*Dh] */
    const void* d_bo;         /* optional */
    const void* d_ffn_norm;   /* [dim] */
    const void* d_w13;        /* [2*I, dim]: rows [0,I) = w1, rows [I,2I) = w3 */
    const void* d_w2;         /* [dim, I] */
} fsb_lm_layer;

typedef struct {
    const void* d_embeddings;           /* [vocab, dim] */
    const void* d_codebook_embeddings;  /* [num_codebooks*codebook_size, dim] */
    const void* d_norm;                 /* [dim] */
    const void* d_head;                 /* [head_rows, dim]: the rows of the (tied) LM head that the
                                           semantic_logit_bias leaves selectable — semantic ids in
                                           order, then <|im_end|> (inference.py:308-320) */
    int head_rows;                      /* (semantic_end_id - semantic_begin_id + 1) + 1 */
    const void* d_freqs;                /* bf16 [kv_len, head_dim/2, 2] (llama.py:1004-1023) */
    const fsb_lm_layer* layers;         /* host array [n_layer] */
    const void* d_fast_embeddings;      /* [codebook_size, fast_dim] */
    const void* d_fast_norm;            /* [fast_dim] */
    const void* d_fast_output;          /* [codebook_size, fast_dim] */
    const void* d_fast_freqs;           /* bf16 [num_codebooks, fast_head_dim/2, 2] */
    const void* d_fast_proj_w;          /* optional [fast_dim, dim] (llama.py:665-668) */
    const void* d_fast_proj_b;          /* optional [fast_dim] */
    const fsb_lm_layer* fast_layers;    /* host array [n_fast_layer] */
} fsb_lm_weights;

/* sample() parameters (inference.py:54-93).  temperature / top_p must already be rounded to the
 * model dtype the way the reference builds them (torch.tensor(v, dtype=bf16), inference.py:303-304). */
typedef struct {
    float temperature;
    float top_p;
    int top_k;
    unsigned long long seed;
} fsb_sampling;

int fsb_lm_create(const fsb_lm_config* cfg, const fsb_lm_weights* w, fsb_lm** out);
void fsb_lm_destroy(fsb_lm* h);

/* Prefill = decode_one_token_ar on whole prompts (inference.py:322-335), for `nseq` sequences packed
 * as `rows` token rows: d_tokens[row][0] = token id, [1..C] = codes; d_row_slot / d_row_pos give the
 * sequence slot and position of each row.  If do_sample != 0 the first frame of each sequence is
 * sampled from its last row (d_last_rows[k], slot d_slots[k]) and stored as output frame 0. */
int fsb_lm_prefill(fsb_lm* h, const int32_t* d_tokens, const int32_t* d_row_slot,
                   const int32_t* d_row_pos, int rows, const int32_t* d_last_rows,
                   const int32_t* d_slots, int nseq, int do_sample, const fsb_sampling* s,
                   void* stream);

/* decode_n_tokens (inference.py:184-238): `nframes` more frames for slots [0, batch). One frame =
 * slow step + constrained sampling (+RAS) + num_codebooks fast steps; captured as a CUDA graph when
 * use_graph != 0. The per-frame <|im_end|> test is kept on the device (finished flags). */
int fsb_lm_decode(fsb_lm* h, int batch, int nframes, const fsb_sampling* s, int use_graph,
                  void* stream);

/* Reset per-slot generation state (frame counters, RAS window, finished flags). */
int fsb_lm_reset(fsb_lm* h, void* stream);

/* Named device buffers of the handle, for reading results and for tests:
 *   "out_tokens" int32 [max_batch][C+1][max_frames], "n_out" int32 [max_batch], "pos" int32 [max_batch],
 *   "finished" int32 [max_batch], "cur_tok" int32 [max_batch][C+1], "ras_window" int32 [max_batch][10],
 *   debug only: "slow_logits" f32 [max_batch][head_rows], "fast_logits" f32 [C][max_batch][codebook_size],
 *   "hidden" bf16 [32][dim], "dbg_x" bf16 [n_layer+1][32][dim] */
int fsb_lm_buffer(fsb_lm* h, const char* name, void** d_ptr, size_t* bytes);

/* Measurement hook for bench.py: launch every weight-streaming GEMM of one decode frame `reps` times
 * (no glue kernels); returns the algorithmic weight bytes and launch count of one repetition. */
int fsb_lm_bench_gemms(fsb_lm* h, int reps, double* weight_bytes_per_rep, int* launches_per_rep,
                       void* stream);

/* ------------------------------------------------------------------------------------------------
 * Unit-test hooks (used by tests/ only)
 * ---------------------------------------------------------------------------------------------- */
/* out[j][i] (fp32, ld = m) = sum_k A[i][k] * B[j][k]; A [m,k], B [n,k] bf16 row-major.
 * bn in {32,64,128,256}; streamk_ctas > 0 uses the decode-style stream-K schedule + partial sums. */
int fsb_op_gemm(const void* d_a, const void* d_b, float* d_out, int m, int n, int k, int bn,
                int streamk_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FISHB200_H */
