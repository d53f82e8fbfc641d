// Persistent per-stack decode kernel (lm_persist.cu): device-side descriptors.
#pragma once
#include "common.cuh"

namespace fsb {

struct PkGemm {            // one weight-streaming GEMM inside the persistent kernel
    CUtensorMap tmA, tmB;  // weights [n_out, K] (box 128 rows), activations [32, K] (box 32 rows)
    const int4* sched;     // stream-K items {tile, kb_begin, kb_end, slot}
    const int* cta_items;  // [grid + 1]
    const int* nparts;     // partial count per 128-feature tile
    int n_out;
    int max_parts;
    int kblocks;  // k-blocks per tile
    int pad;
    long long slot_stride;
};

struct PkLayer {
    PkGemm qkv, wo, w13, w2;
    const __nv_bfloat16 *bqkv, *q_norm, *k_norm, *bo, *ffn_norm, *next_norm;
    __nv_bfloat16 *kcache, *vcache;
};

struct PkArgs {
    const PkLayer* layers;
    int nl, rows, D, H, Hkv, Dh, I, S;
    float eps;
    int bf16_math, qk_norm, kv_only_last, stages;
    int l2_prefetch;  // weight tiles per CTA prefetched into L2 while the producer waits for a barrier
    const int* row_seq;
    const int* row_pos;
    const __nv_bfloat16* freqs;
    __nv_bfloat16 *xres, *xn, *attn, *hbuf;
    float* ws;
    unsigned* bar;  // [2]: phase counter, exit counter
    unsigned long long* trace;  // optional: CTA 0 globaltimer stamps, 4 per GEMM (epilogue done, barrier, consumer done, barrier)
    int trace_max;
};

int pk_init();
size_t pk_scratch_bytes(int H, int Hkv, int Dh, int S);
int launch_stack_persistent(const PkArgs& A, int grid, cudaStream_t st);
// fused qkv post-processing + KV append + attention for decode rows; returns -1 if the head shape has no instance
int launch_prep_attn(const PkArgs& A, const PkLayer& L, int kv_only, cudaStream_t st);

}  // namespace fsb
