"""Model geometries used by bench / smoke / tests.

S2-Pro's config.json is not available offline; the geometry below is the one SURVEY.md §8 derives
(Qwen3-4B slow stack + 4-layer fast stack, 10 codebooks x 4096) and is a parameter everywhere — a real
checkpoint's config.json is parsed by BaseModelArgs.from_pretrained instead.
"""
from __future__ import annotations

from .models.text2semantic.llama import DualARModelArgs

S2PRO_IM_END_ID = 151645
S2PRO_TEXT_VOCAB = 151643


def s2pro_args(max_seq_len: int = 4096, **over) -> DualARModelArgs:
    kw = dict(
        model_type="dual_ar", vocab_size=155776, n_layer=36, n_head=32, dim=2560, intermediate_size=9728,
        n_local_heads=8, head_dim=128, rope_base=1e6, norm_eps=1e-6, max_seq_len=max_seq_len,
        tie_word_embeddings=True, attention_qk_norm=True, codebook_size=4096, num_codebooks=10,
        semantic_begin_id=151678, semantic_end_id=155773, scale_codebook_embeddings=True,
        norm_fastlayer_input=True, n_fast_layer=4, fast_dim=2560, fast_n_head=32, fast_n_local_heads=8,
        fast_head_dim=128, fast_intermediate_size=9728, fast_attention_qk_norm=False,
    )
    kw.update(over)
    return DualARModelArgs(**kw)
