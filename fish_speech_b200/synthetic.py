"""Seeded synthetic checkpoints (reference state_dict naming) generated directly on the GPU.

There are no real S2-Pro weights offline; bench.py / smoke use random weights of the real architecture
(throughput does not depend on the values). LM: N(0, 0.02) like the reference's _init_weights
(llama.py:468-477) with wider output heads; codec: fan-in scaled so activations stay O(1)."""
from __future__ import annotations

import math

import torch


def lm_state_dict(cfg, device, seed: int = 1234, head_gain: float = 4.0) -> dict:
    g = torch.Generator(device=device).manual_seed(seed)
    std = cfg.initializer_range

    def lin(o, i, gain=1.0):
        return (torch.randn(o, i, generator=g, device=device, dtype=torch.float32) * (std * gain)).to(torch.bfloat16)

    def nrm(n):
        return (1.0 + 0.1 * torch.randn(n, generator=g, device=device)).to(torch.bfloat16)

    w = {"embeddings.weight": lin(cfg.vocab_size, cfg.dim, head_gain),
         "codebook_embeddings.weight": lin(cfg.codebook_size * cfg.num_codebooks, cfg.dim)}

    def block(prefix, dim, nh, nkv, hd, inter, qk):
        w[f"{prefix}.attention.wqkv.weight"] = lin((nh + 2 * nkv) * hd, dim)
        w[f"{prefix}.attention.wo.weight"] = lin(dim, nh * hd)
        if qk:
            w[f"{prefix}.attention.q_norm.weight"] = nrm(hd)
            w[f"{prefix}.attention.k_norm.weight"] = nrm(hd)
        w[f"{prefix}.feed_forward.w1.weight"] = lin(inter, dim)
        w[f"{prefix}.feed_forward.w3.weight"] = lin(inter, dim)
        w[f"{prefix}.feed_forward.w2.weight"] = lin(dim, inter)
        w[f"{prefix}.ffn_norm.weight"] = nrm(dim)
        w[f"{prefix}.attention_norm.weight"] = nrm(dim)

    for l in range(cfg.n_layer):
        block(f"layers.{l}", cfg.dim, cfg.n_head, cfg.n_local_heads, cfg.head_dim, cfg.intermediate_size,
              cfg.attention_qk_norm)
    w["norm.weight"] = nrm(cfg.dim)
    if cfg.fast_dim != cfg.dim:
        w["fast_project_in.weight"] = lin(cfg.fast_dim, cfg.dim)
        w["fast_project_in.bias"] = torch.zeros(cfg.fast_dim, device=device, dtype=torch.bfloat16)
    w["fast_embeddings.weight"] = lin(cfg.codebook_size, cfg.fast_dim)
    for l in range(cfg.n_fast_layer):
        block(f"fast_layers.{l}", cfg.fast_dim, cfg.fast_n_head, cfg.fast_n_local_heads, cfg.fast_head_dim,
              cfg.fast_intermediate_size, cfg.fast_attention_qk_norm)
    w["fast_norm.weight"] = nrm(cfg.fast_dim)
    w["fast_output.weight"] = lin(cfg.codebook_size, cfg.fast_dim, head_gain)
    return w


def codec_state_dict(cfg, device, seed: int = 1234) -> dict:
    """cfg: models.dac.modded_dac.CodecConfig. Keys follow the reference module tree
    (modded_dac.py:670-801, rvq.py:204-291; weight-norm as parametrizations.weight.original0/1)."""
    g = torch.Generator(device=device).manual_seed(seed)
    w: dict = {}

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, device=device) * std

    def conv(prefix, cout, cin, k, wn=True, gain=1.0):
        v = rn(cout, cin, k, std=gain / math.sqrt(cin * k))
        if wn:
            w[f"{prefix}.parametrizations.weight.original1"] = v
            w[f"{prefix}.parametrizations.weight.original0"] = v.flatten(1).norm(dim=1).view(cout, 1, 1)
        else:
            w[f"{prefix}.weight"] = v
        w[f"{prefix}.bias"] = rn(cout, std=0.02)

    def convT(prefix, cin, cout, k, stride, wn=True):
        v = rn(cin, cout, k, std=1.0 / math.sqrt(cin * k / stride))
        if wn:
            w[f"{prefix}.parametrizations.weight.original1"] = v
            w[f"{prefix}.parametrizations.weight.original0"] = v.flatten(1).norm(dim=1).view(cin, 1, 1)
        else:
            w[f"{prefix}.weight"] = v
        w[f"{prefix}.bias"] = rn(cout, std=0.02)

    def snake(prefix, c):
        w[f"{prefix}.alpha"] = (1.0 + 0.3 * rn(1, c, 1)).abs().clamp_min(0.2)

    def res_unit(prefix, c):
        snake(f"{prefix}.block.0", c)
        conv(f"{prefix}.block.1.conv", c, c, 7, gain=0.5)
        snake(f"{prefix}.block.2", c)
        conv(f"{prefix}.block.3.conv", c, c, 1, gain=0.5)

    def tfm(prefix, t):
        for l in range(t.n_layer):
            p = f"{prefix}.layers.{l}"
            w[f"{p}.attention.wqkv.weight"] = rn(3 * t.n_head * t.head_dim, t.dim, std=1 / math.sqrt(t.dim))
            w[f"{p}.attention.wo.weight"] = rn(t.dim, t.n_head * t.head_dim, std=1 / math.sqrt(t.dim))
            w[f"{p}.feed_forward.w1.weight"] = rn(t.intermediate_size, t.dim, std=1 / math.sqrt(t.dim))
            w[f"{p}.feed_forward.w3.weight"] = rn(t.intermediate_size, t.dim, std=1 / math.sqrt(t.dim))
            w[f"{p}.feed_forward.w2.weight"] = rn(t.dim, t.intermediate_size, std=1 / math.sqrt(t.intermediate_size))
            for nme in ("ffn_norm.weight", "attention_norm.weight"):
                w[f"{p}.{nme}"] = 1 + 0.1 * rn(t.dim)
            for nme in ("attention_layer_scale.gamma", "ffn_layer_scale.gamma"):
                w[f"{p}.{nme}"] = 0.3 + 0.1 * rn(t.dim)
        w[f"{prefix}.norm.weight"] = 1 + 0.1 * rn(t.dim)

    def convnext(prefix, c):
        w[f"{prefix}.dwconv.conv.weight"] = rn(c, 1, 7, std=1 / math.sqrt(7))
        w[f"{prefix}.dwconv.conv.bias"] = rn(c, std=0.02)
        w[f"{prefix}.norm.weight"] = 1 + 0.1 * rn(c)
        w[f"{prefix}.norm.bias"] = rn(c, std=0.02)
        w[f"{prefix}.pwconv1.weight"] = rn(4 * c, c, std=1 / math.sqrt(c))
        w[f"{prefix}.pwconv1.bias"] = rn(4 * c, std=0.02)
        w[f"{prefix}.pwconv2.weight"] = rn(c, 4 * c, std=1 / math.sqrt(4 * c))
        w[f"{prefix}.pwconv2.bias"] = rn(c, std=0.02)
        w[f"{prefix}.gamma"] = 0.3 + 0.1 * rn(c)

    d = cfg.encoder_dim
    conv("encoder.block.0.conv", d, 1, 7)
    for i, (stride, ntl) in enumerate(zip(cfg.encoder_rates, cfg.encoder_transformer_layers)):
        d *= 2
        p = f"encoder.block.{i + 1}.block"
        for j in range(3):
            res_unit(f"{p}.{j}", d // 2)
        snake(f"{p}.3", d // 2)
        conv(f"{p}.4.conv", d, d // 2, 2 * stride)
        if ntl > 0:
            tfm(f"{p}.5", cfg.enc_tfm(d, ntl))
    nb = len(cfg.encoder_rates) + 1
    snake(f"encoder.block.{nb}", d)
    conv(f"encoder.block.{nb + 1}.conv", cfg.latent_dim, d, 3)
    D = cfg.latent_dim

    def vq(prefix, size):
        conv(f"{prefix}.in_proj", cfg.codebook_dim, D, 1)
        conv(f"{prefix}.out_proj", D, cfg.codebook_dim, 1)
        w[f"{prefix}.codebook.weight"] = rn(size, cfg.codebook_dim)

    vq("quantizer.semantic_quantizer.quantizers.0", cfg.semantic_codebook_size)
    for i in range(cfg.n_codebooks):
        vq(f"quantizer.quantizer.quantizers.{i}", cfg.codebook_size)
    for i, f_ in enumerate(cfg.downsample_factor):
        conv(f"quantizer.downsample.{i}.0.conv", D, D, f_, wn=False)
        convnext(f"quantizer.downsample.{i}.1", D)
    for i, f_ in enumerate(reversed(cfg.downsample_factor)):
        convT(f"quantizer.upsample.{i}.0.conv", D, D, f_, f_, wn=False)
        convnext(f"quantizer.upsample.{i}.1", D)
    tfm("quantizer.pre_module", cfg.quant_tfm)
    tfm("quantizer.post_module", cfg.quant_tfm)
    c = cfg.decoder_dim
    conv("decoder.model.0.conv", c, D, 7)
    cout = c
    for i, stride in enumerate(cfg.decoder_rates):
        cin, cout = cfg.decoder_dim // 2 ** i, cfg.decoder_dim // 2 ** (i + 1)
        p = f"decoder.model.{i + 1}.block"
        snake(f"{p}.0", cin)
        convT(f"{p}.1.conv", cin, cout, 2 * stride, stride)
        for j in range(3):
            res_unit(f"{p}.{2 + j}", cout)
    n = len(cfg.decoder_rates) + 1
    snake(f"decoder.model.{n}", cout)
    conv(f"decoder.model.{n + 1}.conv", 1, cout, 7, gain=0.12)
    return w
