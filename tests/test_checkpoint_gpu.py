"""Checkpoint loading through the reference-shaped entry points, on the GPU:
`init_model(checkpoint_dir)` (inference.py:362-392 -> llama.py:480-594: config.json + model.pth with the
"model." key prefix a Lightning checkpoint carries) and `dac.inference.load_model(config, codec.pth)`
(dac/inference.py:23-47: "generator."-prefixed state dict). The loaded models must behave exactly like
the ones built directly from the same tensors (integer tokens identical to the reference golden; waveform
bitwise equal)."""
import dataclasses
import json
from pathlib import Path

import numpy as np
import pytest
import torch
import yaml

from oracle import codec_oracle as CO
from tests.lm_util import load_golden, model_args

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def test_init_model_from_checkpoint_dir_reproduces_reference_golden(tmp_path):
    from fish_speech_b200.models.text2semantic import inference as inf

    cfg, w, z = load_golden(GOLD / "lm_tiny_greedy.npz")
    args = model_args(cfg)
    conf = dataclasses.asdict(args)
    conf["im_end_id"] = cfg.im_end_id  # no tokenizer files in a synthetic checkpoint
    (tmp_path / "config.json").write_text(json.dumps(conf))
    torch.save({"state_dict": {f"model.{k}": v for k, v in w.items()}}, tmp_path / "model.pth")

    model, decode_one_token = inf.init_model(str(tmp_path), "cuda", torch.bfloat16, compile=False)
    assert decode_one_token is inf.decode_one_token_ar
    assert model.config.num_codebooks == cfg.num_codebooks and next(model.parameters()).dtype == torch.bfloat16
    model.setup_caches(max_batch_size=1, max_seq_len=model.config.max_seq_len, dtype=torch.bfloat16)
    prompt = torch.from_numpy(z["prompt"]).cuda()
    got = inf.generate(model=model, prompt=prompt, max_new_tokens=int(z["new_frames"]), decode_one_token=decode_one_token,
                       temperature=float(z["temperature"]), top_p=float(z["top_p"]), top_k=int(z["top_k"])).cpu()
    assert torch.equal(got.to(torch.int32), torch.from_numpy(z["ref_tokens"]))

    # decode_n_tokens (inference.py:184-238) continues from the prefill's token exactly like generate does
    T = prompt.size(1)
    first = inf.decode_one_token_ar(model, prompt.view(1, cfg.num_codebooks + 1, -1), torch.arange(T, device="cuda"),
                                    float(z["temperature"]), float(z["top_p"]), int(z["top_k"]))
    rest = inf.decode_n_tokens(model, first.view(1, cfg.num_codebooks + 1, -1), torch.tensor([T], device="cuda"),
                               int(z["new_frames"]) - 1, float(z["temperature"]), float(z["top_p"]), int(z["top_k"]))
    seq = torch.cat([first.cpu(), rest.cpu()], dim=1).to(torch.int32)
    assert torch.equal(seq, torch.from_numpy(z["ref_tokens"])[:, T:])


@pytest.mark.parametrize("sharded", [True, False])
def test_init_model_from_safetensors_with_hf_style_keys(tmp_path, sharded):
    """llama.py:549-567: `model.safetensors.index.json` + shards (or one `model.safetensors`) whose keys carry the
    released checkpoints' prefixes (`text_model.model.*`, `audio_decoder.*` -> `_remap_fish_qwen3_omni_keys`,
    llama.py:229-246): same tokens as the reference golden."""
    from safetensors.torch import save_file

    from fish_speech_b200.models.text2semantic import inference as inf

    cfg, w, z = load_golden(GOLD / "lm_tiny_greedy.npz")
    conf = dataclasses.asdict(model_args(cfg))
    conf["im_end_id"] = cfg.im_end_id
    (tmp_path / "config.json").write_text(json.dumps(conf))

    def hf_key(k):
        if k.startswith("fast_"):
            return "audio_decoder." + k[len("fast_"):]
        if k.startswith("codebook_embeddings."):
            return "audio_decoder." + k
        return "text_model.model." + k

    sd = {hf_key(k): v.contiguous() for k, v in w.items()}
    if sharded:
        keys = sorted(sd)
        half = len(keys) // 2
        parts = {"model-00001-of-00002.safetensors": keys[:half], "model-00002-of-00002.safetensors": keys[half:]}
        for fn, ks in parts.items():
            save_file({k: sd[k] for k in ks}, str(tmp_path / fn))
        (tmp_path / "model.safetensors.index.json").write_text(
            json.dumps({"metadata": {}, "weight_map": {k: fn for fn, ks in parts.items() for k in ks}}))
    else:
        save_file(sd, str(tmp_path / "model.safetensors"))
    model, decode_one_token = inf.init_model(str(tmp_path), "cuda", torch.bfloat16, compile=False)
    model.setup_caches(max_batch_size=1, max_seq_len=model.config.max_seq_len, dtype=torch.bfloat16)
    got = inf.generate(model=model, prompt=torch.from_numpy(z["prompt"]).cuda(), max_new_tokens=int(z["new_frames"]),
                       decode_one_token=decode_one_token, temperature=float(z["temperature"]), top_p=float(z["top_p"]),
                       top_k=int(z["top_k"])).cpu()
    assert torch.equal(got.to(torch.int32), torch.from_numpy(z["ref_tokens"]))


def test_codec_load_model_from_checkpoint_file(tmp_path):
    from fish_speech_b200.models.dac import inference as dinf
    from tests.test_codec_gpu import build, rand_codes

    cfg = CO.tiny_config()
    w = CO.make_weights(cfg, seed=5)
    t = cfg.quant_tfm
    y = dict(sample_rate=cfg.sample_rate, encoder_dim=cfg.encoder_dim, encoder_rates=list(cfg.encoder_rates),
             decoder_dim=cfg.decoder_dim, decoder_rates=list(cfg.decoder_rates),
             encoder_transformer_layers=list(cfg.encoder_transformer_layers),
             transformer_general_config=dict(window_size=cfg.enc_tfm_window),
             quantizer=dict(n_codebooks=cfg.n_codebooks, codebook_size=cfg.codebook_size,
                            semantic_codebook_size=cfg.semantic_codebook_size, codebook_dim=cfg.codebook_dim,
                            downsample_factor=list(cfg.downsample_factor),
                            post_module=dict(window_size=t.window_size,
                                             config=dict(n_layer=t.n_layer, n_head=t.n_head, dim=t.dim,
                                                         intermediate_size=t.intermediate_size, head_dim=t.head_dim,
                                                         rope_base=t.rope_base, norm_eps=t.norm_eps))))
    (tmp_path / "tiny_dac.yaml").write_text(yaml.safe_dump(y))
    # a training checkpoint: generator.* are the codec's weights, everything else (discriminators) is dropped
    sd = {f"generator.{k}": v for k, v in w.items()}
    sd["discriminator.dummy"] = torch.zeros(3)
    torch.save({"state_dict": sd}, tmp_path / "codec.pth")

    loaded = dinf.load_model(str(tmp_path / "tiny_dac.yaml"), str(tmp_path / "codec.pth"), device="cuda")
    direct = build(cfg, w)
    assert loaded.sample_rate == cfg.sample_rate and loaded.frame_length == cfg.frame_length
    codes = rand_codes(cfg, 2, 9, 3).cuda()
    assert torch.equal(loaded.from_indices(codes.clone()), direct.from_indices(codes.clone()))
    g = torch.Generator().manual_seed(1)
    audio = (0.1 * torch.randn(1, 1, 3 * cfg.frame_length + 17, generator=g)).cuda()
    c1, l1 = loaded.encode(audio)
    c2, l2 = direct.encode(audio)
    assert torch.equal(c1, c2) and torch.equal(l1, l2) and c1.dtype == torch.int64
