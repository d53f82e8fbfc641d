"""Pin the oracle's stop semantics against the REAL reference (test infrastructure; build container only):

    python -m oracle.make_golden_stop

Two cases of `generate` (inference.py:243-359 -> decode_n_tokens :184-238) on weights whose <|im_end|> head row is a
multiple of the row of the token the (collapsed, random-weight) free run keeps picking:
(a) "first": 1.5x, greedy -> <|im_end|> wins at frame 0, the prefill's token, which the reference's loop does NOT
    test, and again at frame 1, where the loop breaks;
(b) "mid": 1.0x (a coin flip against that token), sampled with top_k=30 under a torch RNG seed for which the first
    <|im_end|> is drawn a few frames in -> the loop breaks there (same RNG stream in reference and oracle).
Stores the reference's outputs in tests/golden/ref_stop_cases.npz (not matched by the GPU suite's lm_*.npz glob:
the CUDA path is checked against the oracle for these cases in tests/test_lm_gpu.py)."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import lm_oracle as O  # noqa: E402
from oracle import ref_stubs as R  # noqa: E402
from oracle.make_golden import make_prompt  # noqa: E402

GOLD = ROOT / "tests" / "golden"
SEED, HEAD_GAIN, T, N = 61, 8.0, 10, 20


def stop_weights(cfg, gain: float):
    w = O.make_weights(cfg, seed=SEED, head_gain=HEAD_GAIN)
    prompt = make_prompt(cfg, SEED, T)
    free = O.generate(O.setup(cfg, w), prompt, 6, temperature=0.7, top_p=0.7, top_k=1, noise=False)
    src = int(free[0, T])
    w["embeddings.weight"][cfg.im_end_id] = (w["embeddings.weight"][src].float() * gain).bfloat16()
    return w, prompt, src


def main():
    R.install()
    from fish_speech.models.text2semantic import inference as ref_inf

    cfg = O.tiny_config()
    out = {}
    for tag, gain, samp in (("first", 1.5, dict(temperature=0.7, top_p=0.7, top_k=1)),
                            ("mid", 1.0, dict(temperature=1.0, top_p=0.9, top_k=30))):
        w, prompt, src = stop_weights(cfg, gain)
        rng = 1
        model = R.reference_lm(cfg, w)  # built BEFORE seeding: module construction draws from the torch RNG
        while True:
            torch.manual_seed(rng)
            ref = ref_inf.generate(model=model, prompt=prompt, max_new_tokens=N, audio_masks=None,
                                   audio_parts=None, **samp).to(torch.int32)
            torch.manual_seed(rng)
            got = O.generate(O.setup(cfg, w), prompt, N, **samp)
            assert torch.equal(ref, got), f"{tag}: oracle differs from the reference"
            hits = (ref[0, T:] == cfg.im_end_id).nonzero().flatten().tolist()
            if tag == "first":
                # greedy: an RNG seed for which the reference's bf16 noise never hits U == 0 (lm_oracle.NOISE)
                clean = O.generate(O.setup(cfg, w), prompt, N, noise=False, **samp)
                if torch.equal(clean.to(torch.int32), ref):
                    break
            elif hits and 3 <= hits[0] <= 12 and ref.shape[1] == T + hits[0] + 1:
                break
            rng += 1
        n_new = ref.shape[1] - T
        print(f"{tag}: <|im_end|> row = {gain} x token {src}; rng seed {rng}; reference generated {n_new} frames, "
              f"im_end at generated indices {hits}")
        out[f"{tag}_src_token"] = src
        out[f"{tag}_gain"] = gain
        out[f"{tag}_ref_tokens"] = ref.numpy()
        out[f"{tag}_rng_seed"] = rng
        for k, v in samp.items():
            out[f"{tag}_{k}"] = v
    np.savez_compressed(GOLD / "ref_stop_cases.npz", weight_seed=SEED, head_gain=HEAD_GAIN,
                        prompt=make_prompt(cfg, SEED, T).numpy(), max_new_tokens=N, **out)
    print("wrote", GOLD / "ref_stop_cases.npz")


if __name__ == "__main__":
    main()
