import sys
sys.path.insert(0, ".")
import numpy as np, torch
from oracle import lm_oracle as O
from tests.lm_util import build_model, load_golden, restricted
from tests.test_lm_gpu import _reference_uniforms, _gen

cfg, w, z = load_golden("tests/golden/lm_tiny_ras.npz")
n = int(z["new_frames"]); T = z["prompt"].shape[1]
tr = []; torch.manual_seed(int(z["rng_seed"]))
ref = O.generate(O.setup(cfg, w), torch.from_numpy(z["prompt"]), n, temperature=float(z["temperature"]), top_p=float(z["top_p"]), top_k=int(z["top_k"]), traces=tr)
for nf in (12, 13):
    model = build_model(cfg, w)
    eng = model.engine
    eng.set_sampler_noise(_reference_uniforms(cfg, int(z["rng_seed"]), n, eng.head_rows))
    got = _gen(model, torch.from_numpy(z["prompt"]), nf, temperature=float(z["temperature"]), top_p=float(z["top_p"]), top_k=int(z["top_k"]))
    f = nf - 1
    print("frames", nf, "last frame got", got[:, T + f].tolist(), "ref", ref[:, T + f].tolist())
    slow = eng.buffer("slow_logits")[0].cpu()
    want = restricted(cfg, tr[f]["slow_logits"])
    print("  slow logits max abs diff", float((slow - want).abs().max()), "argmax", int(slow.argmax()), int(want.argmax()))
    fast = eng.buffer("fast_logits")[: cfg.num_codebooks - 1, 0].cpu()
    for p in range(cfg.num_codebooks - 1):
        wf = tr[f]["fast_logits"][p]
        d = (fast[p] - wf).abs()
        print(f"  fast cb{p+1}: max abs diff {float(d.max()):.4f} at {int(d.argmax())}; n differing {(d > 0).sum().item()}; logit[31] {float(fast[p][31]):.4f} vs {float(wf[31]):.4f}; [73] {float(fast[p][73]):.4f} vs {float(wf[73]):.4f}")
