import sys
sys.path.insert(0, ".")
import torch
from oracle import lm_oracle as O
from tests.lm_util import build_model, make_prompt
from tests.test_scheduler_gpu import _alone, _grown_prompt

cfg = O.tiny_config()
w = O.make_weights(cfg, seed=91, head_gain=8.0)
kw = dict(temperature=0.7, top_p=0.7, top_k=1)
p1 = make_prompt(cfg, 91, 40)
m1, m2, m3 = build_model(cfg, w, debug=False), build_model(cfg, w, debug=False), build_model(cfg, w, debug=False)
prompt = p1
for c in range(3):
    a = _alone(m1, prompt, 7, reuse_prefix=True, **kw)
    b = _alone(m2, prompt, 7, **kw)
    b2 = _alone(m3, prompt, 7, **kw)
    b3 = _alone(m3, prompt, 7, **kw)
    T = prompt.size(1)
    print("chunk", c, "T", T, "reuse==fresh", torch.equal(a, b), "fresh==fresh2", torch.equal(b, b2), "fresh2 rerun", torch.equal(b2, b3))
    if not torch.equal(a, b):
        print(" diff at", (a != b).nonzero().tolist()[:10])
        print(" reuse", a[:, T:].tolist())
        print(" fresh", b[:, T:].tolist())
        ref = O.generate(O.setup(cfg, w), prompt, 7, noise=False, **kw)
        print(" oracle", ref[:, T:].tolist())
    prompt = _grown_prompt(cfg, prompt, b[:, T:], 300 + c, 9)
