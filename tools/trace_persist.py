"""Diagnostic: phase timeline of the persistent slow-stack kernel (CTA 0)."""
import os
import sys
from pathlib import Path

os.environ["FSB_PERSISTENT"] = "1"
os.environ["FSB_PK_TRACE"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from fish_speech_b200 import synthetic
from fish_speech_b200.configs import S2PRO_IM_END_ID, s2pro_args
from fish_speech_b200.models.text2semantic.llama import DualARTransformer

dev = torch.device("cuda", 0)
cfg = s2pro_args(max_seq_len=512, n_layer=6)
w = synthetic.lm_state_dict(cfg, dev)
w["embeddings.weight"][S2PRO_IM_END_ID] = 0
model = DualARTransformer(cfg, w, device=dev, im_end_id=S2PRO_IM_END_ID)
model.setup_caches(32, 512)
eng = model.engine
prompts = [torch.randint(0, 150000, (11, 64), dtype=torch.int32) for _ in range(32)]
for p in prompts:
    p[1:] = 0
sp = eng.sampling(0.7, 0.7, 1, 1)
eng.reset()
eng.prefill([p.to(dev) for p in prompts], list(range(32)), sp)
eng.decode(32, 20, sp, use_graph=True)
torch.cuda.synchronize()
t = eng.buffer("pk_trace").cpu().double()
t = (t - t[0]) / 1e3
names = ["qkv", "wo", "w13", "w2"]
i = 1
for l in range(3):
    for g in range(4):
        e, b1, c, b2 = t[i], t[i + 1], t[i + 2], t[i + 3]
        prev = t[i - 1]
        print(f"L{l} {names[g]:4s} gemm+epi {e - prev:6.2f}us | barrier {b1 - e:5.2f} | consumer {c - b1:6.2f} | barrier {b2 - c:5.2f}   (t={b2:8.2f})")
        i += 4
