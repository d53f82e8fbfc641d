"""Diagnostic: one full-geometry codec decode (and optionally encode) for ncu / timing."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from fish_speech_b200 import synthetic
from fish_speech_b200.models.dac.inference import load_codec_config
from fish_speech_b200.models.dac.modded_dac import DAC

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
cfg = load_codec_config()
dac = DAC(cfg, synthetic.codec_state_dict(cfg, dev), device=dev)
g = torch.Generator().manual_seed(0)
codes = torch.stack([torch.randint(0, 4096, (B, T), generator=g)] + [torch.randint(0, 1024, (B, T), generator=g) for _ in range(9)], 1).cuda()
wav = dac.from_indices(codes.clone())
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(reps):
    wav = dac.from_indices(codes.clone())
ev1.record()
torch.cuda.synchronize()
print(f"decode B={B} T={T}: {ev0.elapsed_time(ev1)/reps:.2f} ms per call, wav {tuple(wav.shape)}")
if "--encode" in sys.argv:
    audio = 0.1 * torch.randn(B, 1, 441000, device=dev)
    c, l = dac.encode(audio)
    torch.cuda.synchronize()
    ev0.record()
    c, l = dac.encode(audio)
    ev1.record()
    torch.cuda.synchronize()
    print(f"encode B={B} 10 s: {ev0.elapsed_time(ev1):.2f} ms, codes {tuple(c.shape)}")
