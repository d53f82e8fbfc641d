"""Diagnostic (not part of the product path): per-CTA timeline of the decode GEMM chain."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from fish_speech_b200 import _lib, synthetic
from fish_speech_b200.configs import S2PRO_IM_END_ID, s2pro_args
from fish_speech_b200.models.text2semantic.llama import DualARTransformer

dev = torch.device("cuda", 0)
cfg = s2pro_args(max_seq_len=512, n_layer=6)
model = DualARTransformer(cfg, synthetic.lm_state_dict(cfg, dev), device=dev, im_end_id=S2PRO_IM_END_ID)
model.setup_caches(32, 512)
L = _lib.lib()
n = 24
trace = torch.zeros(n * 148 * 3, dtype=torch.int64, device=dev)
grid = C.c_int()
st = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    trace.zero_()
    got = L.fsb_lm_trace_gemms(model.engine.h, trace.data_ptr(), n, C.byref(grid), st)
    torch.cuda.synchronize()
t = trace.cpu().view(n, 148, 3).double()
t0 = t[..., 0][t[..., 0] > 0].min()
names = ["qkv", "wo", "w13", "w2"]
prev_end = None
for i in range(got):
    g = t[i, : grid.value]
    start, wait, end = g[:, 0] - t0, g[:, 1] - t0, g[:, 2] - t0
    line = (f"{i:2d} {names[i % 4]:4s} start[min {start.min()/1e3:7.2f} max {start.max()/1e3:7.2f}] "
            f"wait_ret[min {wait.min()/1e3:7.2f}] end[min {end.min()/1e3:7.2f} max {end.max()/1e3:7.2f}] "
            f"dur {(end.max()-start.min())/1e3:6.2f}us")
    if prev_end is not None:
        line += f"  start-prev_end {(start.min()-prev_end)/1e3:6.2f}us"
    prev_end = end.max()
    print(line)
