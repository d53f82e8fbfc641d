"""Probe: is the step GEMM's weight stream bound by HBM or by the SM-side pipeline?  The same GEMM is launched 40 times
back to back (weights L2-resident after the first launch where they fit 126 MB) and compared with the same kind of GEMM
walking through 36 different layers (weights from HBM).  python tools/probe/l2_resident_gemm.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import torch  # noqa: E402

from bench import s2pro_cfg  # noqa: E402
from fish_speech_b200 import _lib, synthetic  # noqa: E402
from fish_speech_b200.configs import S2PRO_IM_END_ID  # noqa: E402
from fish_speech_b200.models.text2semantic.llama import DualARTransformer  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = s2pro_cfg()
w = synthetic.lm_state_dict(cfg, dev)
model = DualARTransformer(cfg, w, device=dev, im_end_id=S2PRO_IM_END_ID)
model.max_rows = 32 * 64
model.setup_caches(max_batch_size=32, max_seq_len=cfg.max_seq_len)
eng = model.engine
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
mb = {"qkv": 31.5, "wo": 21.0, "w1|w3": 99.6, "w2": 49.8}
for kind, name in enumerate(["qkv", "wo", "w1|w3", "w2"]):
    res = {}
    for mode in ("same layer x40", "36 layers"):
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if mode.startswith("same"):
                _lib.check(L.fsb_lm_repeat_step_gemm(eng.h, 3, kind, 40, st))
                n = 40
            else:
                for l in range(cfg.n_layer):
                    _lib.check(L.fsb_lm_repeat_step_gemm(eng.h, l, kind, 1, st))
                n = cfg.n_layer
            e1.record()
            torch.cuda.synchronize()
            res[mode] = e0.elapsed_time(e1) * 1e3 / n
    print(f"{name:6s} {mb[name]:6.1f} MiB: " + ", ".join(f"{k}: {v:.2f} us/launch ({mb[name] * 1.048576 / v:.2f} TB/s)" for k, v in res.items()))
