"""Run a short S2-Pro-geometry generation (prefill + a few decode frames, eager launches) for ncu launch lists.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/frame_launches.csv \
        python tools/profile_frame.py --frames 3
"""
import argparse
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("FSB_NO_GRAPH", "1")

import torch  # noqa: E402

from bench import make_prompts, s2pro_cfg  # noqa: E402
from fish_speech_b200 import synthetic  # noqa: E402
from fish_speech_b200.configs import S2PRO_IM_END_ID  # noqa: E402
from fish_speech_b200.models.text2semantic.llama import DualARTransformer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--layers", type=int, default=0, help="override n_layer (0 = full 36)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = s2pro_cfg()
    if a.layers:
        cfg.n_layer = a.layers
    w = synthetic.lm_state_dict(cfg, dev)
    w["embeddings.weight"][S2PRO_IM_END_ID] = 0
    model = DualARTransformer(cfg, w, device=dev, im_end_id=S2PRO_IM_END_ID)
    model.max_rows = a.batch * 64
    model.setup_caches(max_batch_size=a.batch, max_seq_len=cfg.max_seq_len)
    eng = model.engine
    prompts = [p.to(dev) for p in make_prompts(cfg, a.batch, 42)]
    sp = eng.sampling(0.7, 0.7, 1, 42)
    eng.reset()
    eng.prefill(prompts, list(range(a.batch)), sp, do_sample=True)
    eng.decode(a.batch, a.frames, sp, use_graph=False)
    torch.cuda.synchronize()
    print("tokens", eng.buffer("out_tokens")[0, :, : a.frames + 1].tolist())


if __name__ == "__main__":
    main()
