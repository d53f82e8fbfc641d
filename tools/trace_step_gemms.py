"""Diagnostic: per-CTA timeline of the decode step GEMMs launched back to back, without the consumer kernels between them
(include/fishb200.h fsb_lm_trace_step_gemms); tools/trace_frame.py is the in-frame version.

    python tools/trace_step_gemms.py [--layers 4]
"""
import argparse
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from bench import make_prompts, s2pro_cfg  # noqa: E402
from fish_speech_b200 import _lib, synthetic  # noqa: E402
from fish_speech_b200.configs import S2PRO_IM_END_ID  # noqa: E402
from fish_speech_b200.models.text2semantic.llama import DualARTransformer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=6)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = s2pro_cfg()
    cfg.n_layer = a.layers
    w = synthetic.lm_state_dict(cfg, dev)
    model = DualARTransformer(cfg, w, device=dev, im_end_id=S2PRO_IM_END_ID)
    model.max_rows = 32 * 64
    model.setup_caches(max_batch_size=32, max_seq_len=cfg.max_seq_len)
    eng = model.engine
    sp = eng.sampling(0.7, 0.7, 1, 42)
    eng.reset()
    eng.prefill([p.to(dev) for p in make_prompts(cfg, 32, 42)], list(range(32)), sp, do_sample=True)
    eng.decode(32, 4, sp, use_graph=True)
    n = 4 * a.layers
    L = _lib.lib()
    grid = C.c_int()
    st = torch.cuda.current_stream().cuda_stream
    for rep in range(3):  # the last repetition is reported (warm instruction cache, steady clocks)
        trace = torch.zeros(n, 512, 8, dtype=torch.int64, device=dev)
        got = L.fsb_lm_trace_step_gemms(eng.h, trace.data_ptr(), n, C.byref(grid), st)
        torch.cuda.synchronize()
        assert got == n, _lib.lib().fsb_last_error()
    t = trace.cpu()
    names = ["qkv", "wo", "w13", "w2"]
    g = grid.value
    print(f"grid {g} CTAs; per launch: wall = last end - first start; medians over CTAs, microseconds")
    print("launch  kind  wall   start_spread  wait_returned  first_acc  end    normalisers_ready  first_tile_normalised")
    prev_end = None
    for i in range(n):
        r = t[i, :g].double()
        live = r[:, 0] > 0
        r = r[live]
        t0 = r[:, 0].min()
        med = lambda c: float(((r[:, c] - r[:, 0])[r[:, c] > 0]).median()) / 1e3 if (r[:, c] > 0).any() else float("nan")
        wall = float(r[:, 5].max() - t0) / 1e3
        spread = float(r[:, 0].max() - t0) / 1e3
        gap = "" if prev_end is None else f" gap {float(t0 - prev_end) / 1e3:6.2f}"
        print(f"{i:4d}  {names[i % 4]:4s} {wall:6.2f}  {spread:8.2f}  {med(1):12.2f}  {med(4):8.2f}  {med(5):6.2f}  {med(6):12.2f}  "
              f"{med(7):12.2f}{gap}")
        prev_end = r[:, 5].max()


if __name__ == "__main__":
    main()
