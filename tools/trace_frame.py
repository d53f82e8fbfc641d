"""Diagnostic: in-frame timeline of ONE decode frame at step-GEMM granularity (include/fishb200.h fsb_lm_trace_frame).

Every step GEMM of the frame records per-CTA globaltimer stamps. Per launch this prints
    gap    last CTA of the previous GEMM ended -> this GEMM's dependency wait returned (median CTA): the kernels between
           the two GEMMs (finalize / attention / sampler / embed) plus the completion latency of the boundary
    dep    wait returned -> last CTA ended: operand fetch, normalise, remaining weight stream, accumulator read-out, stores
    early  first CTA started -> wait returned: how long the GEMM was resident (prefetching weights) before it could run
and sums them by kind.  python tools/trace_frame.py [--batch 32] [--markdown]
"""
import argparse
import ctypes as C
import sys
from collections import OrderedDict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from bench import make_prompts, s2pro_cfg  # noqa: E402
from fish_speech_b200 import _lib, synthetic  # noqa: E402
from fish_speech_b200.configs import S2PRO_IM_END_ID  # noqa: E402
from fish_speech_b200.models.text2semantic.llama import DualARTransformer  # noqa: E402


def kinds(cfg):
    """Launch order of decode_one_frame's step GEMMs (csrc/lm_engine.cu)."""
    out = []
    for _ in range(cfg.n_layer):
        out += ["slow qkv", "slow wo", "slow w13", "slow w2"]
    out.append("slow head")
    C_ = cfg.num_codebooks
    for p in range(C_):
        for l in range(cfg.n_fast_layer):
            out.append("fast qkv")
            if p == 0 and l == cfg.n_fast_layer - 1:
                break  # pass 0 only fills the KV cache
            out += ["fast wo", "fast w13", "fast w2"]
        if p > 0:
            out.append("fast head")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--every", type=int, default=0, help="also print every k-th launch")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = s2pro_cfg()
    w = synthetic.lm_state_dict(cfg, dev)
    w["embeddings.weight"][S2PRO_IM_END_ID] = 0
    model = DualARTransformer(cfg, w, device=dev, im_end_id=S2PRO_IM_END_ID)
    B = a.batch
    model.max_rows = B * 64
    model.setup_caches(max_batch_size=B, max_seq_len=cfg.max_seq_len)
    eng = model.engine
    sp = eng.sampling(0.7, 0.7, 1, 42)
    eng.reset()
    eng.prefill([p.to(dev) for p in make_prompts(cfg, B, 42)], list(range(B)), sp, do_sample=True)
    eng.decode(B, 4, sp, use_graph=True)
    names = kinds(cfg)
    n = len(names)
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):  # the last repetition is reported (warm instruction cache, steady clocks)
        trace = torch.zeros(n + 8, 512, 8, dtype=torch.int64, device=dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        atr = torch.zeros(512, 8, dtype=torch.int64, device=dev)
        got = L.fsb_lm_trace_frame(eng.h, B, C.byref(sp), trace.data_ptr(), n + 8, atr.data_ptr(), 512, st)
        ev1.record()
        torch.cuda.synchronize()
        assert got == n, (got, n, L.fsb_last_error())
    t = trace.cpu().double()
    frame_us = ev0.elapsed_time(ev1) * 1e3
    agg = OrderedDict()
    prev_end = None
    first_start = last_end = None
    rows = []
    for i in range(n):
        r = t[i]
        r = r[r[:, 0] > 0]
        s0 = float(r[:, 0].min())
        wait = float(r[:, 1][r[:, 1] > 0].median())
        end = float(r[:, 5].max())
        gap = (wait - prev_end) / 1e3 if prev_end is not None else 0.0
        dep = (end - wait) / 1e3
        early = (wait - s0) / 1e3
        k = agg.setdefault(names[i], [0, 0.0, 0.0, 0.0])
        k[0] += 1; k[1] += gap; k[2] += dep; k[3] += early
        rows.append((i, names[i], gap, dep, early, len(r), wait, end))
        prev_end = end
        first_start = s0 if first_start is None else first_start
        last_end = end
    print(f"batch {B}: eager traced frame {frame_us:.0f} us by CUDA events; first GEMM start -> last GEMM end "
          f"{(last_end - first_start) / 1e3:.0f} us; {n} step GEMMs")
    print("| GEMM | launches | gap us (mean) | dep us (mean) | resident before wait us (mean) | gap total | dep total |")
    print("|---|---|---|---|---|---|---|")
    tg = td = 0.0
    for name, (c, g, d, e) in agg.items():
        print(f"| {name} | {c} | {g / c:.2f} | {d / c:.2f} | {e / c:.2f} | {g:.0f} | {d:.0f} |")
        tg += g; td += d
    print(f"| all | {n} | | | | {tg:.0f} | {td:.0f} |")
    at = atr.cpu().double()
    at = at[at[:, 0] > 0]
    ns, nf = cfg.n_layer, at.shape[0] - cfg.n_layer
    labels = ["start->wait", "wait->partials summed", "->q/k/v finished", "->scores", "->softmax", "->values", "->end"]
    for name, blk in (("slow attention", at[:ns]), ("fast attention", at[ns:])):
        blk = blk[blk[:, 7] > 0]  # KV-only launches stop after the cache write
        if len(blk) == 0:
            continue
        d = (blk[:, 1:] - blk[:, :-1]) / 1e3
        print(f"{name}, CTA (0,0), {len(blk)} launches, mean us: " +
              ", ".join(f"{lab} {float(d[:, i].mean()):.2f}" for i, lab in enumerate(labels)) +
              f"; wait->end {float((blk[:, 7] - blk[:, 1]).mean()) / 1e3:.2f}")
    # boundaries around the slow attention: qkv GEMM's last CTA end -> attention's wait returned; attention end -> wo's wait
    b1 = [float(at[l, 1]) - rows[4 * l][7] for l in range(ns)]
    b2 = [rows[4 * l + 1][6] - float(at[l, 7]) for l in range(ns)]
    print(f"slow attention boundaries, mean us: qkv GEMM end -> attention may run {sum(b1) / ns / 1e3:.2f}; "
          f"attention CTA (0,0) end -> wo GEMM may run {sum(b2) / ns / 1e3:.2f}")
    if a.every:
        for i, nm, g, d, e, ctas, _w, _e in rows[:: a.every]:
            print(f"{i:4d} {nm:10s} gap {g:6.2f} dep {d:6.2f} early {e:6.2f} ctas {ctas}")


if __name__ == "__main__":
    main()
