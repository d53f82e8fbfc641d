"""Diagnostic: per-layer timing of one full-geometry codec decode. Every conv / linear GEMM call is bracketed by
CUDA events (the extra events serialise nothing: one stream), and reported with its shape, useful FLOPs,
activation bytes (bf16 in + out, fp32 where the output is fp32) and the implied TFLOP/s and GB/s."""
import sys
from collections import OrderedDict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from fish_speech_b200 import synthetic
from fish_speech_b200.models.dac.inference import load_codec_config
from fish_speech_b200.models.dac.modded_dac import DAC

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
cfg = load_codec_config()
dac = DAC(cfg, synthetic.codec_state_dict(cfg, dev), device=dev)
g = torch.Generator().manual_seed(0)
codes = torch.stack([torch.randint(0, 4096, (B, T), generator=g)] + [torch.randint(0, 1024, (B, T), generator=g) for _ in range(9)], 1).cuda()
for _ in range(2):
    dac.from_indices(codes.clone())
torch.cuda.synchronize()

records = []
orig = DAC._gemm


def timed(self, cv, x, Bn, T_in, c_in, T_out, out0=None, out1=None, snake=None, resid=None, gamma=None, act=0,
          out_f32=False):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(self, cv, x, Bn, T_in, c_in, T_out, out0=out0, out1=out1, snake=snake, resid=resid, gamma=gamma, act=act,
         out_f32=out_f32)
    e1.record()
    outs = (out0 is not None) + (out1 is not None)
    ob = 4 if out_f32 else 2
    records.append(dict(e0=e0, e1=e1, c_out=cv.c_out, taps=cv.taps, kpad=cv.kpad, c_in_eff=cv.c_in_eff, rows=Bn * T_out,
                        flops=2.0 * Bn * T_out * cv.c_out * cv.taps * cv.c_in_eff,
                        padded=2.0 * Bn * T_out * (-(-cv.c_out // 128) * 128) * cv.taps * cv.kpad,
                        bytes=Bn * T_in * c_in * 2 + outs * Bn * T_out * cv.c_out * ob
                        + (Bn * T_out * cv.c_out * 2 if resid is not None else 0),
                        outs=outs, resid=resid is not None))


DAC._gemm = timed
e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e_all0.record()
dac.from_indices(codes.clone())
e_all1.record()
torch.cuda.synchronize()
total = e_all0.elapsed_time(e_all1)
agg = OrderedDict()
for r in records:
    key = (r["c_out"], r["taps"], r["kpad"], r["rows"], r["outs"], r["resid"])
    a = agg.setdefault(key, dict(n=0, ms=0.0, flops=0.0, padded=0.0, bytes=0.0))
    a["n"] += 1
    a["ms"] += r["e0"].elapsed_time(r["e1"])
    a["flops"] += r["flops"]
    a["padded"] += r["padded"]
    a["bytes"] += r["bytes"]
gemm_ms = sum(a["ms"] for a in agg.values())
print(f"# codec decode per-layer timing, B={B} T={T}: whole call {total:.1f} ms, conv/linear GEMMs {gemm_ms:.1f} ms\n")
print("| c_out | taps | K/tap (padded) | rows | outs | +res | calls | ms | useful TFLOP/s | issued TFLOP/s | GB/s (activations) |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for (c_out, taps, kpad, rows, outs, resid), a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"| {c_out} | {taps} | {kpad} | {rows} | {outs} | {int(resid)} | {a['n']} | {a['ms']:.2f} | "
          f"{a['flops'] / a['ms'] / 1e9:.0f} | {a['padded'] / a['ms'] / 1e9:.0f} | {a['bytes'] / a['ms'] / 1e6:.0f} |")
