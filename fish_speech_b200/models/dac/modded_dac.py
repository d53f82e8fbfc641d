"""Drop-in surface of fish_speech/models/dac/modded_dac.py + rvq.py for inference.

`DAC` keeps the reference's public surface (`encode`, `from_indices`, `decode`, `sample_rate`, `hop_length`,
`frame_length`, `device`, `parameters()`; modded_dac.py:804-946) but holds packed device weights and
walks the reference's module tree issuing one CUDA op per layer through the C-ABI (include/fishb200.h):

  * every Conv1d / ConvTranspose1d / Linear -> fsb_conv_gemm (tcgen05 implicit-im2col GEMM, fused bias /
    Snake / GELU / LayerScale / residual epilogues); weight-norm (parametrizations.weight.original0/1) is
    folded ONCE at load instead of on every forward (modded_dac.py:554-556)
  * channels-last bf16 activations [B][T][C]; fp32 accumulation; fp32 waveform out
  * strided convs (k = m*stride) read the input as [T/stride][stride*C] rows -> m taps of K = stride*C
  * ConvTranspose1d (k = m*stride) is ONE GEMM with stride*C_out output channels and m taps
  * the window-limited transformers reuse the LM glue kernels with a banded (length-aware) attention —
    no T x T mask is ever materialised (modded_dac.py:380-398)
  * encode skips the post_module + upsample pass whose result the reference discards (rvq.py:318-319 vs
    modded_dac.py:919-923)
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
from dataclasses import dataclass, field
from typing import Iterator, Optional

import torch

from ... import _lib


@dataclass
class TfmConfig:
    n_layer: int = 8
    n_head: int = 16
    dim: int = 1024
    intermediate_size: int = 3072
    head_dim: int = 64
    rope_base: float = 10000
    norm_eps: float = 1e-5
    window_size: Optional[int] = 128


@dataclass
class CodecConfig:
    """Values of fish_speech/configs/modded_dac_vq.yaml (see config.load_codec_config)."""

    sample_rate: int = 44100
    encoder_dim: int = 64
    encoder_rates: tuple = (2, 4, 8, 8)
    decoder_dim: int = 1536
    decoder_rates: tuple = (8, 8, 4, 2)
    encoder_transformer_layers: tuple = (0, 0, 0, 4)
    n_codebooks: int = 9
    codebook_size: int = 1024
    semantic_codebook_size: int = 4096
    codebook_dim: int = 8
    downsample_factor: tuple = (2, 2)
    quant_tfm: TfmConfig = field(default_factory=TfmConfig)
    enc_tfm_window: int = 512

    @property
    def latent_dim(self) -> int:
        return self.encoder_dim * (2 ** len(self.encoder_rates))

    @property
    def hop_length(self) -> int:
        return int(math.prod(self.encoder_rates))

    def enc_tfm(self, dim: int, n_layer: int) -> TfmConfig:
        return TfmConfig(n_layer=n_layer, n_head=dim // 64, dim=dim, intermediate_size=dim * 3, head_dim=64,
                         rope_base=10000, norm_eps=1e-5, window_size=self.enc_tfm_window)


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _Conv:
    """A conv / linear lowered to the multi-tap GEMM: packed bf16 weight + fp32 epilogue vectors."""

    __slots__ = ("w", "bias", "c_out", "taps", "kpad", "shifts", "in_group", "out_group", "c_in_eff")

    def __init__(self, w, bias, c_out, taps, kpad, shifts, in_group=1, out_group=1, c_in_eff=0):
        self.w, self.bias, self.c_out, self.taps, self.kpad = w, bias, c_out, taps, kpad
        self.shifts = (C.c_int * len(shifts))(*shifts)
        self.in_group, self.out_group, self.c_in_eff = in_group, out_group, c_in_eff


class _Snake:
    __slots__ = ("alpha", "inv")

    def __init__(self, alpha: torch.Tensor, dev, repeat: int = 1):
        a = alpha.detach().float().reshape(-1).repeat(repeat)
        self.alpha = a.to(dev).contiguous()
        self.inv = (a + 1e-9).reciprocal().to(dev).contiguous()


class DAC:
    def __init__(self, cfg: CodecConfig, state_dict: dict, device="cuda"):
        if not torch.cuda.is_available():
            raise _lib.FsbError("fish_speech_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.lib()
        self.cfg = cfg
        self._device = torch.device(device)
        self.sample_rate = cfg.sample_rate
        self.encoder_rates = list(cfg.encoder_rates)
        self.decoder_rates = list(cfg.decoder_rates)
        self.hop_length = cfg.hop_length
        self.frame_length = self.hop_length * 4  # modded_dac.py:861
        self.latent_dim = cfg.latent_dim
        self._keep: list[torch.Tensor] = []
        self._bufs: dict = {}
        self._graphs: dict = {}   # (B, S, T) -> (CUDAGraph, static index buffer, static waveform) of from_indices
        self._graph_seen: dict = {}
        self._use_graphs = os.environ.get("FSB_CODEC_GRAPH", "1") != "0"
        # decoder ResidualUnits as one kernel each (csrc/codec_resunit.cu); 0 = two conv GEMM launches per unit
        self._fused_units = os.environ.get("FSB_FUSED_RESUNIT", "1") != "0"
        self._lock = threading.RLock()
        self._rope: dict = {}
        self._sd = {k: v.detach().float() for k, v in state_dict.items()}  # folded on the device they live on
        with torch.cuda.device(self._device):
            self._pack()
        del self._sd

    # ---- nn.Module-like surface ----
    @property
    def device(self) -> torch.device:
        return self._device

    def parameters(self) -> Iterator[torch.Tensor]:
        return iter(self._keep)

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    # -------------------------------------------------------------------------------------------
    # weight packing
    # -------------------------------------------------------------------------------------------
    def _dev(self, t: torch.Tensor, dtype) -> torch.Tensor:
        t = t.to(device=self._device, dtype=dtype).contiguous()
        self._keep.append(t)
        return t

    def _eff_weight(self, prefix: str) -> torch.Tensor:
        sd = self._sd
        g = sd.get(f"{prefix}.parametrizations.weight.original0")
        if g is None:
            return sd[f"{prefix}.weight"]
        return torch._weight_norm(sd[f"{prefix}.parametrizations.weight.original1"], g, 0)

    def _conv(self, prefix: str, dilation: int = 1, stride: int = 1) -> _Conv:
        """Conv1d weight [C_out, C_in, k] -> [C_out][taps][kpad]. stride>1 needs k % stride == 0 (the codec
        only uses k = stride and k = 2*stride)."""
        w = self._eff_weight(prefix)
        c_out, c_in, k = w.shape
        if stride == 1:
            kpad = _pad64(c_in)
            m = torch.zeros(c_out, k, kpad, device=w.device)
            m[:, :, :c_in] = w.permute(0, 2, 1)
            shifts = [-(k - 1 - j) * dilation for j in range(k)]
            taps, c_in_eff = k, c_in
        else:
            assert k % stride == 0 and dilation == 1
            taps = k // stride
            c_in_eff = stride * c_in
            kpad = _pad64(c_in_eff)
            m = torch.zeros(c_out, taps, kpad, device=w.device)
            # tap q covers kernel positions j = q*stride + r, r < stride, laid out (r, ci)
            wr = w.permute(0, 2, 1).reshape(c_out, taps, stride * c_in)
            m[:, :, :c_in_eff] = wr
            shifts = [-(taps - 1 - q) for q in range(taps)]
        bias = self._sd.get(f"{prefix}.bias")
        return _Conv(self._dev(m.reshape(c_out, taps * kpad), torch.bfloat16),
                     self._dev(bias, torch.float32) if bias is not None else None,
                     c_out, taps, kpad, shifts, in_group=stride, c_in_eff=c_in_eff)

    def _convT(self, prefix: str, stride: int) -> _Conv:
        """ConvTranspose1d weight [C_in, C_out, k], k = m*stride, right-trimmed by k - stride
        (modded_dac.py:574-580): out[t*s + p] = sum_q W[:, :, p + q*s]^T x[t - q]  -> one GEMM with
        s*C_out output channels (p major) and m taps (shift -q)."""
        w = self._eff_weight(prefix)
        c_in, c_out, k = w.shape
        assert k % stride == 0
        taps = k // stride
        kpad = _pad64(c_in)
        m = torch.zeros(stride * c_out, taps, kpad, device=w.device)
        for q in range(taps):
            # rows (p, co)  <-  W[ci, co, p + q*stride]
            blk = w[:, :, q * stride:(q + 1) * stride]  # [ci, co, p]
            m[:, q, :c_in] = blk.permute(2, 1, 0).reshape(stride * c_out, c_in)
        bias = self._sd.get(f"{prefix}.bias")
        return _Conv(self._dev(m.reshape(stride * c_out, taps * kpad), torch.bfloat16),
                     self._dev(bias.repeat(stride), torch.float32) if bias is not None else None,
                     stride * c_out, taps, kpad, [-q for q in range(taps)], out_group=stride, c_in_eff=c_in)

    def _linear(self, prefix: str, with_bias=True) -> _Conv:
        w = self._sd[f"{prefix}.weight"]
        n, k = w.shape
        kpad = _pad64(k)
        m = torch.zeros(n, kpad, device=w.device)
        m[:, :k] = w
        bias = self._sd.get(f"{prefix}.bias") if with_bias else None
        return _Conv(self._dev(m, torch.bfloat16), self._dev(bias, torch.float32) if bias is not None else None,
                     n, 1, kpad, [0], c_in_eff=k)

    def _snake(self, prefix: str, repeat: int = 1) -> _Snake:
        s = _Snake(self._sd[f"{prefix}.alpha"], self._device, repeat)
        self._keep += [s.alpha, s.inv]
        return s

    def _tfm(self, prefix: str, t: TfmConfig) -> dict:
        bf = torch.bfloat16
        layers = []
        for l in range(t.n_layer):
            p = f"{prefix}.layers.{l}"
            w13 = torch.cat([self._sd[f"{p}.feed_forward.w1.weight"], self._sd[f"{p}.feed_forward.w3.weight"]], 0)
            layers.append(dict(
                wqkv=self._dev(self._sd[f"{p}.attention.wqkv.weight"], bf),
                wo=self._dev(self._sd[f"{p}.attention.wo.weight"], bf),
                w13=self._dev(w13, bf), w2=self._dev(self._sd[f"{p}.feed_forward.w2.weight"], bf),
                attn_norm=self._dev(self._sd[f"{p}.attention_norm.weight"], bf),
                ffn_norm=self._dev(self._sd[f"{p}.ffn_norm.weight"], bf),
                ls_attn=self._dev(self._sd[f"{p}.attention_layer_scale.gamma"], bf),
                ls_ffn=self._dev(self._sd[f"{p}.ffn_layer_scale.gamma"], bf)))
        return dict(cfg=t, layers=layers, norm=self._dev(self._sd[f"{prefix}.norm.weight"], bf))

    def _convnext(self, prefix: str) -> dict:
        f32 = torch.float32
        dw = self._sd[f"{prefix}.dwconv.conv.weight"]  # [C, 1, 7]
        return dict(dw=self._dev(dw[:, 0, :], f32), dw_b=self._dev(self._sd[f"{prefix}.dwconv.conv.bias"], f32),
                    ln_w=self._dev(self._sd[f"{prefix}.norm.weight"], f32), ln_b=self._dev(self._sd[f"{prefix}.norm.bias"], f32),
                    pw1=self._linear(f"{prefix}.pwconv1"), pw2=self._linear(f"{prefix}.pwconv2"),
                    gamma=self._dev(self._sd[f"{prefix}.gamma"], f32), k=dw.shape[-1])

    def _res_unit(self, prefix: str, dilation: int) -> dict:
        return dict(s0=self._snake(f"{prefix}.block.0"), c7=self._conv(f"{prefix}.block.1.conv", dilation=dilation),
                    s1=self._snake(f"{prefix}.block.2"), c1=self._conv(f"{prefix}.block.3.conv"), dil=dilation)

    def _vq_tables(self, prefix: str):
        """out_proj(codebook) tables (fp32 [size, D]) for decode, and in_proj / normalised codebook for encode."""
        w_out = self._eff_weight(f"{prefix}.out_proj")[:, :, 0]  # [D, cd]
        cb = self._sd[f"{prefix}.codebook.weight"]  # [size, cd]
        tab = cb @ w_out.t() + self._sd[f"{prefix}.out_proj.bias"]
        w_in = self._eff_weight(f"{prefix}.in_proj")[:, :, 0]  # [cd, D]
        return tab, w_in, self._sd[f"{prefix}.in_proj.bias"], torch.nn.functional.normalize(cb)

    def _pack(self):
        cfg = self.cfg
        f32 = torch.float32
        # ---- quantizer ----
        prefixes = ["quantizer.semantic_quantizer.quantizers.0"] + [
            f"quantizer.quantizer.quantizers.{i}" for i in range(cfg.n_codebooks)]
        tabs, w_in, b_in, cbn, sizes = [], [], [], [], []
        for p in prefixes:
            tab, wi, bi, cn = self._vq_tables(p)
            tabs.append(self._dev(tab, f32))
            w_in.append(wi)
            b_in.append(bi)
            cbn.append(cn)
            sizes.append(tab.shape[0])
        self.n_stage = len(prefixes)
        self.vq_tabs = tabs
        self.vq_tab_ptrs = self._dev(torch.tensor([t.data_ptr() for t in tabs], dtype=torch.int64), torch.int64)
        self.vq_sizes = self._dev(torch.tensor(sizes, dtype=torch.int32), torch.int32)
        self.vq_in_w = self._dev(torch.stack(w_in), f32)  # [S, cd, D]
        self.vq_in_b = self._dev(torch.stack(b_in), f32)
        self.vq_cbn = self._dev(torch.cat(cbn, 0), f32)
        offs = [0]
        for s in sizes[:-1]:
            offs.append(offs[-1] + s)
        self.vq_cb_off = self._dev(torch.tensor(offs, dtype=torch.int32), torch.int32)
        self.down = [dict(conv=self._conv(f"quantizer.downsample.{i}.0.conv", stride=f),
                          cnx=self._convnext(f"quantizer.downsample.{i}.1"), f=f)
                     for i, f in enumerate(cfg.downsample_factor)]
        self.up = [dict(conv=self._convT(f"quantizer.upsample.{i}.0.conv", f),
                        cnx=self._convnext(f"quantizer.upsample.{i}.1"), f=f)
                   for i, f in enumerate(reversed(cfg.downsample_factor))]
        self.pre_tfm = self._tfm("quantizer.pre_module", cfg.quant_tfm)
        self.post_tfm = self._tfm("quantizer.post_module", cfg.quant_tfm)
        # ---- decoder ----
        self.dec_in = self._conv("decoder.model.0.conv")
        self.dec_blocks = []
        for i, s in enumerate(cfg.decoder_rates):
            p = f"decoder.model.{i + 1}.block"
            res = [self._res_unit(f"{p}.{2 + j}", d) for j, d in enumerate((1, 3, 9))]
            up_snake = _SnakeView(res[0]["s0"], s)  # Snake params in the (p, co) order of the convT GEMM
            self._keep += [up_snake.alpha, up_snake.inv]
            self.dec_blocks.append(dict(s_in=self._snake(f"{p}.0"), up=self._convT(f"{p}.1.conv", s), stride=s,
                                        res=res, up_snake=up_snake))
        n = len(cfg.decoder_rates) + 1
        self.dec_out_snake = self._snake(f"decoder.model.{n}")
        wf = self._eff_weight(f"decoder.model.{n + 1}.conv")  # [1, C, 7]
        self.dec_out_w = self._dev(wf[0].t(), f32)  # [K][C]
        self.dec_out_b = float(self._sd[f"decoder.model.{n + 1}.conv.bias"][0])
        self.dec_out_k = wf.shape[-1]
        # ---- encoder ----
        w0 = self._eff_weight("encoder.block.0.conv")  # [C0, 1, 7]
        self.enc_in_w = self._dev(w0[:, 0, :], f32)
        self.enc_in_b = self._dev(self._sd["encoder.block.0.conv.bias"], f32)
        self.enc_blocks = []
        d = cfg.encoder_dim
        for i, (s, ntl) in enumerate(zip(cfg.encoder_rates, cfg.encoder_transformer_layers)):
            d *= 2
            p = f"encoder.block.{i + 1}.block"
            self.enc_blocks.append(dict(
                res=[self._res_unit(f"{p}.{j}", dil) for j, dil in enumerate((1, 3, 9))],
                s_out=self._snake(f"{p}.3"), down=self._conv(f"{p}.4.conv", stride=s), stride=s, dim=d,
                tfm=self._tfm(f"{p}.5", cfg.enc_tfm(d, ntl)) if ntl > 0 else None))
        nb = len(cfg.encoder_rates) + 1
        self.enc_out_snake = self._snake(f"encoder.block.{nb}")
        self.enc_out = self._conv(f"encoder.block.{nb + 1}.conv")

    # -------------------------------------------------------------------------------------------
    # op helpers
    # -------------------------------------------------------------------------------------------
    def _buf(self, name: str, numel: int, dtype=torch.bfloat16) -> torch.Tensor:
        """Named persistent workspace (stable device pointers keep the GEMM plan cache small)."""
        b = self._bufs.get(name)
        if b is None or b.numel() < numel or b.dtype != dtype:
            if b is not None:
                self._graphs.clear()  # captured graphs point into the buffer that is being replaced
                self._graph_seen.clear()
            b = torch.empty(max(numel, 1), dtype=dtype, device=self._device)
            self._bufs[name] = b
        return b

    def _gemm(self, cv: _Conv, x: torch.Tensor, B: int, T_in: int, c_in: int, T_out: int, out0=None, out1=None,
              snake: Optional[_Snake] = None, resid=None, gamma=None, act=0, out_f32=False):
        """x: flat bf16 buffer viewed as [B][T_in][c_in] rows."""
        vp = lambda t: None if t is None else t.data_ptr()
        _lib.check(self.lib.fsb_conv_gemm(
            x.data_ptr(), B, T_in, c_in, c_in, T_in * c_in, cv.w.data_ptr(), cv.c_out, cv.taps, cv.kpad, cv.shifts,
            T_out, vp(cv.bias), vp(gamma), vp(resid), act, vp(out0), vp(out1),
            snake.alpha.data_ptr() if snake is not None else None,
            snake.inv.data_ptr() if snake is not None else None, int(out_f32), _stream()))

    def _rope_table(self, T: int, head_dim: int, base: float) -> torch.Tensor:
        key = (head_dim, base)
        t = self._rope.get(key)
        if t is None or t.shape[0] < T:
            n = max(T, 1024)
            inv = 1.0 / (base ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
            ang = torch.outer(torch.arange(n), inv)
            cis = torch.polar(torch.ones_like(ang), ang)
            new = torch.stack([cis.real, cis.imag], dim=-1).to(torch.bfloat16).to(self._device)  # modded_dac.py:442-452
            if t is not None:
                self._graphs.clear()  # captured graphs point into the table that is being replaced
                self._graph_seen.clear()
            t = new
            self._rope[key] = t
        return t

    def _transformer(self, tf: dict, x: torch.Tensor, B: int, T: int, tag: str, pos0: int = 0,
                     kv_cache: Optional[list] = None, cache_len: int = 0) -> torch.Tensor:
        """WindowLimitedTransformer.forward on x = flat [B*T, D] bf16 (updated in place as the residual
        stream); returns the final-normed tensor (a workspace view).
        Incremental use (DecodeStream): the T rows are positions pos0 .. pos0+T-1 of a longer sequence whose earlier
        K/V live in `kv_cache` = per layer (K, V) buffers [B][H][cache_len][Dh]; the window attention reads them."""
        t: TfmConfig = tf["cfg"]
        D, H, Dh, I = t.dim, t.n_head, t.head_dim, t.intermediate_size
        rows = B * T
        L = self.lib
        st = _stream()
        S = cache_len if kv_cache is not None else T
        ws = self._buf("tf_ws", rows * max(3 * H * Dh, 2 * I, D), torch.float32)
        n = self._buf("tf_n", rows * D)
        q = self._buf("tf_q", rows * H * Dh)
        kc = self._buf("tf_k", rows * H * Dh)
        vc = self._buf("tf_v", rows * H * Dh)
        att = self._buf("tf_att", rows * H * Dh)
        hb = self._buf("tf_h", rows * I)
        # (sequence, position) of every row: two persistent int32 buffers refilled per call (a per-(B, T) cache
        # would grow without bound over a dataset-scale bulk encode, where almost every batch has a new shape)
        seq = self._buf("tf_seq", rows, torch.int32)[:rows]
        pos = self._buf("tf_pos", rows, torch.int32)[:rows]
        ar = torch.arange(rows, dtype=torch.int32, device=self._device)
        torch.div(ar, T, rounding_mode="floor", out=seq)
        torch.remainder(ar, T, out=pos)
        if pos0:
            pos += pos0
        freqs = self._rope_table(pos0 + T, Dh, t.rope_base)
        window = t.window_size if t.window_size is not None else 0
        layers = tf["layers"]
        chk = _lib.check
        chk(L.fsb_resid_scale_norm(None, 0, None, x.data_ptr(), None, layers[0]["attn_norm"].data_ptr(), n.data_ptr(),
                                   rows, D, t.norm_eps, 0, st))
        for l, lw in enumerate(layers):
            if kv_cache is not None:
                kc, vc = kv_cache[l]
            chk(L.fsb_linear_f32(n.data_ptr(), rows, D, lw["wqkv"].data_ptr(), 3 * H * Dh, ws.data_ptr(), st))
            chk(L.fsb_qkv_rope(ws.data_ptr(), rows, H, H, Dh, freqs.data_ptr(), seq.data_ptr(), pos.data_ptr(),
                               q.data_ptr(), kc.data_ptr(), vc.data_ptr(), S, st))
            chk(L.fsb_window_attn(q.data_ptr(), kc.data_ptr(), vc.data_ptr(), seq.data_ptr(), pos.data_ptr(), rows,
                                  H, H, Dh, S, window, att.data_ptr(), st))
            chk(L.fsb_linear_f32(att.data_ptr(), rows, H * Dh, lw["wo"].data_ptr(), D, ws.data_ptr(), st))
            chk(L.fsb_resid_scale_norm(ws.data_ptr(), D, lw["ls_attn"].data_ptr(), x.data_ptr(), x.data_ptr(),
                                       lw["ffn_norm"].data_ptr(), n.data_ptr(), rows, D, t.norm_eps, 0, st))
            chk(L.fsb_linear_f32(n.data_ptr(), rows, D, lw["w13"].data_ptr(), 2 * I, ws.data_ptr(), st))
            chk(L.fsb_swiglu_f32(ws.data_ptr(), rows, I, hb.data_ptr(), st))
            chk(L.fsb_linear_f32(hb.data_ptr(), rows, I, lw["w2"].data_ptr(), D, ws.data_ptr(), st))
            nxt = layers[l + 1]["attn_norm"] if l + 1 < len(layers) else tf["norm"]
            chk(L.fsb_resid_scale_norm(ws.data_ptr(), D, lw["ls_ffn"].data_ptr(), x.data_ptr(), x.data_ptr(),
                                       nxt.data_ptr(), n.data_ptr(), rows, D, t.norm_eps, 0, st))
        return n

    def _convnext_block(self, cn: dict, x: torch.Tensor, B: int, T: int, Cc: int, tag: str):
        """ConvNeXtBlock (rvq.py:173-191), in place on x (flat [B][T][C])."""
        y = self._buf("cnx_y", B * T * Cc)
        h = self._buf("cnx_h", B * T * 4 * Cc)
        _lib.check(self.lib.fsb_dwconv_ln(x.data_ptr(), cn["dw"].data_ptr(), cn["dw_b"].data_ptr(), cn["ln_w"].data_ptr(),
                                          cn["ln_b"].data_ptr(), B, T, Cc, cn["k"], 1e-6, y.data_ptr(), _stream()))
        self._gemm(cn["pw1"], y, B, T, Cc, T, out0=h, act=1)
        self._gemm(cn["pw2"], h, B, T, 4 * Cc, T, out0=x, gamma=cn["gamma"], resid=x)

    # -------------------------------------------------------------------------------------------
    # decode
    # -------------------------------------------------------------------------------------------
    @torch.inference_mode()
    def from_indices(self, indices: torch.Tensor) -> torch.Tensor:
        """modded_dac.py:925-927: codes [B, 1+n_codebooks, T] -> waveform [B, 1, T*frame_length] (fp32).
        Like the reference (rvq.py:354-359) the caller's tensor is clamped in place.
        The ~190 launches of one decode are captured into a CUDA graph the second time a (B, T) shape is seen and
        replayed from then on: the layer walk below costs more host time than the GPU needs for short utterances."""
        cfg = self.cfg
        # one caller at a time per DAC object: the layers share persistent workspaces (the reference's nn.Module
        # allocates per call and may be driven from two host threads, e.g. TTSInferenceEngine + a batch encoder)
        with self._lock, torch.cuda.device(self._device):
            indices[:, 0] = torch.clamp(indices[:, 0], max=cfg.semantic_codebook_size - 1)
            indices[:, 1:] = torch.clamp(indices[:, 1:], max=cfg.codebook_size - 1)
            B, S, T = indices.shape
            key = (B, S, T)
            hit = self._graphs.get(key)
            if hit is not None:
                graph, g_idx, g_wav = hit
                g_idx.copy_(indices.reshape(-1))
                graph.replay()
                return g_wav.clone().view(B, 1, -1)
            seen = self._graph_seen.get(key, 0)
            self._graph_seen[key] = seen + 1
            if not self._use_graphs or seen == 0 or len(self._graphs) >= 4:
                # first sight of a shape (warms the plan cache and sizes the workspaces), or capture disabled / full
                idx = indices.to(device=self._device, dtype=torch.int32).contiguous()
                wav = self._from_indices_body(idx, B, S, T)
                self._idx_keepalive = idx
                return wav.view(B, 1, -1)
            g_idx = torch.empty(B * S * T, dtype=torch.int32, device=self._device)
            g_idx.copy_(indices.reshape(-1))
            torch.cuda.current_stream().synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):  # the LM worker thread keeps running
                g_wav = self._from_indices_body(g_idx.view(B, S, T), B, S, T)
            self._graphs[key] = (graph, g_idx, g_wav)
            graph.replay()
            return g_wav.clone().view(B, 1, -1)

    def _from_indices_body(self, idx: torch.Tensor, B: int, S: int, T: int) -> torch.Tensor:
        """idx int32 [B, S, T] (clamped) -> waveform [B, T*frame_length]: rvq.decode + Decoder.forward."""
        D = self.cfg.latent_dim
        z = self._buf("q_z", B * T * D)
        _lib.check(self.lib.fsb_codebook_sum(idx.data_ptr(), self.vq_tab_ptrs.data_ptr(), self.vq_sizes.data_ptr(),
                                             S, B, T, D, z.data_ptr(), _stream()))
        z = self._transformer(self.post_tfm, z, B, T, "post")
        Tc = T
        cur = z
        for i, u in enumerate(self.up):
            out = self._buf(f"up_{i}", B * Tc * u["f"] * D)
            self._gemm(u["conv"], cur, B, Tc, D, Tc, out0=out)
            Tc *= u["f"]
            self._convnext_block(u["cnx"], out, B, Tc, D, f"up{i}")
            cur = out
        return self._decoder(cur, B, Tc)

    def open_decode_stream(self, batch: int = 1, max_frames: int = 4096, conv_context: int = 16) -> "DecodeStream":
        """Incremental `from_indices`: push code frames as the LM emits them, get their samples back.
        The codec is causal (rvq.py:395-398 asserts it), so what a frame sounds like never depends on later frames."""
        return DecodeStream(self, batch, max_frames, conv_context)

    @torch.inference_mode()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """modded_dac.py:929-946: latent [B, D, T'] (channels first) -> waveform [B, 1, T'*hop]."""
        with self._lock, torch.cuda.device(self._device):
            B, D, Tc = z.shape
            buf = self._buf("dec_zin", B * Tc * D)
            buf[: B * Tc * D].view(B, Tc, D).copy_(z.to(self._device).transpose(1, 2))
            return self._decoder(buf, B, Tc).view(B, 1, -1)

    def _unit_is_fusable(self, ru: dict, c: int) -> bool:
        c7, c1 = ru["c7"], ru["c1"]
        return (self._fused_units and c7.taps == 7 and c1.taps == 1 and c7.bias is not None and c1.bias is not None
                and c7.kpad == _pad64(c) and c1.kpad == _pad64(c) and bool(self.lib.fsb_res_unit_supported(c)))

    def _decoder(self, z: torch.Tensor, B: int, T: int) -> torch.Tensor:
        """Decoder.forward (modded_dac.py:760-801) on z = flat [B][T][latent] bf16."""
        cfg = self.cfg
        c = cfg.decoder_dim
        # largest activation: after the last transposed conv
        t_full = T * cfg.hop_length
        numel = max(B * (T * math.prod(cfg.decoder_rates[: i + 1])) * (cfg.decoder_dim // 2 ** (i + 1))
                    for i in range(len(cfg.decoder_rates)))
        numel = max(numel, B * T * c)
        bufs = [self._buf(f"dec_{k}", numel) for k in range(3)]
        free = [0, 1, 2]

        def take():
            return free.pop(0)

        a = take()
        first = self.dec_blocks[0]
        self._gemm(self.dec_in, z, B, T, cfg.latent_dim, T, out1=bufs[a], snake=first["s_in"])
        Tc, cin = T, c
        for bi, blk in enumerate(self.dec_blocks):
            s = blk["stride"]
            cout = cin // 2
            x, a2 = take(), take()
            # transposed conv: raw -> x (residual stream), Snake(res0.block.0) -> a2
            self._gemm(blk["up"], bufs[a], B, Tc, cin, Tc, out0=bufs[x], out1=bufs[a2], snake=blk["up_snake"])
            free.append(a)
            Tc *= s
            a = a2
            for j, ru in enumerate(blk["res"]):
                m = take()
                last_unit = j == 2
                if not last_unit:
                    nxt = blk["res"][j + 1]["s0"]
                elif bi + 1 < len(self.dec_blocks):
                    nxt = self.dec_blocks[bi + 1]["s_in"]
                else:
                    nxt = self.dec_out_snake
                keep_raw = not last_unit
                if self._unit_is_fusable(ru, cout):
                    # Snake -> conv7 -> Snake -> conv1 -> + x in one kernel; the activated result goes to the spare buffer
                    # (neighbouring tiles still read their halo from bufs[a])
                    c7, c1 = ru["c7"], ru["c1"]
                    _lib.check(self.lib.fsb_res_unit(
                        bufs[a].data_ptr(), bufs[x].data_ptr(), B, Tc, cout, ru["dil"], c7.w.data_ptr(), c7.bias.data_ptr(),
                        ru["s1"].alpha.data_ptr(), ru["s1"].inv.data_ptr(), c1.w.data_ptr(), c1.bias.data_ptr(),
                        bufs[x].data_ptr() if keep_raw else None, bufs[m].data_ptr(), nxt.alpha.data_ptr(),
                        nxt.inv.data_ptr(), _stream()))
                    free.append(a)
                    a = m
                    continue
                self._gemm(ru["c7"], bufs[a], B, Tc, cout, Tc, out1=bufs[m], snake=ru["s1"])
                self._gemm(ru["c1"], bufs[m], B, Tc, cout, Tc, out0=bufs[x] if keep_raw else None, out1=bufs[a],
                           snake=nxt, resid=bufs[x])
                free.append(m)
            free.append(x)
            cin = cout
        wav = torch.empty(B, t_full, dtype=torch.float32, device=self._device)
        _lib.check(self.lib.fsb_final_conv_tanh(bufs[a].data_ptr(), self.dec_out_w.data_ptr(), self.dec_out_b, B, Tc, cin,
                                                self.dec_out_k, wav.data_ptr(), _stream()))
        return wav

    # -------------------------------------------------------------------------------------------
    # encode
    # -------------------------------------------------------------------------------------------
    @torch.inference_mode()
    def encode(self, audio_data: torch.Tensor, audio_lengths: Optional[torch.Tensor] = None,
               n_quantizers: Optional[int] = None, **kwargs):
        """modded_dac.py:874-923: audio [B,1,N] or [B,N] -> (codes int64 [B, 1+n_codebooks, T], lens [B])."""
        cfg = self.cfg
        with self._lock, torch.cuda.device(self._device):
            if audio_data.ndim == 2:
                audio_data = audio_data.unsqueeze(1)
            length = audio_data.shape[-1]
            right_pad = math.ceil(length / self.frame_length) * self.frame_length - length
            wav = torch.nn.functional.pad(audio_data.to(device=self._device, dtype=torch.float32), (0, right_pad))
            if audio_lengths is None:
                audio_lengths = torch.LongTensor([length + right_pad]).to(self._device)
            B, _, N = wav.shape
            wav = wav.reshape(B, N).contiguous()
            d0 = cfg.encoder_dim
            numel = B * N * d0
            bufs = [self._buf(f"enc_{k}", numel) for k in range(3)]
            x, a, m = 0, 1, 2
            r0 = self.enc_blocks[0]["res"][0]
            _lib.check(self.lib.fsb_first_conv(wav.data_ptr(), self.enc_in_w.data_ptr(), self.enc_in_b.data_ptr(),
                                               r0["s0"].alpha.data_ptr(), r0["s0"].inv.data_ptr(), B, N, d0,
                                               self.enc_in_w.shape[1], bufs[x].data_ptr(), bufs[a].data_ptr(), _stream()))
            Tc, c = N, d0
            for bi, blk in enumerate(self.enc_blocks):
                for j, ru in enumerate(blk["res"]):
                    self._gemm(ru["c7"], bufs[a], B, Tc, c, Tc, out1=bufs[m], snake=ru["s1"])
                    nxt = blk["res"][j + 1]["s0"] if j < 2 else blk["s_out"]
                    self._gemm(ru["c1"], bufs[m], B, Tc, c, Tc, out0=bufs[x] if j < 2 else None, out1=bufs[a],
                               snake=nxt, resid=bufs[x])
                s, cout = blk["stride"], blk["dim"]
                Tn = Tc // s
                # strided conv reads the activation as [Tc/s][s*c] rows
                last = bi + 1 == len(self.enc_blocks)
                if blk["tfm"] is not None:
                    self._gemm(blk["down"], bufs[a], B, Tn, s * c, Tn, out0=bufs[x])
                    n_out = self._transformer(blk["tfm"], bufs[x], B, Tn, f"enc{bi}")
                    sn = self.enc_out_snake if last else self.enc_blocks[bi + 1]["res"][0]["s0"]
                    _lib.check(self.lib.fsb_snake(n_out.data_ptr(), sn.alpha.data_ptr(), sn.inv.data_ptr(),
                                                  B * Tn * cout, cout, bufs[a].data_ptr(), _stream()))
                    bufs[x][: B * Tn * cout].copy_(n_out[: B * Tn * cout])
                else:
                    sn = self.enc_out_snake if last else self.enc_blocks[bi + 1]["res"][0]["s0"]
                    self._gemm(blk["down"], bufs[a], B, Tn, s * c, Tn, out0=bufs[x], out1=bufs[m], snake=sn)
                    a, m = m, a
                Tc, c = Tn, cout
            z = self._buf("enc_z", B * Tc * cfg.latent_dim)
            self._gemm(self.enc_out, bufs[a], B, Tc, c, Tc, out0=z)
            # ---- quantizer.forward up to the codes (rvq.py:293-317) ----
            D = cfg.latent_dim
            cur = z
            for i, dn in enumerate(self.down):
                f = dn["f"]
                Tn = Tc // f
                out = self._buf(f"down_{i}", B * Tn * D)
                self._gemm(dn["conv"], cur, B, Tn, f * D, Tn, out0=out)
                self._convnext_block(dn["cnx"], out, B, Tn, D, f"down{i}")
                cur, Tc = out, Tn
            zq = self._transformer(self.pre_tfm, cur, B, Tc, "pre")
            S = self.n_stage if n_quantizers is None else min(self.n_stage, 1 + int(n_quantizers))
            codes = torch.empty(B, S, Tc, dtype=torch.int32, device=self._device)
            _lib.check(self.lib.fsb_vq_encode(zq.data_ptr(), self.vq_in_w.data_ptr(), self.vq_in_b.data_ptr(),
                                              self.vq_cbn.data_ptr(), self.vq_cb_off.data_ptr(), self.vq_sizes.data_ptr(),
                                              self.vq_tab_ptrs.data_ptr(), S, cfg.codebook_dim, B, Tc, D,
                                              codes.data_ptr(), _stream()))
            indices_lens = torch.ceil(audio_lengths.to(self._device) / self.frame_length).long()
            return codes.long(), indices_lens


class DecodeStream:
    """Streaming state of one `from_indices` over a growing code sequence (one per utterance batch).

    * post-transformer (8 layers, causal window 128): K/V of every layer stay in a cache; a push computes only the new
      frames at their absolute positions (the kernels are the ones the whole-sequence path uses: row positions + cache
      length are arguments) -- exact.
    * upsample + decoder convolutions: causal with a finite receptive field (< 10 frames: dilated k7 convs at 4..512
      steps per frame, ConvNeXt k7 at 2 and 4 steps per frame); the last `conv_context` frames of the transformer
      OUTPUT are kept and re-run in front of the new frames, and only the new frames' samples are returned -- exact
      as long as conv_context covers the receptive field (tests compare with the one-shot decode).
    """

    def __init__(self, dac: "DAC", batch: int, max_frames: int, conv_context: int):
        self.dac, self.B, self.S, self.ctx = dac, int(batch), int(max_frames), int(conv_context)
        t: TfmConfig = dac.post_tfm["cfg"]
        dev = dac.device
        n = self.B * t.n_head * self.S * t.head_dim
        self.kv = [(torch.zeros(n, dtype=torch.bfloat16, device=dev), torch.zeros(n, dtype=torch.bfloat16, device=dev))
                   for _ in dac.post_tfm["layers"]]
        self.pos = 0
        self.tail: Optional[torch.Tensor] = None  # [B, <=ctx, D] transformer output of the latest frames

    @torch.inference_mode()
    def push(self, indices: torch.Tensor) -> torch.Tensor:
        """codes [B, 1+n_codebooks, k] of the next k frames -> waveform [B, 1, k * frame_length] (fp32)."""
        dac, cfg = self.dac, self.dac.cfg
        B, S_, k = indices.shape
        if B != self.B:
            raise ValueError(f"stream was opened for batch {self.B}, got {B}")
        if k == 0:
            return torch.zeros(B, 1, 0, dtype=torch.float32, device=dac.device)
        if self.pos + k > self.S:
            raise ValueError(f"decode stream capacity exceeded ({self.pos} + {k} > {self.S} frames)")
        with dac._lock, torch.cuda.device(dac.device):
            idx = indices.to(device=dac.device, dtype=torch.int32).clone()
            idx[:, 0].clamp_(max=cfg.semantic_codebook_size - 1)  # rvq.py:354-359
            idx[:, 1:].clamp_(max=cfg.codebook_size - 1)
            idx = idx.contiguous()
            D = cfg.latent_dim
            z = dac._buf("st_z", B * k * D)
            _lib.check(dac.lib.fsb_codebook_sum(idx.data_ptr(), dac.vq_tab_ptrs.data_ptr(), dac.vq_sizes.data_ptr(), S_, B,
                                                k, D, z.data_ptr(), _stream()))
            zn = dac._transformer(dac.post_tfm, z, B, k, "post", pos0=self.pos, kv_cache=self.kv, cache_len=self.S)
            new = zn[: B * k * D].view(B, k, D)
            seq = new if self.tail is None else torch.cat([self.tail, new], dim=1)
            c = seq.shape[1] - k  # context frames in front of the new ones
            Tc = seq.shape[1]
            cur = dac._buf("st_in", B * Tc * D)
            cur[: B * Tc * D].view(B, Tc, D).copy_(seq)
            for i, u in enumerate(dac.up):
                out = dac._buf(f"up_{i}", B * Tc * u["f"] * D)
                dac._gemm(u["conv"], cur, B, Tc, D, Tc, out0=out)
                Tc *= u["f"]
                dac._convnext_block(u["cnx"], out, B, Tc, D, f"up{i}")
                cur = out
            wav = dac._decoder(cur, B, Tc).view(B, 1, -1)
            self.tail = seq[:, -self.ctx:].clone() if self.ctx > 0 else None
            self.pos += k
            dac._idx_keepalive = idx
            return wav[:, :, c * dac.frame_length:]


class _SnakeView:
    """Snake parameters repeated for the (p, co) channel order of a transposed-conv GEMM."""

    def __init__(self, s: _Snake, repeat: int):
        self.alpha = s.alpha.repeat(repeat).contiguous()
        self.inv = s.inv.repeat(repeat).contiguous()
