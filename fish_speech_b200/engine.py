"""Host-side wrapper of the Dual-AR engine handle (fsb_lm_*).  PyTorch is used only for device memory
(weights, index tensors) and streams; all compute happens in libfishb200.so."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import torch

from . import _lib

MAX_DECODE_BATCH = 32


class _DevView:
    """Expose a raw device pointer through __cuda_array_interface__ so torch can view it (zero copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {
            "shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None,
        }


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def rope_table(seq_len: int, n_elem: int, base: float) -> torch.Tensor:
    """bf16 cos/sin table exactly as the reference builds it (llama.py:1004-1023): the table is an
    input of the kernels, so it is produced by the same torch ops the reference uses."""
    inv = 1.0 / (base ** (torch.arange(0, n_elem, 2)[: n_elem // 2].float() / n_elem))
    ang = torch.outer(torch.arange(seq_len), inv)
    cis = torch.polar(torch.ones_like(ang), ang)
    return torch.stack([cis.real, cis.imag], dim=-1).to(torch.bfloat16)


def interleave_w13(w1: torch.Tensor, w3: torch.Tensor) -> torch.Tensor:
    """Fused gate|up projection in the row order the step GEMM's SwiGLU epilogue expects (include/fishb200.h,
    d_w13): per 128-row tile t, rows 32w+l (l<16) = w1[64t+16w+l] and rows 32w+16+l = w3[64t+16w+l]; the
    hidden size is padded to a multiple of 64 with zero rows."""
    I, D = w1.shape
    Ip = (I + 63) // 64 * 64
    if Ip != I:
        pad = torch.zeros(Ip - I, D, dtype=w1.dtype, device=w1.device)
        w1, w3 = torch.cat([w1, pad]), torch.cat([w3, pad])
    g = w1.reshape(Ip // 16, 1, 16, D)
    u = w3.reshape(Ip // 16, 1, 16, D)
    return torch.cat([g, u], dim=1).reshape(2 * Ip, D).contiguous()


def bf16_round(v: float) -> float:
    """Sampling scalars are bf16 tensors in the reference (inference.py:303-304)."""
    return float(torch.tensor(v, dtype=torch.bfloat16).float())


class LmEngine:
    """One model replica on one GPU: weights (bf16, device), KV caches and workspaces (library-owned)."""

    def __init__(self, config, state_dict: dict, device, im_end_id: int, max_batch: int = 1,
                 kv_len: Optional[int] = None, max_rows: int = 2048, max_frames: Optional[int] = None,
                 debug: bool = False):
        if not torch.cuda.is_available():
            raise _lib.FsbError("fish_speech_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.lib()
        self.cfg = config
        self.im_end_id = int(im_end_id)
        self.device = torch.device(device)
        kv_len = int(kv_len or config.max_seq_len)
        kv_len = min(kv_len, config.max_seq_len)
        max_frames = int(max_frames or kv_len)
        self.max_batch, self.kv_len, self.max_rows, self.max_frames = max_batch, kv_len, max_rows, max_frames
        self._keep = []  # device tensors the library points into
        dev = self.device

        def put(t: torch.Tensor) -> int:
            t = t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
            self._keep.append(t)
            return t.data_ptr()

        sd = state_dict
        c = config

        def layer_struct(prefix: str) -> _lib.LmLayer:
            L = _lib.LmLayer()
            L.d_attn_norm = put(sd[f"{prefix}.attention_norm.weight"])
            L.d_wqkv = put(sd[f"{prefix}.attention.wqkv.weight"])
            L.d_bqkv = put(sd[f"{prefix}.attention.wqkv.bias"]) if f"{prefix}.attention.wqkv.bias" in sd else None
            L.d_q_norm = put(sd[f"{prefix}.attention.q_norm.weight"]) if f"{prefix}.attention.q_norm.weight" in sd else None
            L.d_k_norm = put(sd[f"{prefix}.attention.k_norm.weight"]) if f"{prefix}.attention.k_norm.weight" in sd else None
            L.d_wo = put(sd[f"{prefix}.attention.wo.weight"])
            L.d_bo = put(sd[f"{prefix}.attention.wo.bias"]) if f"{prefix}.attention.wo.bias" in sd else None
            L.d_ffn_norm = put(sd[f"{prefix}.ffn_norm.weight"])
            # fused gate|up projection, rows interleaved so that SwiGLU runs in the GEMM epilogue
            L.d_w13 = put(interleave_w13(sd[f"{prefix}.feed_forward.w1.weight"], sd[f"{prefix}.feed_forward.w3.weight"]))
            L.d_w2 = put(sd[f"{prefix}.feed_forward.w2.weight"])
            return L

        self._layers = (_lib.LmLayer * c.n_layer)(*[layer_struct(f"layers.{i}") for i in range(c.n_layer)])
        self._fast_layers = (_lib.LmLayer * c.n_fast_layer)(
            *[layer_struct(f"fast_layers.{i}") for i in range(c.n_fast_layer)])
        head_src = sd["embeddings.weight"] if c.tie_word_embeddings else sd["output.weight"]
        # Only the semantic ids and <|im_end|> survive the reference's -inf logit bias
        # (inference.py:308-320): the head is restricted to those rows.
        head = torch.cat([head_src[c.semantic_begin_id: c.semantic_end_id + 1],
                          head_src[self.im_end_id: self.im_end_id + 1]], 0)
        W = _lib.LmWeights()
        W.d_embeddings = put(sd["embeddings.weight"])
        W.d_codebook_embeddings = put(sd["codebook_embeddings.weight"])
        W.d_norm = put(sd["norm.weight"])
        W.d_head = put(head)
        W.head_rows = head.shape[0]
        W.d_freqs = put(rope_table(kv_len, c.head_dim, c.rope_base))
        W.layers = self._layers
        W.d_fast_embeddings = put(sd["fast_embeddings.weight"])
        W.d_fast_norm = put(sd["fast_norm.weight"])
        W.d_fast_output = put(sd["fast_output.weight"])
        W.d_fast_freqs = put(rope_table(c.num_codebooks, c.fast_head_dim, c.rope_base))
        if "fast_project_in.weight" in sd:
            W.d_fast_proj_w = put(sd["fast_project_in.weight"])
            W.d_fast_proj_b = put(sd["fast_project_in.bias"])
        W.fast_layers = self._fast_layers
        self.head_rows = head.shape[0]

        K = _lib.LmConfig()
        K.dim, K.n_layer, K.n_head, K.n_kv_head, K.head_dim, K.intermediate = (
            c.dim, c.n_layer, c.n_head, c.n_local_heads, c.head_dim, c.intermediate_size)
        (K.fast_dim, K.n_fast_layer, K.fast_n_head, K.fast_n_kv_head, K.fast_head_dim, K.fast_intermediate) = (
            c.fast_dim, c.n_fast_layer, c.fast_n_head, c.fast_n_local_heads, c.fast_head_dim, c.fast_intermediate_size)
        K.vocab_size, K.codebook_size, K.num_codebooks = c.vocab_size, c.codebook_size, c.num_codebooks
        K.semantic_begin_id, K.semantic_end_id, K.im_end_id = c.semantic_begin_id, c.semantic_end_id, self.im_end_id
        K.norm_eps = c.norm_eps
        K.qk_norm, K.fast_qk_norm = int(c.attention_qk_norm), int(c.fast_attention_qk_norm)
        K.scale_codebook_embeddings = int(c.scale_codebook_embeddings)
        K.norm_fastlayer_input = int(getattr(c, "norm_fastlayer_input", False))
        K.max_batch, K.kv_len, K.max_rows, K.max_frames, K.debug = max_batch, kv_len, max_rows, max_frames, int(debug)
        self._K, self._W = K, W
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fsb_lm_create(C.byref(K), C.byref(W), C.byref(h)))
        self.h = h
        self.debug = debug
        self._views = {}
        self._ctx_bound = 0  # upper bound of every slot's length, tracked on the host
        self.slot_control = False
        self._slot_tokens: dict = {}  # slot -> token rows [C+1, T] (CPU int32) whose K/V positions [0, T) are valid
        self.rows_reused = 0
        self.rows_prefilled = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.fsb_lm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -------------------------------------------------------------------------------------------
    def buffer(self, name: str) -> torch.Tensor:
        """Zero-copy torch view of a named device buffer of the handle."""
        if name in self._views:
            return self._views[name]
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        _lib.check(self.lib.fsb_lm_buffer(self.h, name.encode(), C.byref(ptr), C.byref(nbytes)))
        c = self.cfg
        C1 = c.num_codebooks + 1
        shapes = {
            "out_tokens": ((self.max_batch, C1, self.max_frames), "<i4"),
            "n_out": ((self.max_batch,), "<i4"),
            "pos": ((32,), "<i4"),
            "finished": ((self.max_batch,), "<i4"),
            "cur_tok": ((self.max_batch, C1), "<i4"),
            "ras_window": ((self.max_batch, 10), "<i4"),
            "slot_state": ((self.max_batch,), "<i4"),
            "slot_limit": ((self.max_batch,), "<i4"),
            "slot_temperature": ((self.max_batch,), "<f4"),
            "slot_top_p": ((self.max_batch,), "<f4"),
            "slot_top_k": ((self.max_batch,), "<i4"),
            "slot_seed": ((self.max_batch,), "<i8"),
            "slow_logits": ((self.max_batch, self.head_rows), "<f4"),
            "fast_logits": ((c.num_codebooks, self.max_batch, c.codebook_size), "<f4"),
        }
        shape, ts = shapes[name]
        with torch.cuda.device(self.device):
            t = torch.as_tensor(_DevView(ptr.value, shape, ts), device=self.device)
        if ts == "<u2":
            t = t.view(torch.bfloat16)
        self._views[name] = t
        return t

    def sampling(self, temperature: float, top_p: float, top_k: int, seed: int = 0) -> _lib.Sampling:
        s = _lib.Sampling()
        s.temperature = bf16_round(temperature)
        s.top_p = bf16_round(top_p)
        s.top_k = int(top_k)
        s.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        return s

    def _grow_bound(self, n: int):
        n = min(int(n), self.kv_len)
        if n > self._ctx_bound:
            self._ctx_bound = n
        _lib.check(self.lib.fsb_lm_set_context_bound(self.h, self._ctx_bound))

    def set_context_bound(self, n: int):
        """For callers that place tokens at explicit positions (decode_one_token_ar / decode_n_tokens)."""
        self._grow_bound(n)

    def set_context_bound_exact(self, n: int):
        """For a scheduler that knows every live slot's length: the bound may also shrink."""
        self._ctx_bound = 0
        self._grow_bound(max(1, n))

    def set_slot_control(self, enable: bool):
        """Per-slot sampling parameters / stop rule / RNG stream (include/fishb200.h fsb_lm_set_slot_control)."""
        _lib.check(self.lib.fsb_lm_set_slot_control(self.h, int(bool(enable))))
        self.slot_control = bool(enable)

    def set_sampler_noise(self, u: Optional[torch.Tensor]):
        """Test hook (include/fishb200.h fsb_lm_set_sampler_noise): u fp32 [frames][draws][ld] on the device."""
        if u is None:
            self._noise = None
            _lib.check(self.lib.fsb_lm_set_sampler_noise(self.h, None, 0, 0))
            return
        u = u.to(device=self.device, dtype=torch.float32).contiguous()
        self._noise = u
        _lib.check(self.lib.fsb_lm_set_sampler_noise(self.h, u.data_ptr(), u.shape[1], u.shape[2]))

    def reset(self):
        self._ctx_bound = 0
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fsb_lm_reset(self.h, _stream()))

    def prefill(self, prompts: Sequence[torch.Tensor], slots: Sequence[int], sp: Optional[_lib.Sampling],
                start_pos: Optional[Sequence[int]] = None, do_sample: bool = True) -> None:
        """prompts[k]: integer tensor [C+1, T_k] (row 0 token ids, rows 1..C codes) for slot slots[k]."""
        C1 = self.cfg.num_codebooks + 1
        start_pos = list(start_pos) if start_pos is not None else [0] * len(prompts)
        for p, s0 in zip(prompts, start_pos):
            if p.shape[0] != C1:
                raise ValueError(f"prompt must have {C1} rows, got {tuple(p.shape)}")
            if s0 + p.shape[1] > self.kv_len:
                raise ValueError(f"Input sequence length {s0 + p.shape[1]} exceeds the KV cache ({self.kv_len})")
        # group whole sequences into passes of <= max_rows rows; a longer sequence is split in time
        groups, cur, cur_rows = [], [], 0
        for k, p in enumerate(prompts):
            T = p.shape[1]
            t0 = 0
            while T - t0 > self.max_rows:  # time-split: no sampling on non-final pieces
                if cur:
                    groups.append(cur)
                    cur, cur_rows = [], 0
                groups.append([(k, t0, t0 + self.max_rows, False)])
                t0 += self.max_rows
            if cur_rows + (T - t0) > self.max_rows:
                groups.append(cur)
                cur, cur_rows = [], 0
            cur.append((k, t0, T, True))
            cur_rows += T - t0
        if cur:
            groups.append(cur)
        dev = self.device
        for k, s0 in enumerate(start_pos):
            if s0 == 0:
                self._slot_tokens.pop(int(slots[k]), None)  # overwritten from position 0: the old record is void
        self._grow_bound(max(s0 + p.shape[1] for p, s0 in zip(prompts, start_pos)) + 1)
        with torch.cuda.device(dev):
            for grp in groups:
                toks, rslot, rpos, last, gsl = [], [], [], [], []
                rows = 0
                final = all(f for (_, _, _, f) in grp)
                for (k, a, b, _) in grp:
                    p = prompts[k][:, a:b]
                    toks.append(p.t().to(device=dev, dtype=torch.int32))
                    n = b - a
                    rslot.append(torch.full((n,), int(slots[k]), dtype=torch.int32))
                    rpos.append(torch.arange(start_pos[k] + a, start_pos[k] + b, dtype=torch.int32))
                    rows += n
                    last.append(rows - 1)
                    gsl.append(int(slots[k]))
                d_tok = torch.cat(toks).contiguous()
                d_slot = torch.cat(rslot).to(dev)
                d_pos = torch.cat(rpos).to(dev)
                d_last = torch.tensor(last, dtype=torch.int32, device=dev)
                d_gsl = torch.tensor(gsl, dtype=torch.int32, device=dev)
                _lib.check(self.lib.fsb_lm_prefill(
                    self.h, d_tok.data_ptr(), d_slot.data_ptr(), d_pos.data_ptr(), rows, d_last.data_ptr(),
                    d_gsl.data_ptr(), len(grp), int(do_sample and final), C.byref(sp) if sp is not None else None,
                    _stream()))
                # index tensors must outlive the asynchronous kernels that read them
                self._inflight = (d_tok, d_slot, d_pos, d_last, d_gsl)

    # ---- prefix KV reuse (include/fishb200.h fsb_lm_copy_kv) -----------------------------------
    MIN_REUSE = 16  # shorter shared prefixes are not worth a separate prefill pass

    @staticmethod
    def _common_prefix(a: torch.Tensor, b: torch.Tensor) -> int:
        n = min(a.shape[1], b.shape[1])
        if n == 0:
            return 0
        diff = (a[:, :n] != b[:, :n]).any(dim=0).nonzero()
        return int(diff[0]) if len(diff) else n

    def prefill_reusing(self, prompts: Sequence[torch.Tensor], slots: Sequence[int], sp: Optional[_lib.Sampling],
                        do_sample: bool = True) -> list[int]:
        """`prefill` of full prompts that skips every row whose K/V the cache already holds: the longest prefix
        a prompt shares with what its own slot was last prefilled with, or with another slot that is not being
        re-prefilled in this call (copied over with fsb_lm_copy_kv). The reference re-prefills the whole growing
        conversation for every chunk of `generate_long` (inference.py:611-721); prefill is row-independent, so
        the reused K/V is bit-identical to what a full prefill would write. Returns the reused length per prompt."""
        host = [p.detach().to("cpu", torch.int32) for p in prompts]
        busy = set(int(s) for s in slots)
        reused, pieces, starts = [], [], []
        with torch.cuda.device(self.device):
            for p, s in zip(host, slots):
                s = int(s)
                best, donor = 0, s
                for d, toks in self._slot_tokens.items():
                    if d != s and d in busy:
                        continue
                    n = self._common_prefix(p, toks)
                    if n > best:
                        best, donor = n, d
                best = min(best, p.shape[1] - 1)  # at least the last row is computed: its logits are sampled
                if best < self.MIN_REUSE:
                    best = 0
                elif donor != s:
                    _lib.check(self.lib.fsb_lm_copy_kv(self.h, donor, s, best, _stream()))
                reused.append(best)
                starts.append(best)
            for p, b in zip(prompts, reused):
                pieces.append(p[:, b:])
        self.prefill(pieces, slots, sp, start_pos=starts, do_sample=do_sample)
        for p, s in zip(host, slots):
            self._slot_tokens[int(s)] = p
        self.rows_reused += sum(reused)
        self.rows_prefilled += sum(p.shape[1] - b for p, b in zip(host, reused))
        return reused

    def decode(self, batch: int, nframes: int, sp: Optional[_lib.Sampling], use_graph: bool = True) -> None:
        import os

        if os.environ.get("FSB_NO_GRAPH") == "1":  # diagnostic: eager launches instead of graph replays
            use_graph = False
        if not self.slot_control and self._ctx_bound - 1 + nframes > self.kv_len:
            raise ValueError(f"decode: {nframes} more frames from position {self._ctx_bound - 1} run past the KV "
                             f"cache ({self.kv_len} positions)")
        self._grow_bound(self._ctx_bound + nframes)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.fsb_lm_decode(self.h, batch, nframes, C.byref(sp) if sp is not None else None,
                                               int(use_graph), _stream()))
