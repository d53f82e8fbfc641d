"""Continuous batching over the engine's KV slots.

The reference serves ONE request at a time: `launch_thread_safe_queue` (inference.py:748-799) pops a
request, runs `generate` (prefill, then the `decode_n_tokens` loop with a host-side <|im_end|> test per
frame, inference.py:184-238) to completion, then pops the next one. Here up to `max_slots` requests
share every decode frame: a request is prefilled into a free slot while the others are mid-utterance,
its stop rule runs on the device, and its slot is handed to the next waiting request as soon as it
finishes. The kernels are batch-invariant and a request's random stream depends on its own seed and
frame index only, so every request returns exactly the tokens `generate` would return for it alone.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass, field
from typing import Callable, Optional

import torch

from .engine import bf16_round


@dataclass(eq=False)  # identity semantics: requests hold tensors
class SlotRequest:
    """One `generate` call waiting for / occupying a slot."""

    prompt: torch.Tensor  # [C+1, T] integer
    max_new_tokens: int
    temperature: float = 1.0
    top_p: float = 0.9
    top_k: int = 30
    seed: int = 0
    reuse_prefix: bool = False  # prefill only the rows whose K/V no slot already holds (LmEngine.prefill_reusing)
    on_done: Optional[Callable[["SlotRequest"], None]] = None
    # streaming: called after a poll with the codes [C, k] of frames that are final, in order; over the
    # whole request exactly result[1:, T:-1] (what generate_long keeps, inference.py:708), before on_done
    on_frames: Optional[Callable[["SlotRequest", torch.Tensor], None]] = None
    tag: object = None
    # filled by the batcher
    result: Optional[torch.Tensor] = None  # [C+1, T+n] like generate()
    error: Optional[BaseException] = None
    done: threading.Event = field(default_factory=threading.Event)
    slot: int = -1
    _limit: int = 0
    _emitted: int = 0


class ContinuousBatcher:
    def __init__(self, model, max_slots: int = 32, frames_per_poll: int = 8, min_free_to_admit: int = 1):
        """`frames_per_poll`: decode frames between two looks at the device-side slot states (a finished slot
        idles for at most that many frames). `min_free_to_admit`: admissions wait until that many slots are
        free (or nothing is running) — every admission is one extra prefill pass + first frame that streams
        all weights for the new rows only, so grouping them trades admission latency for throughput."""
        from .models.text2semantic.inference import _ensure_engine

        if not 1 <= max_slots <= 32:
            raise ValueError("max_slots must be in 1..32")
        self.model = model
        self.cfg = model.config
        self.eng = _ensure_engine(model, max_slots)
        self.max_slots = max_slots
        self.frames_per_poll = int(frames_per_poll)
        self.min_free_to_admit = max(1, int(min_free_to_admit))
        self.waiting: list[SlotRequest] = []
        self.active: dict[int, SlotRequest] = {}
        self._len: dict[int, int] = {}  # slot -> upper bound of its context length
        self.frames_run = 0  # decode frames launched (each serves every active slot)
        self.slot_frames = 0  # sum over frames of the active slots (occupancy numerator)
        eng = self.eng
        eng.reset()
        getattr(eng, "_slot_tokens", {}).clear()  # idle rows of the frames below write position 0 of unused slots
        eng.set_slot_control(True)
        self._state = eng.buffer("slot_state")
        self._limit = eng.buffer("slot_limit")
        self._temp = eng.buffer("slot_temperature")
        self._top_p = eng.buffer("slot_top_p")
        self._top_k = eng.buffer("slot_top_k")
        self._seed = eng.buffer("slot_seed")
        self._n_out = eng.buffer("n_out")
        self._pos = eng.buffer("pos")
        self._ras = eng.buffer("ras_window")
        self._out = eng.buffer("out_tokens")

    def close(self):
        """Give the engine back to the one-request-at-a-time entry points."""
        self.eng.set_slot_control(False)
        self.eng.reset()

    # ------------------------------------------------------------------------------------------
    def submit(self, req: SlotRequest) -> SlotRequest:
        """Queue a request. Everything that can be wrong with ONE request is rejected here, with a ValueError the
        caller can hand back to that request alone; what reaches the engine later cannot fail per request."""
        cfg = self.cfg
        p = req.prompt
        if not isinstance(p, torch.Tensor) or p.ndim != 2 or p.size(0) != cfg.num_codebooks + 1 or p.size(1) < 1:
            raise ValueError(f"prompt must be an integer tensor [{cfg.num_codebooks + 1}, T>=1], got "
                             f"{tuple(p.shape) if isinstance(p, torch.Tensor) else type(p)}")
        if p.is_floating_point() or p.dtype == torch.bool:
            raise ValueError(f"prompt must hold integer token ids / codes, got {p.dtype}")
        # the reference asserts 0 < top_p <= 1 and 0 < temperature < 2 (inference.py:541-542); top_k >= 1 keeps
        # rank 0 only at 1. The sampler materialises at most 256 ranks: a larger top_k behaves like 256.
        if not (0.0 < float(req.top_p) <= 1.0):
            raise ValueError(f"top_p must be in (0, 1], got {req.top_p}")
        if not (0.0 < float(req.temperature) < 2.0):
            raise ValueError(f"temperature must be in (0, 2), got {req.temperature}")
        if int(req.top_k) < 1:
            raise ValueError(f"top_k must be >= 1, got {req.top_k}")
        T = int(req.prompt.size(1))
        if T >= cfg.max_seq_len:  # inference.py:262-265
            raise ValueError(f"Input sequence length {T} exceeds max_seq_len {cfg.max_seq_len}")
        n = int(req.max_new_tokens) if req.max_new_tokens else cfg.max_seq_len - T
        n = min(n, cfg.max_seq_len - T, self.eng.max_frames, self.eng.kv_len - T)
        if n < 1:
            raise ValueError("no room for new tokens")
        req._limit = n
        self.waiting.append(req)
        return req

    def idle(self) -> bool:
        return not self.waiting and not self.active

    def _admit(self):
        free = [s for s in range(self.max_slots) if s not in self.active]
        if not free or not self.waiting:
            return
        if self.active and len(free) < min(self.min_free_to_admit, len(self.waiting)):
            return
        batch, slots = [], []
        budget = self.eng.max_rows  # one prefill pass worth of rows per step keeps decode latency bounded
        while free and self.waiting:
            T = int(self.waiting[0].prompt.size(1))
            if batch and T > budget:
                break
            req = self.waiting.pop(0)
            s = free.pop(0)
            req.slot = s
            batch.append(req)
            slots.append(s)
            budget -= T
        idx = torch.tensor(slots, dtype=torch.long, device=self._state.device)

        def put(buf, vals, dtype):
            buf[idx] = torch.tensor(vals, dtype=dtype, device=buf.device)

        put(self._limit, [r._limit for r in batch], torch.int32)
        put(self._temp, [bf16_round(r.temperature) for r in batch], torch.float32)
        put(self._top_p, [bf16_round(r.top_p) for r in batch], torch.float32)
        put(self._top_k, [int(r.top_k) for r in batch], torch.int32)
        put(self._seed, [int(r.seed) & 0x7FFFFFFFFFFFFFFF for r in batch], torch.int64)
        self._n_out[idx] = 0
        self._ras[idx] = 0
        self._state[idx] = 1
        for r in batch:
            self.active[r.slot] = r
            self._len[r.slot] = int(r.prompt.size(1)) + 1
        self.eng.set_context_bound_exact(max(self._len.values()) + 1)
        if any(r.reuse_prefix for r in batch) and hasattr(self.eng, "prefill_reusing"):
            self.eng.prefill_reusing([r.prompt for r in batch], slots, None, do_sample=True)
        else:
            self.eng.prefill([r.prompt for r in batch], slots, None, do_sample=True)

    def _retire(self) -> list[SlotRequest]:
        st = torch.stack([self._state.to(torch.int32), self._n_out]).cpu()  # one small D2H copy (synchronises)
        for s, req in self.active.items():
            # frames of a streaming request leave as soon as they are final: all but the newest one (the last frame of
            # a request -- <|im_end|> or the one at the budget -- is never part of what generate_long keeps)
            if req.on_frames is None:
                continue
            n = int(st[1, s])
            if n - 1 > req._emitted:
                req.on_frames(req, self._out[s, 1:, req._emitted: n - 1].cpu())
                req._emitted = n - 1
        finished = []
        for s, req in list(self.active.items()):
            if int(st[0, s]) != 2:
                continue
            n = int(st[1, s])
            gen = self._out[s, :, :n].to(req.prompt.dtype)
            req.result = torch.cat([req.prompt.to(gen.device), gen], dim=1)
            finished.append(req)
            del self.active[s]
            del self._len[s]
        if finished:
            idx = torch.tensor([r.slot for r in finished], dtype=torch.long, device=self._state.device)
            self._state[idx] = 0
            self._pos[idx] = -1  # parked: the decode frames neither attend for this row nor touch its K/V (which a
            #                      later request with the same prompt prefix may reuse)
        for req in finished:
            req.done.set()
            if req.on_done is not None:
                req.on_done(req)
        return finished

    def step(self) -> list[SlotRequest]:
        """Admit waiting requests into free slots, run `frames_per_poll` frames, retire what finished."""
        self._admit()
        if not self.active:
            return []
        k = self.frames_per_poll
        for s in self._len:
            self._len[s] += k
        self.eng.set_context_bound_exact(min(max(self._len.values()) + 1, self.eng.kv_len))
        # every slot row is decoded: idle and frozen rows cost one attention position and nothing else
        # (the GEMMs stream the weights once for 32 activation rows whatever the number of live ones)
        self.eng.decode(self.max_slots, k, None, use_graph=True)
        self.frames_run += k
        self.slot_frames += k * len(self.active)
        return self._retire()

    def run(self) -> None:
        while not self.idle():
            self.step()
