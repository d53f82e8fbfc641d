"""Host-side protocol of the drop-in boundary (no GPU): the LM worker queue
(launch_thread_safe_queue, inference.py:736-799) and TTSInferenceEngine.inference
(inference_engine/__init__.py:40-142) with the device work replaced by stand-ins."""
import queue

import numpy as np
import pytest
import torch

import fish_speech_b200.models.text2semantic.inference as inf
from fish_speech_b200.inference_engine import TTSInferenceEngine
from fish_speech_b200.inference_engine.schema import ServeTTSRequest
from fish_speech_b200.models.dac.modded_dac import DAC


class _FakeModel:
    class config:
        max_seq_len = 64

    dtype = torch.bfloat16

    def setup_caches(self, **kw):
        self.caches = kw


def _fake_generate_long(*, model, decode_one_token, text, **kw):
    if text == "boom":
        raise ValueError("bad request")
    yield inf.GenerateResponse(action="sample", codes=torch.zeros(10, 3, dtype=torch.long), text=text)
    yield inf.GenerateResponse(action="next")


def test_worker_queue_protocol(monkeypatch):
    monkeypatch.setattr(inf, "init_model", lambda *a, **k: (_FakeModel(), inf.decode_one_token_ar))
    monkeypatch.setattr(inf, "generate_long", _fake_generate_long)
    q = inf.launch_thread_safe_queue("ckpt", "cpu", torch.bfloat16)
    rq = queue.Queue()
    q.put(inf.GenerateRequest(request=dict(text="hello"), response_queue=rq))
    a, b = rq.get(timeout=10), rq.get(timeout=10)
    assert a.status == "success" and a.response.action == "sample" and a.response.codes.shape == (10, 3)
    assert b.status == "success" and b.response.action == "next"
    q.put(inf.GenerateRequest(request=dict(text="boom"), response_queue=rq))
    e = rq.get(timeout=10)
    assert e.status == "error" and isinstance(e.response, ValueError)
    q.put(None)  # shutdown sentinel


def _fake_dac():
    d = object.__new__(DAC)  # no GPU: bypass __init__, keep isinstance(x, DAC) true (vq_manager.py:19)
    d.sample_rate = 44100
    d._device = torch.device("cpu")
    d.from_indices = lambda codes: torch.full((codes.shape[0], 1, codes.shape[-1] * 2048), 0.25)
    return d


def _serve(q):
    while True:
        item = q.get()
        if item is None:
            return
        text = item.request["text"]
        if text == "fail":
            item.response_queue.put(inf.WrappedGenerateResponse(status="error", response=RuntimeError("lm died")))
            continue
        for _ in range(2):
            item.response_queue.put(inf.WrappedGenerateResponse(
                status="success", response=inf.GenerateResponse(action="sample", codes=torch.zeros(10, 2, dtype=torch.long))))
        item.response_queue.put(inf.WrappedGenerateResponse(status="success", response=inf.GenerateResponse(action="next")))


def test_tts_engine_streaming_and_errors():
    import threading

    q = queue.Queue()
    threading.Thread(target=_serve, args=(q,), daemon=True).start()
    eng = TTSInferenceEngine(q, _fake_dac(), torch.bfloat16, compile=False)
    res = list(eng.inference(ServeTTSRequest(text="hi", streaming=True, seed=3)))
    assert [r.code for r in res] == ["header", "segment", "segment", "final"]
    assert res[0].audio[0] == 44100 and res[-1].audio[1].shape == (2 * 2 * 2048,)
    assert res[-1].audio[1].dtype == np.float32 and np.allclose(res[-1].audio[1], 0.25)
    res = list(eng.inference(ServeTTSRequest(text="hi")))
    assert [r.code for r in res] == ["final"]
    res = list(eng.inference(ServeTTSRequest(text="fail")))
    assert res[0].code == "error" and isinstance(res[0].error, RuntimeError)
    q.put(None)


def test_wav_chunk_header_is_riff():
    from fish_speech_b200.inference_engine.utils import wav_chunk_header

    h = wav_chunk_header(44100)
    assert h[:4] == b"RIFF" and h[8:12] == b"WAVE"


def test_generate_long_validates_sampling_args():
    with pytest.raises(AssertionError):
        next(inf.generate_long(model=None, device="cpu", decode_one_token=None, text="x", top_p=0.0))
    with pytest.raises(AssertionError):
        next(inf.generate_long(model=None, device="cpu", decode_one_token=None, text="x", temperature=2.5))


class _FakeBatcher:
    """Stands in for scheduler.ContinuousBatcher: a request finishes after ceil(limit / 4) steps and its
    result is the prompt followed by `max_new_tokens` copies of its seed's low byte."""

    def __init__(self, slots):
        self.slots, self.waiting, self.active, self.max_active, self.closed = slots, [], [], 0, False

    def submit(self, r):
        self.waiting.append(r)

    def idle(self):
        return not self.waiting and not self.active

    def step(self):
        while self.waiting and len(self.active) < self.slots:
            r = self.waiting.pop(0)
            self.active.append([r, (r.max_new_tokens + 3) // 4])
        self.max_active = max(self.max_active, len(self.active))
        for e in list(self.active):
            e[1] -= 1
            if e[1] <= 0:
                self.active.remove(e)
                r = e[0]
                r.result = torch.cat([r.prompt, torch.full((r.prompt.size(0), r.max_new_tokens), 7)], dim=1)
                r.on_done(r)

    def close(self):
        self.closed = True


def _fake_plan(*, model, text, chunks=2, fail=False, **kw):
    """Shape of _generate_long_plan: per chunk one generate request, then a sample response; 'next' at the end."""
    for c in range(chunks):
        y = yield ("generate", dict(prompt=torch.full((3, 2 + c), len(text)), max_new_tokens=4 * (c + 1) + len(text),
                                    audio_masks=None, audio_parts=None, temperature=0.7, top_p=0.7, top_k=30))
        if fail:
            raise RuntimeError("boom")
        yield ("response", inf.GenerateResponse(action="sample", codes=y[1:, 2 + c:], text=f"{text}#{c}"))
    yield ("response", inf.GenerateResponse(action="next"))


def test_slot_scheduler_worker_interleaves_requests(monkeypatch):
    """serve_requests: several queued requests advance together, each response queue sees the reference's
    sequence (sample per chunk, then next), an exception reaches only its own request, None shuts down."""
    import queue as Q

    monkeypatch.setattr(inf, "_generate_long_plan", _fake_plan)
    fb = _FakeBatcher(slots=3)
    q = Q.Queue()
    outs = []
    for i, (text, fail) in enumerate([("a", False), ("bbbbbb", False), ("cc", True), ("ddd", False)]):
        rq = Q.Queue()
        outs.append(rq)
        q.put(inf.GenerateRequest(request=dict(text=text, fail=fail, chunks=2), response_queue=rq))
    q.put(None)
    inf.serve_requests(_FakeModel(), q, 3, batcher=fb)
    assert fb.closed and fb.max_active == 3  # three requests really shared the scheduler steps

    def drain(rq):
        items = []
        while not rq.empty():
            items.append(rq.get_nowait())
        return items

    for i, text in [(0, "a"), (1, "bbbbbb"), (3, "ddd")]:
        items = drain(outs[i])
        assert [it.status for it in items] == ["success"] * 3
        assert [it.response.action for it in items] == ["sample", "sample", "next"]
        assert [it.response.text for it in items[:2]] == [f"{text}#0", f"{text}#1"]
        assert items[0].response.codes.shape == (2, 4 + len(text))
        assert items[1].response.codes.shape == (2, 8 + len(text))
    bad = drain(outs[2])
    assert len(bad) == 1 and bad[0].status == "error" and isinstance(bad[0].response, RuntimeError)


def test_generate_long_drives_the_same_plan(monkeypatch):
    """generate_long = the plan driven one `generate` at a time (inference.py:611-721)."""
    monkeypatch.setattr(inf, "_generate_long_plan", _fake_plan)
    calls = []

    def fake_generate(*, model, prompt, max_new_tokens, seed, **kw):
        calls.append(seed)
        return torch.cat([prompt, torch.zeros(prompt.size(0), max_new_tokens, dtype=prompt.dtype)], dim=1)

    monkeypatch.setattr(inf, "generate", fake_generate)
    out = list(inf.generate_long(model=_FakeModel(), decode_one_token=None, text="xy", chunks=3))
    assert [r.action for r in out] == ["sample"] * 3 + ["next"]
    assert len(set(calls)) == 3  # every generate call gets its own Philox key


class _FakeEngine:
    """Host-visible behaviour of LmEngine under slot control, on CPU tensors: prefill writes the first frame of
    every admitted slot, a decode frame appends one token per live slot and freezes a slot at its limit."""

    def __init__(self, slots, kv_len=64, max_frames=32, max_rows=16):
        self.max_batch, self.kv_len, self.max_frames, self.max_rows = slots, kv_len, max_frames, max_rows
        self.device = torch.device("cpu")
        z = lambda *s, dt=torch.int32: torch.zeros(*s, dtype=dt)
        self.bufs = dict(slot_state=z(slots), slot_limit=z(slots), slot_temperature=z(slots, dt=torch.float32),
                         slot_top_p=z(slots, dt=torch.float32), slot_top_k=z(slots), slot_seed=z(slots, dt=torch.int64),
                         n_out=z(slots), pos=z(32), ras_window=z(slots, 10), out_tokens=z(slots, 3, max_frames))
        self.slot_control = False
        self.prefills, self.decodes, self.bounds = [], 0, []

    def buffer(self, name):
        return self.bufs[name]

    def reset(self):
        for k in ("slot_state", "n_out", "pos"):
            self.bufs[k].zero_()

    def set_slot_control(self, on):
        self.slot_control = bool(on)

    def set_context_bound_exact(self, n):
        self.bounds.append(n)

    def _frame(self, s):
        b = self.bufs
        if int(b["slot_state"][s]) != 1:
            return
        f = int(b["n_out"][s])
        b["out_tokens"][s, :, f] = 100 * s + f
        b["n_out"][s] = f + 1
        if f + 1 >= int(b["slot_limit"][s]):
            b["slot_state"][s] = 2
        else:
            b["pos"][s] += 1

    def prefill(self, prompts, slots, sp, start_pos=None, do_sample=True):
        assert sp is None and self.slot_control
        self.prefills.append(list(slots))
        for p, s in zip(prompts, slots):
            self.bufs["pos"][s] = p.shape[1]
            self._frame(s)

    def decode(self, batch, nframes, sp, use_graph=True):
        assert sp is None and batch == self.max_batch
        self.decodes += nframes
        for _ in range(nframes):
            for s in range(batch):
                self._frame(s)


class _SchedModel:
    def __init__(self, eng, max_seq_len=64):
        self.engine, self.max_batch_size = eng, eng.max_batch
        self.config = type("C", (), dict(max_seq_len=max_seq_len, num_codebooks=2))()


def test_continuous_batcher_host_logic():
    """scheduler.ContinuousBatcher against a fake engine: budgets (generate's max_new_tokens clamp, inference.py:
    270-279), slot reuse, grouped admissions, retirement order, engine handed back on close."""
    from fish_speech_b200.scheduler import ContinuousBatcher, SlotRequest

    eng = _FakeEngine(slots=2)
    model = _SchedModel(eng)
    b = ContinuousBatcher(model, max_slots=2, frames_per_poll=4)
    assert eng.slot_control
    order = []
    mk = lambda T, n, tag: SlotRequest(prompt=torch.full((3, T), tag), max_new_tokens=n, tag=tag,
                                       on_done=lambda r: order.append(r.tag))
    reqs = [b.submit(mk(5, 3, 0)), b.submit(mk(7, 9, 1)), b.submit(mk(4, 1, 2)), b.submit(mk(6, 0, 3))]
    assert reqs[3]._limit == min(64 - 6, eng.max_frames)  # max_new_tokens=0 -> as many as fit
    with pytest.raises(ValueError):
        b.submit(mk(64, 4, 9))  # prompt as long as max_seq_len (inference.py:262-265)
    b.run()
    assert b.idle() and sorted(order) == [0, 1, 2, 3] and order[0] == 0  # the 3-frame request leaves first
    assert eng.prefills[0] == [0, 1] and all(len(p) <= 2 for p in eng.prefills)  # two slots, reused afterwards
    assert sum(len(p) for p in eng.prefills) == 4
    for r, n in zip(reqs, [3, 9, 1, reqs[3]._limit]):
        T = r.prompt.size(1)
        assert r.result.shape == (3, T + n) and torch.equal(r.result[:, :T], r.prompt)
        assert r.result[0, T:].tolist() == [100 * r.slot + f for f in range(n)]  # its own frames, in order
    assert 0 < b.slot_frames <= b.frames_run * 2 and max(eng.bounds) <= eng.kv_len + 1
    b.close()
    assert not eng.slot_control and int(eng.bufs["slot_state"].sum()) == 0

    # grouped admissions: with min_free_to_admit=2 a single freed slot waits for the second one
    eng2 = _FakeEngine(slots=2)
    b2 = ContinuousBatcher(_SchedModel(eng2), max_slots=2, frames_per_poll=2, min_free_to_admit=2)
    for tag, n in enumerate([2, 8, 4, 4]):
        b2.submit(mk(3, n, tag))
    b2.run()
    assert eng2.prefills == [[0, 1], [0, 1]]  # the short request's slot was not refilled on its own


class _FakeCodec:
    """encode() of a DAC: codes[b, c, t] = c + 10 * (sample count of clip b) for t < ceil(len / hop)."""
    sample_rate, hop = 100, 10
    device = torch.device("cpu")

    def __init__(self):
        self.batches = []

    def encode(self, audios, lengths):
        B, _, N = audios.shape
        self.batches.append((B, N, lengths.tolist()))
        T = -(-N // self.hop)
        codes = torch.arange(3).view(1, 3, 1).expand(B, 3, T) + 10 * lengths.view(B, 1, 1)
        return codes.to(torch.int64), -(-lengths // self.hop)


def _write_wav(path, n, sr=100, nch=1):
    import wave

    with wave.open(str(path), "wb") as f:
        f.setnchannels(nch)
        f.setsampwidth(2)
        f.setframerate(sr)
        f.writeframes((np.arange(n * nch) % 100).astype("<i2").tobytes())


def test_bulk_encode_pipeline(tmp_path, monkeypatch):
    """fish_speech_b200/bulk_encode.py (the job of tools/vqgan/extract_vq.py): shard stride, skip of finished files,
    length-grouped padded batches, per-file trimming of the codes, unreadable files skipped, stereo down-mix."""
    from fish_speech_b200 import bulk_encode as be

    lens = [95, 12, 40, 41, 230, 13, 39, 11]
    for i, n in enumerate(lens):
        _write_wav(tmp_path / f"a{i}.wav", n, nch=2 if i == 2 else 1)
    (tmp_path / "broken.wav").write_bytes(b"not a wav")
    (tmp_path / "notes.txt").write_text("x")
    np.save(tmp_path / "a1.npy", np.zeros((3, 2)))  # already done -> skipped
    files = be.list_audio_files(tmp_path)
    assert [f.name for f in files] == sorted([f"a{i}.wav" for i in range(8)] + ["broken.wav"])

    monkeypatch.setenv("SLURM_PROCID", "1")
    monkeypatch.setenv("SLURM_NTASKS", "2")
    assert be.worker_identity() == (1, 2)
    todo_all = be.pending_files(files, 0, 1)
    assert tmp_path / "a1.wav" not in todo_all and len(todo_all) == 8
    assert be.pending_files(files, 1, 2) == todo_all[1::2]

    # batch planning: similar lengths share a batch, caps respected, every index exactly once
    plan = be.plan_batches([95, 40, 41, 230, 13, 39, 11], batch_size=3, max_padded_samples=300)
    assert sorted(i for b in plan for i in b) == list(range(7))
    assert all(len(b) <= 3 for b in plan) and [3] in plan  # the 230-sample clip cannot share a 300-sample budget
    assert any(set(b) == {1, 2, 5} or set(b) >= {1, 2} for b in plan)

    codec = _FakeCodec()
    n, secs = be.encode_files(todo_all, codec, batch_size=3, max_batch_seconds=3.0)
    assert n == 7 and abs(secs - sum(lens[i] for i in (0, 2, 3, 4, 5, 6, 7)) / 100) < 1e-6  # broken.wav skipped
    for i in (0, 2, 3, 4, 5, 6, 7):
        got = np.load(tmp_path / f"a{i}.npy")
        assert got.shape == (3, -(-lens[i] // 10)) and got.dtype == np.int64
        assert (got[:, 0] == np.arange(3) + 10 * lens[i]).all()  # its own codes, trimmed to its own length
    assert np.load(tmp_path / "a1.npy").shape == (3, 2)  # untouched
    assert not (tmp_path / "broken.npy").exists() and not list(tmp_path.glob("*.tmp"))
    assert all(B <= 3 and B * N <= 300 for B, N, _ in codec.batches)
    # everything is done now: a second pass finds nothing
    assert be.pending_files(be.list_audio_files(tmp_path), 0, 1) == [tmp_path / "broken.wav"]


def test_extract_vq_spawn_plan():
    import importlib.util
    from pathlib import Path

    spec = importlib.util.spec_from_file_location("extract_vq", Path(__file__).resolve().parent.parent / "tools/vqgan/extract_vq.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    plan = mod.spawn_plan(4, ["0", "3"])
    assert [p["CUDA_VISIBLE_DEVICES"] for p in plan] == ["0", "3", "0", "3"]
    assert [p["SLURM_PROCID"] for p in plan] == ["0", "1", "2", "3"] and {p["SLURM_NTASKS"] for p in plan} == {"4"}


def test_slot_scheduler_worker_reports_engine_failure(monkeypatch):
    """If the batcher itself fails (a CUDA error, say) every request in flight receives an error response, and so
    does every request that arrives afterwards: the worker keeps answering instead of dying with callers blocked on
    their queues (the reference's worker catches per request and keeps serving, inference.py:748-799). `None` ends it;
    a close() that fails on the broken engine does not mask anything."""
    import queue as Q
    import threading

    monkeypatch.setattr(inf, "_generate_long_plan", _fake_plan)

    class Boom(_FakeBatcher):
        def step(self):
            raise RuntimeError("device lost")

        def close(self):
            self.closed = True
            raise RuntimeError("reset failed too")

    fb = Boom(slots=2)
    q, outs = Q.Queue(), [Q.Queue(), Q.Queue()]
    for rq in outs:
        q.put(inf.GenerateRequest(request=dict(text="ab", chunks=1), response_queue=rq))
    t = threading.Thread(target=inf.serve_requests, args=(_FakeModel(), q, 2), kwargs=dict(batcher=fb), daemon=True)
    t.start()
    for rq in outs:
        item = rq.get(timeout=20)
        assert item.status == "error" and "device lost" in str(item.response)
    late = Q.Queue()
    q.put(inf.GenerateRequest(request=dict(text="late", chunks=1), response_queue=late))
    item = late.get(timeout=20)
    assert item.status == "error" and "device lost" in str(item.response)
    q.put(None)
    t.join(timeout=20)
    assert not t.is_alive() and fb.closed


def test_continuous_batcher_rejects_bad_requests_up_front():
    """Per-request errors surface in submit() (so the worker can answer that request alone), not in the engine."""
    from fish_speech_b200.scheduler import ContinuousBatcher, SlotRequest

    b = ContinuousBatcher(_SchedModel(_FakeEngine(slots=2)), max_slots=2)
    ok = dict(prompt=torch.zeros(3, 4, dtype=torch.long), max_new_tokens=4)
    for bad in (dict(top_k=0), dict(top_p=0.0), dict(top_p=1.5), dict(temperature=0.0), dict(temperature=2.0),
                dict(prompt=torch.zeros(2, 4, dtype=torch.long)), dict(prompt=torch.zeros(3, 4)),
                dict(prompt=torch.zeros(3, 0, dtype=torch.long))):
        with pytest.raises(ValueError):
            b.submit(SlotRequest(**{**ok, **bad}))
    b.submit(SlotRequest(**ok))
    assert len(b.waiting) == 1
