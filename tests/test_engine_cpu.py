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
