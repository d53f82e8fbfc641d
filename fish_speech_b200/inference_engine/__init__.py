"""fish_speech/inference_engine/__init__.py surface: TTSInferenceEngine(llama_queue, decoder_model,
precision, compile).inference(req) -> Generator[InferenceResult] (:22-142), send_Llama_request (:144),
get_audio_segment (:179). The LM worker thread and this caller's codec calls drive one GPU from two host
threads, exactly as in the reference; each uses its own handle / stream."""
from __future__ import annotations

import queue
from typing import Generator

import numpy as np
import torch

from ..models.dac.modded_dac import DAC
from ..models.text2semantic.inference import GenerateRequest, GenerateResponse, WrappedGenerateResponse
from .reference_loader import ReferenceLoader
from .schema import ServeTTSRequest
from .utils import InferenceResult, wav_chunk_header
from .vq_manager import VQManager

try:
    from loguru import logger
except Exception:  # pragma: no cover
    import logging

    logger = logging.getLogger("fish_speech_b200")


def set_seed(seed: int):
    """fish_speech/utils/utils.py:120-134."""
    import random

    if seed < 0:
        seed = -seed
    if seed > (1 << 31):
        seed = 1 << 31
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


class TTSInferenceEngine(ReferenceLoader, VQManager):
    def __init__(self, llama_queue: queue.Queue, decoder_model: DAC, precision: torch.dtype, compile: bool) -> None:
        super().__init__()
        self.llama_queue = llama_queue
        self.decoder_model = decoder_model
        self.precision = precision
        self.compile = compile

    @torch.inference_mode()
    def inference(self, req: ServeTTSRequest) -> Generator[InferenceResult, None, None]:
        ref_id = req.reference_id
        prompt_tokens, prompt_texts = [], []
        if ref_id is not None:
            prompt_tokens, prompt_texts = self.load_by_id(ref_id, req.use_memory_cache)
        elif req.references:
            prompt_tokens, prompt_texts = self.load_by_hash(req.references, req.use_memory_cache)
        if req.seed is not None:
            set_seed(req.seed)
            logger.warning(f"set seed: {req.seed}")
        response_queue = self.send_Llama_request(req, prompt_tokens, prompt_texts)
        sample_rate = self.decoder_model.sample_rate
        if req.streaming:
            yield InferenceResult(code="header", audio=(sample_rate, np.array(wav_chunk_header(sample_rate=sample_rate))),
                                  error=None)
        segments = []
        while True:
            wrapped: WrappedGenerateResponse = response_queue.get()
            if wrapped.status == "error":
                yield InferenceResult(code="error", audio=None,
                                      error=wrapped.response if isinstance(wrapped.response, Exception)
                                      else Exception("Unknown error"))
                break
            if not isinstance(wrapped.response, GenerateResponse):
                raise TypeError(f"Expected GenerateResponse, got {type(wrapped.response).__name__}")
            result: GenerateResponse = wrapped.response
            if result.action != "next":
                segment = self.get_audio_segment(result)
                if req.streaming:
                    yield InferenceResult(code="segment", audio=(sample_rate, segment), error=None)
                segments.append(segment)
            else:
                break
        if len(segments) == 0:
            yield InferenceResult(code="error", audio=None,
                                  error=RuntimeError("No audio generated, please check the input text."))
        else:
            yield InferenceResult(code="final", audio=(sample_rate, np.concatenate(segments, axis=0)), error=None)
        return None

    def send_Llama_request(self, req: ServeTTSRequest, prompt_tokens: list, prompt_texts: list) -> queue.Queue:
        request = dict(
            device=self.decoder_model.device, max_new_tokens=req.max_new_tokens, text=req.text, top_p=req.top_p,
            repetition_penalty=req.repetition_penalty, temperature=req.temperature, compile=self.compile,
            iterative_prompt=req.chunk_length > 0, chunk_length=req.chunk_length, prompt_tokens=prompt_tokens,
            prompt_text=prompt_texts)
        response_queue = queue.Queue()
        self.llama_queue.put(GenerateRequest(request=request, response_queue=response_queue))
        return response_queue

    def get_audio_segment(self, result: GenerateResponse) -> np.ndarray:
        # the codec computes in bf16 with fp32 accumulation regardless of `precision` (the reference wraps
        # this call in torch.autocast(bf16), inference_engine/__init__.py:185-189)
        segment = self.decode_vq_tokens(codes=result.codes)
        return segment.float().cpu().numpy()
