"""What `TTSInferenceEngine.inference` yields, and the streaming WAV preamble (fish_speech/inference_engine/utils.py:
`InferenceResult`, `wav_chunk_header`)."""
import struct
from dataclasses import dataclass
from typing import Literal, Optional, Tuple

import numpy as np


@dataclass
class InferenceResult:
    """code "header": audio = (sample_rate, header bytes as uint8 array); "segment" / "final": (sample_rate, float
    waveform); "error": `error` holds the exception."""

    code: Literal["header", "segment", "error", "final"]
    audio: Optional[Tuple[int, np.ndarray]]
    error: Optional[Exception]


def wav_chunk_header(sample_rate: int = 44100, bit_depth: int = 16, channels: int = 1) -> bytes:
    """The 44 bytes a PCM WAV file starts with, for a stream whose length is not known yet (data size 0): a client
    that receives this and then raw little-endian PCM chunks plays the stream."""
    frame_bytes = channels * bit_depth // 8
    fmt = struct.pack("<IHHIIHH", 16, 1, channels, sample_rate, sample_rate * frame_bytes, frame_bytes, bit_depth)
    return b"RIFF" + struct.pack("<I", 4 + (4 + len(fmt)) + 8) + b"WAVE" + b"fmt " + fmt + b"data" + struct.pack("<I", 0)


def pcm16(waveform: np.ndarray) -> bytes:
    """Float waveform in [-1, 1] -> little-endian 16-bit PCM bytes (what follows `wav_chunk_header` on the wire;
    tools/server/inference.py:32-33 does this per streamed segment)."""
    x = np.clip(np.asarray(waveform, dtype=np.float32), -1.0, 1.0)
    return (x * 32767.0).astype("<i2").tobytes()
