"""Reference audio -> VQ prompt codes, with the two caches of the reference's loader
(fish_speech/inference_engine/reference_loader.py: `load_by_id` :62 keyed by a library id under ./references,
`load_by_hash` :99 keyed by the sha256 of the uploaded audio, `load_audio` :133).  Same entry points, arguments and
error text; built around one cached-encode helper, and audio decoding goes through this package's own readers
(fish_speech_b200/bulk_encode.py) instead of a torchaudio backend.  Library management (add / delete of ids) stays
with the reference's serving shell."""
from __future__ import annotations

import re
from hashlib import sha256
from pathlib import Path
from typing import Callable, Hashable, Literal

import torch

AUDIO_EXTENSIONS = {".mp3", ".wav", ".flac", ".ogg", ".m4a", ".wma", ".aac", ".aiff", ".aif", ".aifc"}
LIBRARY = Path("references")
_ID_OK = re.compile(r"^[a-zA-Z0-9\-_ ]+$")

try:
    from loguru import logger
except Exception:  # pragma: no cover
    import logging

    logger = logging.getLogger("fish_speech_b200")


def _is_reference_clip(p: Path) -> bool:
    return p.is_file() and p.suffix.lower() in AUDIO_EXTENSIONS


class ReferenceLoader:
    """Mixin of TTSInferenceEngine: `encode_reference` (VQManager) turns audio bytes into codes; results are memoised
    per library id and per audio hash unless the request says `use_memory_cache="off"`."""

    def __init__(self) -> None:
        self.ref_by_id: dict = {}
        self.ref_by_hash: dict = {}
        self.encode_reference: Callable

    # -- one code path for both caches ---------------------------------------------------------
    def _cached(self, cache: dict, key: Hashable, use_cache: str, make: Callable[[], tuple]) -> tuple[tuple, bool]:
        """(value, came_from_cache). `make` runs the codec; its result replaces whatever the cache held."""
        if use_cache != "off" and key in cache:
            return cache[key], True
        value = make()
        cache[key] = value
        return value, False

    def _encode(self, audio: bytes):
        return self.encode_reference(reference_audio=audio, enable_reference_audio=True)

    # -- the reference's entry points ----------------------------------------------------------
    def load_by_id(self, id: str, use_cache: Literal["on", "off"]) -> tuple[list, list]:
        """All clips of ./references/<id>/ (audio + same-named .lab transcript) as (codes list, texts list)."""
        if len(id) > 255 or not _ID_OK.match(id):
            raise ValueError("Reference ID contains invalid characters or is too long. "
                             "Only alphanumeric, hyphens, underscores, and spaces are allowed.")
        folder = LIBRARY / id
        folder.mkdir(parents=True, exist_ok=True)

        def encode_folder():
            clips = [p for p in folder.rglob("*") if _is_reference_clip(p)]
            return ([self._encode(p.read_bytes()) for p in clips],
                    [p.with_suffix(".lab").read_text(encoding="utf-8") for p in clips])

        (tokens, texts), hit = self._cached(self.ref_by_id, id, use_cache, encode_folder)
        if hit:
            logger.info("Use same references")
        return tokens, texts

    def load_by_hash(self, references: list, use_cache: Literal["on", "off"]) -> tuple[list, list]:
        """Uploaded references (objects with .audio bytes and .text): each is encoded once per distinct audio."""
        tokens, texts, any_hit = [], [], False
        for ref in references:
            (tok, txt), hit = self._cached(self.ref_by_hash, sha256(ref.audio).hexdigest(), use_cache,
                                           lambda ref=ref: (self._encode(ref.audio), ref.text))
            tokens.append(tok)
            texts.append(txt)
            any_hit |= hit
        if any_hit:
            logger.info("Use same references")
        return tokens, texts

    def load_audio(self, reference_audio, sr: int):
        """Mono float32 numpy waveform at `sr` from a path or from encoded bytes (reference_loader.py:133-160)."""
        from ..bulk_encode import decode_audio_bytes, read_audio, resample

        is_path = isinstance(reference_audio, (str, Path)) and len(str(reference_audio)) <= 255 \
            and Path(reference_audio).exists()
        wav, file_sr = read_audio(Path(reference_audio)) if is_path else decode_audio_bytes(bytes(reference_audio))
        return resample(wav.to(torch.float32), file_sr, sr).numpy()

    def list_reference_ids(self) -> list[str]:
        """Ids of the library that hold at least one clip with its transcript."""
        if not LIBRARY.exists():
            return []
        return sorted(d.name for d in LIBRARY.iterdir()
                      if d.is_dir() and any(_is_reference_clip(p) and p.with_suffix(".lab").exists() for p in d.iterdir()))
