"""fish_speech/inference_engine/reference_loader.py surface: reference audio -> VQ codes with the
id / sha256 caches (load_by_id :62, load_by_hash :99, load_audio :133). Reference-library file management
(add / delete) stays with the reference's serving shell."""
from __future__ import annotations

import io
from hashlib import sha256
from pathlib import Path
from typing import Callable, Literal, Tuple

import torch

AUDIO_EXTENSIONS = {".mp3", ".wav", ".flac", ".ogg", ".m4a", ".wma", ".aac", ".aiff", ".aif", ".aifc"}

try:
    from loguru import logger
except Exception:  # pragma: no cover
    import logging

    logger = logging.getLogger("fish_speech_b200")


class ReferenceLoader:
    def __init__(self) -> None:
        self.ref_by_id: dict = {}
        self.ref_by_hash: dict = {}
        self.encode_reference: Callable
        self.backend = "soundfile"

    def load_by_id(self, id: str, use_cache: Literal["on", "off"]) -> Tuple:
        import re

        if not re.match(r"^[a-zA-Z0-9\-_ ]+$", id) or len(id) > 255:
            raise ValueError("Reference ID contains invalid characters or is too long. "
                             "Only alphanumeric, hyphens, underscores, and spaces are allowed.")
        ref_folder = Path("references") / id
        ref_folder.mkdir(parents=True, exist_ok=True)
        ref_audios = [p for p in ref_folder.rglob("*") if p.suffix.lower() in AUDIO_EXTENSIONS]
        if use_cache == "off" or id not in self.ref_by_id:
            prompt_tokens = [self.encode_reference(reference_audio=p.read_bytes(), enable_reference_audio=True)
                             for p in ref_audios]
            prompt_texts = [p.with_suffix(".lab").read_text(encoding="utf-8") for p in ref_audios]
            self.ref_by_id[id] = (prompt_tokens, prompt_texts)
        else:
            logger.info("Use same references")
            prompt_tokens, prompt_texts = self.ref_by_id[id]
        return prompt_tokens, prompt_texts

    def load_by_hash(self, references: list, use_cache: Literal["on", "off"]) -> Tuple:
        hashes = [sha256(ref.audio).hexdigest() for ref in references]
        cache_used = False
        prompt_tokens, prompt_texts = [], []
        for i, ref in enumerate(references):
            if use_cache == "off" or hashes[i] not in self.ref_by_hash:
                prompt_tokens.append(self.encode_reference(reference_audio=ref.audio, enable_reference_audio=True))
                prompt_texts.append(ref.text)
                self.ref_by_hash[hashes[i]] = (prompt_tokens[-1], ref.text)
            else:
                tok, txt = self.ref_by_hash[hashes[i]]
                prompt_tokens.append(tok)
                prompt_texts.append(txt)
                cache_used = True
        if cache_used:
            logger.info("Use same references")
        return prompt_tokens, prompt_texts

    def load_audio(self, reference_audio, sr: int):
        import torchaudio

        if len(reference_audio) > 255 or not Path(reference_audio).exists():
            reference_audio = io.BytesIO(reference_audio)
        waveform, original_sr = torchaudio.load(reference_audio, backend=self.backend)
        if waveform.shape[0] > 1:
            waveform = torch.mean(waveform, dim=0, keepdim=True)
        if original_sr != sr:
            waveform = torchaudio.transforms.Resample(orig_freq=original_sr, new_freq=sr)(waveform)
        return waveform.squeeze().numpy()

    def list_reference_ids(self) -> list[str]:
        base = Path("references")
        if not base.exists():
            return []
        ids = []
        for d in base.iterdir():
            if d.is_dir() and any(p.suffix.lower() in AUDIO_EXTENSIONS and p.with_suffix(".lab").exists()
                                  for p in d.iterdir()):
                ids.append(d.name)
        return sorted(ids)
