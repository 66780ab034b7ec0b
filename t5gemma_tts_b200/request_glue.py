"""Host-side request assembly around the hot path, mirroring inference_tts_utils.inference_one_sample
(/root/reference/inference_tts_utils.py:229-286 prompt/y_sep/text/tgt_y_lens, :323-354 _strip_sep_and_eos) so that a
batched front door can feed `T5GemmaVoiceEngine.inference_tts_batch` with exactly what the bs=1 reference glue would
have produced per utterance.  Tokenisers (text SentencePiece, XCodec2) stay outside: this takes/returns ids."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .engine import GenerationRequest


def build_request(cfg, target_text_ids: Sequence[int], target_seconds: float, prompt_codes: Optional[Sequence[int]] = None,
                  prefix_text_ids: Optional[Sequence[int]] = None, top_k=30, top_p=0.9, min_p=0.0, temperature=0.8,
                  stop_repetition: int = 3, silence_tokens: Optional[Sequence[int]] = None) -> GenerationRequest:
    """cfg: the reference config / model_args (y_sep_token, x_sep_token, encodec_sr, add_eos_to_text, add_bos_to_text,
    parallel_pattern).  prompt_codes: XCodec2 codes of the reference speech (None/empty = no voice prompt)."""
    g = lambda n, d=None: getattr(cfg, n, d)
    codec_sr = int(g("encodec_sr", 50))
    prompt = np.asarray(prompt_codes if prompt_codes is not None else [], dtype=np.int64).reshape(-1)
    y_sep = g("y_sep_token", None)
    if y_sep is not None and prompt.size > 0:                       # :229-242 only with a reference prompt
        prompt = np.concatenate([prompt, np.array([int(y_sep)], dtype=np.int64)])
    text = list(int(t) for t in target_text_ids)
    if prefix_text_ids:                                             # :259-264
        x_sep = g("x_sep_token", None)
        text = list(int(t) for t in prefix_text_ids) + ([int(x_sep)] if x_sep is not None else []) + text
    if g("add_eos_to_text", 0):
        text.append(int(g("add_eos_to_text")))
    if g("add_bos_to_text", 0):
        text = [int(g("add_bos_to_text"))] + text
    extra = 2 if int(g("parallel_pattern", 0) or 0) != 0 else 0     # :281-286 (effective_delay_inc = 0 for one codebook)
    tgt = int(prompt.size + codec_sr * target_seconds + extra)
    return GenerationRequest(text_ids=np.asarray(text, dtype=np.int64), prompt_ids=prompt, target_total=tgt,
                             prompt_frames=int(prompt.size), top_k=top_k, top_p=top_p, min_p=min_p,
                             temperature=temperature, stop_repetition=stop_repetition,
                             silence_tokens=list(silence_tokens or []))


def strip_sep_and_eos(frames: np.ndarray, sep_token: Optional[int], eos_token: Optional[int]) -> np.ndarray:
    """Single-codebook equivalent of inference_tts_utils.py:323-354: drops y_sep / eos ids, keeps order.
    frames [1,1,T] (or [T]) int64 -> [1,1,T'] (or [T'])."""
    a = np.asarray(frames)
    flat = a.reshape(-1)
    keep = np.ones(flat.shape, dtype=bool)
    if sep_token is not None:
        keep &= flat != sep_token
    if eos_token is not None:
        keep &= flat != eos_token
    out = flat[keep]
    return out.reshape(1, 1, -1) if a.ndim == 3 else out


def generate_batch(engine, requests: List[GenerationRequest], chunk_steps: int = 32):
    """Batched front door: list of requests -> list of (concat_frames, gen_frames) with sep/eos stripped, i.e. what
    inference_one_sample hands to the codec for every utterance."""
    cfg = engine.cfg
    outs = engine.inference_tts_batch(requests, chunk_steps=chunk_steps)
    res = []
    for concat, gen in outs:
        res.append((strip_sep_and_eos(concat.numpy(), cfg.y_sep_token, cfg.stop_token),
                    strip_sep_and_eos(gen.numpy(), cfg.y_sep_token, cfg.stop_token)))
    return res
