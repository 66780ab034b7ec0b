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


def inference_batch(engine, model_args, text_tokenizer, audio_tokenizer, items, decode_config, tokenize_audio_fn=None,
                    normalize_text_fn=None, chunk_steps: int = 32, return_frames: bool = False):
    """Multi-utterance front door: what inference_tts_utils.inference_one_sample (`:141-379`) does for ONE utterance, done
    for a list so that batch sizes > 1 are reachable from the CLI / UI code.  `items` is a sequence of dicts with the
    per-utterance arguments of inference_one_sample: `target_text`, `target_generation_length` (seconds) and optionally
    `audio_fn` (reference speech; None / "" / "none" = no voice prompt), `prompt_end_frame`, `prefix_transcript`, `lang`.
    `decode_config` carries the same keys (`top_k`, `top_p`, `min_p`, `temperature`, `stop_repetition`, `silence_tokens`,
    `codec_sr`).  Tokenisers stay the reference's objects: `text_tokenizer.encode(text, add_special_tokens=False)`,
    `audio_tokenizer.decode(frames)`, and `tokenize_audio_fn(audio_tokenizer, audio_fn, offset=0, num_frames=...)`
    (= data.tokenizer.tokenize_audio); `normalize_text_fn(text, lang) -> (text, lang)` is the reference's
    `normalize_text_with_lang` when Japanese normalisation is wanted.
    Returns a list of `(concat_sample, gen_sample)` (plus the stripped frames with `return_frames`), in item order."""
    import torch
    silence = decode_config.get("silence_tokens", [])
    if isinstance(silence, str):
        import ast
        silence = ast.literal_eval(silence)
    if int(getattr(model_args, "n_codebooks", 1)) != 1:
        raise ValueError("XCodec2 backend supports only n_codebooks=1.")
    cfg = SimpleNamespaceView(model_args, encodec_sr=int(decode_config.get("codec_sr", getattr(model_args, "encodec_sr", 50))))
    requests, has_ref = [], []

    def encode_text(text):
        if isinstance(text, list):
            text = " ".join(text)
        return list(text_tokenizer.encode(text.strip(), add_special_tokens=False))

    for it in items:
        audio_fn = it.get("audio_fn")
        ref = audio_fn is not None and str(audio_fn).lower() not in {"", "none", "null"}
        prompt = None
        if ref:
            if tokenize_audio_fn is None:
                from data.tokenizer import tokenize_audio as tokenize_audio_fn      # the reference's own helper
            pef = int(it.get("prompt_end_frame", -1))
            frames = tokenize_audio_fn(audio_tokenizer, audio_fn, offset=0, num_frames=pef if pef > 0 else -1)
            prompt = torch.as_tensor(frames).reshape(-1).cpu().numpy()
        text, lang = it["target_text"], it.get("lang")
        prefix = it.get("prefix_transcript")
        if normalize_text_fn is not None:
            text, lang = normalize_text_fn(text, lang)
            if prefix:
                prefix, _ = normalize_text_fn(prefix, lang)
        requests.append(build_request(cfg, encode_text(text), float(it["target_generation_length"]), prompt_codes=prompt,
                                      prefix_text_ids=encode_text(prefix) if prefix else None,
                                      top_k=decode_config["top_k"], top_p=decode_config["top_p"], min_p=decode_config.get("min_p", 0.0),
                                      temperature=decode_config["temperature"], stop_repetition=decode_config.get("stop_repetition", 3),
                                      silence_tokens=silence))
        has_ref.append(ref)
    outs = generate_batch(engine, requests, chunk_steps=chunk_steps)
    res = []
    for (concat, gen), ref in zip(outs, has_ref):
        concat_t, gen_t = torch.from_numpy(np.ascontiguousarray(concat)), torch.from_numpy(np.ascontiguousarray(gen))
        gen_sample = audio_tokenizer.decode(gen_t)
        concat_sample = audio_tokenizer.decode(concat_t) if ref else gen_sample
        res.append((concat_sample, gen_sample, concat_t, gen_t) if return_frames else (concat_sample, gen_sample))
    return res


class SimpleNamespaceView:
    """model_args with a few fields overridden (the front door takes codec_sr from decode_config, not from the model)."""

    def __init__(self, base, **over):
        self._base, self._over = base, over

    def __getattr__(self, name):
        over = object.__getattribute__(self, "_over")
        if name in over:
            return over[name]
        return getattr(object.__getattribute__(self, "_base"), name)
