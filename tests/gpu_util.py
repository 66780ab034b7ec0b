"""Shared helpers for the -m gpu tests (oracle/ is used here only as the checker)."""
from types import SimpleNamespace

import numpy as np
import torch

from oracle import fixtures
from oracle.t5gemma_voice_oracle import Oracle, OracleConfig
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine, GenerationRequest

_engines = {}


def engine_for(name: str, **sizing) -> T5GemmaVoiceEngine:
    key = (name, tuple(sorted(sizing.items())))
    if key not in _engines:
        _, sd, meta = fixtures.load_model_fixture(name)
        ns = SimpleNamespace(**meta)
        sz = dict(max_slots=1, max_text_len=64, max_dec_len=512, max_prefill_tokens=512)
        sz.update(sizing)
        _engines[key] = T5GemmaVoiceEngine.from_state_dict(ns, sd, **sz)
    return _engines[key]


def bf16_round_oracle(name: str) -> Oracle:
    """Oracle whose >=2-D weights are rounded to bf16 (what the engine stores); isolates kernel logic
    from the weight-quantisation error."""
    cfg, sd, _ = fixtures.load_model_fixture(name)
    sd2 = {k: (v.to(torch.bfloat16).float() if v.dim() >= 2 else v) for k, v in sd.items()}
    return Oracle(cfg, sd2)


def rel_err(got: np.ndarray, ref: np.ndarray) -> float:
    """max-abs error / abs-max of the reference tensor (SURVEY.md section 7 normalisation)."""
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))
