"""EngineConfig: the frozen set of constants the hot path reads.

Derived from the reference's T5GemmaVoiceConfig (hf_export/configuration_t5gemma_voice.py:50-151) or
its argparse namespace (config.py:47-240); special ids derive from the audio vocabulary exactly as
config.py:224-228 does (empty=V, eog=V+1, pad=V+2, eos=V+3, y_sep=V+4)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional


def _get(obj: Any, name: str, default=None):
    if isinstance(obj, dict):
        return obj.get(name, default)
    return getattr(obj, name, default)


@dataclass
class EngineConfig:
    hidden: int = 2304
    inter: int = 9216
    n_enc_layers: int = 26
    n_dec_layers: int = 26
    n_heads: int = 8
    n_kv_heads: int = 4
    head_dim: int = 256
    sliding_window: int = 4096
    query_pre_attn_scalar: float = 256.0
    attn_softcap: Optional[float] = 50.0     # None when attn_implementation != "eager" (sdpa drops it)
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0
    text_vocab: int = 256000
    audio_vocab: int = 65536
    n_special: int = 5
    special_first: int = 0
    progress_scale: float = 2000.0
    encodec_sr: int = 50
    extra_cutoff: float = 5.0
    text_guard_frames_per_token: int = 0
    eos: int = -1                              # models/t5gemma.py:861-863: eos if eos > 0 else eog
    eog: int = -1
    empty_token: int = -1
    y_sep_token: int = -1
    enc_layer_types: List[str] = field(default_factory=list)
    dec_layer_types: List[str] = field(default_factory=list)
    # engine sizing
    max_slots: int = 1
    max_text_len: int = 512
    max_dec_len: int = 2560
    max_prefill_tokens: int = 2560
    kv_page_tokens: int = 32              # one 32-token attention tile = one page = one TMA box per 64-dim slab (16 also supported)

    def __post_init__(self):
        V = self.audio_vocab
        if self.empty_token < 0: self.empty_token = V
        if self.eog < 0: self.eog = V + 1
        if self.eos < 0: self.eos = V + 3
        if self.y_sep_token < 0: self.y_sep_token = V + 4
        if not self.enc_layer_types:
            self.enc_layer_types = ["sliding_attention" if (i + 1) % 2 else "full_attention"
                                    for i in range(self.n_enc_layers)]
        if not self.dec_layer_types:
            self.dec_layer_types = ["sliding_attention" if (i + 1) % 2 else "full_attention"
                                    for i in range(self.n_dec_layers)]

    @property
    def n_audio_tokens(self) -> int:
        return self.audio_vocab + self.n_special

    @property
    def stop_token(self) -> int:
        return self.eos if self.eos > 0 else self.eog

    @staticmethod
    def from_reference(cfg: Any, **sizing) -> "EngineConfig":
        """cfg: T5GemmaVoiceConfig-like object (attributes or dict) carrying t5_config_dict."""
        t5 = _get(cfg, "t5_config_dict")
        if t5 is None:
            raise ValueError("reference config has no t5_config_dict (backbone geometry)")
        enc, dec = t5["encoder"], t5["decoder"]
        attn_impl = _get(cfg, "attn_implementation", "eager")
        rope = dec.get("rope_parameters") or {}
        avs = _get(cfg, "audio_vocab_size", 65536)
        if isinstance(avs, (list, tuple)):
            avs = avs[0]
        if int(_get(cfg, "n_codebooks", 1)) != 1:
            raise ValueError("XCodec2 inference expects n_codebooks=1.")
        return EngineConfig(
            hidden=dec["hidden_size"], inter=dec["intermediate_size"],
            n_enc_layers=enc["num_hidden_layers"], n_dec_layers=dec["num_hidden_layers"],
            n_heads=dec["num_attention_heads"], n_kv_heads=dec["num_key_value_heads"], head_dim=dec["head_dim"],
            sliding_window=dec["sliding_window"], query_pre_attn_scalar=float(dec["query_pre_attn_scalar"]),
            attn_softcap=(dec.get("attn_logit_softcapping") if attn_impl == "eager" else None),
            rms_eps=float(dec.get("rms_norm_eps", 1e-6)),
            rope_theta=float(rope.get("rope_theta", dec.get("rope_theta", 10000.0))),
            text_vocab=enc["vocab_size"], audio_vocab=int(avs), n_special=int(_get(cfg, "n_special", 5)),
            special_first=int(_get(cfg, "special_first", 0)),
            progress_scale=float(_get(cfg, "progress_scale", 2000.0)), encodec_sr=int(_get(cfg, "encodec_sr", 50)),
            extra_cutoff=float(_get(cfg, "extra_cutoff", 5.0)),
            text_guard_frames_per_token=int(_get(cfg, "text_guard_frames_per_token", 0)),
            eos=int(_get(cfg, "eos", -1)), eog=int(_get(cfg, "eog", -1)),
            empty_token=int(_get(cfg, "empty_token", -1)), y_sep_token=int(_get(cfg, "y_sep_token", -1)),
            enc_layer_types=list(enc.get("layer_types") or []), dec_layer_types=list(dec.get("layer_types") or []),
            **sizing)
