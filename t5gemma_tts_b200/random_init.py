"""Seeded random-init weights of the reference architecture (no checkpoint download is possible).

Distributions follow what the reference actually produces: backbone linears and
text embedding N(0, 0.02) (HF _init_weights), audio_embedding N(0, 1) (nn.Embedding default), predict_layer
Kaiming-uniform U(+-1/sqrt(d)) weights and biases (nn.Linear default).  RMSNorm weights are NOT left at the HF
init value 0 (gain 1+w = 1 everywhere hides a swapped or dropped gain): like a trained checkpoint, every norm tensor
gets its own N(0, 0.3) draw (`norm_std`; 0 restores the HF init).
Tensors are produced one at a time on `device` so a 2b-2b model never needs 21 GB of host memory."""
from __future__ import annotations

import math
from typing import Iterator, Tuple

import torch

from .config import EngineConfig


def iter_random_state_dict(cfg: EngineConfig, seed: int = 0, device="cuda", dtype=torch.bfloat16,
                           norm_std: float = 0.3) -> Iterator[Tuple[str, torch.Tensor]]:
    g = torch.Generator(device=device).manual_seed(seed)
    d, I = cfg.hidden, cfg.inter
    QD, KD = cfg.n_heads * cfg.head_dim, cfg.n_kv_heads * cfg.head_dim
    V = cfg.n_audio_tokens

    def normal(shape, std):
        return (torch.randn(shape, device=device, generator=g, dtype=torch.float32) * std).to(dtype)

    def uniform(shape, bound):
        return ((torch.rand(shape, device=device, generator=g, dtype=torch.float32) * 2 - 1) * bound).to(dtype)

    yield "backbone.model.encoder.embed_tokens.weight", normal((cfg.text_vocab, d), 0.02)
    yield "backbone.model.encoder.norm.weight", normal((d,), norm_std)
    yield "backbone.model.decoder.norm.weight", normal((d,), norm_std)
    for side, n_layers in (("encoder", cfg.n_enc_layers), ("decoder", cfg.n_dec_layers)):
        for l in range(n_layers):
            p = f"backbone.model.{side}.layers.{l}."
            yield p + "self_attn.q_proj.weight", normal((QD, d), 0.02)
            yield p + "self_attn.k_proj.weight", normal((KD, d), 0.02)
            yield p + "self_attn.v_proj.weight", normal((KD, d), 0.02)
            yield p + "self_attn.o_proj.weight", normal((d, QD), 0.02)
            yield p + "mlp.gate_proj.weight", normal((I, d), 0.02)
            yield p + "mlp.up_proj.weight", normal((I, d), 0.02)
            yield p + "mlp.down_proj.weight", normal((d, I), 0.02)
            norms = ["pre_self_attn_layernorm", "post_self_attn_layernorm", "pre_feedforward_layernorm",
                     "post_feedforward_layernorm"]
            if side == "decoder":
                yield p + "cross_attn.q_proj.weight", normal((QD, d), 0.02)
                yield p + "cross_attn.k_proj.weight", normal((KD, d), 0.02)
                yield p + "cross_attn.v_proj.weight", normal((KD, d), 0.02)
                yield p + "cross_attn.o_proj.weight", normal((d, QD), 0.02)
                norms += ["pre_cross_attn_layernorm", "post_cross_attn_layernorm"]
            for n in norms:
                yield p + n + ".weight", normal((d,), norm_std)
    yield "audio_embedding.0.weight", normal((V, d), 1.0)
    b = 1.0 / math.sqrt(d)
    yield "predict_layer.0.0.weight", uniform((d, d), b)
    yield "predict_layer.0.0.bias", uniform((d,), b)
    yield "predict_layer.0.2.weight", uniform((V, d), b)
    yield "predict_layer.0.2.bias", uniform((V,), b)
