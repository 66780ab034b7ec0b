"""Helpers to load the committed golden fixtures (tests/golden) -- test infrastructure only."""
from __future__ import annotations

import json
import os
from types import SimpleNamespace

import numpy as np
import torch

from .t5gemma_voice_oracle import Oracle, OracleConfig

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_model_fixture(name: str):
    """Returns (OracleConfig, state_dict of fp32 torch tensors, meta dict)."""
    z = np.load(os.path.join(GOLDEN, f"model_{name}.npz"), allow_pickle=False)
    meta = json.loads(str(z["__meta__"]))
    sd = {k: torch.from_numpy(z[k]) for k in z.files if k != "__meta__"}
    ns = SimpleNamespace(t5_config_dict=meta["t5_config_dict"], attn_implementation=meta["attn_implementation"],
                         audio_vocab_size=meta["audio_vocab_size"], n_special=meta["n_special"],
                         progress_scale=meta["progress_scale"], encodec_sr=meta["encodec_sr"],
                         extra_cutoff=meta["extra_cutoff"],
                         text_guard_frames_per_token=meta["text_guard_frames_per_token"])
    return OracleConfig.from_reference_config(ns), sd, meta


def load_case(name: str):
    z = np.load(os.path.join(GOLDEN, f"case_{name}.npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def load_oracle(name: str) -> Oracle:
    cfg, sd, _ = load_model_fixture(name)
    return Oracle(cfg, sd)


def load_sampler_cases():
    z = np.load(os.path.join(GOLDEN, "sampler.npz"), allow_pickle=False)
    meta = json.loads(str(z["__meta__"]))
    return [(z[f"logits_{i}"], z[f"keep_{i}"], z[f"probs_{i}"], m) for i, m in enumerate(meta)]
