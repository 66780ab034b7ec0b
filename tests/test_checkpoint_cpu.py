"""Checkpoint readers (HF safetensors dir / .pth bundle) on synthetic checkpoints built from the golden fixture."""
import argparse
import json
import os

import pytest
import torch

from oracle import fixtures
from t5gemma_tts_b200 import checkpoint as ck
from t5gemma_tts_b200.config import EngineConfig


def _write_hf_dir(tmp_path, shards=1):
    from safetensors.torch import save_file
    _, sd, meta = fixtures.load_model_fixture("tinyA_eager")
    d = tmp_path / "hf"
    d.mkdir()
    cfg = dict(meta, model_type="t5gemma_voice", eos=103, eog=101, empty_token=100, y_sep_token=104)
    (d / "config.json").write_text(json.dumps(cfg))
    sd = {k: v.contiguous() for k, v in sd.items()}
    if shards == 1:
        save_file(sd, str(d / "model.safetensors"))
    else:
        keys = sorted(sd)
        half = len(keys) // 2
        wm = {}
        for i, ks in enumerate((keys[:half], keys[half:])):
            fn = f"model-{i+1:05d}-of-00002.safetensors"
            save_file({k: sd[k] for k in ks}, str(d / fn))
            wm.update({k: fn for k in ks})
        (d / "model.safetensors.index.json").write_text(json.dumps({"weight_map": wm}))
    return str(d), sd


@pytest.mark.parametrize("shards", [1, 2])
def test_hf_dir_roundtrip(tmp_path, shards):
    d, sd = _write_hf_dir(tmp_path, shards)
    cfg = ck.load_hf_config(d)
    ec = EngineConfig.from_reference(cfg)
    assert ec.hidden == 64 and ec.stop_token == 103 and ec.attn_softcap == 50.0
    got = dict(ck.iter_hf_tensors(d))
    assert set(got) == set(sd)
    for k in sd:
        assert torch.equal(got[k], sd[k])


def test_hf_dir_errors(tmp_path):
    d = tmp_path / "x"
    d.mkdir()
    (d / "config.json").write_text(json.dumps({"audio_vocab_size": 100}))
    with pytest.raises(ValueError):
        ck.load_hf_config(str(d))
    (d / "config.json").write_text(json.dumps({"t5_config_dict": {}}))
    with pytest.raises(FileNotFoundError):
        list(ck.iter_hf_tensors(str(d)))


def test_pth_bundle(tmp_path):
    _, sd, meta = fixtures.load_model_fixture("tinyA_eager")
    args = argparse.Namespace(audio_vocab_size=100, n_special=5, n_codebooks=1, encodec_sr=50, progress_scale=2000.0,
                              extra_cutoff=5, attn_implementation="eager", eos=103, eog=101, empty_token=100,
                              y_sep_token=104)
    p = tmp_path / "bundle.pth"
    torch.save({"model": sd, "args": args}, str(p))
    cfg, sd2 = ck.load_pth_bundle(str(p), t5_config_dict=meta["t5_config_dict"])
    ec = EngineConfig.from_reference(cfg)
    assert ec.n_audio_tokens == 105 and ec.stop_token == 103 and ec.n_dec_layers == 3
    assert set(sd2) == set(sd)


def test_pth_bundle_names_missing_backbone_geometry(tmp_path):
    """A bundle trained on a non-default backbone carries only its hub name (models/t5gemma.py builds the backbone from
    args.t5gemma_model_name): without t5_config_dict the loader must say so instead of failing later on tensor shapes."""
    _, sd, meta = fixtures.load_model_fixture("tinyA_eager")
    args = argparse.Namespace(audio_vocab_size=100, t5gemma_model_name="google/t5gemma-b-b-ul2")
    p = tmp_path / "bundle_bb.pth"
    torch.save({"model": sd, "args": args}, str(p))
    with pytest.raises(ValueError, match="t5gemma-b-b-ul2"):
        ck.load_pth_bundle(str(p))
    cfg, _ = ck.load_pth_bundle(str(p), t5_config_dict=meta["t5_config_dict"])
    assert EngineConfig.from_reference(cfg).hidden == 64
    # the 2b-2b names fall back to transformers' default geometry
    args.t5gemma_model_name = "google/t5gemma-2b-2b-ul2"
    torch.save({"model": sd, "args": args}, str(p))
    cfg, _ = ck.load_pth_bundle(str(p))
    assert EngineConfig.from_reference(cfg).hidden == 2304


def test_pth_bundle_with_unmerged_lora_adapters(tmp_path):
    """`use_lora=1` bundles keep PEFT's wrapped names (base_layer / lora_A / lora_B under backbone.base_model.model.);
    the loader returns the plain reference keys with W + (alpha / r) * B @ A, i.e. what merge_and_unload() produces
    (scripts/export_t5gemma_voice_hf_lora.py)."""
    _, sd, meta = fixtures.load_model_fixture("tinyA_eager")
    g = torch.Generator().manual_seed(0)
    r, alpha = 4, 8
    lora_sd, want = {}, {}
    targets = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")
    for k, v in sd.items():
        if k.startswith("backbone.") and k.endswith(".weight") and k.split(".")[-2] in targets:
            pre = "backbone.base_model.model." + k[len("backbone."):-len(".weight")]
            A, B = torch.randn(r, v.shape[1], generator=g) * 0.1, torch.randn(v.shape[0], r, generator=g) * 0.1
            lora_sd[pre + ".base_layer.weight"] = v
            lora_sd[pre + ".lora_A.default.weight"] = A
            lora_sd[pre + ".lora_B.default.weight"] = B
            want[k] = v + (alpha / r) * (B @ A)
        elif k.startswith("backbone."):
            lora_sd["backbone.base_model.model." + k[len("backbone."):]] = v
            want[k] = v
        else:
            lora_sd[k] = v
            want[k] = v
    args = argparse.Namespace(audio_vocab_size=100, use_lora=1, lora_r=r, lora_alpha=alpha)
    p = tmp_path / "bundle_lora.pth"
    torch.save({"model": lora_sd, "args": args}, str(p))
    cfg, got = ck.load_pth_bundle(str(p), t5_config_dict=meta["t5_config_dict"])
    assert set(got) == set(want)
    for k in want:
        assert torch.allclose(got[k], want[k], atol=1e-6), k
    n_merged = sum(1 for k in want if not torch.equal(want[k], sd[k]))
    assert n_merged == 7 * 3 + 7 * 3 + 4 * 3       # 7 projections per layer side (3+3 layers) + 4 cross projections per decoder layer
