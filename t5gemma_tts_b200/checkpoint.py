"""Checkpoint readers for the two on-disk formats the reference produces.

* HF directory written by scripts/export_t5gemma_voice_hf.py:152-175: `config.json` (T5GemmaVoiceConfig incl.
  `t5_config_dict`) + `model.safetensors` or sharded `model-0000x-of-0000y.safetensors` with
  `model.safetensors.index.json`; keys `backbone.model.{encoder,decoder}.*`, `audio_embedding.0.weight`,
  `predict_layer.0.{0,2}.{weight,bias}` (the `encoder_module.*`/`decoder_module.*` aliases are dropped at save time,
  hf_export/modeling_t5gemma_voice.py:497-506).
* `.pth` bundle `{model, args, ...}` loaded by inference_commandline.py:121-156 (`torch.load(weights_only=True)`); bundles
  saved with `use_lora=1` carry PEFT adapter keys, which are merged into the base weights here (`merge_lora_state_dict`),
  like scripts/export_t5gemma_voice_hf_lora.py does before it writes the HF directory.

Tensors are yielded one at a time so a 2b-2b checkpoint streams to the GPU without a full host copy."""
from __future__ import annotations

import glob
import json
import os
from types import SimpleNamespace
from typing import Any, Dict, Iterator, Tuple

import torch


def load_hf_config(model_dir: str) -> SimpleNamespace:
    with open(os.path.join(model_dir, "config.json"), "r", encoding="utf-8") as f:
        d = json.load(f)
    if d.get("t5_config_dict") is None:
        raise ValueError(f"{model_dir}/config.json has no t5_config_dict (backbone geometry); cannot build offline")
    d.setdefault("n_codebooks", 1)
    return SimpleNamespace(**d)


def iter_hf_tensors(model_dir: str, device: str = "cpu") -> Iterator[Tuple[str, torch.Tensor]]:
    from safetensors import safe_open
    index = os.path.join(model_dir, "model.safetensors.index.json")
    if os.path.exists(index):
        with open(index) as f:
            files = sorted(set(json.load(f)["weight_map"].values()))
    else:
        files = sorted(os.path.basename(p) for p in glob.glob(os.path.join(model_dir, "*.safetensors")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {model_dir}")
    for fn in files:
        with safe_open(os.path.join(model_dir, fn), framework="pt", device=device) as f:
            for k in f.keys():
                yield k, f.get_tensor(k)


def merge_lora_state_dict(sd: Dict[str, torch.Tensor], lora_alpha: float, lora_r: int) -> Dict[str, torch.Tensor]:
    """Un-merged LoRA bundle (`use_lora=1`, models/t5gemma.py:552-600: `self.backbone = get_peft_model(self.backbone, cfg)`)
    -> plain reference keys.  PEFT names a wrapped Linear `<prefix>.base_layer.weight` with adapters
    `<prefix>.lora_A.<adapter>.weight [r, in]` / `<prefix>.lora_B.<adapter>.weight [out, r]` and inserts `base_model.model.`
    after the wrapped module's attribute (`backbone.`); the merged weight is `W + (lora_alpha / r) * B @ A`, which is what
    scripts/export_t5gemma_voice_hf_lora.py obtains with `merge_and_unload()`.  Keys without adapters pass through."""
    scale = float(lora_alpha) / float(lora_r)
    out: Dict[str, torch.Tensor] = {}

    def plain(k: str) -> str:
        return k.replace("backbone.base_model.model.", "backbone.", 1)

    for k, v in sd.items():
        if ".lora_A." in k or ".lora_B." in k or ".lora_dropout" in k or ".lora_embedding_" in k:
            continue
        if k.endswith(".base_layer.weight") or k.endswith(".base_layer.bias"):
            prefix, leaf = k.rsplit(".base_layer.", 1)
            w = v
            if leaf == "weight":
                a_keys = [x for x in sd if x.startswith(prefix + ".lora_A.") and x.endswith(".weight")]
                for ak in a_keys:
                    bk = ak.replace(".lora_A.", ".lora_B.")
                    if bk not in sd:
                        raise KeyError(f"LoRA bundle: {ak} has no matching {bk}")
                    w = w.float() + scale * (sd[bk].float() @ sd[ak].float())
            out[plain(prefix) + "." + leaf] = w.to(v.dtype) if torch.is_floating_point(v) else w
        else:
            out[plain(k)] = v
    return out


def load_pth_bundle(path: str, t5_config_dict: Dict[str, Any] | None = None):
    """Returns (config-like namespace, state_dict) from a training bundle.  The bundle's `args` carry the TTS
    constants but not the backbone geometry (the trainer loads it from the hub by name), so `t5_config_dict`
    defaults to transformers' T5GemmaConfig() = 2b-2b (HF:configuration_t5gemma.py:68-99)."""
    import argparse
    with torch.serialization.safe_globals([argparse.Namespace]):
        ckpt = torch.load(path, map_location="cpu", weights_only=True)
    sd = ckpt["model"] if "model" in ckpt else ckpt
    args = ckpt.get("args", None)
    a = vars(args) if args is not None and not isinstance(args, dict) else dict(args or {})
    if any(".lora_A." in k for k in sd):          # un-merged LoRA bundle (inference_commandline.py:153 loads it non-strictly)
        sd = merge_lora_state_dict(sd, a.get("lora_alpha", 32), a.get("lora_r", 16))
    if t5_config_dict is None:
        # the reference builds the backbone from args.t5gemma_model_name (hub name); offline, only the 2b-2b default
        # geometry is known -- any other backbone must come from the local HF cache or from the caller
        name = str(a.get("t5gemma_model_name") or "")
        if name and "2b-2b" not in name:
            try:
                from transformers import AutoConfig
                t5_config_dict = AutoConfig.from_pretrained(name, local_files_only=True).to_dict()
            except Exception as ex:
                raise ValueError(f"{path}: the bundle was trained on backbone '{name}', whose geometry is neither stored in the "
                                 "bundle nor found in the local Hugging Face cache; pass t5_config_dict=<its config dict>") from ex
        else:
            from transformers.models.t5gemma import T5GemmaConfig
            t5_config_dict = T5GemmaConfig().to_dict()
    V = int(a.get("audio_vocab_size", 65536) if not isinstance(a.get("audio_vocab_size"), (list, tuple)) else a["audio_vocab_size"][0])
    cfg = dict(t5_config_dict=t5_config_dict, attn_implementation=a.get("attn_implementation", "eager"),
               audio_vocab_size=V, n_special=int(a.get("n_special", 5)), n_codebooks=int(a.get("n_codebooks", 1)),
               special_first=int(a.get("special_first", 0)), encodec_sr=a.get("encodec_sr", 50),
               progress_scale=a.get("progress_scale", 2000.0), extra_cutoff=a.get("extra_cutoff", 5.0),
               text_guard_frames_per_token=int(a.get("text_guard_frames_per_token", 0)),
               empty_token=a.get("empty_token", V), eog=a.get("eog", V + 1), eos=a.get("eos", V + 3),
               y_sep_token=a.get("y_sep_token", V + 4), x_sep_token=a.get("x_sep_token", 255999))
    return SimpleNamespace(**cfg), sd
