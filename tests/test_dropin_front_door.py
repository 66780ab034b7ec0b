"""Drop-in proof against the reference's REAL caller: `inference_tts_utils.inference_one_sample`
(/root/reference/inference_tts_utils.py:141-379) is loaded unmodified (from /root/reference in the build container, from the
git-ignored copy baseline/_ref/ on the GPU box) and driven with stub text / audio tokenizers.

  * CPU: a recording `model` captures what the front door hands to `model.inference_tts`; `request_glue.build_request`
    must assemble exactly the same request, and `strip_sep_and_eos` must equal the front door's nested
    `_strip_sep_and_eos` on what comes back.
  * GPU: the engine is passed as `model` (with `eng.args` as `model_args`) and must behave like the reference model
    passed through the same call: identical prompt handling in `concat_frames`, same utterance length (time-budget
    rule), same greedy tokens for the first steps, tensors on the caller's device.
"""
import importlib.util
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import fixtures, ref_loader
from t5gemma_tts_b200.request_glue import build_request, strip_sep_and_eos

FRONT_DOOR = os.path.join(ref_loader.REFERENCE_ROOT, "inference_tts_utils.py")
needs_reference = pytest.mark.skipif(not os.path.isfile(FRONT_DOOR), reason="reference front door not available")


class StubTextTokenizer:
    """text -> ids: one id per whitespace-separated integer (the SentencePiece model is not part of the hot path)."""

    def encode(self, text, add_special_tokens=False):
        assert add_special_tokens is False
        return [int(t) for t in text.split()]


class StubAudioTokenizer:
    """XCodec2 stand-in: `encode` side is replaced by the stub `tokenize_audio` below, `decode` returns the frames."""

    def __init__(self, prompt_codes, device="cpu"):
        self.prompt_codes = torch.as_tensor(prompt_codes, dtype=torch.long).view(1, 1, -1)
        self.device = torch.device(device)
        self.decoded = []

    def decode(self, frames):
        self.decoded.append(frames.detach().cpu().clone())
        return frames.detach().float().cpu()           # "waveform"


def load_front_door():
    """Imports the reference file as it is, with stand-ins for the modules that pull in the codec / ASR stacks
    (data.tokenizer -> xcodec2, duration_estimator -> language detection, torchaudio when absent)."""
    if "ref_inference_tts_utils" in sys.modules:
        return sys.modules["ref_inference_tts_utils"]
    tok = types.ModuleType("data.tokenizer")
    tok.AudioTokenizer = StubAudioTokenizer
    tok.tokenize_audio = lambda tokenizer, audio_fn, offset=-1, num_frames=-1: tokenizer.prompt_codes.clone()
    data = types.ModuleType("data")
    data.tokenizer = tok
    dur = types.ModuleType("duration_estimator")
    dur.detect_language = lambda text: "en"
    saved = {k: sys.modules.get(k) for k in ("data", "data.tokenizer", "duration_estimator", "torchaudio")}
    sys.modules.update({"data": data, "data.tokenizer": tok, "duration_estimator": dur})
    try:
        import torchaudio  # noqa: F401
    except Exception:
        sys.modules["torchaudio"] = types.ModuleType("torchaudio")
    try:
        spec = importlib.util.spec_from_file_location("ref_inference_tts_utils", FRONT_DOOR)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    sys.modules["ref_inference_tts_utils"] = mod
    return mod


DECODE = dict(top_k=1, top_p=1.0, min_p=0.0, temperature=1.0, stop_repetition=-1, codec_sr=50, silence_tokens="[]",
              sample_batch_size=1)


class RecordingModel:
    def __init__(self, ret_gen):
        self.calls, self.ret_gen = [], ret_gen

    def inference_tts(self, x, x_lens, y, **kw):
        self.calls.append((x.clone(), x_lens.clone(), y.clone(), dict(kw)))
        gen = torch.as_tensor(self.ret_gen, dtype=torch.long).view(1, 1, -1)
        return torch.cat([y.transpose(1, 2), gen], dim=2), gen


@needs_reference
@pytest.mark.parametrize("with_prompt,prefix", [(True, "7 8"), (True, None), (False, None)])
def test_front_door_request_equals_build_request(with_prompt, prefix):
    fd = load_front_door()
    args = SimpleNamespace(n_codebooks=1, empty_token=100, y_sep_token=104, x_sep_token=255999, eos=103, eog=101,
                           add_eos_to_text=0, add_bos_to_text=0, parallel_pattern=0, encodec_sr=50)
    prompt = np.array([3, 1, 4, 1, 5, 9, 2, 6])
    audio_tok = StubAudioTokenizer(prompt)
    ret_gen = [11, 104, 12, 13, 103]                                    # a y_sep and the eos inside the generated part
    model = RecordingModel(ret_gen)
    out = fd.inference_one_sample(model, args, StubTextTokenizer(), audio_tok, "ref.wav" if with_prompt else None,
                                  "21 22 23", "en", "cpu", dict(DECODE, top_k=30, top_p=0.9, temperature=0.8, stop_repetition=3,
                                                                silence_tokens="[5, 6]"),
                                  prompt_end_frame=-1, target_generation_length=1.5, prefix_transcript=prefix, quiet=True,
                                  return_frames=True)
    (x, x_lens, y, kw), = model.calls
    req = build_request(args, [21, 22, 23], 1.5, prompt_codes=prompt if with_prompt else None,
                        prefix_text_ids=[7, 8] if prefix else None, top_k=30, top_p=0.9, temperature=0.8, stop_repetition=3,
                        silence_tokens=[5, 6])
    assert x[0].tolist() == req.text_ids.tolist() and int(x_lens[0]) == len(req.text_ids)
    assert y.shape[0] == 1 and y.shape[2] == 1 and y[0, :, 0].tolist() == req.prompt_ids.tolist()
    assert int(kw["tgt_y_lens"][0]) == req.target_total and kw["prompt_frames"] == req.prompt_frames
    assert (kw["top_k"], kw["top_p"], kw["min_p"], kw["temperature"], kw["stop_repetition"]) == \
           (req.top_k, req.top_p, req.min_p, req.temperature, req.stop_repetition)
    assert list(kw["silence_tokens"]) == list(req.silence_tokens)
    # what the front door hands to the codec == strip_sep_and_eos of what the model returned
    concat_sample, gen_sample, concat_frames, gen_frames = out
    want_gen = strip_sep_and_eos(np.array(ret_gen).reshape(1, 1, -1), 104, 103)
    want_concat = strip_sep_and_eos(np.concatenate([req.prompt_ids, ret_gen]).reshape(1, 1, -1), 104, 103)
    assert np.array_equal(gen_frames.numpy(), want_gen) and np.array_equal(concat_frames.numpy(), want_concat)
    assert np.array_equal(audio_tok.decoded[-1].numpy(), want_gen)


@needs_reference
@pytest.mark.gpu
@pytest.mark.parametrize("with_prompt", [True, False])
def test_engine_is_a_drop_in_for_inference_one_sample(with_prompt):
    from tests.gpu_util import engine_for
    fd = load_front_door()
    name = "tinyA_eager"
    _, sd, meta = fixtures.load_model_fixture(name)
    # the reference model holds the bf16-rounded weights the engine stores, so only activation rounding differs
    sd16 = {k: (v.to(torch.bfloat16).float() if v.dim() >= 2 else v) for k, v in sd.items()}
    ref = ref_loader.build_reference_model_from_tensors(
        meta["t5_config_dict"], sd16.items(), audio_vocab=meta["audio_vocab_size"], n_special=meta["n_special"],
        attn_implementation=meta["attn_implementation"], progress_scale=meta["progress_scale"],
        extra_cutoff=meta["extra_cutoff"], text_guard_frames_per_token=meta["text_guard_frames_per_token"])
    eng = engine_for(name)
    c = fixtures.load_case("tinyA_eager_prompt")
    text = " ".join(str(int(t)) for t in c["x"][0])
    prompt = c["y"][0, :, 0]

    def run(model, model_args, device):
        tok = StubAudioTokenizer(prompt, device)
        torch.manual_seed(0)
        out = fd.inference_one_sample(model, model_args, StubTextTokenizer(), tok, "ref.wav" if with_prompt else None, text,
                                      "en", device, DECODE, prompt_end_frame=-1, target_generation_length=0.2, quiet=True,
                                      return_frames=True)
        return out, tok

    (r_cs, r_gs, r_cf, r_gf), r_tok = run(ref, ref.args if hasattr(ref, "args") else ref.config, "cpu")
    (e_cs, e_gs, e_cf, e_gf), e_tok = run(eng, eng.args, "cuda")
    # same shapes / dtypes / sep+eos handling, prompt part identical, same utterance length (time budget)
    assert e_cf.dtype == r_cf.dtype == torch.long and e_gf.shape == r_gf.shape and e_cf.shape == r_cf.shape
    n_prompt = len(prompt) if with_prompt else 0
    assert torch.equal(e_cf[0, 0, :n_prompt], r_cf[0, 0, :n_prompt]) and e_cf.shape[2] == n_prompt + e_gf.shape[2]
    y_sep, eos = eng.cfg.y_sep_token, eng.cfg.stop_token
    assert not bool(((e_cf == y_sep) | (e_cf == eos)).any())
    # free-running greedy: the first tokens agree (bf16 activations may flip a near-tied argmax later on)
    n = e_gf.shape[2]
    first_div = next((i for i in range(n) if int(e_gf[0, 0, i]) != int(r_gf[0, 0, i])), n)
    assert first_div >= 8, (first_div, e_gf[0, 0, :12], r_gf[0, 0, :12])
    # the codec stub received the stripped frames in both runs
    assert torch.equal(e_tok.decoded[-1], e_gf) and torch.equal(r_tok.decoded[-1], r_gf)
    assert e_gs.shape == r_gs.shape and e_cs.shape == r_cs.shape


@needs_reference
@pytest.mark.gpu
def test_batched_front_door_equals_one_sample_per_utterance():
    """request_glue.inference_batch (bs > 1 reachable from the CLI / UI code) must hand the codec, for every utterance, exactly
    what the reference's inference_one_sample hands it when the same engine is driven one utterance at a time (the bs <= 4
    decode path is row-independent bit for bit)."""
    from tests.gpu_util import engine_for
    from t5gemma_tts_b200.request_glue import inference_batch
    fd = load_front_door()
    eng = engine_for("tinyA_eager", max_slots=3, max_prefill_tokens=1024)
    c = fixtures.load_case("tinyA_eager_prompt")
    text = " ".join(str(int(t)) for t in c["x"][0])
    prompt = c["y"][0, :, 0]
    items = [dict(audio_fn="ref.wav", target_text=text, target_generation_length=0.2),
             dict(audio_fn=None, target_text="5 6 7 8 9", target_generation_length=0.3),
             dict(audio_fn="ref.wav", target_text="11 12 13", target_generation_length=0.1, prefix_transcript="3 4")]
    tok = StubAudioTokenizer(prompt, "cuda")
    args = SimpleNamespace(**vars(eng.args))
    args.x_sep_token = 7                      # the production id (255999) is outside the tiny model's text vocabulary
    batch = inference_batch(eng, args, StubTextTokenizer(), tok, items, DECODE,
                            tokenize_audio_fn=lambda t, fn, offset=-1, num_frames=-1: t.prompt_codes.clone(), return_frames=True)
    for it, (b_cs, b_gs, b_cf, b_gf) in zip(items, batch):
        tok1 = StubAudioTokenizer(prompt, "cuda")
        _, _, o_cf, o_gf = fd.inference_one_sample(eng, args, StubTextTokenizer(), tok1, it["audio_fn"], it["target_text"], "en",
                                                   "cuda", DECODE, prompt_end_frame=-1,
                                                   target_generation_length=it["target_generation_length"],
                                                   prefix_transcript=it.get("prefix_transcript"), quiet=True, return_frames=True)
        assert torch.equal(b_gf, o_gf) and torch.equal(b_cf, o_cf)
        assert b_gs.shape == o_gf.shape
