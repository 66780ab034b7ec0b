"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Imports /root/reference through oracle/ref_loader.py, builds seeded tiny random-init models
(SURVEY.md section 8d "tiny config"), and records:
  * model_<name>.npz : the reference state_dict (fp32) + config scalars
  * case_<name>.npz  : inputs, reference encoder states, teacher-forced logits (uncached decoder
                       pass + predict_layer), greedy inference_tts tokens and the per-step logits
                       seen by predict_layer during that run
  * sampler.npz      : logits rows + kwargs -> finite mask of the reference top_k_top_p_filtering
The fixtures pin oracle/ (tests/test_oracle_golden.py) and, on the GPU, the CUDA engine.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def randomize_norm_gains(model, seed):
    """HF init leaves every RMSNorm weight at 0 (gain 1+w = 1), which hides a swapped or dropped gain.  Real
    checkpoints carry distinct non-zero gains, so each norm tensor gets its own seeded N(0, 0.3) draw (the six norms
    of a PMDecoderLayer, models/t5gemma.py:205-243, the four of an encoder layer and the two final norms)."""
    g = torch.Generator().manual_seed(seed)
    n = 0
    with torch.no_grad():
        for k, p in model.named_parameters():
            if k.startswith(("encoder_module.", "decoder_module.")):
                continue
            if k.endswith("layernorm.weight") or k.endswith(".norm.weight"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
                n += 1
    return n


def save_model(name, model, t5d, attn_impl):
    sd = {k: v.detach().float().numpy() for k, v in model.state_dict().items()
          if not (k.startswith("encoder_module.") or k.startswith("decoder_module."))}
    meta = dict(t5_config_dict=t5d, attn_implementation=attn_impl,
                audio_vocab_size=int(model.config.audio_vocab_size), n_special=int(model.config.n_special),
                progress_scale=float(model.config.progress_scale), encodec_sr=float(model.config.encodec_sr),
                extra_cutoff=float(model.config.extra_cutoff),
                text_guard_frames_per_token=int(model.config.text_guard_frames_per_token))
    np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), __meta__=json.dumps(meta), **sd)


def run_case(model, x, y, tgt, prompt_frames, top_k=1):
    cfg = model.config
    S = x.shape[1]
    x_lens = torch.tensor([S])
    step_logits = []
    hook = model.predict_layer[0].register_forward_hook(
        lambda m, i, o: step_logits.append(o.detach().reshape(-1, o.shape[-1])[-1].clone()))
    res, gen = model.inference_tts(x, x_lens, y, None if tgt is None else torch.tensor([tgt]), top_k=top_k, top_p=1.0,
                                   temperature=1.0, prompt_frames=prompt_frames)
    hook.remove()
    step_logits = torch.stack(step_logits)
    # encoder states + teacher-forced logits along BOS ++ prompt ++ generated[:-1]
    with torch.inference_mode():
        enc_pos = model._build_position_ids(x_lens, S, x.device)
        mem = model.encoder_module(input_ids=x, attention_mask=torch.ones_like(x),
                                   position_ids=enc_pos).last_hidden_state
        dec_ids = torch.cat([torch.tensor([cfg.empty_token]), y[0, :, 0], gen[0, 0, :-1]])
        if tgt is not None:
            est_total = max(tgt + 1, y.shape[1] + 1)
        else:                                   # models/t5gemma.py:925-933 (progress_lookahead_secs = 2.0)
            est_total = max(int(y.shape[1] + 1 + int(cfg.encodec_sr) * 2.0), y.shape[1] + 1)
        T = dec_ids.shape[0]
        # positions exactly as the generate loop produced them: prefill formula for the prompt part,
        # python-float step formula afterwards (models/t5gemma.py:945-950,1086-1099)
        npre = y.shape[1] + 1
        pos = torch.zeros(T, dtype=torch.float32)
        pos[:npre] = torch.arange(npre, dtype=torch.float32) / max(1, est_total - 1) * model.progress_scale
        for t in range(npre, T):
            pos[t] = min(float(t) / max(1, est_total - 1) * model.progress_scale, model.progress_scale)
        pos = pos[None]
        emb = model.audio_embedding[0](dec_ids[None])
        out = model.decoder_module(inputs_embeds=emb, attention_mask=torch.ones(1, T, dtype=torch.long),
                                   encoder_hidden_states=mem, encoder_attention_mask=torch.ones_like(x),
                                   use_cache=False, position_ids=pos, pm_decoder_position_ids=pos,
                                   pm_encoder_position_ids=enc_pos)
        tf_logits = model.predict_layer[0](out.last_hidden_state)[0]
    return dict(x=x.numpy(), y=y.numpy(), tgt=np.int64(-1 if tgt is None else tgt), prompt_frames=np.int64(prompt_frames),
                memory=mem[0].numpy(), dec_ids=dec_ids.numpy(), dec_pos=pos[0].numpy(),
                tf_logits=tf_logits.numpy(), gen=gen.numpy(), res=res.numpy(),
                step_logits=step_logits.numpy(), est_total=np.int64(est_total))


def silence_case(model):
    """Silence-repetition penalty (models/t5gemma.py:999-1011,1050-1054) exercised through the reference's own
    sample_helper: a forward hook replaces the head output by crafted logits so that greedy decoding repeats token 7."""
    V = model.predict_layer[0][2].out_features
    base = (torch.randn(V, generator=torch.Generator().manual_seed(99)) * 0.1)
    base[7], base[9], base[11] = 5.0, 3.0, -4.0
    hook = model.predict_layer[0].register_forward_hook(lambda m, i, o: base.clone().reshape(1, 1, V).expand(o.shape[0], o.shape[1], V).clone())
    x = torch.randint(2, 500, (1, 9), generator=torch.Generator().manual_seed(5))
    y = torch.zeros((1, 0, 1), dtype=torch.long)
    res, gen = model.inference_tts(x, torch.tensor([9]), y, torch.tensor([20]), top_k=1, top_p=1.0, temperature=1.0,
                                   prompt_frames=0, stop_repetition=2, silence_tokens=[7, 11])
    hook.remove()
    return dict(x=x.numpy(), base_logits=base.numpy(), gen=gen.numpy(), tgt=np.int64(20), stop_repetition=np.int64(2),
                silence_tokens=np.array([7, 11], dtype=np.int64))


def glue_cases():
    """_strip_sep_and_eos (inference_tts_utils.py:323-354, a nested function) extracted by AST and run on random frames."""
    import ast
    from typing import Optional
    src = open(os.path.join(ref_loader.REFERENCE_ROOT, "inference_tts_utils.py")).read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "_strip_sep_and_eos")
    ns = {"torch": torch, "Optional": Optional}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref_strip", "exec"), ns)
    rng = np.random.default_rng(3)
    out = {}
    for i in range(6):
        T = int(rng.integers(1, 40))
        fr = rng.integers(95, 105, (1, 1, T))
        fr[0, 0, -1] = 103
        if i % 2:
            fr[0, 0, T // 2] = 104
        out[f"in_{i}"] = fr
        out[f"out_{i}"] = ns["_strip_sep_and_eos"](torch.from_numpy(fr), 104, 103).numpy()
    np.savez_compressed(os.path.join(OUT, "glue_strip.npz"), **out)


def sampler_cases():
    U = ref_loader.load_reference_sampling()
    rows = []
    # SURVEY.md Appendix A known-answer cases
    L = [2.0, 1.0, 1.0, 0.0, -1.0, 3.0]
    ka = [dict(top_k=2, top_p=1.0), dict(top_k=3, top_p=1.0), dict(top_k=0, top_p=0.5),
          dict(top_k=0, top_p=0.7), dict(top_k=4, top_p=0.9), dict(top_k=0, top_p=0.9),
          dict(top_k=-100, top_p=1.0), dict(top_k=2, top_p=0.5, min_p=0.2), dict(top_k=1, top_p=0.9)]
    for kw in ka:
        rows.append((np.array(L, dtype=np.float32), kw, 1.0))
    g = torch.Generator().manual_seed(7)
    for V, scale in [(105, 1.0), (105, 4.0), (1000, 2.0), (4099, 3.0)]:
        for kw in [dict(top_k=30, top_p=0.9), dict(top_k=30, top_p=1.0), dict(top_k=5, top_p=0.5),
                   dict(top_k=1, top_p=1.0), dict(top_k=50, top_p=0.95), dict(top_k=-100, top_p=1.0),
                   dict(top_k=0, top_p=0.8), dict(top_k=40, top_p=0.9, min_p=0.05), dict(top_k=100, top_p=0.3)]:
            for T in (1.0, 0.8):
                lg = (torch.randn(V, generator=g) * scale).numpy().astype(np.float32)
                rows.append((lg, kw, T))
    out = {}
    meta = []
    for i, (lg, kw, T) in enumerate(rows):
        t = torch.from_numpy(lg.copy())
        if T != 1.0:
            t = t / T
        f = U.top_k_top_p_filtering(t.clone(), top_k=kw.get("top_k", 0), top_p=kw.get("top_p", 1.0),
                                    min_p=kw.get("min_p", 0.0))
        out[f"logits_{i}"] = lg
        out[f"keep_{i}"] = torch.isfinite(f).numpy()
        out[f"probs_{i}"] = torch.softmax(f, -1).numpy()
        meta.append(dict(kw, temperature=T))
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), __meta__=json.dumps(meta), **out)
    print("sampler cases:", len(rows))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    gen = torch.Generator().manual_seed(1234)
    # tiny A: 3 layers (sliding/full/sliding), window 8, eager softcap
    t5a = ref_loader.tiny_t5_config_dict(layers=3)
    for name, impl in (("tinyA_eager", "eager"), ("tinyA_sdpa", "sdpa")):
        model = ref_loader.build_reference_model(t5a, audio_vocab=100, attn_implementation=impl, seed=0)
        print(name, "norm tensors randomised:", randomize_norm_gains(model, seed=100))
        save_model(name, model, t5a, impl)
        x = torch.randint(2, 500, (1, 12), generator=gen)
        y = torch.randint(0, 100, (1, 5, 1), generator=gen)
        c = run_case(model, x, y, tgt=5 + 10, prompt_frames=5)
        np.savez_compressed(os.path.join(OUT, f"case_{name}_prompt.npz"), **c)
        print(name, "prompt gen len", c["gen"].shape)
        x = torch.randint(2, 500, (1, 20), generator=gen)
        y = torch.zeros((1, 0, 1), dtype=torch.long)
        c = run_case(model, x, y, tgt=12, prompt_frames=0)
        np.savez_compressed(os.path.join(OUT, f"case_{name}_noprompt.npz"), **c)
        print(name, "noprompt gen len", c["gen"].shape)
        if name == "tinyA_eager":
            sc = silence_case(model)
            np.savez_compressed(os.path.join(OUT, "case_tinyA_eager_silence.npz"), **sc)
            print("silence gen head", sc["gen"][0, 0, :24])
    # tgt_y_lens=None (models/t5gemma.py:896-933): no time budget, est_total from the 2 s lookahead; the text guard
    # (text_guard_frames_per_token=3) is what ends the utterance.  Same weights as tinyA_eager; a top_k LIST as well
    # (models/t5gemma.py:991-994) -- greedy at every step but the list form goes through the schedule path.
    model = ref_loader.build_reference_model(t5a, audio_vocab=100, attn_implementation="eager", seed=0,
                                             text_guard_frames_per_token=3)
    randomize_norm_gains(model, seed=100)
    x = torch.randint(2, 500, (1, 14), generator=gen)
    y = torch.randint(0, 100, (1, 4, 1), generator=gen)
    c = run_case(model, x, y, tgt=None, prompt_frames=4, top_k=[1, 1, 1])
    c["text_guard_frames_per_token"] = np.int64(3)
    np.savez_compressed(os.path.join(OUT, "case_tinyA_eager_notarget.npz"), **c)
    print("notarget gen len", c["gen"].shape)
    # tiny B: wider heads (head_dim 32, 4 layers, GQA 4/2), softcap strongly binding (cap 5)
    t5b = ref_loader.tiny_t5_config_dict(hidden=128, inter=256, layers=4, heads=4, kv_heads=2, head_dim=32,
                                         window=16, qpas=32, softcap=5.0)
    model = ref_loader.build_reference_model(t5b, audio_vocab=200, attn_implementation="eager", seed=1)
    print("tinyB norm tensors randomised:", randomize_norm_gains(model, seed=101))
    save_model("tinyB_eager", model, t5b, "eager")
    x = torch.randint(2, 500, (1, 33), generator=gen)
    y = torch.randint(0, 200, (1, 21, 1), generator=gen)
    c = run_case(model, x, y, tgt=21 + 30, prompt_frames=21)
    np.savez_compressed(os.path.join(OUT, "case_tinyB_eager_prompt.npz"), **c)
    print("tinyB gen len", c["gen"].shape)
    sampler_cases()
    glue_cases()


if __name__ == "__main__":
    main()
