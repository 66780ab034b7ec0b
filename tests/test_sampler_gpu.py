"""Sampler S: the CUDA kernel must be BIT-EXACT against oracle/sampler_oracle.py given identical fp32
logits and identical uniform draws (north_star), at tiny and at the full XCodec2 vocabulary (65541)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import fixtures, sampler_oracle
from tests.gpu_util import engine_for
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine

pytestmark = pytest.mark.gpu


def _full_vocab_engine():
    cfg = EngineConfig(hidden=64, inter=128, n_enc_layers=1, n_dec_layers=1, n_heads=4, n_kv_heads=2, head_dim=16,
                       sliding_window=8, query_pre_attn_scalar=16, text_vocab=32, audio_vocab=65536,
                       max_slots=1, max_text_len=16, max_dec_len=64, max_prefill_tokens=64)
    return T5GemmaVoiceEngine(cfg)      # weights not needed for t5g_sample


PARAMS = [dict(top_k=30, top_p=0.9, temperature=0.8), dict(top_k=30, top_p=1.0, temperature=1.0),
          dict(top_k=1, top_p=1.0, temperature=1.0), dict(top_k=5, top_p=0.5, temperature=1.3),
          dict(top_k=-100, top_p=1.0, temperature=1.0), dict(top_k=-100, top_p=1.0, temperature=0.7),
          dict(top_k=40, top_p=0.9, min_p=0.05, temperature=1.0), dict(top_k=1000, top_p=0.95, temperature=0.9),
          dict(top_k=50, top_p=0.3, temperature=0.8), dict(top_k=0, top_p=1.0, min_p=0.002, temperature=1.0)]


def _run(eng, V, eos, scale, n_rows, seed):
    rng = np.random.default_rng(seed)
    logits = (rng.standard_normal((n_rows, V)) * scale).astype(np.float32)
    # a few exact ties around the top to exercise tie handling
    logits[:, 7] = logits[:, 3]
    rows = []
    for i in range(n_rows):
        p = dict(PARAMS[i % len(PARAMS)])
        p.update(u=float(np.float32(rng.random())), cur_num_gen=int(rng.integers(0, 30)), prompt_offset=6,
                 target_total=40, n_text=10)
        p["current_length"] = p["prompt_offset"] + p["cur_num_gen"]
        if i % 7 == 0:
            logits[i, eos] = 50.0            # argmax == eos: forced stop once eos is no longer suppressed
        rows.append(p)
    want_tok, want_amax, want_logits = [], [], logits.copy()
    for i, p in enumerate(rows):
        tok, det = sampler_oracle.sample_step(want_logits[i], eos=eos, cur_num_gen=p["cur_num_gen"],
                                              current_length=p["current_length"], prompt_offset=p["prompt_offset"],
                                              target_total=p["target_total"], top_k=p.get("top_k", -100),
                                              top_p=p.get("top_p", 1.0), min_p=p.get("min_p", 0.0),
                                              temperature=p.get("temperature", 1.0), u=p["u"], x_len=p["n_text"],
                                              return_detail=True)
        want_tok.append(tok)
        want_amax.append(det["argmax"])
    dl = torch.from_numpy(logits).cuda()
    tok, amax = eng.sample(dl, rows)
    assert np.array_equal(tok, np.array(want_tok)), (tok, want_tok)
    assert np.array_equal(amax, np.array(want_amax))
    # in-place eos edits are visible to the caller exactly like the reference's logits_adjust
    assert np.array_equal(dl.cpu().numpy(), want_logits)


def test_sampler_bit_exact_tiny_vocab():
    eng = engine_for("tinyA_eager")
    _run(eng, eng.cfg.n_audio_tokens, eng.cfg.stop_token, 2.0, 200, 0)


def test_sampler_bit_exact_full_vocab():
    eng = _full_vocab_engine()
    _run(eng, 65541, 65539, 0.2, 60, 1)       # random-init-like flat logits (std 0.2)
    _run(eng, 65541, 65539, 3.0, 60, 2)       # peaked logits
    eng.close()


def test_sampler_known_answers_appendix_a():
    """SURVEY.md Appendix A (reference top_k_top_p_filtering, fp32 CPU): survivors determine which ids can
    ever be drawn; sweep u over [0,1) and check the drawn set equals the reference survivor set."""
    eng = engine_for("tinyA_eager")
    V = eng.cfg.n_audio_tokens
    base = np.full(V, -30.0, dtype=np.float32)
    base[:6] = [2.0, 1.0, 1.0, 0.0, -1.0, 3.0]
    cases = [(dict(top_k=2, top_p=1.0), {0, 5}), (dict(top_k=3, top_p=1.0), {0, 1, 2, 5}),
             (dict(top_k=4, top_p=0.9), {0, 1, 5}), (dict(top_k=1, top_p=0.9), {5}),
             (dict(top_k=2, top_p=0.5, min_p=0.2), {0, 5})]
    us = np.linspace(0, 0.999, 64, dtype=np.float32)
    for kw, want in cases:
        rows = [dict(kw, temperature=1.0, u=float(u), cur_num_gen=20, current_length=40, prompt_offset=6,
                     target_total=400) for u in us]
        dl = torch.from_numpy(np.tile(base, (len(us), 1))).cuda()
        tok, _ = eng.sample(dl, rows)
        assert set(tok.tolist()) == want, (kw, set(tok.tolist()))


LARGE_PARAMS = [dict(top_k=0, top_p=0.9, temperature=1.0), dict(top_k=-100, top_p=0.5, temperature=0.8),
                dict(top_k=2000, top_p=1.0, temperature=1.0), dict(top_k=5000, top_p=0.95, temperature=1.2),
                dict(top_k=0, top_p=0.999, temperature=1.0), dict(top_k=30, top_p=0.9, min_p=1e-5, temperature=1.0)]


def _run_large(eng, V, eos, scale, n_rows, seed, ties=False):
    rng = np.random.default_rng(seed)
    logits = (rng.standard_normal((n_rows, V)) * scale).astype(np.float32)
    if ties:
        logits = np.round(logits * 4) / 4          # massive ties: > CAND_CAP equal values around any threshold
    rows, want = [], []
    for i in range(n_rows):
        p = dict(LARGE_PARAMS[i % len(LARGE_PARAMS)])
        p.update(u=float(np.float32(rng.random())), cur_num_gen=20, current_length=40, prompt_offset=6, target_total=400)
        rows.append(p)
        want.append(sampler_oracle.sample_step(logits[i].copy(), eos=eos, cur_num_gen=20, current_length=40,
                                               prompt_offset=6, target_total=400, top_k=p.get("top_k", -100),
                                               top_p=p.get("top_p", 1.0), min_p=p.get("min_p", 0.0),
                                               temperature=p.get("temperature", 1.0), u=p["u"]))
    tok, _ = eng.sample(torch.from_numpy(logits).cuda(), rows)
    assert np.array_equal(tok, np.array(want)), (tok, want)


def test_sampler_general_path_bit_exact():
    """Nucleus sampling without top-k, top_k > 1024, min-p with many survivors, and > 2048 ties at the k-th value all
    take the full-sort path (stable LSD radix sort in global scratch + the oracle's sequential sums)."""
    eng = _full_vocab_engine()
    _run_large(eng, 65541, 65539, 0.2, 18, 11)
    _run_large(eng, 65541, 65539, 3.0, 12, 12)
    _run_large(eng, 65541, 65539, 1.0, 6, 13, ties=True)
    eng.close()
    small = engine_for("tinyA_eager")
    _run_large(small, small.cfg.n_audio_tokens, small.cfg.stop_token, 2.0, 48, 14)


def test_sampler_silence_penalty_bit_exact():
    """Silence-repetition penalty + state update (models/t5gemma.py:999-1011,1050-1054) vs the numpy oracle."""
    eng = engine_for("tinyA_eager")
    V, eos = eng.cfg.n_audio_tokens, eng.cfg.stop_token
    silence = [7, 11, 42]
    eng.set_sample_silence(silence, stop_repetition=2)
    rng = np.random.default_rng(3)
    n = 96
    logits = (rng.standard_normal((n, V)) * 2).astype(np.float32)
    rows, want = [], []
    for i in range(n):
        prev = [7, 11, 42, 5, -1][i % 5]
        consec = int(rng.integers(0, 7))
        logits[i, 7] = 6.0 if i % 2 else -6.0           # both branches of the penalty (divide / multiply)
        p = dict(top_k=[1, 30][i % 2], top_p=0.9, temperature=1.0, u=float(np.float32(rng.random())), cur_num_gen=20,
                 current_length=40, prompt_offset=6, target_total=400, prev_token=prev, consec_silence_count=consec)
        rows.append(p)
        lg = logits[i].copy()
        tok = sampler_oracle.sample_step(lg, eos=eos, cur_num_gen=20, current_length=40, prompt_offset=6, target_total=400,
                                         top_k=p["top_k"], top_p=0.9, temperature=1.0, u=p["u"], prev_token=prev,
                                         consec_silence_count=consec, stop_repetition=2, silence_tokens=tuple(silence))
        want.append((tok, lg))
    dl = torch.from_numpy(logits).cuda()
    tok, _ = eng.sample(dl, rows)
    got_l = dl.cpu().numpy()
    for i in range(n):
        assert int(tok[i]) == want[i][0], i
        assert np.array_equal(got_l[i], want[i][1]), i       # the in-place edit is bit-identical
    eng.set_sample_silence([], 3)
