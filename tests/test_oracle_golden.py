"""Pins oracle/ against the committed outputs of the unmodified reference (tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import fixtures, sampler_oracle

CASES = [("tinyA_eager", "tinyA_eager_prompt"), ("tinyA_eager", "tinyA_eager_noprompt"),
         ("tinyA_sdpa", "tinyA_sdpa_prompt"), ("tinyA_sdpa", "tinyA_sdpa_noprompt"),
         ("tinyB_eager", "tinyB_eager_prompt")]
TOL = 2e-5   # fp32 restatement vs fp32 reference, max-abs


@pytest.mark.parametrize("model,case", CASES)
def test_encoder_and_teacher_forced_logits(model, case):
    orc = fixtures.load_oracle(model)
    c = fixtures.load_case(case)
    x = torch.from_numpy(c["x"])[0]
    mem = orc.encoder(x)
    assert np.abs(mem.numpy() - c["memory"]).max() <= TOL
    cross = orc.cross_kv(mem)
    cache = [None] * orc.cfg.n_dec_layers
    dec_ids = torch.from_numpy(c["dec_ids"])
    hid = orc.decoder(orc.embed_audio(dec_ids), torch.from_numpy(c["dec_pos"]), cache, cross)
    logits = orc.head(hid).numpy()
    assert np.abs(logits - c["tf_logits"]).max() <= TOL


@pytest.mark.parametrize("model,case", CASES)
def test_greedy_generate_matches_reference(model, case):
    orc = fixtures.load_oracle(model)
    c = fixtures.load_case(case)
    x = torch.from_numpy(c["x"])
    y = torch.from_numpy(c["y"])
    res, gen, logits = orc.inference_tts(x, torch.tensor([x.shape[1]]), y, torch.tensor([int(c["tgt"])]),
                                         top_k=1, top_p=1.0, temperature=1.0,
                                         prompt_frames=int(c["prompt_frames"]), return_logits=True)
    assert gen.shape == tuple(c["gen"].shape)
    assert np.array_equal(gen.numpy(), c["gen"])
    assert np.array_equal(res.numpy(), c["res"])
    # per-step logits (cached decode path) -- eos column is edited in place by the reference
    ref = c["step_logits"].copy()
    got = logits.numpy()
    eos = orc.cfg.eos
    ref[:, eos] = 0
    got[:, eos] = 0
    assert np.abs(got - ref).max() <= 5e-5


def test_positions_match_reference_formulas():
    orc = fixtures.load_oracle("tinyA_eager")
    c = fixtures.load_case("tinyA_eager_prompt")
    npre = c["y"].shape[1] + 1
    est = int(c["est_total"])
    pos = orc.decoder_prefill_positions(npre, est).numpy()
    assert np.array_equal(pos, c["dec_pos"][:npre])
    for t in range(npre, c["dec_pos"].shape[0]):
        assert np.float32(orc.decoder_step_position(t + 1, est)) == c["dec_pos"][t]
    assert c["dec_pos"][-1] == np.float32(2000.0)      # saturates once generation overruns the target


def test_sampler_survivor_sets_match_reference():
    n = 0
    for logits, keep, probs, kw in fixtures.load_sampler_cases():
        T = kw["temperature"]
        z = (logits / np.float32(T)).astype(np.float32) if T != 1.0 else logits
        surv, filtered = sampler_oracle.filter_survivors(z, kw.get("top_k", 0), kw.get("top_p", 1.0), kw.get("min_p", 0.0))
        got = np.zeros(z.shape[0], dtype=bool)
        got[surv] = True
        assert np.array_equal(got, keep), kw
        # oracle's survivor probabilities agree with the reference softmax over survivors
        if filtered:
            zs = z[surv]
            e = sampler_oracle.det_exp(zs - zs[0])
            p = e / e.sum()
            assert np.abs(p - probs[surv]).max() < 1e-6
        n += 1
    assert n >= 81


def test_det_exp_accuracy():
    x = -np.abs(np.random.default_rng(0).standard_normal(100000).astype(np.float32)) * 20
    got = sampler_oracle.det_exp(x).astype(np.float64)
    ref = np.exp(x.astype(np.float64))
    rel = (np.abs(got - ref) / ref)[x >= -86.0]
    assert rel.max() < 5e-7
    assert (got[x < -86.0] == 0).all()
    assert sampler_oracle.det_exp(np.float32(0.0)) == np.float32(1.0)
    assert sampler_oracle.det_exp(np.float32(-100.0)) == np.float32(0.0)


def test_draw_is_inverse_cdf():
    rng = np.random.default_rng(1)
    z = rng.standard_normal(105).astype(np.float32)
    surv, filt = sampler_oracle.filter_survivors(z, 5, 1.0, 0.0)
    assert filt and len(surv) == 5
    assert sampler_oracle.draw(z, surv, filt, 0.0) == surv[0]
    assert sampler_oracle.draw(z, surv, filt, 0.9999999) == surv[-1]
    # unfiltered path covers the whole vocabulary
    s2, f2 = sampler_oracle.filter_survivors(z, -100, 1.0, 0.0)
    assert not f2 and len(s2) == 105
    counts = np.zeros(105)
    for u in rng.random(4000):
        counts[sampler_oracle.draw(z, s2, f2, u)] += 1
    p = np.exp(z - z.max()); p /= p.sum()
    assert np.abs(counts / 4000 - p).max() < 0.03


def test_stop_rules():
    V, eos = 105, 103
    base = np.zeros(V, dtype=np.float32)
    base[7] = 5.0
    kw = dict(eos=eos, prompt_offset=6, target_total=15, top_k=1, top_p=1.0, temperature=1.0, u=0.3)
    # first generated token: eos forced to -1e9 then -10000 (both edits applied, the later wins)
    lg = base.copy(); lg[eos] = 100.0
    assert sampler_oracle.sample_step(lg, cur_num_gen=0, current_length=6, **kw) == 7
    assert lg[eos] == np.float32(-10000.0)
    # after 10 tokens eos is no longer suppressed: argmax==eos forces stop
    lg = base.copy(); lg[eos] = 100.0
    assert sampler_oracle.sample_step(lg, cur_num_gen=11, current_length=17, **kw) == eos
    # time budget: cur_num_gen > target_total - prompt_offset + 250
    lg = base.copy()
    assert sampler_oracle.sample_step(lg, cur_num_gen=259, current_length=265, **kw) == 7
    lg = base.copy()
    assert sampler_oracle.sample_step(lg, cur_num_gen=260, current_length=266, **kw) == eos


def test_silence_repetition_penalty_matches_reference():
    """models/t5gemma.py:999-1011,1050-1054 pinned through the reference's own sample_helper (crafted head logits)."""
    orc = fixtures.load_oracle("tinyA_eager")
    c = fixtures.load_case("tinyA_eager_silence")
    x = torch.from_numpy(c["x"])
    base = c["base_logits"]
    res, gen = orc.inference_tts(x, torch.tensor([x.shape[1]]), torch.zeros((1, 0, 1), dtype=torch.long),
                                 torch.tensor([int(c["tgt"])]), top_k=1, top_p=1.0, temperature=1.0, prompt_frames=0,
                                 stop_repetition=int(c["stop_repetition"]), silence_tokens=c["silence_tokens"].tolist(),
                                 logits_hook=lambda lg, step: base.copy())
    assert np.array_equal(gen.numpy(), c["gen"])
    assert gen[0, 0, :10].tolist() == [7, 7, 7, 7, 9, 7, 7, 7, 7, 9]


def test_no_target_fallback_and_top_k_list_match_reference():
    """tgt_y_lens=None (models/t5gemma.py:896-933): est_total from the 2 s lookahead, no time budget, the text guard
    (3 frames per text token) ends the utterance; top_k given as a list (models/t5gemma.py:991-994)."""
    cfg, sd, meta = fixtures.load_model_fixture("tinyA_eager")
    c = fixtures.load_case("tinyA_eager_notarget")
    cfg.text_guard_frames_per_token = int(c["text_guard_frames_per_token"])
    from oracle.t5gemma_voice_oracle import Oracle
    orc = Oracle(cfg, sd)
    x, y = torch.from_numpy(c["x"]), torch.from_numpy(c["y"])
    assert int(c["tgt"]) == -1 and int(c["est_total"]) == y.shape[1] + 1 + 100
    res, gen, logits = orc.inference_tts(x, torch.tensor([x.shape[1]]), y, None, top_k=[1, 1, 1], top_p=1.0,
                                         temperature=1.0, prompt_frames=int(c["prompt_frames"]), return_logits=True)
    assert np.array_equal(gen.numpy(), c["gen"])
    assert gen.shape[-1] == 3 * x.shape[1] + 2          # effective_length > 3 * n_text fires at the 44th token
    ref, got = c["step_logits"].copy(), logits.numpy()
    ref[:, orc.cfg.eos] = 0
    got[:, orc.cfg.eos] = 0
    assert np.abs(got - ref).max() <= 5e-5


@pytest.mark.parametrize("model", ["tinyA_eager", "tinyA_sdpa", "tinyB_eager"])
def test_fixture_norm_gains_are_distinct_and_nonzero(model):
    """A swapped / dropped RMSNorm gain must be visible: every norm tensor of every fixture is a different non-zero draw."""
    _, sd, _ = fixtures.load_model_fixture(model)
    norms = {k: v.numpy() for k, v in sd.items() if k.endswith("layernorm.weight") or k.endswith(".norm.weight")}
    assert len(norms) >= 32
    keys = sorted(norms)
    for k in keys:
        assert np.abs(norms[k]).mean() > 0.1, k
    for i in range(len(keys)):
        for j in range(i + 1, len(keys)):
            assert np.abs(norms[keys[i]] - norms[keys[j]]).max() > 0.1, (keys[i], keys[j])
