"""GPU parity tests: the CUDA engine (through the C ABI) against the committed reference outputs
(tests/golden) and the CPU oracle.  Tolerances:
  * vs the fp32 reference fixtures: max-abs-err / abs-max <= 2e-2 (north_star's bf16 tolerance)
  * vs the oracle evaluated with bf16-rounded weights: <= 1.5e-2 (kernel logic; bf16 K/V cache and prefill GEMM inputs; activations stay fp32 on the
    decode path, bf16 on the prefill GEMM inputs)
  * greedy agreement, teacher-forced along the reference sequence: >= 99 % of steps
"""
import numpy as np
import pytest
import torch

from oracle import fixtures
from tests.gpu_util import engine_for, bf16_round_oracle, rel_err
from t5gemma_tts_b200 import GenerationRequest

pytestmark = pytest.mark.gpu

CASES = [("tinyA_eager", "tinyA_eager_prompt"), ("tinyA_eager", "tinyA_eager_noprompt"),
         ("tinyA_sdpa", "tinyA_sdpa_prompt"), ("tinyA_sdpa", "tinyA_sdpa_noprompt"),
         ("tinyB_eager", "tinyB_eager_prompt")]
TOL_REF = 2e-2
TOL_BF16W = 1.5e-2


def _request(c, **kw):
    return GenerationRequest(text_ids=c["x"][0], prompt_ids=c["y"][0, :, 0], target_total=int(c["tgt"]),
                             prompt_frames=int(c["prompt_frames"]), **kw)


@pytest.mark.parametrize("model,case", CASES)
def test_prefill_matches_reference(model, case):
    eng = engine_for(model)
    c = fixtures.load_case(case)
    req = _request(c, top_k=1)
    eng.prefill([req], [0])
    S = c["x"].shape[1]
    mem = eng.read_memory(0, S)
    assert rel_err(mem, c["memory"]) <= TOL_REF
    npre = c["y"].shape[1] + 1
    logits = eng.prefill_logits(0, npre)
    assert rel_err(logits, c["tf_logits"][:npre]) <= TOL_REF
    # kernel logic at tight tolerance: same computation with bf16-rounded weights on the CPU
    orc = bf16_round_oracle(model)
    mem_o = orc.encoder(torch.from_numpy(c["x"][0]))
    assert rel_err(mem, mem_o.numpy()) <= TOL_BF16W
    cross = orc.cross_kv(mem_o)
    cache = [None] * orc.cfg.n_dec_layers
    dec_ids = torch.from_numpy(c["dec_ids"][:npre])
    hid = orc.decoder(orc.embed_audio(dec_ids), torch.from_numpy(c["dec_pos"][:npre]), cache, cross)
    assert rel_err(logits, orc.head(hid).numpy()) <= TOL_BF16W
    assert rel_err(eng.read_last_hidden(0), hid[-1].numpy()) <= TOL_BF16W
    eng.release(0)


@pytest.mark.parametrize("model,case", CASES)
def test_decode_teacher_forced_matches_reference(model, case):
    """Steps the CUDA-graph decode loop one token at a time along the reference's own greedy sequence
    and compares the logits seen by the sampler at every step (cached-decode path of the reference)."""
    eng = engine_for(model)
    c = fixtures.load_case(case)
    gen = c["gen"][0, 0]
    req = _request(c, top_k=1, forced_tokens=gen)
    eng.prefill([req], [0])
    eos = eng.cfg.stop_token
    worst = 0.0
    for step in range(len(gen)):
        eng.decode(1)
        st = eng.poll()[0]
        assert st.n_generated == step + 1
        got = eng.read_logits(0)
        ref = c["step_logits"][step].copy()
        got[eos] = 0.0
        ref[eos] = 0.0
        worst = max(worst, rel_err(got, ref))
    assert worst <= TOL_REF, worst
    st = eng.poll()[0]
    assert st.finished == 1 and st.active == 0
    toks = eng.read_tokens(0)
    assert np.array_equal(toks, gen)                       # forced sequence incl. the stop-rule eos
    picks = eng.read_picks(0)
    agree = float((picks[:-1] == gen[:-1]).mean())         # last step is the time-budget eos, not an argmax
    assert agree >= 0.99, agree
    eng.release(0)


@pytest.mark.parametrize("model,case", CASES[:1] + CASES[4:])
def test_inference_tts_contract(model, case):
    """Drop-in call: same signature, shapes, dtypes, device; greedy tokens vs the reference run."""
    eng = engine_for(model)
    c = fixtures.load_case(case)
    x = torch.from_numpy(c["x"]).cuda()
    y = torch.from_numpy(c["y"]).cuda()
    res, gen = eng.inference_tts(x, torch.tensor([x.shape[1]]).cuda(), y, torch.tensor([int(c["tgt"])]).cuda(),
                                 top_k=1, top_p=1.0, temperature=1.0, prompt_frames=int(c["prompt_frames"]))
    assert res.dtype == torch.long and gen.dtype == torch.long and res.device == x.device
    assert gen.dim() == 3 and gen.shape[:2] == (1, 1) and res.shape == (1, 1, y.shape[1] + gen.shape[2])
    assert int(gen[0, 0, -1]) == eng.cfg.stop_token
    assert torch.equal(res[0, 0, : y.shape[1]], y[0, :, 0])
    # random-init runs end at the time-budget cutoff: same length as the reference
    assert gen.shape[2] == c["gen"].shape[2]
    # free-running greedy: report agreement (bf16 weights may flip near-tied argmaxes and then diverge)
    n = min(gen.shape[2], c["gen"].shape[2])
    first_div = next((i for i in range(n) if int(gen[0, 0, i]) != int(c["gen"][0, 0, i])), n)
    assert first_div >= 8, first_div


def test_inference_tts_errors():
    eng = engine_for("tinyA_eager")
    x = torch.randint(2, 500, (2, 5)).cuda()
    with pytest.raises(AssertionError):
        eng.inference_tts(x, torch.tensor([5, 5]), torch.zeros(2, 0, 1, dtype=torch.long), torch.tensor([10, 10]))
    old = getattr(eng.args, "n_codebooks", 1)
    eng.args.n_codebooks = 2
    with pytest.raises(ValueError):
        eng.inference_tts(x[:1], torch.tensor([5]), torch.zeros(1, 0, 1, dtype=torch.long), torch.tensor([10]))
    eng.args.n_codebooks = old


def test_batched_rows_equal_single_runs():
    """bs>1 has no reference path; each row must equal the bs=1 run of that request (Appendix B.13)."""
    eng1 = engine_for("tinyA_eager")
    eng3 = engine_for("tinyA_eager", max_slots=3)
    rng = np.random.default_rng(5)
    reqs = []
    for i in range(5):
        S = int(rng.integers(4, 30))
        Tp = int(rng.integers(0, 12))
        u = torch.rand(400, generator=torch.Generator().manual_seed(100 + i)).cuda()
        reqs.append(GenerationRequest(text_ids=rng.integers(2, 500, S), prompt_ids=rng.integers(0, 100, Tp),
                                      target_total=Tp + int(rng.integers(5, 40)), top_k=20, top_p=0.9,
                                      temperature=0.8, uniforms=u, max_new_tokens=int(rng.integers(20, 60))))
    single = [eng1.generate([r])[0] for r in reqs]
    batched = eng3.generate(reqs, chunk_steps=7)
    for a, b in zip(single, batched):
        assert np.array_equal(a, b)
    assert all(s[-1] == eng1.cfg.stop_token for s in single)


def test_continuous_batching_keeps_per_slot_schedules_and_silence_lists():
    """More requests than rows, every request with its own top_k LIST (models/t5gemma.py:991-994) and silence-token
    list: admissions happen while other rows are still decoding, and must not disturb their schedules / lists (each
    slot owns a fixed region of the engine's int pool).  Every row must equal the bs=1 run of that request."""
    eng1 = engine_for("tinyA_eager")
    eng3 = engine_for("tinyA_eager", max_slots=3)
    rng = np.random.default_rng(17)
    reqs = []
    for i in range(8):
        S, Tp = int(rng.integers(4, 30)), int(rng.integers(0, 12))
        u = torch.rand(400, generator=torch.Generator().manual_seed(300 + i)).cuda()
        sched = [int(k) for k in rng.integers(1, 40, int(rng.integers(1, 30)))]
        sil = [int(t) for t in rng.choice(100, int(rng.integers(1, 6)), replace=False)]
        reqs.append(GenerationRequest(text_ids=rng.integers(2, 500, S), prompt_ids=rng.integers(0, 100, Tp),
                                      target_total=Tp + int(rng.integers(5, 40)), top_k=sched, top_p=0.95, temperature=1.3,
                                      uniforms=u, max_new_tokens=int(rng.integers(10, 70)), stop_repetition=1,
                                      silence_tokens=sil))
    single = [eng1.generate([r])[0] for r in reqs]
    batched = eng3.generate(reqs, chunk_steps=5)
    for a, b in zip(single, batched):
        assert np.array_equal(a, b)
    # the schedules matter: the same requests with a constant top_k give different tokens
    import dataclasses
    flat = [eng1.generate([dataclasses.replace(r, top_k=1)])[0] for r in reqs]
    assert any(not np.array_equal(a, b) for a, b in zip(single, flat))


def test_streaming_chunks_concatenate_to_the_full_result():
    """generate_stream / inference_tts(on_chunk=...) hand out token chunks every chunk_steps decode steps;
    their concatenation is exactly the non-streaming result."""
    eng = engine_for("tinyA_eager", max_slots=3)
    rng = np.random.default_rng(23)
    reqs = []
    for i in range(5):
        u = torch.rand(400, generator=torch.Generator().manual_seed(500 + i)).cuda()
        reqs.append(GenerationRequest(text_ids=rng.integers(2, 500, 9 + i), prompt_ids=rng.integers(0, 100, 3 * i),
                                      target_total=3 * i + 20, top_k=10, temperature=0.9, uniforms=u,
                                      max_new_tokens=25 + 7 * i))
    full = eng.generate(reqs, chunk_steps=16)
    parts = {i: [] for i in range(len(reqs))}
    n_chunks = {i: 0 for i in range(len(reqs))}
    finished = set()
    for i, new, fin in eng.generate_stream(reqs, chunk_steps=6):
        assert i not in finished
        parts[i].append(new)
        n_chunks[i] += 1
        if fin:
            finished.add(i)
    assert finished == set(range(len(reqs)))
    for i, f in enumerate(full):
        assert np.array_equal(np.concatenate(parts[i]), f)
        assert max(len(p) for p in parts[i]) <= 6 and n_chunks[i] >= -(-len(f) // 6)
    assert max(n_chunks.values()) >= 4
    # drop-in call with a streaming callback
    c = fixtures.load_case("tinyA_eager_prompt")
    x, y = torch.from_numpy(c["x"]).cuda(), torch.from_numpy(c["y"]).cuda()
    got = []
    eng1 = engine_for("tinyA_eager")
    res, gen = eng1.inference_tts(x, torch.tensor([x.shape[1]]), y, torch.tensor([int(c["tgt"])]), top_k=1,
                                  prompt_frames=int(c["prompt_frames"]), chunk_steps=50, on_chunk=lambda t, fin: got.append((t, fin)))
    assert len(got) >= 5 and got[-1][1] and not got[0][1]
    assert np.array_equal(np.concatenate([t for t, _ in got]), gen[0, 0].cpu().numpy())


def test_top_k_list_is_applied_per_step():
    """top_k as a per-step list: at every step the engine's pick equals the numpy sampler oracle run on the engine's
    own logits with kk = top_k[min(len-1, cur_num_gen)] and the same uniform draw (bit-exact sampler contract)."""
    from oracle import sampler_oracle
    eng = engine_for("tinyA_eager")
    c = fixtures.load_case("tinyA_eager_prompt")
    gen = c["gen"][0, 0][:40]
    sched = [1, 50, 3, 100, 2, 7, 30]
    u = torch.rand(64, generator=torch.Generator().manual_seed(9))
    req = _request(c, top_k=sched, top_p=0.9, temperature=0.7, uniforms=u.cuda(), forced_tokens=gen, max_new_tokens=40)
    eng.prefill([req], [0])
    npre = c["y"].shape[1] + 1
    want = []
    for step in range(39):
        eng.decode(1)
        eng.poll()
        lg = eng.read_logits(0)        # logits as the sampler left them (eos edits applied in place, idempotent)
        want.append(sampler_oracle.sample_step(lg.copy(), eos=eng.cfg.stop_token, cur_num_gen=step, current_length=npre + step,
                                               prompt_offset=int(c["prompt_frames"]) + 1, target_total=int(c["tgt"]),
                                               top_k=sched[min(len(sched) - 1, step)], top_p=0.9, temperature=0.7,
                                               u=float(u[step]), x_len=c["x"].shape[1]))
    picks = eng.read_picks(0)[:39]
    assert np.array_equal(picks, np.asarray(want)), (picks, want)
    assert len(set(picks.tolist())) > 5                 # not a degenerate greedy run
    eng.release(0)


def test_no_target_fallback_matches_reference():
    """tgt_y_lens=None (models/t5gemma.py:896-933): est_total = current_length + 2 s lookahead, no time budget; the
    text guard (3 frames per text token) stops the run exactly where the reference stops.  Teacher-forced logits along
    the reference sequence, then the drop-in call with tgt_y_lens=None."""
    from types import SimpleNamespace
    from t5gemma_tts_b200 import T5GemmaVoiceEngine
    _, sd, meta = fixtures.load_model_fixture("tinyA_eager")
    c = fixtures.load_case("tinyA_eager_notarget")
    meta = dict(meta, text_guard_frames_per_token=int(c["text_guard_frames_per_token"]))
    eng = T5GemmaVoiceEngine.from_state_dict(SimpleNamespace(**meta), sd, max_slots=1, max_text_len=64, max_dec_len=512,
                                             max_prefill_tokens=512)
    gen = c["gen"][0, 0]
    req = GenerationRequest(text_ids=c["x"][0], prompt_ids=c["y"][0, :, 0], target_total=None,
                            prompt_frames=int(c["prompt_frames"]), top_k=[1, 1, 1], forced_tokens=gen)
    eng.prefill([req], [0])
    eos = eng.cfg.stop_token
    worst = 0.0
    for step in range(len(gen)):
        eng.decode(1)
        eng.poll()
        got, ref = eng.read_logits(0), c["step_logits"][step].copy()
        got[eos] = ref[eos] = 0.0
        worst = max(worst, rel_err(got, ref))
    assert worst <= TOL_REF, worst
    st = eng.poll()[0]
    assert st.finished == 1 and st.n_generated == len(gen)
    assert np.array_equal(eng.read_tokens(0), gen)
    eng.release(0)
    x, y = torch.from_numpy(c["x"]).cuda(), torch.from_numpy(c["y"]).cuda()
    res, g = eng.inference_tts(x, torch.tensor([x.shape[1]]).cuda(), y, None, top_k=[1, 1, 1], prompt_frames=int(c["prompt_frames"]))
    assert g.shape[2] == len(gen) and int(g[0, 0, -1]) == eos
    assert float((g[0, 0].cpu().numpy() == gen).mean()) >= 0.9
    eng.close()


def test_sliding_window_binds():
    """tiny configs use window 8 / 16 so the window mask is exercised by every decode test; check the
    decode attention really limits the context by comparing against a wide-window engine."""
    from types import SimpleNamespace
    import copy
    _, sd, meta = fixtures.load_model_fixture("tinyA_eager")
    meta2 = copy.deepcopy(meta)
    for side in ("encoder", "decoder"):
        meta2["t5_config_dict"][side]["sliding_window"] = 4096
    from t5gemma_tts_b200 import T5GemmaVoiceEngine
    wide = T5GemmaVoiceEngine.from_state_dict(SimpleNamespace(**meta2), sd, max_slots=1, max_text_len=64,
                                              max_dec_len=512, max_prefill_tokens=512)
    c = fixtures.load_case("tinyA_eager_prompt")
    gen = c["gen"][0, 0]
    req = _request(c, top_k=1, forced_tokens=gen)
    wide.prefill([req], [0])
    wide.decode(40)
    wide.poll()
    got = wide.read_logits(0)
    ref = c["step_logits"][39]
    eos = wide.cfg.stop_token
    got[eos] = ref[eos] = 0
    assert rel_err(got, ref) > 1e-3      # differs from the windowed reference
    wide.close()


def test_batched_tensor_core_step_matches_reference():
    """max_slots > 4 switches the decode step to the tcgen05 GEMM path (M = batch rows).  Three requests in
    non-contiguous slots of an 8-row engine, teacher-forced along the reference sequences; every row's logits
    must match the reference's cached-decode logits within the bf16 tolerance."""
    eng = engine_for("tinyA_eager", max_slots=8, max_prefill_tokens=1024)
    cases = [fixtures.load_case("tinyA_eager_prompt"), fixtures.load_case("tinyA_eager_noprompt"),
             fixtures.load_case("tinyA_eager_prompt")]
    slots = [5, 0, 2]
    reqs = [_request(c, top_k=1, forced_tokens=c["gen"][0, 0]) for c in cases]
    eng.prefill(reqs, slots)
    eos = eng.cfg.stop_token
    n_steps = max(len(c["gen"][0, 0]) for c in cases)
    worst = 0.0
    for step in range(n_steps):
        eng.decode(1)
        st = eng.poll()
        for c, s in zip(cases, slots):
            if step < len(c["gen"][0, 0]):
                assert st[s].n_generated == step + 1
                got = eng.read_logits(s)
                ref = c["step_logits"][step].copy()
                got[eos] = ref[eos] = 0.0
                worst = max(worst, rel_err(got, ref))
    assert worst <= TOL_REF, worst
    st = eng.poll()
    for c, s in zip(cases, slots):
        assert st[s].finished == 1
        assert np.array_equal(eng.read_tokens(s), c["gen"][0, 0])
        picks = eng.read_picks(s)
        assert float((picks[:-1] == c["gen"][0, 0][:-1]).mean()) >= 0.99
        eng.release(s)
    assert st[1].active == 0 and st[1].n_generated == 0      # untouched rows stay idle


def test_from_pretrained_hf_dir(tmp_path):
    """HF-format directory (config.json + sharded safetensors) -> engine -> same prefill logits as the fixture."""
    import json
    from safetensors.torch import save_file
    from t5gemma_tts_b200 import T5GemmaVoiceEngine
    _, sd, meta = fixtures.load_model_fixture("tinyB_eager")
    d = tmp_path / "hf"
    d.mkdir()
    (d / "config.json").write_text(json.dumps(dict(meta, model_type="t5gemma_voice")))
    keys = sorted(sd)
    wm = {}
    for i, ks in enumerate((keys[: len(keys) // 2], keys[len(keys) // 2:])):
        fn = f"model-{i + 1:05d}-of-00002.safetensors"
        save_file({k: sd[k].contiguous() for k in ks}, str(d / fn))
        wm.update({k: fn for k in ks})
    (d / "model.safetensors.index.json").write_text(json.dumps({"weight_map": wm}))
    eng = T5GemmaVoiceEngine.from_pretrained(str(d), max_slots=1, max_text_len=64, max_dec_len=512, max_prefill_tokens=512)
    c = fixtures.load_case("tinyB_eager_prompt")
    eng.prefill([_request(c, top_k=1)], [0])
    npre = c["y"].shape[1] + 1
    assert rel_err(eng.prefill_logits(0, npre), c["tf_logits"][:npre]) <= TOL_REF
    assert eng.config.audio_vocab_size == 200          # reference-facing config passes through
    eng.close()


@pytest.mark.parametrize("max_slots,env", [(1, {}), (8, {}), (3, {"T5G_GEMV_PAIR": "0"}), (8, {"T5G_ATTN_CHUNK": "32"}),
                                           (8, {"T5G_ATTN_CHUNK": "0"}), (8, {"T5G_ATTN_TMA": "0", "T5G_ATTN_CHUNK": "32"}),
                                           (8, {"GEOM": "4x1x64", "T5G_ATTN_CHUNK": "32"}), (8, {"GEOM": "2x2x128", "PT": "16"}),
                                           (8, {"PT": "16", "T5G_ATTN_CHUNK": "32"}), (1, {"PT": "16"}),
                                           (3, {"GEOM": "4x1x64"})],
                         ids=["single", "batched-mma", "three-rows-unpaired", "batched-mma-4-chunks", "batched-mma-unchunked",
                              "batched-cp.async-4-chunks", "batched-G4-D64-chunks", "batched-G1-D128-pages16", "batched-pages16-chunks",
                              "single-pages16", "three-rows-G4-D64"])
def test_head_dim_256_decode_matches_oracle(max_slots, env, monkeypatch):
    """The production head geometry (head_dim 256, 2 query heads per KV head) on a narrow 2+2-layer model with a
    sliding window of 48: exercises the D=256 instantiations of both decode attention kernels (CUDA-core for
    max_slots <= 4, cp.async + mma.sync tile kernel for batched rows) with contexts that cross the 32-token tile
    and the window, teacher-forced along the oracle's greedy sequences; logits within the bf16 tolerance.  The third
    case runs three rows through the GEMV path with o_proj and the cross q projection as two kernels; the last two run the
    batched attention with 32-key chunks (up to 4 chunks per row and kv head, merged by the last CTA to arrive; the text
    of 70 tokens gives the cross-attention 3 chunks) and with chunking off (one CTA per row and kv head); then the `cp.async`
    front end of the tile kernel, and the other head geometries the TMA front end is instantiated for (4 query heads per kv
    head at head_dim 64, one at head_dim 128)."""
    env = dict(env)
    nh, nkv, hd = (int(v) for v in env.pop("GEOM", "2x1x256").split("x"))   # other head geometries of the attention kernels
    page_tokens = int(env.pop("PT", "32"))                                  # KV page size (two TMA boxes per tile at 16)
    for k, v in env.items():
        monkeypatch.setenv(k, v)                     # read by t5g_create
    from oracle.t5gemma_voice_oracle import Oracle, OracleConfig
    from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine
    from t5gemma_tts_b200.random_init import iter_random_state_dict
    cfg = EngineConfig(hidden=512, inter=1024, n_enc_layers=2, n_dec_layers=2, n_heads=nh, n_kv_heads=nkv, head_dim=hd,
                       query_pre_attn_scalar=float(hd), sliding_window=48, text_vocab=300, audio_vocab=400,
                       max_slots=max_slots, max_text_len=96, max_dec_len=512, max_prefill_tokens=1024,
                       kv_page_tokens=page_tokens)
    sd = {k: v.float().cpu() for k, v in iter_random_state_dict(cfg, seed=3, device="cuda")}
    ocfg = OracleConfig(hidden=cfg.hidden, inter=cfg.inter, n_enc_layers=cfg.n_enc_layers, n_dec_layers=cfg.n_dec_layers,
                        n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads, head_dim=cfg.head_dim,
                        sliding_window=cfg.sliding_window, query_pre_attn_scalar=cfg.query_pre_attn_scalar,
                        attn_softcap=cfg.attn_softcap, text_vocab=cfg.text_vocab, audio_vocab=cfg.audio_vocab,
                        n_special=cfg.n_special)
    orc = Oracle(ocfg, sd)
    eng = T5GemmaVoiceEngine(cfg)
    eng.load_state_dict(iter_random_state_dict(cfg, seed=3, device="cuda"))
    rng = np.random.default_rng(11)
    shapes = [(40, 0), (70, 40), (33, 75)] if max_slots > 1 else [(70, 40)]     # (text tokens, prompt tokens)
    slots = ([5, 0, 2] if max_slots > 4 else [2, 0, 1]) if max_slots > 1 else [0]
    N_NEW = 40
    refs, reqs = [], []
    for S, P in shapes:
        x = torch.from_numpy(rng.integers(2, cfg.text_vocab, S))[None]
        y = torch.from_numpy(rng.integers(0, cfg.audio_vocab, P))[None, :, None]
        if P:
            y = torch.cat([y, torch.tensor([[[cfg.y_sep_token]]])], dim=1)
        tgt = torch.tensor([y.shape[1] + 100])
        with torch.no_grad():
            _, gen, logits = orc.inference_tts(x, torch.tensor([S]), y, tgt, top_k=1, prompt_frames=y.shape[1],
                                               max_new_tokens=N_NEW, return_logits=True)
        refs.append((gen[0, 0].numpy(), logits.numpy()))
        reqs.append(GenerationRequest(text_ids=x[0].numpy(), prompt_ids=y[0, :, 0].numpy(), target_total=int(tgt[0]),
                                      prompt_frames=y.shape[1], top_k=1, forced_tokens=gen[0, 0].numpy()))
    eng.prefill(reqs, slots)
    eos = cfg.stop_token
    worst = 0.0
    for step in range(N_NEW):
        eng.decode(1)
        eng.poll()
        for (gen, logits), s in zip(refs, slots):
            if step < len(gen):
                got = eng.read_logits(s)
                ref = logits[step].copy()
                got[eos] = ref[eos] = 0.0
                worst = max(worst, rel_err(got, ref))
    assert worst <= TOL_REF, worst
    for (gen, _), s in zip(refs, slots):
        assert np.array_equal(eng.read_tokens(s)[:len(gen)], gen)
    eng.close()


@pytest.mark.parametrize("max_slots", [1, 64], ids=["single-row-gemv", "64-row-tcgen05"])
def test_production_width_two_layers_matches_oracle(max_slots):
    """2b-2b WIDTH (hidden 2304, 8 query / 4 kv heads of 256, MLP 9216, audio vocabulary 65 536) with 2 + 2 layers, so that the
    fp32 CPU oracle stays a few seconds: a 300-token voice prompt (prefill through the CTA-pair GEMM and the single-pass
    tcgen05 attention), then 32 teacher-forced decode steps through the production-width decode kernels (K = 2304 / 9216
    GEMVs and the paired projection at one row; tcgen05 GEMMs, chunked TMA attention and the full-vocabulary sampler at 64
    rows).  Encoder states and every step's logits within the bf16 tolerance.  (`bench.py` `parity_2b` does the same over
    all 26 + 26 layers.)"""
    from oracle.t5gemma_voice_oracle import Oracle, OracleConfig
    from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine
    from t5gemma_tts_b200.random_init import iter_random_state_dict
    cfg = EngineConfig(hidden=2304, inter=9216, n_enc_layers=2, n_dec_layers=2, n_heads=8, n_kv_heads=4, head_dim=256,
                       query_pre_attn_scalar=256.0, sliding_window=4096, text_vocab=1000, audio_vocab=65536,
                       max_slots=max_slots, max_text_len=128, max_dec_len=1024, max_prefill_tokens=2048)
    sd = {k: v.float().cpu() for k, v in iter_random_state_dict(cfg, seed=5, device="cuda")}
    ocfg = OracleConfig(hidden=cfg.hidden, inter=cfg.inter, n_enc_layers=cfg.n_enc_layers, n_dec_layers=cfg.n_dec_layers,
                        n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads, head_dim=cfg.head_dim,
                        sliding_window=cfg.sliding_window, query_pre_attn_scalar=cfg.query_pre_attn_scalar,
                        attn_softcap=cfg.attn_softcap, text_vocab=cfg.text_vocab, audio_vocab=cfg.audio_vocab,
                        n_special=cfg.n_special)
    orc = Oracle(ocfg, sd)
    eng = T5GemmaVoiceEngine(cfg)
    eng.load_state_dict(iter_random_state_dict(cfg, seed=5, device="cuda"))
    rng = np.random.default_rng(7)
    S, P, N_NEW = 96, 300, 32
    x = torch.from_numpy(rng.integers(2, cfg.text_vocab, S))[None]
    y = torch.cat([torch.from_numpy(rng.integers(0, cfg.audio_vocab, P))[None, :, None], torch.tensor([[[cfg.y_sep_token]]])], dim=1)
    tgt = torch.tensor([y.shape[1] + 200])
    with torch.no_grad():
        _, gen, logits = orc.inference_tts(x, torch.tensor([S]), y, tgt, top_k=1, prompt_frames=y.shape[1],
                                           max_new_tokens=N_NEW, return_logits=True)
        mem = orc.encoder(x[0]).numpy()
    gen, logits = gen[0, 0].numpy(), logits.numpy()
    req = GenerationRequest(text_ids=x[0].numpy(), prompt_ids=y[0, :, 0].numpy(), target_total=int(tgt[0]),
                            prompt_frames=y.shape[1], top_k=1, forced_tokens=gen)
    slot = 37 if max_slots > 1 else 0
    eng.prefill([req], [slot])
    assert rel_err(eng.read_memory(slot, S), mem) <= TOL_REF
    eos = cfg.stop_token
    worst = 0.0
    for step in range(len(gen)):
        eng.decode(1)
        eng.poll()
        got, ref = eng.read_logits(slot), logits[step].copy()
        got[eos] = ref[eos] = 0.0
        worst = max(worst, rel_err(got, ref))
    assert worst <= TOL_REF, worst
    assert np.array_equal(eng.read_tokens(slot)[:len(gen)], gen)
    eng.close()
