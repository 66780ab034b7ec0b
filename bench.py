#!/usr/bin/env python
"""bench.py -- T5Gemma-TTS token-generation hot path on B200 (contract: see the task statement).

A "step" is one utterance through the hot path (encoder prefill -> decoder prefill over the voice prompt ->
autoregressive decode until the reference's stop rules fire).  Workload at every N = BASELINE.json configs[1]:
T5Gemma-TTS-2b-2b random-init, bf16 weights, batch 1 per GPU, 64-token text, 150-token voice prompt + 10 s
target (500 XCodec2 tokens; random-init never emits EOS, so the time-budget rule stops at 751 tokens exactly as
the reference does), top_k=30 / top_p=0.9 / T=0.8.  N>1: one engine replica per GPU, requests sharded, no
collective on the data path (NCCL only gathers token counts and timings) -> "scaling": "weak".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-cpu-baseline]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = ("configs[1]: T5Gemma-TTS-2b-2b random-init (seed 0), bs=1/GPU, 64-token text, 150-token voice prompt "
            "(+y_sep) + 10 s target (500 tok), top_k=30 top_p=0.9 T=0.8")
N_TEXT, N_PROMPT, TARGET_TOKENS = 64, 150, 500


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(seed: int, cfg):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(2, min(255000, cfg.text_vocab), (1, N_TEXT), generator=g)
    prompt = torch.randint(0, cfg.audio_vocab, (1, N_PROMPT, 1), generator=g)
    y = torch.cat([prompt, torch.full((1, 1, 1), cfg.y_sep_token, dtype=torch.long)], dim=1)   # inference_tts_utils.py:229-242
    tgt = torch.tensor([y.shape[1] + TARGET_TOKENS])
    return x, torch.tensor([N_TEXT]), y, tgt


def engine_config():
    from t5gemma_tts_b200 import EngineConfig
    return EngineConfig(max_slots=1, max_text_len=128, max_dec_len=1024, max_prefill_tokens=1024)


def oracle_from_engine_weights(cfg, device):
    """fp32 CPU oracle holding exactly the (bf16-representable) weights the engine was given."""
    from oracle.t5gemma_voice_oracle import Oracle, OracleConfig
    from t5gemma_tts_b200.random_init import iter_random_state_dict
    sd = {k: v.float().cpu() for k, v in iter_random_state_dict(cfg, seed=0, device=device)}
    ocfg = OracleConfig(hidden=cfg.hidden, inter=cfg.inter, n_enc_layers=cfg.n_enc_layers, n_dec_layers=cfg.n_dec_layers,
                        n_heads=cfg.n_heads, n_kv_heads=cfg.n_kv_heads, head_dim=cfg.head_dim,
                        sliding_window=cfg.sliding_window, query_pre_attn_scalar=cfg.query_pre_attn_scalar,
                        attn_softcap=cfg.attn_softcap, text_vocab=cfg.text_vocab, audio_vocab=cfg.audio_vocab,
                        n_special=cfg.n_special)
    return Oracle(ocfg, sd)


def cpu_port_run(orc, x, y, tgt, n_tokens: int, cores: int):
    """Times the CPU port (oracle) on a bounded sample: prefill once (untimed), then n_tokens greedy decode
    steps of the hot loop (head -> argmax -> embed -> 26 layers).  Returns tokens/s and first-step logits."""
    torch.set_num_threads(cores)
    c = orc.cfg
    with torch.no_grad():
        mem = orc.encoder(x[0])
        cross = orc.cross_kv(mem)
        cache = [None] * c.n_dec_layers
        dec_ids = torch.cat([torch.tensor([c.empty_token]), y[0, :, 0]])
        est_total = int(tgt[0]) + 1
        hid = orc.decoder(orc.embed_audio(dec_ids), orc.decoder_prefill_positions(len(dec_ids), est_total), cache, cross)
        last = hid[-1:]
        cur = len(dec_ids)
        first_logits = None
        t0 = time.perf_counter()
        for i in range(n_tokens):
            logits = orc.head(last)[0]
            if first_logits is None:
                first_logits = logits.clone()
            tok = int(torch.argmax(logits))
            cur += 1
            p = orc.decoder_step_position(cur, est_total)
            last = orc.decoder(orc.embed_audio(torch.tensor([tok])), torch.tensor([p]), cache, cross)
        dt = time.perf_counter() - t0
    return n_tokens / dt, first_logits.numpy(), mem.numpy()


def time_dominant_kernel(eng, cfg, dev):
    """gate|up GEMV (the largest single kernel of the decode step) timed live with CUDA events on the launching
    stream; weights rotate over 7 buffers (595 MB > L2) so every launch is HBM-sourced."""
    import ctypes as C
    from t5gemma_tts_b200 import lib as L
    N, K = 2 * cfg.inter, cfg.hidden
    ws = [(torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16) for _ in range(7)]
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def run(n):
        for i in range(n):
            L.check(eng.lib, eng.lib.t5g_debug_gemv_gateup(eng._h, C.c_void_p(ws[i % 7].data_ptr()), st))
    run(14)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(56)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000.0 / 56
    del ws
    return N * K * 2, us


def bench_bs64(cfg_mod, dev):
    """BASELINE.json configs[2]: batched decode, 64 ragged requests (2-20 s targets) on one engine with 64 rows."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_batched as bb
    from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine
    from t5gemma_tts_b200.random_init import iter_random_state_dict
    cfg = EngineConfig(max_slots=64, max_text_len=128, max_dec_len=1536, max_prefill_tokens=8192)
    eng = T5GemmaVoiceEngine(cfg, device=dev)
    eng.load_state_dict(iter_random_state_dict(cfg, seed=0, device=dev))
    reqs = bb.make_requests(64, cfg, seed=0)
    eng.generate(reqs[:4], chunk_steps=8)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = eng.generate(reqs, chunk_steps=32)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    toks = sum(len(o) for o in outs)
    eng.prefill(reqs, list(range(64)))
    eng.decode(8)
    eng.poll()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.decode(32)
    e1.record()
    eng.poll()
    step_ms = e0.elapsed_time(e1) / 32
    eng.close()
    return {"workload": "configs[2]: 64 ragged requests (S~U[32,96], 50% with 150-token prompt, 2-20 s targets), 64 engine rows",
            "tokens": toks, "seconds": dt, "tokens_per_s_whole_job": toks / dt, "full_batch_step_ms": step_ms,
            "tokens_per_s_full_batch": 64 / step_ms * 1000.0}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cfg = engine_config()
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    orc = oracle_from_engine_weights(cfg, dev)
    x, x_lens, y, tgt = make_inputs(1234, cfg)
    n_tok = 8
    vals = []
    for i in range(args.warmup + args.steps):
        tps, _, _ = cpu_port_run(orc, x, y, tgt, n_tok, cores)
        if i >= args.warmup:
            vals.append(tps)
    v = float(np.mean(vals))
    out = {"impl": "reference", "metric": "audio_tokens_per_sec", "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * n_tok / v, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD},
           "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": cores, "kind": "port",
                            "sample": f"per step: {n_tok} greedy decode tokens after an untimed prefill "
                                      "(oracle/ fp32 torch port of models/t5gemma.py; the Python reference cannot travel)"},
           "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-tokens", type=int, default=24)
    ap.add_argument("--no-extra", action="store_true", help="skip the bs=64 (configs[2]) extra measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    assert args.warmup >= 3 or os.environ.get("T5G_BENCH_ALLOW_SHORT"), "timing rules: W >= 3"
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from t5gemma_tts_b200 import T5GemmaVoiceEngine
    from t5gemma_tts_b200.random_init import iter_random_state_dict
    cfg = engine_config()
    eng = T5GemmaVoiceEngine(cfg, device=dev)
    eng.load_state_dict(iter_random_state_dict(cfg, seed=0, device=dev))
    torch.manual_seed(1 + rank)                       # seed_everything(seed=1) semantics for the uniform draws

    # ---- inputs: pinned host tensors (e2e) ----
    reqs = [make_inputs(1234 + rank * 1000 + i, cfg) for i in range(args.warmup + args.steps)]
    reqs = [tuple(t.pin_memory() for t in r) for r in reqs]
    kw = dict(top_k=30, top_p=0.9, temperature=0.8)

    def run_one(r):
        x, xl, y, tgt = r
        xd, yd = x.to(dev, non_blocking=True), y.to(dev, non_blocking=True)      # H2D inside the timed region
        res, gen = eng.inference_tts(xd, xl, yd, tgt, prompt_frames=y.shape[1], chunk_steps=64, **kw)
        return int(gen.cpu().shape[-1])                                            # D2H of the result

    for r in reqs[: args.warmup]:
        run_one(r)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev0.record()
    n_tokens, dev_ms, decode_ms, prefill_ms, step_tokens = 0, 0.0, 0.0, 0.0, []
    for r in reqs[args.warmup:]:
        n = run_one(r)
        n_tokens += n
        step_tokens.append(n)
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    dev_total_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - l0
    clk = clocks.stop()

    # ---- device-resident measurement ("value"): same utterances, ids already on the GPU, timed with the
    # engine's own CUDA events (prefill + every decode call), host polling gaps excluded ----
    tm = []
    for r in reqs[args.warmup:]:
        x, xl, y, tgt = r
        xd, yd = x.to(dev), y.to(dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res, gen = eng.inference_tts(xd, xl, yd, tgt, prompt_frames=y.shape[1], chunk_steps=1024, **kw)
        e1.record()
        torch.cuda.synchronize()
        t = eng.timings()
        tm.append((gen.shape[-1], e0.elapsed_time(e1), t[0] + t[2], t[3]))
    toks_dev = sum(t[0] for t in tm)
    ms_dev = sum(t[1] for t in tm)
    prefill_ms = float(np.mean([t[2] for t in tm]))
    # decode-only step time: last decode call covers all remaining steps at chunk 1024
    ms_per_token = float(np.mean([t[3] / max(1, t[0]) for t in tm]))

    stats = torch.tensor([float(n_tokens), wall, float(toks_dev), ms_dev / 1000.0, float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        allst = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allst, stats)
        allst = torch.stack(allst).cpu().numpy()
    else:
        allst = stats.cpu().numpy()[None]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tot_tokens, max_wall = allst[:, 0].sum(), allst[:, 1].max()
    tot_tokens_dev, max_dev = allst[:, 2].sum(), allst[:, 3].max()
    peak, peak_src = load_peaks()
    w_bytes = eng.weight_bytes_per_step()
    kv_tok = eng.kv_bytes_per_token()
    mean_ctx = (N_PROMPT + 2) + np.mean(step_tokens) / 2.0
    alg_bytes = w_bytes + (mean_ctx + N_TEXT) * kv_tok
    achieved = alg_bytes / (ms_per_token * 1e-3) / 1e9
    out = {
        "metric": "audio_tokens_per_sec", "value": float(tot_tokens_dev / max_dev), "unit": "tokens/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(1000.0 * max_wall / args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "inputs larger than L2 (4.85 GB of weights streamed per token)",
                   "tokens_per_utterance": float(np.mean(step_tokens)), "parallelism": f"replicas x{world}"},
        "ms_per_token_bs1": ms_per_token, "prefill_ms": prefill_ms,
        "real_time_factor": float((tot_tokens / 50.0) / max_wall),
        "e2e": {"value": float(tot_tokens / max_wall), "unit": "tokens/s",
                "h2d_bytes_per_step": int((N_TEXT + N_PROMPT + 1) * 8),
                "d2h_bytes_per_step": int(np.mean(step_tokens) * 4 + (N_PROMPT + 1 + np.mean(step_tokens)) * 8)},
        "gpu_launches": int(allst[:, 4].sum()),
        "clocks": clk,
        "roofline": None,
        "step_roofline": {"bound": "hbm", "achieved": float(achieved), "peak": peak, "unit": "GB/s",
                          "frac": float(achieved / peak), "peak_source": peak_src,
                          "what": "whole decode step (186 kernel launches, 4 steps per CUDA-graph replay): (weights + KV bytes) / ms_per_token",
                          "algorithmic_bytes_per_step": float(alg_bytes)},
    }
    kbytes, kus = time_dominant_kernel(eng, cfg, dev)
    out["roofline"] = {"bound": "hbm", "achieved": float(kbytes / kus / 1e3), "peak": peak, "unit": "GB/s",
                       "frac": float(kbytes / kus / 1e3 / peak), "traffic": 85002240.0, "peak_source": peak_src,
                       "kernel": "gemv_kernel<1,P_RES_NORM,E_GEGLU> (gate|up projection, 26 launches per decode step, "
                                 "largest single kernel: 27 % of the step)",
                       "algorithmic_bytes_per_launch": float(kbytes), "us_per_launch": float(kus),
                       "traffic_source": "ncu --set full dram__bytes_read.sum+write.sum per launch (profiles/r1_decode_step_summary.md)",
                       "timing": "CUDA events on the launching stream, 56 back-to-back launches over 7 rotating weight buffers (595 MB > L2)"}
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        orc = oracle_from_engine_weights(cfg, dev)
        x, xl, y, tgt = reqs[0]
        tps, first_logits, mem = cpu_port_run(orc, x, y, tgt, args.cpu_tokens, cores)
        out["cpu_baseline"] = {"value": float(tps), "unit": "tokens/s", "cores": cores, "kind": "port",
                               "sample": f"{args.cpu_tokens} greedy decode tokens after an untimed prefill of the same "
                                         "utterance (oracle/ fp32 torch port; weights identical to the engine's)"}
        # full-size parity on the same weights: encoder states + first-step logits
        from t5gemma_tts_b200 import GenerationRequest
        rq = GenerationRequest(text_ids=x[0].numpy(), prompt_ids=y[0, :, 0].numpy(), target_total=int(tgt[0]),
                               prompt_frames=y.shape[1], top_k=1)
        eng.prefill([rq], [0])
        m = eng.read_memory(0, N_TEXT)
        eng.decode(1)
        eng.poll()
        lg = eng.read_logits(0)
        eos = cfg.stop_token
        lg[eos] = first_logits[eos] = 0.0
        out["parity_2b"] = {"encoder_states_rel_err": float(np.abs(m - mem).max() / np.abs(mem).max()),
                            "first_step_logits_rel_err": float(np.abs(lg - first_logits).max() / np.abs(first_logits).max()),
                            "argmax_equal": bool(int(np.argmax(lg)) == int(np.argmax(first_logits))),
                            "tolerance": 2e-2}
        eng.release(0)
    else:
        out["cpu_baseline"] = None
    if world == 1 and not args.no_extra:
        try:
            eng.close()
            out["extra_bs64"] = bench_bs64(cfg, dev)
        except Exception as ex:          # the extra must never take the headline line down
            out["extra_bs64"] = {"error": repr(ex)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
