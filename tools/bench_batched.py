"""BASELINE.json configs[2]: batched decode bs=64 with ragged target durations 2-20 s on 1 x B200."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine, GenerationRequest  # noqa: E402
from t5gemma_tts_b200.random_init import iter_random_state_dict  # noqa: E402


def make_requests(n, cfg, seed=0, max_new=0):
    rng = np.random.default_rng(seed)
    reqs = []
    for i in range(n):
        S = int(rng.integers(32, 97))
        has_prompt = bool(rng.integers(0, 2))
        prompt = np.concatenate([rng.integers(0, cfg.audio_vocab, 150), [cfg.y_sep_token]]) if has_prompt else np.zeros(0, np.int64)
        secs = float(rng.uniform(2, 20))
        reqs.append(GenerationRequest(text_ids=rng.integers(2, 255000, S), prompt_ids=prompt,
                                      target_total=len(prompt) + int(50 * secs), prompt_frames=len(prompt),
                                      top_k=30, top_p=0.9, temperature=0.8, max_new_tokens=max_new))
    return reqs


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--slots", type=int, default=64)
    ap.add_argument("--requests", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=32)
    ap.add_argument("--max-new", type=int, default=0)
    a = ap.parse_args()
    cfg = EngineConfig(max_slots=a.slots, max_text_len=128, max_dec_len=1536, max_prefill_tokens=8192)
    eng = T5GemmaVoiceEngine(cfg)
    eng.load_state_dict(iter_random_state_dict(cfg, seed=0, device="cuda"))
    torch.manual_seed(0)
    reqs = make_requests(a.requests, cfg, max_new=a.max_new)
    eng.generate(reqs[:4], chunk_steps=8)          # warm-up (graph capture, lazy attributes)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = eng.generate(reqs, chunk_steps=a.chunk)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    toks = sum(len(o) for o in outs)
    # steady-state step time with all rows busy
    eng.prefill(reqs[: a.slots], list(range(a.slots)))
    eng.decode(8); eng.poll()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.decode(32); e1.record(); eng.poll()
    step_ms = e0.elapsed_time(e1) / 32
    print(json.dumps({"slots": a.slots, "requests": a.requests, "tokens": toks, "seconds": dt, "tokens_per_s": toks / dt,
                      "full_batch_step_ms": step_ms, "full_batch_tokens_per_s": a.slots / step_ms * 1000,
                      "prefill_ms": eng.timings()[0] + eng.timings()[2]}))
