"""BASELINE.json configs[4]: long-form -- 512-token text encoder prefill + 30 s target (1500 tokens) at bs=16."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine, GenerationRequest  # noqa: E402
from t5gemma_tts_b200.random_init import iter_random_state_dict  # noqa: E402

B, S, TGT = 16, 512, 1500
cfg = EngineConfig(max_slots=B, max_text_len=S, max_dec_len=TGT + 300, max_prefill_tokens=B * S)
eng = T5GemmaVoiceEngine(cfg)
eng.load_state_dict(iter_random_state_dict(cfg, seed=0, device="cuda"))
rng = np.random.default_rng(5)
reqs = [GenerationRequest(text_ids=rng.integers(2, 255000, S), prompt_ids=np.zeros(0, np.int64), target_total=TGT,
                          prompt_frames=0, top_k=30, top_p=0.9, temperature=0.8) for _ in range(B)]
torch.manual_seed(0)
eng.prefill(reqs, list(range(B)))          # warm-up (lazy attributes)
for s in range(B):
    eng.release(s)
torch.cuda.synchronize()
t0 = time.perf_counter()
eng.prefill(reqs, list(range(B)))
torch.cuda.synchronize()
t_prefill = time.perf_counter() - t0
tm = eng.timings()
eng.decode(8); eng.poll()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.decode(64); e1.record(); eng.poll()
step_ms = e0.elapsed_time(e1) / 64
enc_flops = 2 * 2024517888 * B * S + 4 * B * S * S * 2048 * 26
print(json.dumps({"workload": "configs[4]: bs=16, 512-token text, 30 s target", "encoder_prefill_ms": tm[0],
                  "decoder_prefill_ms": tm[2], "prefill_wall_ms": t_prefill * 1e3,
                  "encoder_tflops": enc_flops / (tm[0] * 1e-3) / 1e12,
                  "decode_step_ms_bs16_early_ctx": step_ms, "decode_tokens_per_s": B / step_ms * 1e3}))
