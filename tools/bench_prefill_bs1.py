"""TTFT of the bs=1 voice-clone config: 64-token text + 151-token prompt prefill (encoder, cross-KV, decoder prompt)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine, GenerationRequest  # noqa: E402
from t5gemma_tts_b200.random_init import iter_random_state_dict  # noqa: E402
cfg = EngineConfig(max_slots=1, max_text_len=128, max_dec_len=1536, max_prefill_tokens=1024)
eng = T5GemmaVoiceEngine(cfg)
eng.load_state_dict(iter_random_state_dict(cfg, seed=0, device="cuda"))
rng = np.random.default_rng(1)
prompt = np.concatenate([rng.integers(0, 65536, 150), [cfg.y_sep_token]])
rq = GenerationRequest(text_ids=rng.integers(2, 255000, 64), prompt_ids=prompt, target_total=651, prompt_frames=151,
                       top_k=30, top_p=0.9, temperature=0.8)
ts = []
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.prefill([rq], [0]); torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
    eng.release(0)
tm = eng.timings()
print(f"prefill wall ms: median {np.median(ts[2:]):.2f} (min {min(ts[2:]):.2f}); device encoder+cross {tm[0]:.2f} decoder {tm[2]:.2f}")
