"""configs[4] prefill alone (16 x 512 text tokens), for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine, GenerationRequest  # noqa: E402
from t5gemma_tts_b200.random_init import iter_random_state_dict  # noqa: E402

B, S = 16, 512
cfg = EngineConfig(max_slots=B, max_text_len=S, max_dec_len=512, max_prefill_tokens=B * S)
eng = T5GemmaVoiceEngine(cfg)
eng.load_state_dict(iter_random_state_dict(cfg, seed=0, device="cuda"))
rng = np.random.default_rng(5)
reqs = [GenerationRequest(text_ids=rng.integers(2, 255000, S), prompt_ids=np.zeros(0, np.int64), target_total=100,
                          prompt_frames=0, top_k=30, top_p=0.9, temperature=0.8) for _ in range(B)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(n):
    eng.prefill(reqs, list(range(B)))
    torch.cuda.synchronize()
    print("timings", eng.timings())
    for s in range(B):
        eng.release(s)
