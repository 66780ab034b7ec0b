"""Driver for ncu launch lists / DRAM traffic of ONE batched decode step at bench-like contexts: 64 rows, 250-token prompts
(context 250 + text U[32,128]), prefill, two warm-up steps, one captured step (graphs off so every kernel is a launch)."""
import os
import sys

import numpy as np
import torch

os.environ.setdefault("T5G_GRAPH", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine, GenerationRequest  # noqa: E402
from t5gemma_tts_b200.random_init import iter_random_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
P = int(sys.argv[2]) if len(sys.argv) > 2 else 250
cfg = EngineConfig(max_slots=B, max_text_len=128, max_dec_len=1024, max_prefill_tokens=max(8192, (P + 8) * B))
eng = T5GemmaVoiceEngine(cfg)
eng.load_state_dict(iter_random_state_dict(cfg, seed=0, device="cuda"))
rng = np.random.default_rng(3)
reqs = [GenerationRequest(text_ids=rng.integers(2, 255000, int(rng.integers(32, 129))), prompt_ids=rng.integers(0, cfg.audio_vocab, P),
                          target_total=P + 300, prompt_frames=P, top_k=30, top_p=0.9, temperature=0.8) for _ in range(B)]
eng.prefill(reqs, list(range(B)))
eng.decode(2); eng.poll()
torch.cuda.synchronize()
l0 = eng.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.decode(1); e1.record(); eng.poll()
print("step ms (graphs off)", e0.elapsed_time(e1), "launches in the step", eng.launch_count() - l0)
