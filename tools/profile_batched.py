"""Short driver for ncu: batched (bs=64) decode steps, 2b-2b random-init."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_batched import make_requests
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine
from t5gemma_tts_b200.random_init import iter_random_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = EngineConfig(max_slots=B, max_text_len=128, max_dec_len=1536, max_prefill_tokens=8192)
eng = T5GemmaVoiceEngine(cfg)
eng.load_state_dict(iter_random_state_dict(cfg, seed=0, device="cuda"))
reqs = make_requests(B, cfg)
eng.prefill(reqs, list(range(B)))
eng.decode(3); eng.poll()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.decode(3); e1.record(); eng.poll()
print("ms/step", e0.elapsed_time(e1) / 3)
