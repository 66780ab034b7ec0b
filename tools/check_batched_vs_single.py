"""Full-size consistency: the same requests through the bs<=4 GEMV step and through the batched (tcgen05 GEMM + tile
attention) step of a 64-row engine, teacher-forced along the single-row engine's greedy tokens; reports the per-step
logits difference (both paths are bf16 approximations of the same fp32 reference, tolerance 2e-2 each)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine, GenerationRequest  # noqa: E402
from t5gemma_tts_b200.random_init import iter_random_state_dict  # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.default_rng(7)
specs = [(64, 150), (40, 0), (96, 75)]
reqs = []
for S, P in specs:
    prompt = np.concatenate([rng.integers(0, 65536, P), [65540]]) if P else np.zeros(0, np.int64)
    reqs.append(dict(text_ids=rng.integers(2, 255000, S), prompt_ids=prompt, target_total=len(prompt) + 400,
                     prompt_frames=len(prompt), top_k=1))

def run(max_slots, slots, forced=None):
    cfg = EngineConfig(max_slots=max_slots, max_text_len=128, max_dec_len=1024, max_prefill_tokens=2048)
    eng = T5GemmaVoiceEngine(cfg)
    eng.load_state_dict(iter_random_state_dict(cfg, seed=0, device="cuda"))
    rs = [GenerationRequest(**r, **({"forced_tokens": forced[i]} if forced is not None else {})) for i, r in enumerate(reqs)]
    logits = [[] for _ in rs]
    if max_slots == 1:
        toks = []
        for i, r in enumerate(rs):
            eng.prefill([r], [0])
            for _ in range(STEPS):
                eng.decode(1); eng.poll(); logits[i].append(eng.read_logits(0).copy())
            toks.append(eng.read_tokens(0)[:STEPS].copy()); eng.release(0)
    else:
        eng.prefill(rs, slots)
        for _ in range(STEPS):
            eng.decode(1); eng.poll()
            for i, s in enumerate(slots):
                logits[i].append(eng.read_logits(s).copy())
        toks = [eng.read_tokens(s)[:STEPS].copy() for s in slots]
    eng.close()
    return toks, logits

t1, l1 = run(1, [0])
t64, l64 = run(64, [5, 40, 63], forced=t1)
eos = 65539
worst, agree = 0.0, []
for i in range(len(reqs)):
    for a, b in zip(l1[i], l64[i]):
        a = a.copy(); b = b.copy(); a[eos] = b[eos] = 0.0
        worst = max(worst, float(np.abs(a - b).max() / np.abs(a).max()))
    am1 = np.array([np.argmax(np.where(np.arange(len(x)) == eos, -1e30, x)) for x in l1[i]])
    am64 = np.array([np.argmax(np.where(np.arange(len(x)) == eos, -1e30, x)) for x in l64[i]])
    agree.append(float((am1 == am64).mean()))
print(f"steps {STEPS} x {len(reqs)} requests: max rel logits diff single-row vs 64-row engine {worst:.3e}; argmax agreement {agree}")
