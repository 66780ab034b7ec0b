"""Short driver for ncu: batched (bs=64) decode steps, 2b-2b random-init."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_batched import make_requests
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine
from t5gemma_tts_b200.random_init import iter_random_state_dict
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = EngineConfig(max_slots=B, max_text_len=128, max_dec_len=1536, max_prefill_tokens=max(8192, 920 * B))
eng = T5GemmaVoiceEngine(cfg)
eng.load_state_dict(iter_random_state_dict(cfg, seed=0, device="cuda"))
reqs = make_requests(B, cfg)
if len(sys.argv) > 2 and sys.argv[2] == "spread":
    # contexts like the middle of a continuous-batching job: prompts U[0, 900) tokens, every row alive for the whole run
    import numpy as np
    from t5gemma_tts_b200 import GenerationRequest
    rng = np.random.default_rng(3)
    reqs = []
    for i in range(B):
        P = int(rng.integers(0, 900))
        reqs.append(GenerationRequest(text_ids=rng.integers(2, 255000, int(rng.integers(32, 129))), prompt_ids=rng.integers(0, cfg.audio_vocab, P),
                                      target_total=P + 300, prompt_frames=P, top_k=30, top_p=0.9, temperature=0.8))
eng.prefill(reqs, list(range(B)))
eng.decode(8); eng.poll()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.decode(24); e1.record(); eng.poll()
print("ms/step", e0.elapsed_time(e1) / 24)
if os.environ.get("T5G_TRACE") == "1":
    import ctypes as C
    import numpy as np
    from t5gemma_tts_b200 import lib as L
    eng.decode(1); eng.poll()
    b = np.zeros(1024, dtype=np.uint64); e_ = np.zeros(1024, dtype=np.uint64); n = C.c_int(0)
    L.check(eng.lib, eng.lib.t5g_debug_trace(eng._h, b.ctypes.data_as(C.POINTER(C.c_uint64)), e_.ctypes.data_as(C.POINTER(C.c_uint64)), 1024, C.byref(n)))
    names = ["hnorm", "fc1", "fc2", "sampler"] + ["norm1", "qkv", "sattn", "o", "norm2", "qc", "cattn", "oc", "norm3", "gu", "down"] * 26
    n = min(n.value, len(names))
    b, e_ = b[:n].astype(np.int64), e_[:n].astype(np.int64)
    print("step span us", (e_.max() - b[0]) / 1000.0)
    agg = {}
    for i in range(n):
        dur = (e_[i] - b[i]) / 1000.0
        gap = (b[i] - e_[i - 1]) / 1000.0 if i else 0.0
        a = agg.setdefault(names[i], [0, 0.0, 0.0]); a[0] += 1; a[1] += dur; a[2] += gap
    for nm, (c, d, g) in agg.items():
        print(f"{nm:8s} n={c:3d} body avg {d/c:7.2f} us   gap avg {g/c:6.2f} us   total {(d+g):8.1f} us")
    # in-kernel checkpoints of layer 5's self-attention (kv head 0, split 0 CTA of each row), relative to the kernel's begin
    bb = np.zeros(1024, dtype=np.uint64); ee = np.zeros(1024, dtype=np.uint64); nn = C.c_int(0)
    L.check(eng.lib, eng.lib.t5g_debug_trace(eng._h, bb.ctypes.data_as(C.POINTER(C.c_uint64)), ee.ctypes.data_as(C.POINTER(C.c_uint64)), 1024, C.byref(nn)))
    k_sattn5 = 4 + 5 * 11 + 2
    kb0 = int(bb[k_sattn5])
    pr = bb[300:300 + min(B, 64) * 11].astype(np.int64).reshape(-1, 11)
    lens = [int(r.n_dec) if hasattr(r, "n_dec") else -1 for r in reqs]
    print("probe: row ctx | resident, pre-wait done, post-wait, q in smem, q frags, tile0, tile1, tile2, loop end, merged, exit (us rel. kernel begin)")
    order = np.argsort(pr[:, 10])[::-1]
    for r in list(order[:6]) + list(order[-3:]):
        if pr[r, 0] > 0 and pr[r, 0] < 2**62:
            ctx = len(reqs[r].prompt_ids) if hasattr(reqs[r], "prompt_ids") else -1
            print(f"row {r:2d} ctx~{ctx + 8:4d} |", " ".join(f"{(v - kb0) / 1000.0:7.2f}" if 0 < v < 2**62 else "      -" for v in pr[r]))
    k_gu5 = 4 + 5 * 11 + 9
    g0 = int(bb[k_gu5]); gp = bb[1016:1024].astype(np.int64)
    lab = ["entry", "post-wait", "loads issued", "acc ready", "epilogue done", "mma: first stage full", "mma: half", "mma: all issued"]
    print("gate|up CTA probe (us rel. kernel begin):", ", ".join(f"{l} {(v - g0) / 1000.0:.2f}" for l, v in zip(lab, gp) if 0 < v < 2**62))
