"""Short driver for ncu: 2b-2b random-init engine, one prefill, then a few decode steps (bs=1, config[1] shape)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import engine_config, make_inputs_bs1 as make_inputs  # noqa: E402
from t5gemma_tts_b200 import T5GemmaVoiceEngine, GenerationRequest  # noqa: E402
from t5gemma_tts_b200.random_init import iter_random_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--ctx", type=int, default=150)
ap.add_argument("--warm", type=int, default=0, help="decode steps before the timed ones (longer context)")
a = ap.parse_args()
cfg = engine_config(1)
eng = T5GemmaVoiceEngine(cfg)
eng.load_state_dict(iter_random_state_dict(cfg, seed=0, device="cuda"))
x, xl, y, tgt = make_inputs(1234, cfg)
rq = GenerationRequest(text_ids=x[0].numpy(), prompt_ids=y[0, : a.ctx + 1, 0].numpy(), target_total=int(tgt[0]),
                       prompt_frames=a.ctx + 1, top_k=30, top_p=0.9, temperature=0.8)
eng.prefill([rq], [0])
if a.warm > 0:
    eng.decode(a.warm)
    eng.poll()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.decode(a.steps)
eng.poll()
e0.record()
eng.decode(a.steps)
e1.record()
eng.poll()
print("ms/step", e0.elapsed_time(e1) / a.steps, "launches", eng.launch_count())
if os.environ.get("T5G_TRACE") == "1":
    import ctypes as C
    import numpy as np
    from t5gemma_tts_b200 import lib as L
    eng.decode(1); eng.poll()
    b = np.zeros(1024, dtype=np.uint64); e_ = np.zeros(1024, dtype=np.uint64); n = C.c_int(0)
    L.check(eng.lib, eng.lib.t5g_debug_trace(eng._h, b.ctypes.data_as(C.POINTER(C.c_uint64)), e_.ctypes.data_as(C.POINTER(C.c_uint64)), 1024, C.byref(n)))
    n = n.value
    b, e_ = b[:n].astype(np.int64), e_[:n].astype(np.int64)
    t0 = b[0]
    per_layer = ["qkv", "sattn", "o", "qc", "cattn", "oc", "gu", "down"]
    if (n - 3) % 7 == 0 and (n - 3) % 8 != 0:
        if os.environ.get("T5G_FUSE_XATTN") == "1":
            per_layer = ["qkv", "sattn", "o", "qc", "xattn+oc", "gu", "down"]     # cross-attention fused into o_proj
        else:
            per_layer = ["qkv", "sattn", "o+qc", "cattn", "oc", "gu", "down"]     # o_proj + cross q_proj in one kernel
    names = ["head1", "head2", "sampler"] + per_layer * 26
    if n == 4:
        names = ["head1", "head2", "sampler", "layers(persistent)"]
    print("step span us", (e_.max() - t0) / 1000.0)
    agg = {}
    for i in range(n):
        nm = names[i] if i < len(names) else str(i)
        dur = (e_[i] - b[i]) / 1000.0                     # post-wait body duration
        gap = (b[i] - e_[i - 1]) / 1000.0 if i else 0.0   # previous kernel's exit -> this kernel past its wait
        a = agg.setdefault(nm, [0, 0.0, 0.0]); a[0] += 1; a[1] += dur; a[2] += gap
    for nm, (c, d, g) in agg.items():
        print(f"{nm:8s} n={c:3d} body avg {d/c:7.2f} us   gap avg {g/c:6.2f} us   total {(d+g):8.1f} us")
    sp = np.zeros(1024, dtype=np.uint64); se = np.zeros(1024, dtype=np.uint64); nn = C.c_int(0)
    L.check(eng.lib, eng.lib.t5g_debug_trace(eng._h, sp.ctypes.data_as(C.POINTER(C.c_uint64)), se.ctypes.data_as(C.POINTER(C.c_uint64)), 1024, C.byref(nn)))
    if n == 4:   # phase timestamps of CTA 0 for layer 5 of the persistent kernel
        pp = sp[900:918].astype(np.int64)
        pp = sp[900:916].astype(np.int64)
        lab = ["norm+qkv", "sattn", "load+o", "norm+qc", "cattn", "load+oc", "norm+gu", "act+down"]
        print("persistent kernel, layer 5, CTA 0 (us):", ", ".join(f"{lab[i]} {(pp[2 * i + 1] - pp[2 * i]) / 1000.0:.2f}" for i in range(8)),
              f"| 8 stages {(pp[15] - pp[0]) / 1000.0:.2f}")
        fp = sp[932:996].astype(np.int64)
        fp = fp[(fp > 0) & (fp < 2**62)]
        print("fine checkpoints (us since layer start):", " ".join(f"{(t - pp[0]) / 1000.0:.2f}" for t in fp))
    pr = sp[1000:1012].astype(np.int64)
    lab = ["post-wait", "slot staged", "eos edits", "logits pass", "argmax", "lower bound", "candidates", "top-k done", "drawn", "state updated", "exit"]
    print("sampler probes (us after post-wait):", ", ".join(f"{l} {(pr[i] - pr[0]) / 1000.0:.2f}" for i, l in enumerate(lab) if 0 < pr[i] < 2**62), f"| candidates {int(pr[11])}")
