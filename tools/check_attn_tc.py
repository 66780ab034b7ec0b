"""tcgen05 prefill attention vs a torch fp32 reference (and vs the CUDA-core kernel); run under `timeout`."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine, lib as L  # noqa: E402


def ref_attention(q, k, v, q_off, k_off, Hq, Hkv, D, causal, window, scale, softcap):
    out = torch.zeros_like(q, dtype=torch.float32)
    G = Hq // Hkv
    for s in range(len(q_off) - 1):
        qs = q[q_off[s]:q_off[s + 1]].float().view(-1, Hq, D).transpose(0, 1)        # [Hq, Lq, D]
        ks = k[k_off[s]:k_off[s + 1]].float().view(-1, Hkv, D).transpose(0, 1).repeat_interleave(G, 0)
        vs = v[k_off[s]:k_off[s + 1]].float().view(-1, Hkv, D).transpose(0, 1).repeat_interleave(G, 0)
        sc = qs @ ks.transpose(1, 2) * scale
        if softcap > 0:
            sc = torch.tanh(sc / softcap) * softcap
        Lq, Lk = qs.shape[1], ks.shape[1]
        qi = torch.arange(Lq, device=q.device)[:, None]
        ki = torch.arange(Lk, device=q.device)[None, :]
        ok = torch.ones(Lq, Lk, dtype=torch.bool, device=q.device)
        if causal:
            ok &= ki <= qi
            if window > 0:
                ok &= ki > qi - window
        elif window > 0:
            ok &= (qi - ki).abs() <= window
        sc = sc.masked_fill(~ok[None], float("-inf"))
        pr = torch.softmax(sc, -1)
        o = (pr @ vs).transpose(0, 1).reshape(Lq, Hq * D)
        out[q_off[s]:q_off[s + 1]] = o
    return out


def run(D, Hq, Hkv, q_lens, k_lens, causal, window, softcap, seed=0, qscale=1.0):
    cfg = EngineConfig(hidden=256, inter=512, n_enc_layers=1, n_dec_layers=1, n_heads=Hq, n_kv_heads=Hkv, head_dim=D,
                       query_pre_attn_scalar=float(D), text_vocab=64, audio_vocab=64, max_slots=1, max_text_len=64,
                       max_dec_len=64, max_prefill_tokens=max(sum(q_lens), sum(k_lens)) + 64)
    eng = T5GemmaVoiceEngine(cfg)
    g = torch.Generator(device="cuda").manual_seed(seed)
    Tq, Tk = sum(q_lens), sum(k_lens)
    q = (torch.randn(Tq, Hq * D, device="cuda", generator=g) * qscale).to(torch.bfloat16)   # qscale > 1: score spread that forces running-max raises
    k = torch.randn(Tk, Hkv * D, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(Tk, Hkv * D, device="cuda", generator=g).to(torch.bfloat16)
    q_off = np.concatenate([[0], np.cumsum(q_lens)]).astype(np.int32)
    k_off = np.concatenate([[0], np.cumsum(k_lens)]).astype(np.int32)
    seg_of = np.concatenate([np.full(n, i, dtype=np.int32) for i, n in enumerate(q_lens)])
    dq, dk, ds = torch.from_numpy(q_off).cuda(), torch.from_numpy(k_off).cuda(), torch.from_numpy(seg_of).cuda()
    ref = ref_attention(q, k, v, q_off, k_off, Hq, Hkv, D, causal, window, D ** -0.5, softcap)
    res = {}
    for impl in (0, 1):
        out = torch.full((Tq, Hq * D), float("nan"), device="cuda", dtype=torch.bfloat16)
        torch.cuda.synchronize()
        def call():
            L.check(eng.lib, eng.lib.t5g_debug_attn_prefill(
                eng._h, C.c_void_p(q.data_ptr()), C.c_void_p(k.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(dq.data_ptr()),
                C.c_void_p(dk.data_ptr()), C.c_void_p(ds.data_ptr()), len(q_lens), Tq, Tk, max(q_lens), int(causal), window,
                float(softcap), C.c_void_p(out.data_ptr()), impl, None))
        call()
        torch.cuda.synchronize()
        err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            call()
        e1.record()
        torch.cuda.synchronize()
        res[impl] = (err, e0.elapsed_time(e1) * 200)
    eng.close()
    print(f"D={D} Hq={Hq}/{Hkv} q={q_lens[:3]}.. k={k_lens[:3]}.. causal={causal} win={window} cap={softcap}: "
          f"simt err {res[0][0]:.2e} {res[0][1]:.0f} us | tcgen05 err {res[1][0]:.2e} {res[1][1]:.0f} us", flush=True)
    return res


CASES = {
    "a": lambda: run(256, 8, 4, [5], [5], False, 0, 0.0),
    "b": lambda: run(64, 4, 2, [70, 130], [70, 130], False, 0, 50.0),
    "c": lambda: run(128, 4, 2, [5], [5], False, 0, 0.0),
    "d": lambda: run(256, 8, 4, [5], [5], False, 0, 50.0),
    "e": lambda: run(256, 2, 1, [5], [5], False, 0, 0.0),
    "f": lambda: run(64, 4, 2, [70], [70], False, 0, 0.0),
    "g": lambda: run(64, 4, 2, [130], [60], False, 0, 0.0),
    "h": lambda: run(64, 4, 2, [5, 5], [5, 5], False, 0, 0.0),
    "i": lambda: run(64, 4, 2, [64], [64], False, 0, 0.0),
    "j": lambda: run(64, 4, 2, [5], [65], False, 0, 0.0),
}

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in CASES:
        CASES[sys.argv[1]]()
        sys.exit(0)
    small = len(sys.argv) > 1 and sys.argv[1] == "small"
    run(64, 4, 2, [5], [5], False, 0, 0.0)
    run(256, 8, 4, [70, 130], [70, 130], False, 0, 50.0)
    if not small:
        run(256, 8, 4, [152, 33, 300], [152, 33, 300], True, 0, 50.0)
        run(256, 8, 4, [200, 90], [200, 90], True, 40, 50.0)
        run(256, 8, 4, [300], [300], False, 37, 0.0)
        run(256, 8, 4, [152, 20], [64, 96], False, 0, 50.0)          # cross attention shapes
        run(128, 4, 4, [257], [257], True, 0, 5.0)
        run(256, 8, 4, [300, 200], [300, 200], True, 0, 0.0, 0, 8.0)     # wide score spread: O rescale path
        run(256, 8, 4, [300], [300], False, 0, 50.0, 0, 30.0)
        run(256, 8, 4, [512] * 16, [512] * 16, False, 4096, 50.0)    # configs[4] encoder layer
