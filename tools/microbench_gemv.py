"""GEMV micro-benchmark: back-to-back launches over distinct weight matrices (total > L2), CUDA events."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine, lib as L  # noqa: E402

cfg = EngineConfig(hidden=64, inter=128, n_enc_layers=1, n_dec_layers=1, n_heads=4, n_kv_heads=2, head_dim=16,
                   text_vocab=32, audio_vocab=64, max_slots=1, max_text_len=16, max_dec_len=64, max_prefill_tokens=64)
eng = T5GemmaVoiceEngine(cfg)
for (N, K) in [(4096, 2304), (2304, 2048), (18432, 2304), (2304, 9216), (65664, 2304)]:
    nbuf = max(2, int(600e6 // (N * K * 2)))
    ws = [(torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16) for _ in range(nbuf)]
    x = torch.randn(1, K, device="cuda")
    out = torch.empty(1, N, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    def run(n):
        for i in range(n):
            L.check(eng.lib, eng.lib.t5g_debug_gemv(eng._h, C.c_void_p(x.data_ptr()), C.c_void_p(ws[i % nbuf].data_ptr()),
                                                    C.c_void_p(out.data_ptr()), 1, N, K, C.c_void_p(st)))
    run(nbuf)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 4 * nbuf
    e0.record(); run(n); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / n
    print(f"N={N} K={K}: {us:.2f} us/launch  {N*K*2/us/1e3:.0f} GB/s  (nbuf {nbuf})")
