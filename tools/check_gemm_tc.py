"""tcgen05 GEMMs vs torch fp64 reference + timing (run under `timeout`).  argv[1]: number of shapes, argv[2]: "epi" also
checks the bf16 / GeGLU epilogues on the prefill shapes."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine, lib as L  # noqa: E402

cfg = EngineConfig(hidden=64, inter=128, n_enc_layers=1, n_dec_layers=1, n_heads=4, n_kv_heads=2, head_dim=16,
                   text_vocab=32, audio_vocab=64, max_slots=1, max_text_len=16, max_dec_len=64, max_prefill_tokens=64)
eng = T5GemmaVoiceEngine(cfg)
shapes = [(16, 128, 64), (5, 100, 128), (64, 2304, 2304), (64, 18432, 2304), (64, 2304, 9216), (152, 4096, 2304),
          (1024, 2304, 2048), (8192, 2304, 2304), (8192, 4096, 2304), (8192, 2304, 2048), (8192, 18432, 2304), (8192, 2304, 9216),
          (300, 65664, 2304), (33, 4096, 2304), (2500, 1000, 576), (257, 18432, 64)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
EPI = {"f32": 0, "geglu": 1, "bf16": 4}


def run_one(M, N, K, epi, reps=10):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    if epi == "f32":
        out = torch.full((M, N), float("nan"), device="cuda")
    elif epi == "bf16":
        out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    else:
        out = torch.full((M, N // 2), float("nan"), device="cuda", dtype=torch.bfloat16)
    torch.cuda.synchronize()

    def run():
        L.check(eng.lib, eng.lib.t5g_debug_gemm(eng._h, C.c_void_p(a.data_ptr()), C.c_void_p(w.data_ptr()),
                                                C.c_void_p(out.data_ptr()), M, N, K, 1 | (EPI[epi] << 8), None))
    run()
    torch.cuda.synchronize()
    ref = (a.float() @ w.float().t()).double()
    if epi == "geglu":
        ref = torch.nn.functional.gelu(ref[:, 0::2], approximate="tanh") * ref[:, 1::2]
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / reps
    print(f"M={M} N={N} K={K} {epi}: rel err {err:.2e}  {us:.1f} us  {2*M*N*K/us/1e6:.1f} TFLOP/s  {N*K*2/us/1e3:.0f} GB/s(weights)", flush=True)
    return err


bad = 0
for (M, N, K) in shapes:
    bad += run_one(M, N, K, "f32") > 2e-5
    if len(sys.argv) > 2 and sys.argv[2] == "epi" and M > 128:
        bad += run_one(M, N, K, "bf16") > 6e-3
        if N % 2 == 0:
            bad += run_one(M, N, K, "geglu") > 8e-3
print("FAILED" if bad else "ok", flush=True)
sys.exit(1 if bad else 0)
