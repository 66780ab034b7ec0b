"""Kernel-level numerics against plain torch fp32 references of the same op."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests.gpu_util import engine_for
from t5gemma_tts_b200 import lib as L

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,N,K", [(1, 4096, 2304), (1, 2304, 9216), (2, 2304, 2048), (4, 1000, 2304), (3, 65, 64)])
def test_gemv_vs_torch(B, N, K):
    eng = engine_for("tinyA_eager")
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + N)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    x = torch.randn(B, K, device="cuda", generator=g)
    out = torch.empty(B, N, device="cuda")
    torch.cuda.synchronize()
    L.check(eng.lib, eng.lib.t5g_debug_gemv(eng._h, C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()),
                                            C.c_void_p(out.data_ptr()), B, N, K, None))
    torch.cuda.synchronize()
    ref = x.double() @ w.double().t()
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err        # fp32 accumulation of exact bf16*fp32 products


@pytest.mark.parametrize("M,N,K", [(64, 256, 128), (152, 4096, 2304), (7, 100, 64), (300, 2304, 9216)])
def test_gemm_simt_vs_torch(M, N, K):
    eng = engine_for("tinyA_eager")
    g = torch.Generator(device="cuda").manual_seed(M + N)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda")
    torch.cuda.synchronize()
    L.check(eng.lib, eng.lib.t5g_debug_gemm(eng._h, C.c_void_p(a.data_ptr()), C.c_void_p(w.data_ptr()),
                                            C.c_void_p(out.data_ptr()), M, N, K, 0, None))
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t()
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err


@pytest.mark.parametrize("M,N,K", [(16, 128, 64), (5, 100, 128), (64, 2304, 2304), (64, 18432, 2304), (100, 1000, 576), (128, 2304, 9216), (152, 4096, 2304),
                                   (1000, 2304, 9216), (33, 65664, 256)])
def test_gemm_tcgen05_vs_torch(M, N, K):
    """tcgen05/TMEM/TMA GEMM (incl. TMA out-of-bounds tiles and the split-K reduction) vs torch."""
    eng = engine_for("tinyA_eager")
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    out = torch.full((M, N), float("nan"), device="cuda")
    torch.cuda.synchronize()
    L.check(eng.lib, eng.lib.t5g_debug_gemm(eng._h, C.c_void_p(a.data_ptr()), C.c_void_p(w.data_ptr()),
                                            C.c_void_p(out.data_ptr()), M, N, K, 1, None))
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t()
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err


@pytest.mark.parametrize("M,N,K", [(300, 2304, 2304), (1000, 4608, 256), (257, 18432, 64), (8192, 2304, 512)])
@pytest.mark.parametrize("epi", ["f32", "bf16", "geglu"])
def test_gemm_cta_pair_epilogues_vs_torch(M, N, K, epi):
    """Prefill shapes (more than 128 tokens, tile grid larger than the machine) run on the persistent CTA-pair kernel
    (tcgen05.mma.cta_group::2, two TMEM accumulators): fp32, bf16 and fused GeGLU epilogues, ragged token / feature tails."""
    eng = engine_for("tinyA_eager")
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    code = {"f32": 0, "geglu": 1, "bf16": 4}[epi]
    out = torch.full((M, N // 2 if epi == "geglu" else N), float("nan"), device="cuda",
                     dtype=torch.float32 if epi == "f32" else torch.bfloat16)
    torch.cuda.synchronize()
    L.check(eng.lib, eng.lib.t5g_debug_gemm(eng._h, C.c_void_p(a.data_ptr()), C.c_void_p(w.data_ptr()),
                                            C.c_void_p(out.data_ptr()), M, N, K, 1 | (code << 8), None))
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t()
    if epi == "geglu":
        ref = torch.nn.functional.gelu(ref[:, 0::2], approximate="tanh") * ref[:, 1::2]
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < (2e-5 if epi == "f32" else 8e-3), err     # bf16 outputs: one rounding (2^-9) + tanh.approx in the GeGLU


ATTN_CASES = [
    # D, Hq, Hkv, q_lens, k_lens, causal, window, softcap
    (64, 4, 2, [5], [5], False, 0, 0.0),
    (64, 4, 2, [5, 5], [5, 5], False, 0, 0.0),                 # second request at an unaligned token offset
    (64, 4, 2, [70, 130], [70, 130], False, 0, 50.0),
    (256, 8, 4, [152, 33, 300], [152, 33, 300], True, 0, 50.0),
    (256, 8, 4, [200, 90], [200, 90], True, 40, 50.0),         # causal sliding window
    (256, 8, 4, [300], [300], False, 37, 0.0),                 # bidirectional window
    (256, 8, 4, [152, 20], [64, 96], False, 0, 50.0),          # cross-attention shapes
    (128, 4, 4, [257], [257], True, 0, 5.0),
    (256, 8, 4, [300, 200], [300, 200], True, 0, 0.0, 0, 8.0),   # wide score spread: running-max raises rescale O in TMEM
    (256, 8, 4, [300], [300], False, 0, 50.0, 0, 30.0),          # scores saturating the softcap
    (64, 4, 2, [200], [700], False, 0, 0.0, 1, 6.0),             # 11 key blocks, K ring wraps several times
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", ATTN_CASES, ids=[f"D{c[0]}-q{'_'.join(map(str, c[3]))}-c{int(c[5])}w{c[6]}" + ("-wide" if len(c) > 8 else "") for c in ATTN_CASES])
def test_prefill_attention_vs_torch(case):
    """Both prefill attention kernels (CUDA-core varlen and the tcgen05/TMA one) against an fp32 torch softmax(QK^T)V with
    the reference's masks (modeling_t5gemma_voice.py make_attention_mask / sliding window) and softcap."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from check_attn_tc import run
    res = run(*case)
    TOL = 2e-2     # bf16 inputs / bf16 probabilities, relative to max |out|
    assert res[0][0] < TOL, f"CUDA-core kernel err {res[0][0]}"
    assert res[1][0] < TOL, f"tcgen05 kernel err {res[1][0]}"
