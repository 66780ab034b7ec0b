"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, struct layouts match the header, config derivation, and the product path refuses to run
without CUDA (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import t5gemma_tts_b200 as pkg                      # noqa: E402
from t5gemma_tts_b200 import lib as L               # noqa: E402
from t5gemma_tts_b200.config import EngineConfig    # noqa: E402


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(L.LIB_PATH):
        from t5gemma_tts_b200 import build
        build.build()
    return L.load_library()


def test_library_exports_every_header_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "t5gtts.h")).read()
    declared = set(re.findall(r"\b(t5g_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(L.SYMBOLS), (declared ^ set(L.SYMBOLS))
    for name in declared:
        assert hasattr(built_lib, name), name
    assert built_lib.t5g_abi_version() == L.T5G_ABI_VERSION


def test_struct_layouts_match_header(tmp_path):
    src = tmp_path / "sz.cpp"
    src.write_text('#include "%s/include/t5gtts.h"\n#include <cstdio>\n#include <cstddef>\n'
                   'int main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n", sizeof(T5GConfig), sizeof(T5GRequest), '
                   'sizeof(T5GSampleRow), sizeof(T5GSlotState), offsetof(T5GRequest, uniforms), '
                   'offsetof(T5GRequest, forced_tokens), offsetof(T5GConfig, max_slots));}' % ROOT)
    exe = tmp_path / "sz"
    subprocess.check_call(["g++", str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(L.T5GConfig), ctypes.sizeof(L.T5GRequest), ctypes.sizeof(L.T5GSampleRow),
            ctypes.sizeof(L.T5GSlotState), L.T5GRequest.uniforms.offset, L.T5GRequest.forced_tokens.offset,
            L.T5GConfig.max_slots.offset]
    assert got == want


def test_last_error_and_null_handling(built_lib):
    rc = built_lib.t5g_finalize_weights(None)
    assert rc == -1
    assert b"null" in built_lib.t5g_last_error()


def test_engine_config_special_ids_and_layer_types():
    c = EngineConfig()
    assert (c.empty_token, c.eog, c.eos, c.y_sep_token) == (65536, 65537, 65539, 65540)   # config.py:224-228
    assert c.n_audio_tokens == 65541 and c.stop_token == 65539
    assert c.dec_layer_types[0] == "sliding_attention" and c.dec_layer_types[1] == "full_attention"
    c2 = EngineConfig(eos=0)            # models/t5gemma.py:861-863: eos<=0 -> stop on eog
    assert c2.stop_token == c2.eog


def test_engine_config_from_reference_fixture():
    from oracle import fixtures
    from types import SimpleNamespace
    _, _, meta = fixtures.load_model_fixture("tinyA_sdpa")
    ns = SimpleNamespace(**meta)
    c = EngineConfig.from_reference(ns, max_slots=2)
    assert c.attn_softcap is None and c.hidden == 64 and c.n_dec_layers == 3 and c.max_slots == 2
    _, _, meta = fixtures.load_model_fixture("tinyA_eager")
    assert EngineConfig.from_reference(SimpleNamespace(**meta)).attn_softcap == 50.0
    with pytest.raises(ValueError):
        EngineConfig.from_reference(SimpleNamespace(**dict(meta, n_codebooks=2)))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU path"):
        pkg.T5GemmaVoiceEngine(EngineConfig(hidden=64, inter=128, n_enc_layers=1, n_dec_layers=1, n_heads=4,
                                            n_kv_heads=2, head_dim=16, text_vocab=32, audio_vocab=16))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "t5gemma_tts_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
