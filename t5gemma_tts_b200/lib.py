"""ctypes binding of libt5gtts.so (include/t5gtts.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libt5gtts.so")
T5G_ABI_VERSION = 1
T5G_MAX_LAYERS = 64
T5G_F32, T5G_BF16, T5G_F16 = 0, 1, 2


class T5GError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libt5gtts error {code}: {msg}")
        self.code = code


class T5GConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("hidden", C.c_int32), ("inter", C.c_int32), ("n_enc_layers", C.c_int32),
        ("n_dec_layers", C.c_int32), ("n_heads", C.c_int32), ("n_kv_heads", C.c_int32), ("head_dim", C.c_int32),
        ("sliding_window", C.c_int32), ("text_vocab", C.c_int32), ("n_audio_tokens", C.c_int32),
        ("eos_token", C.c_int32), ("encodec_sr", C.c_int32), ("text_guard_frames_per_token", C.c_int32),
        ("attn_scale", C.c_float), ("attn_softcap", C.c_float), ("rms_eps", C.c_float), ("rope_theta", C.c_float),
        ("progress_scale", C.c_float), ("extra_cutoff", C.c_float),
        ("enc_layer_sliding", C.c_uint8 * T5G_MAX_LAYERS), ("dec_layer_sliding", C.c_uint8 * T5G_MAX_LAYERS),
        ("max_slots", C.c_int32), ("max_text_len", C.c_int32), ("max_dec_len", C.c_int32),
        ("max_prefill_tokens", C.c_int32), ("kv_page_tokens", C.c_int32), ("reserved0", C.c_int32),
    ]


class T5GSampling(C.Structure):
    _fields_ = [("top_k", C.c_int32), ("top_p", C.c_float), ("min_p", C.c_float), ("temperature", C.c_float)]


class T5GRequest(C.Structure):
    _fields_ = [
        ("slot", C.c_int32), ("n_text", C.c_int32), ("text_ids", C.POINTER(C.c_int64)),
        ("n_dec", C.c_int32), ("dec_ids", C.POINTER(C.c_int64)),
        ("target_total", C.c_int32), ("prompt_frames", C.c_int32), ("max_new_tokens", C.c_int32),
        ("sampling", T5GSampling),
        ("top_k_schedule", C.POINTER(C.c_int32)), ("n_top_k_schedule", C.c_int32),
        ("uniforms", C.c_void_p), ("n_uniforms", C.c_int32),
        ("forced_tokens", C.POINTER(C.c_int32)), ("n_forced", C.c_int32),
        ("silence_tokens", C.POINTER(C.c_int32)), ("n_silence", C.c_int32), ("stop_repetition", C.c_int32),
    ]


class T5GSlotState(C.Structure):
    _fields_ = [("active", C.c_int32), ("finished", C.c_int32), ("n_generated", C.c_int32), ("cur_len", C.c_int32)]


class T5GSampleRow(C.Structure):
    _fields_ = [("sampling", T5GSampling), ("u", C.c_float), ("cur_num_gen", C.c_int32),
                ("current_length", C.c_int32), ("prompt_offset", C.c_int32), ("target_total", C.c_int32),
                ("n_text", C.c_int32), ("prev_token", C.c_int32), ("consec_silence_count", C.c_int32)]


# every symbol include/t5gtts.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "t5g_create": (C.c_int, [C.POINTER(T5GConfig), C.c_int, C.POINTER(_P)]),
    "t5g_destroy": (C.c_int, [_P]),
    "t5g_load_tensor": (C.c_int, [_P, C.c_char_p, _P, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_int]),
    "t5g_finalize_weights": (C.c_int, [_P]),
    "t5g_prefill": (C.c_int, [_P, C.POINTER(T5GRequest), C.c_int, _P]),
    "t5g_decode": (C.c_int, [_P, C.c_int, _P]),
    "t5g_poll": (C.c_int, [_P, C.POINTER(T5GSlotState), _P]),
    "t5g_read_tokens": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int), _P]),
    "t5g_read_picks": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int), _P]),
    "t5g_release_slot": (C.c_int, [_P, C.c_int]),
    "t5g_read_memory": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), _P]),
    "t5g_read_last_hidden": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), _P]),
    "t5g_read_logits": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), _P]),
    "t5g_prefill_logits": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), _P]),
    "t5g_sample": (C.c_int, [_P, _P, C.POINTER(T5GSampleRow), C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P]),
    "t5g_sample_set_silence": (C.c_int, [_P, C.POINTER(C.c_int32), C.c_int, C.c_int]),
    "t5g_launch_count": (C.c_int64, [_P]),
    "t5g_weight_bytes_per_step": (C.c_int64, [_P]),
    "t5g_kv_bytes_per_token": (C.c_int64, [_P]),
    "t5g_get_timings": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "t5g_get_counters": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "t5g_last_error": (C.c_char_p, []),
    "t5g_abi_version": (C.c_int, []),
    "t5g_debug_trace": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_int)]),
    "t5g_debug_gemm": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "t5g_debug_attn_prefill": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_float, _P, C.c_int, _P]),
    "t5g_debug_gemv_gateup": (C.c_int, [_P, _P, _P]),
    "t5g_debug_gemv": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P]),
}

_lib = None


def load_library(path: str | None = None):
    """Loads libt5gtts.so and types every exported symbol.  Raises if the library is absent:
    the product has no CPU path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(f"{p} not found: build it with `python -m t5gemma_tts_b200.build` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.t5g_abi_version() != T5G_ABI_VERSION:
        raise ImportError(f"libt5gtts ABI {lib.t5g_abi_version()} != binding {T5G_ABI_VERSION}")
    if path is None:
        _lib = lib
    return lib


def check(lib, rc: int):
    if rc != 0:
        raise T5GError(rc, (lib.t5g_last_error() or b"").decode("utf-8", "replace"))
