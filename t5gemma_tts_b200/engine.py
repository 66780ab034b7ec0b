"""T5GemmaVoiceEngine -- drop-in for the reference model object on the token-generation path.

Mirrors (same names, argument meaning, return shapes/dtypes, error behaviour):
  * T5GemmaVoiceModel.inference_tts                       /root/reference/models/t5gemma.py:835-1129
  * T5GemmaVoiceForConditionalGeneration.inference_tts    /root/reference/hf_export/modeling_t5gemma_voice.py:565-862
and adds `inference_tts_batch` (the reference asserts batch_size == 1, models/t5gemma.py:865).

All arithmetic runs in libt5gtts.so (hand-written sm_100a CUDA behind include/t5gtts.h).  PyTorch is
used for device memory of the caller-visible tensors, the RNG stream (`torch.rand` from the global
generator, so `seed_everything(seed)` keeps its meaning) and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import logging
import math
from collections import deque
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import lib as L
from .config import EngineConfig

_DTYPES = {torch.float32: L.T5G_F32, torch.bfloat16: L.T5G_BF16, torch.float16: L.T5G_F16}


@dataclass
class GenerationRequest:
    """One utterance for inference_tts_batch (fields = the per-call arguments of inference_tts)."""
    text_ids: Sequence[int]                       # x[0, :x_lens[0]]
    prompt_ids: Sequence[int]                     # y[0, :, 0] (may be empty), before the special_first shift
    target_total: Optional[int]                   # tgt_y_lens[0]; None = tgt_y_lens None (no time budget, models/t5gemma.py:896-933)
    prompt_frames: Optional[int] = None           # kwargs["prompt_frames"], default len(prompt_ids)
    top_k: Union[int, List[int]] = -100
    top_p: float = 1.0
    min_p: float = 0.0
    temperature: float = 1.0
    max_new_tokens: int = 0                       # 0 = reference stop rules only
    uniforms: Optional[torch.Tensor] = None       # explicit U[0,1) draws (CUDA fp32); default torch.rand
    forced_tokens: Optional[Sequence[int]] = None # teacher forcing (parity tests)
    stop_repetition: int = 3                      # inference_tts(stop_repetition=...)
    silence_tokens: Optional[Sequence[int]] = None   # inference_tts(silence_tokens=...), [] in every shipped caller


class T5GemmaVoiceEngine:
    def __init__(self, config: EngineConfig, ref_config: Any = None, device: Union[str, torch.device] = "cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("T5GemmaVoiceEngine needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = L.load_library()
        self.cfg = config
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        # reference-facing attributes (inference_tts_utils.py:163-172 reads model_args.*; CLI reads model.config)
        self.config = ref_config if ref_config is not None else SimpleNamespace(**self._args_dict())
        self.args = self.config
        if ref_config is not None and not all(hasattr(ref_config, k) for k in ("y_sep_token", "eos", "empty_token")):
            # a partial config (e.g. a dict of the fields EngineConfig needs): the front door still finds every
            # model_args field it reads, derived from the engine's own special-token layout
            merged = self._args_dict()
            merged.update({k: v for k, v in vars(ref_config).items() if not k.startswith("_")} if hasattr(ref_config, "__dict__") else {})
            self.args = SimpleNamespace(**merged)
        self.training = False
        self._h = C.c_void_p()
        c = L.T5GConfig()
        c.abi_version = L.T5G_ABI_VERSION
        c.hidden, c.inter = config.hidden, config.inter
        c.n_enc_layers, c.n_dec_layers = config.n_enc_layers, config.n_dec_layers
        c.n_heads, c.n_kv_heads, c.head_dim = config.n_heads, config.n_kv_heads, config.head_dim
        c.sliding_window, c.text_vocab, c.n_audio_tokens = config.sliding_window, config.text_vocab, config.n_audio_tokens
        c.eos_token, c.encodec_sr = config.stop_token, int(config.encodec_sr)
        c.text_guard_frames_per_token = config.text_guard_frames_per_token
        c.attn_scale = float(config.query_pre_attn_scalar) ** -0.5
        c.attn_softcap = float(config.attn_softcap) if config.attn_softcap else 0.0
        c.rms_eps, c.rope_theta = config.rms_eps, config.rope_theta
        c.progress_scale, c.extra_cutoff = config.progress_scale, config.extra_cutoff
        for i, t in enumerate(config.enc_layer_types):
            c.enc_layer_sliding[i] = 1 if t == "sliding_attention" else 0
        for i, t in enumerate(config.dec_layer_types):
            c.dec_layer_sliding[i] = 1 if t == "sliding_attention" else 0
        c.max_slots, c.max_text_len, c.max_dec_len = config.max_slots, config.max_text_len, config.max_dec_len
        c.max_prefill_tokens, c.kv_page_tokens = config.max_prefill_tokens, config.kv_page_tokens
        rc = self.lib.t5g_create(C.byref(c), self.device.index, C.byref(self._h))
        if rc != 0:
            msg = (self.lib.t5g_last_error() or b"").decode()
            if self._h:
                self.lib.t5g_destroy(self._h)
                self._h = C.c_void_p()
            raise L.T5GError(rc, msg)
        self._keep: Dict[int, Any] = {}          # per-slot objects borrowed by the engine (uniforms)

    # ------------------------------------------------------------------ construction helpers
    def _args_dict(self) -> Dict[str, Any]:
        c = self.cfg
        return dict(n_codebooks=1, audio_vocab_size=c.audio_vocab, n_special=c.n_special, empty_token=c.empty_token,
                    eog=c.eog, eos=c.eos, audio_pad_token=c.audio_vocab + 2, y_sep_token=c.y_sep_token,
                    x_sep_token=255999, special_first=c.special_first, encodec_sr=c.encodec_sr,
                    progress_scale=c.progress_scale, extra_cutoff=c.extra_cutoff, use_pm_rope=1,
                    text_input_type="text", add_eos_to_text=0, add_bos_to_text=0, parallel_pattern=0,
                    audio_max_length=40.0, text_guard_frames_per_token=c.text_guard_frames_per_token)

    @classmethod
    def from_reference(cls, module: Any, device="cuda:0", **sizing) -> "T5GemmaVoiceEngine":
        """Builds the engine from a live reference module (either twin) and copies its weights."""
        ref_cfg = getattr(module, "config", None)
        if ref_cfg is None or getattr(ref_cfg, "t5_config_dict", None) is None:
            raise ValueError("module.config must be a T5GemmaVoiceConfig carrying t5_config_dict")
        eng = cls(EngineConfig.from_reference(ref_cfg, **sizing), ref_config=ref_cfg, device=device)
        eng.load_state_dict(module.state_dict())
        return eng

    @classmethod
    def from_state_dict(cls, ref_config: Any, state_dict: Dict[str, torch.Tensor], device="cuda:0", **sizing):
        eng = cls(EngineConfig.from_reference(ref_config, **sizing), ref_config=ref_config, device=device)
        eng.load_state_dict(state_dict)
        return eng

    @classmethod
    def from_pretrained(cls, model_dir: str, device="cuda:0", **sizing) -> "T5GemmaVoiceEngine":
        """HF-format directory (config.json + safetensors) as written by scripts/export_t5gemma_voice_hf.py;
        replaces AutoModelForSeq2SeqLM.from_pretrained(dir, trust_remote_code=True) (inference_commandline_hf.py:102-107)."""
        from . import checkpoint as ck
        ref_cfg = ck.load_hf_config(model_dir)
        eng = cls(EngineConfig.from_reference(ref_cfg, **sizing), ref_config=ref_cfg, device=device)
        eng.load_state_dict(ck.iter_hf_tensors(model_dir, device="cpu"))
        return eng

    @classmethod
    def from_pth(cls, path: str, device="cuda:0", t5_config_dict=None, **sizing) -> "T5GemmaVoiceEngine":
        """Training bundle {model, args} as loaded by inference_commandline.py:121-156."""
        from . import checkpoint as ck
        ref_cfg, sd = ck.load_pth_bundle(path, t5_config_dict)
        return cls.from_state_dict(ref_cfg, sd, device=device, **sizing)

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        """Reference state_dict keys.  Tensors may live on CPU or on this GPU."""
        items = state_dict.items() if hasattr(state_dict, "items") else state_dict   # dict or (name, tensor) iterable
        for name, t in items:
            if not torch.is_tensor(t) or not torch.is_floating_point(t):
                continue
            if t.dtype not in _DTYPES:
                t = t.float()
            t = t.detach().contiguous()
            on_dev = 1 if t.is_cuda else 0
            if t.is_cuda and t.device != self.device:
                t = t.to(self.device)
            shape = (C.c_int64 * t.dim())(*t.shape)
            if t.is_cuda:
                torch.cuda.current_stream(self.device).synchronize()
            L.check(self.lib, self.lib.t5g_load_tensor(self._h, name.encode(), C.c_void_p(t.data_ptr()), _DTYPES[t.dtype],
                                                       t.dim(), shape, on_dev))
        if strict:
            L.check(self.lib, self.lib.t5g_finalize_weights(self._h))
        return self

    def finalize(self):
        L.check(self.lib, self.lib.t5g_finalize_weights(self._h))

    # ------------------------------------------------------------------ nn.Module-ish surface
    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def half(self):
        return self

    def bfloat16(self):
        return self

    def parameters(self):
        return iter(())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.t5g_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ low-level calls
    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def max_new_tokens_of(self, req: GenerationRequest, n_dec: int) -> int:
        """Upper bound on generated tokens implied by the time-budget stop rule (models/t5gemma.py:1042-1046)."""
        pf = len(req.prompt_ids) if req.prompt_frames is None else req.prompt_frames
        if req.target_total is None or req.target_total < 0:
            # the reference has no bound without a target; the slot's KV capacity is the engine's
            n = max(1, self.cfg.max_dec_len - n_dec)
        else:
            lim = math.floor(req.target_total - (pf + 1) + int(self.cfg.encodec_sr) * self.cfg.extra_cutoff)
            n = max(1, lim + 2)
        if req.max_new_tokens > 0:
            n = min(n, req.max_new_tokens)
        return n

    def prefill(self, reqs: Sequence[GenerationRequest], slots: Sequence[int]):
        c = self.cfg
        arr = (L.T5GRequest * len(reqs))()
        keep = []
        for i, (r, slot) in enumerate(zip(reqs, slots)):
            text = np.ascontiguousarray(np.asarray(r.text_ids, dtype=np.int64))
            prompt = np.asarray(r.prompt_ids, dtype=np.int64).reshape(-1)
            if c.special_first:
                prompt = prompt + c.n_special
            dec = np.ascontiguousarray(np.concatenate([np.array([c.empty_token], dtype=np.int64), prompt]))
            pf = len(prompt) if r.prompt_frames is None else int(r.prompt_frames)
            max_new = self.max_new_tokens_of(r, len(dec))
            u = r.uniforms
            if u is None:
                u = torch.rand(max_new, device=self.device, dtype=torch.float32)
            else:
                u = u.to(device=self.device, dtype=torch.float32).contiguous()
            q = arr[i]
            q.slot, q.n_text, q.n_dec = slot, len(text), len(dec)
            q.text_ids = text.ctypes.data_as(C.POINTER(C.c_int64))
            q.dec_ids = dec.ctypes.data_as(C.POINTER(C.c_int64))
            q.target_total = -1 if r.target_total is None or r.target_total < 0 else int(r.target_total)
            q.prompt_frames, q.max_new_tokens = pf, int(r.max_new_tokens)
            sched = None
            if isinstance(r.top_k, (list, tuple)):
                sched = np.ascontiguousarray(np.asarray(r.top_k, dtype=np.int32))
                q.top_k_schedule = sched.ctypes.data_as(C.POINTER(C.c_int32))
                q.n_top_k_schedule = len(sched)
                q.sampling.top_k = int(sched[0])
            else:
                q.top_k_schedule = None
                q.n_top_k_schedule = 0
                q.sampling.top_k = int(r.top_k)
            q.sampling.top_p, q.sampling.min_p, q.sampling.temperature = float(r.top_p), float(r.min_p), float(r.temperature)
            q.uniforms = C.c_void_p(u.data_ptr())
            q.n_uniforms = u.numel()
            forced = None
            if r.forced_tokens is not None and len(r.forced_tokens) > 0:
                forced = np.ascontiguousarray(np.asarray(r.forced_tokens, dtype=np.int32))
                q.forced_tokens = forced.ctypes.data_as(C.POINTER(C.c_int32))
                q.n_forced = len(forced)
            else:
                q.forced_tokens = None
                q.n_forced = 0
            sil = None
            if r.silence_tokens is not None and len(r.silence_tokens) > 0:
                sil = np.ascontiguousarray(np.asarray(r.silence_tokens, dtype=np.int32))
                q.silence_tokens = sil.ctypes.data_as(C.POINTER(C.c_int32))
                q.n_silence = len(sil)
            else:
                q.silence_tokens = None
                q.n_silence = 0
            q.stop_repetition = int(r.stop_repetition)
            keep.append((text, dec, sched, forced, sil))
            self._keep[slot] = u
        torch.cuda.current_stream(self.device).synchronize()     # uniforms were produced on torch's stream
        L.check(self.lib, self.lib.t5g_prefill(self._h, arr, len(reqs), self._stream()))

    def decode(self, steps: int):
        L.check(self.lib, self.lib.t5g_decode(self._h, int(steps), self._stream()))

    def poll(self) -> List[L.T5GSlotState]:
        st = (L.T5GSlotState * self.cfg.max_slots)()
        L.check(self.lib, self.lib.t5g_poll(self._h, st, self._stream()))
        return list(st)

    def read_tokens(self, slot: int) -> np.ndarray:
        buf = np.zeros(self.cfg.max_dec_len, dtype=np.int32)
        n = C.c_int(0)
        L.check(self.lib, self.lib.t5g_read_tokens(self._h, slot, buf.ctypes.data_as(C.POINTER(C.c_int32)), len(buf),
                                                   C.byref(n), self._stream()))
        return buf[: n.value].copy()

    def read_picks(self, slot: int) -> np.ndarray:
        buf = np.zeros(self.cfg.max_dec_len, dtype=np.int32)
        n = C.c_int(0)
        L.check(self.lib, self.lib.t5g_read_picks(self._h, slot, buf.ctypes.data_as(C.POINTER(C.c_int32)), len(buf),
                                                  C.byref(n), self._stream()))
        return buf[: n.value].copy()

    def release(self, slot: int):
        L.check(self.lib, self.lib.t5g_release_slot(self._h, slot))
        self._keep.pop(slot, None)

    def read_memory(self, slot: int, n_text: int) -> np.ndarray:
        out = np.zeros((n_text, self.cfg.hidden), dtype=np.float32)
        L.check(self.lib, self.lib.t5g_read_memory(self._h, slot, out.ctypes.data_as(C.POINTER(C.c_float)), self._stream()))
        return out

    def read_last_hidden(self, slot: int) -> np.ndarray:
        out = np.zeros(self.cfg.hidden, dtype=np.float32)
        L.check(self.lib, self.lib.t5g_read_last_hidden(self._h, slot, out.ctypes.data_as(C.POINTER(C.c_float)), self._stream()))
        return out

    def read_logits(self, slot: int) -> np.ndarray:
        out = np.zeros(self.cfg.n_audio_tokens, dtype=np.float32)
        L.check(self.lib, self.lib.t5g_read_logits(self._h, slot, out.ctypes.data_as(C.POINTER(C.c_float)), self._stream()))
        return out

    def prefill_logits(self, slot: int, n_dec: int) -> np.ndarray:
        out = np.zeros((n_dec, self.cfg.n_audio_tokens), dtype=np.float32)
        L.check(self.lib, self.lib.t5g_prefill_logits(self._h, slot, out.ctypes.data_as(C.POINTER(C.c_float)), self._stream()))
        return out

    def set_sample_silence(self, silence_tokens: Sequence[int], stop_repetition: int = 3):
        arr = np.ascontiguousarray(np.asarray(list(silence_tokens), dtype=np.int32))
        L.check(self.lib, self.lib.t5g_sample_set_silence(self._h, arr.ctypes.data_as(C.POINTER(C.c_int32)), len(arr), int(stop_repetition)))

    def sample(self, logits: torch.Tensor, rows: Sequence[dict]) -> Tuple[np.ndarray, np.ndarray]:
        """Standalone sampler S on CUDA fp32 logits [n, n_audio_tokens] (edited in place)."""
        assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous()
        n = logits.shape[0]
        assert logits.shape[1] == self.cfg.n_audio_tokens and len(rows) == n
        arr = (L.T5GSampleRow * n)()
        for i, r in enumerate(rows):
            a = arr[i]
            a.sampling.top_k, a.sampling.top_p = int(r.get("top_k", -100)), float(r.get("top_p", 1.0))
            a.sampling.min_p, a.sampling.temperature = float(r.get("min_p", 0.0)), float(r.get("temperature", 1.0))
            a.u = float(r.get("u", 0.5))
            a.cur_num_gen, a.current_length = int(r["cur_num_gen"]), int(r["current_length"])
            a.prompt_offset, a.target_total, a.n_text = int(r["prompt_offset"]), int(r["target_total"]), int(r.get("n_text", 1))
            a.prev_token, a.consec_silence_count = int(r.get("prev_token", -1)), int(r.get("consec_silence_count", 0))
        tok = np.zeros(n, dtype=np.int32)
        amax = np.zeros(n, dtype=np.int32)
        torch.cuda.current_stream(self.device).synchronize()
        L.check(self.lib, self.lib.t5g_sample(self._h, C.c_void_p(logits.data_ptr()), arr, n,
                                              tok.ctypes.data_as(C.POINTER(C.c_int32)),
                                              amax.ctypes.data_as(C.POINTER(C.c_int32)), self._stream()))
        return tok, amax

    def launch_count(self) -> int:
        return int(self.lib.t5g_launch_count(self._h))

    def weight_bytes_per_step(self) -> int:
        return int(self.lib.t5g_weight_bytes_per_step(self._h))

    def kv_bytes_per_token(self) -> int:
        return int(self.lib.t5g_kv_bytes_per_token(self._h))

    def timings(self) -> List[float]:
        out = (C.c_float * 4)()
        L.check(self.lib, self.lib.t5g_get_timings(self._h, out))
        return list(out)

    def counters(self) -> Dict[str, float]:
        """Running totals since creation: CUDA-event ms of all prefill / decode calls, decode steps, launches."""
        out = (C.c_double * 8)()
        L.check(self.lib, self.lib.t5g_get_counters(self._h, out))
        return dict(prefill_ms=out[0], decode_ms=out[1], decode_steps=int(out[2]), prefill_calls=int(out[3]),
                    launches=int(out[4]), kernels_per_step=int(out[5]))

    # ------------------------------------------------------------------ generation
    def generate(self, requests: Sequence[GenerationRequest], chunk_steps: int = 32) -> List[np.ndarray]:
        """Continuous batching over max_slots rows.  Returns the generated ids (incl. final eos) per request."""
        results: List[Optional[np.ndarray]] = [None] * len(requests)
        for idx, toks, done in self.generate_stream(requests, chunk_steps=chunk_steps, stream_partial=False):
            if done:
                results[idx] = toks
        return results  # type: ignore

    def generate_stream(self, requests: Sequence[GenerationRequest], chunk_steps: int = 32, stream_partial: bool = True):
        """Generator form of `generate`: after every `chunk_steps` decode steps yields
        (request index, new token ids since the last yield as int64, finished flag) for every running request, so a
        consumer (XCodec2 decode, data/tokenizer.py:117-123) can overlap with generation.  The concatenation of a
        request's chunks is exactly what `generate` returns for it.  Also accumulates self.stats (row-steps and the KV
        tokens they read) for the roofline accounting of bench.py."""
        c = self.cfg
        pending = deque(range(len(requests)))
        free = deque(range(c.max_slots))
        running: Dict[int, Tuple[int, int]] = {}        # slot -> (request index, max_new)
        n_done: Dict[int, int] = {}
        n_sent: Dict[int, int] = {}
        stats = self.stats = getattr(self, "stats", None) or dict(row_steps=0, kv_token_reads=0, decode_steps=0)
        while pending or running:
            admit, slots, tok_e, tok_d = [], [], 0, 0
            while pending and free:
                r = requests[pending[0]]
                ne, nd = len(r.text_ids), len(r.prompt_ids) + 1
                if admit and (tok_e + ne > c.max_prefill_tokens or tok_d + nd > c.max_prefill_tokens):
                    break
                admit.append(pending.popleft())
                slots.append(free.popleft())
                tok_e += ne
                tok_d += nd
            if admit:
                self.prefill([requests[i] for i in admit], slots)
                for i, s in zip(admit, slots):
                    running[s] = (i, self.max_new_tokens_of(requests[i], len(requests[i].prompt_ids) + 1))
                    n_done[s] = 0
                    n_sent[s] = 0
            remaining = max(mx - n_done[s] for s, (_, mx) in running.items())
            steps = max(1, min(chunk_steps, remaining))
            self.decode(steps)
            states = self.poll()
            stats["decode_steps"] += steps
            for s in list(running.keys()):
                i = running[s][0]
                dn = states[s].n_generated - n_done[s]
                # step j of this chunk attended to (BOS + prompt + tokens so far) self keys and n_text cross keys
                ctx0 = len(requests[i].prompt_ids) + 1 + n_done[s] + len(requests[i].text_ids)
                stats["row_steps"] += dn
                stats["kv_token_reads"] += dn * ctx0 + dn * (dn - 1) // 2
                n_done[s] = states[s].n_generated
                fin = bool(states[s].finished)
                if fin or stream_partial:
                    toks = self.read_tokens(s).astype(np.int64)
                    new = toks[n_sent[s]:] if stream_partial else toks
                    n_sent[s] = len(toks)
                    if fin or len(new):
                        yield i, new, fin
                if fin:
                    self.release(s)
                    del running[s]
                    free.append(s)

    @torch.inference_mode()
    def inference_tts(self, x: torch.Tensor, x_lens: torch.Tensor, y: torch.Tensor, tgt_y_lens: torch.Tensor,
                      top_k: Union[int, List[int]] = -100, top_p: float = 1.0, min_p: float = 0.0,
                      temperature: float = 1.0, stop_repetition: int = 3, silence_tokens: List[int] = None,
                      multi_trial: List[int] = None, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        """Same contract as the reference (models/t5gemma.py:835-1129): x [1,S] i64, x_lens [1], y [1,Tp,1] i64,
        tgt_y_lens [1] -> (res [1,1,Tp+Tg] i64, gen [1,1,Tg] i64) on x.device; gen's last id is eos."""
        if getattr(self.args, "n_codebooks", 1) != 1:
            raise ValueError("XCodec2 inference expects n_codebooks=1.")
        if multi_trial:
            logging.warning("multi_trial is unsupported and will be ignored.")
        batch_size = x.shape[0]
        assert batch_size == 1, "Current implementation only supports batch size 1."
        S = int(x_lens[0].item())
        text = x[0, :S].detach().cpu().numpy()
        yv = y.transpose(2, 1).contiguous()                 # [B,1,T]
        y_len = yv.shape[-1]
        prompt = yv[0, 0].detach().cpu().numpy()
        req = GenerationRequest(text_ids=text, prompt_ids=prompt,
                                target_total=None if tgt_y_lens is None else int(tgt_y_lens[0].item()),
                                prompt_frames=kwargs.get("prompt_frames", y_len), top_k=top_k, top_p=top_p, min_p=min_p,
                                temperature=temperature, max_new_tokens=int(kwargs.get("max_new_tokens", 0) or 0),
                                stop_repetition=int(stop_repetition), silence_tokens=list(silence_tokens or []),
                                uniforms=kwargs.get("uniforms"))
        on_chunk = kwargs.get("on_chunk")            # optional streaming callback: on_chunk(new_ids int64 [n], finished)
        if on_chunk is None:
            gen = self.generate([req], chunk_steps=int(kwargs.get("chunk_steps", 32)))[0]
        else:
            parts = []
            for _, new, fin in self.generate_stream([req], chunk_steps=int(kwargs.get("chunk_steps", 32))):
                parts.append(new)
                on_chunk(new - self.cfg.n_special if self.cfg.special_first else new, fin)
            gen = np.concatenate(parts) if parts else np.zeros(0, np.int64)
        if self.cfg.special_first:
            gen = gen - self.cfg.n_special
        gen_t = torch.from_numpy(gen).to(device=x.device, dtype=torch.long)[None, :]       # [1,Tg]
        res = torch.cat([yv[0].to(torch.long), gen_t], dim=1).unsqueeze(0)
        assert res.shape == torch.Size((1, 1, y_len + gen_t.shape[1]))
        return res, gen_t.unsqueeze(0)

    @torch.inference_mode()
    def inference_tts_batch(self, requests: Sequence[GenerationRequest], chunk_steps: int = 32):
        """Batched extension: list of requests -> list of (res [1,1,Tp+Tg], gen [1,1,Tg]) int64 CPU tensors;
        every row equals the bs=1 call of that request."""
        outs = self.generate(requests, chunk_steps=chunk_steps)
        res = []
        for r, g in zip(requests, outs):
            if self.cfg.special_first:
                g = g - self.cfg.n_special
            gt = torch.from_numpy(g).to(torch.long)[None, None, :]
            pt = torch.as_tensor(np.asarray(r.prompt_ids, dtype=np.int64))[None, None, :]
            res.append((torch.cat([pt, gt], dim=2), gt))
        return res
