"""CPU oracle: fp32 restatement of the T5Gemma-TTS token-generation hot path.

TEST INFRASTRUCTURE ONLY.  Importable solely from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs.  The product path
(t5gemma_tts_b200/) never imports this file and has no CPU fallback.

Parity status: PINNED against the reference run in the build container --
oracle/make_golden.py imports the unmodified reference (hf_export twin of
models/t5gemma.py) on seeded tiny configs and commits its encoder states,
teacher-forced logits, generated token sequences and sampler survivor sets to
tests/golden/*.npz; tests/test_oracle_golden.py replays them through this file
(fp32 max-abs error <= 2e-5 on states/logits, token sequences identical).
The reference itself ships no tests or golden vectors.

The arithmetic restated here lives in a third-party dependency of the
reference: `transformers` (pinned 4.57.3 in /root/reference/requirements.txt:15;
5.5.0 installed and used to generate the fixtures), module
transformers.models.t5gemma.modeling_t5gemma ("HF:" below), plus the
reference-owned PM-RoPE / audio head / sampling / stop rules.

torch (CPU, fp32) is used purely as the array library.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import sampler_oracle


@dataclass
class OracleConfig:
    hidden: int = 2304
    inter: int = 9216
    n_enc_layers: int = 26
    n_dec_layers: int = 26
    n_heads: int = 8
    n_kv_heads: int = 4
    head_dim: int = 256
    sliding_window: int = 4096
    query_pre_attn_scalar: float = 256.0
    attn_softcap: Optional[float] = 50.0      # None under attn_implementation="sdpa" 
    rms_eps: float = 1e-6
    rope_theta: float = 10000.0
    text_vocab: int = 256000
    audio_vocab: int = 65536                  # without specials
    n_special: int = 5
    progress_scale: float = 2000.0
    encodec_sr: int = 50
    extra_cutoff: float = 5.0
    text_guard_frames_per_token: int = 0
    enc_layer_types: List[str] = field(default_factory=list)
    dec_layer_types: List[str] = field(default_factory=list)

    def __post_init__(self):
        # HF:configuration_t5gemma.py:96-99 -- sliding on even layers, full on odd
        if not self.enc_layer_types:
            self.enc_layer_types = ["sliding_attention" if (i + 1) % 2 else "full_attention"
                                    for i in range(self.n_enc_layers)]
        if not self.dec_layer_types:
            self.dec_layer_types = ["sliding_attention" if (i + 1) % 2 else "full_attention"
                                    for i in range(self.n_dec_layers)]

    # special ids: /root/reference/config.py:224-228
    @property
    def n_audio_tokens(self): return self.audio_vocab + self.n_special
    @property
    def empty_token(self): return self.audio_vocab
    @property
    def eog(self): return self.audio_vocab + 1
    @property
    def audio_pad_token(self): return self.audio_vocab + 2
    @property
    def eos(self): return self.audio_vocab + 3
    @property
    def y_sep_token(self): return self.audio_vocab + 4

    @staticmethod
    def from_reference_config(cfg) -> "OracleConfig":
        """cfg: a reference T5GemmaVoiceConfig (hf_export/configuration_t5gemma_voice.py:50-151)."""
        d = cfg.t5_config_dict
        e, dd = d["encoder"], d["decoder"]
        attn_impl = getattr(cfg, "attn_implementation", "eager")
        return OracleConfig(
            hidden=dd["hidden_size"], inter=dd["intermediate_size"],
            n_enc_layers=e["num_hidden_layers"], n_dec_layers=dd["num_hidden_layers"],
            n_heads=dd["num_attention_heads"], n_kv_heads=dd["num_key_value_heads"],
            head_dim=dd["head_dim"], sliding_window=dd["sliding_window"],
            query_pre_attn_scalar=float(dd["query_pre_attn_scalar"]),
            attn_softcap=(dd.get("attn_logit_softcapping") if attn_impl == "eager" else None),
            rms_eps=dd.get("rms_norm_eps", 1e-6),
            rope_theta=float((dd.get("rope_parameters") or {}).get("rope_theta", dd.get("rope_theta", 10000.0))),
            text_vocab=e["vocab_size"], audio_vocab=int(cfg.audio_vocab_size), n_special=int(cfg.n_special),
            progress_scale=float(cfg.progress_scale), encodec_sr=int(cfg.encodec_sr),
            extra_cutoff=float(cfg.extra_cutoff),
            text_guard_frames_per_token=int(cfg.text_guard_frames_per_token),
            enc_layer_types=list(e.get("layer_types") or []), dec_layer_types=list(dd.get("layer_types") or []),
        )


# ----------------------------------------------------------------------------
# Blocks (HF:models/t5gemma/modeling_t5gemma.py)
# ----------------------------------------------------------------------------

def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """HF:modeling_t5gemma.py:66-74 -- x*rsqrt(mean(x^2)+eps)*(1+w), fp32."""
    x = x.float()
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * (1.0 + w.float())


def rope_cos_sin(pos: torch.Tensor, head_dim: int, theta: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """HF:modeling_t5gemma.py:118-161 -- inv_freq[i]=theta^(-2i/d); angles fp32 from FLOAT positions.
    pos [T] fp32 -> cos,sin [T, head_dim] (halves duplicated)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = pos.float()[:, None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """HF:modeling_t5gemma.py:164-194 -- half-split rotation. x [H,T,D], cos/sin [T,D]."""
    d = x.shape[-1] // 2
    rot = torch.cat((-x[..., d:], x[..., :d]), dim=-1)
    return x * cos[None] + rot * sin[None]


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: Optional[torch.Tensor],
              scaling: float, softcap: Optional[float]) -> torch.Tensor:
    """HF:modeling_t5gemma.py:209-240 (eager). q [Hq,Tq,D], k/v [Hkv,Tk,D], mask bool [Tq,Tk] (True=attend).
    Returns [Tq, Hq*D]."""
    n_rep = q.shape[0] // k.shape[0]
    k = k.repeat_interleave(n_rep, dim=0)
    v = v.repeat_interleave(n_rep, dim=0)
    s = torch.matmul(q, k.transpose(1, 2)) * scaling
    if softcap is not None:
        s = torch.tanh(s / softcap) * softcap
    if mask is not None:
        s = s + torch.where(mask, 0.0, torch.finfo(torch.float32).min)[None]
    p = F.softmax(s, dim=-1, dtype=torch.float32)
    o = torch.matmul(p, v)                                  # [Hq,Tq,D]
    return o.transpose(0, 1).reshape(o.shape[1], -1)


def mlp(x, wg, wu, wd):
    """HF:modeling_t5gemma.py:92-96 -- down(gelu_tanh(gate(x))*up(x))."""
    return F.linear(F.gelu(F.linear(x, wg), approximate="tanh") * F.linear(x, wu), wd)


class Oracle:
    """Holds the reference state_dict (names as in the reference's HF twin) in fp32."""

    def __init__(self, cfg: OracleConfig, state_dict: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.w = {k: v.detach().to(torch.float32).contiguous() for k, v in state_dict.items()
                  if torch.is_floating_point(v)}
        self.scaling = cfg.query_pre_attn_scalar ** -0.5

    # ---- positions (models/t5gemma.py:609-624, 945-950, 1086-1099) ----
    def encoder_positions(self, length: int, max_len: Optional[int] = None) -> torch.Tensor:
        max_len = length if max_len is None else max_len
        pos = torch.arange(max_len, dtype=torch.float32)
        denom = torch.tensor(float(max(length, 2)), dtype=torch.float32) - 1.0
        out = pos / denom * self.cfg.progress_scale
        out[length:] = 0.0
        return out

    def decoder_prefill_positions(self, cur_len: int, est_total: int) -> torch.Tensor:
        base = torch.arange(cur_len, dtype=torch.float32)
        return base / max(1, est_total - 1) * self.cfg.progress_scale

    def decoder_step_position(self, current_length: int, est_total: int) -> float:
        v = float(current_length - 1) / max(1, est_total - 1) * self.cfg.progress_scale   # python float64
        v = min(v, self.cfg.progress_scale)
        return float(np.float32(v))

    # ---- encoder (HF:modeling_t5gemma.py:664-718, 426-452) ----
    def encoder(self, ids: torch.Tensor) -> torch.Tensor:
        """ids int64 [S] (one utterance, all valid) -> memory [S, hidden] fp32."""
        c, w = self.cfg, self.w
        S = ids.shape[0]
        h = w["backbone.model.encoder.embed_tokens.weight"][ids] * torch.tensor(c.hidden ** 0.5, dtype=torch.float32)
        pos = self.encoder_positions(S)
        cos, sin = rope_cos_sin(pos, c.head_dim, c.rope_theta)
        qi = torch.arange(S)[:, None]
        ki = torch.arange(S)[None, :]
        win_mask = (qi - ki).abs() <= c.sliding_window
        for l in range(c.n_enc_layers):
            p = f"backbone.model.encoder.layers.{l}."
            x = rmsnorm(h, w[p + "pre_self_attn_layernorm.weight"], c.rms_eps)
            q = F.linear(x, w[p + "self_attn.q_proj.weight"]).view(S, c.n_heads, c.head_dim).transpose(0, 1)
            k = F.linear(x, w[p + "self_attn.k_proj.weight"]).view(S, c.n_kv_heads, c.head_dim).transpose(0, 1)
            v = F.linear(x, w[p + "self_attn.v_proj.weight"]).view(S, c.n_kv_heads, c.head_dim).transpose(0, 1)
            q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
            m = win_mask if c.enc_layer_types[l] == "sliding_attention" else None
            a = attention(q, k, v, m, self.scaling, c.attn_softcap)
            a = F.linear(a, w[p + "self_attn.o_proj.weight"])
            h = h + rmsnorm(a, w[p + "post_self_attn_layernorm.weight"], c.rms_eps)
            x = rmsnorm(h, w[p + "pre_feedforward_layernorm.weight"], c.rms_eps)
            f = mlp(x, w[p + "mlp.gate_proj.weight"], w[p + "mlp.up_proj.weight"], w[p + "mlp.down_proj.weight"])
            h = h + rmsnorm(f, w[p + "post_feedforward_layernorm.weight"], c.rms_eps)
        return rmsnorm(h, w["backbone.model.encoder.norm.weight"], c.rms_eps)

    # ---- cross K/V, computed once (models/t5gemma.py:117-149) ----
    def cross_kv(self, memory: torch.Tensor) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        c, w = self.cfg, self.w
        S = memory.shape[0]
        cos, sin = rope_cos_sin(self.encoder_positions(S), c.head_dim, c.rope_theta)
        out = []
        for l in range(c.n_dec_layers):
            p = f"backbone.model.decoder.layers.{l}.cross_attn."
            k = F.linear(memory, w[p + "k_proj.weight"]).view(S, c.n_kv_heads, c.head_dim).transpose(0, 1)
            v = F.linear(memory, w[p + "v_proj.weight"]).view(S, c.n_kv_heads, c.head_dim).transpose(0, 1)
            out.append((apply_rope(k, cos, sin), v))
        return out

    # ---- decoder over q_len new tokens with a KV cache (HF:748-828; models/t5gemma.py:183-243) ----
    def decoder(self, emb: torch.Tensor, pos: torch.Tensor, cache: List[Optional[Tuple[torch.Tensor, torch.Tensor]]],
                cross: List[Tuple[torch.Tensor, torch.Tensor]]) -> torch.Tensor:
        """emb [T,hidden] raw audio embeddings of the new tokens; pos [T] fp32 PM positions;
        cache: per-layer (K,V) [Hkv,past,D] post-RoPE (full history; the sliding window is applied by mask,
        equivalent to DynamicSlidingWindowLayer, HF:cache_utils.py:166-223).  Mutated in place."""
        c, w = self.cfg, self.w
        T = emb.shape[0]
        h = emb * torch.tensor(c.hidden ** 0.5, dtype=torch.float32)
        cos, sin = rope_cos_sin(pos, c.head_dim, c.rope_theta)
        past = 0 if cache[0] is None else cache[0][0].shape[1]
        qi = (past + torch.arange(T))[:, None]
        ki = torch.arange(past + T)[None, :]
        causal = ki <= qi
        sliding = causal & (ki > qi - c.sliding_window)
        for l in range(c.n_dec_layers):
            p = f"backbone.model.decoder.layers.{l}."
            x = rmsnorm(h, w[p + "pre_self_attn_layernorm.weight"], c.rms_eps)
            q = F.linear(x, w[p + "self_attn.q_proj.weight"]).view(T, c.n_heads, c.head_dim).transpose(0, 1)
            k = F.linear(x, w[p + "self_attn.k_proj.weight"]).view(T, c.n_kv_heads, c.head_dim).transpose(0, 1)
            v = F.linear(x, w[p + "self_attn.v_proj.weight"]).view(T, c.n_kv_heads, c.head_dim).transpose(0, 1)
            q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
            if cache[l] is not None:
                k = torch.cat([cache[l][0], k], dim=1)
                v = torch.cat([cache[l][1], v], dim=1)
            cache[l] = (k, v)
            m = sliding if c.dec_layer_types[l] == "sliding_attention" else causal
            a = attention(q, k, v, m, self.scaling, c.attn_softcap)
            a = F.linear(a, w[p + "self_attn.o_proj.weight"])
            h = h + rmsnorm(a, w[p + "post_self_attn_layernorm.weight"], c.rms_eps)

            x = rmsnorm(h, w[p + "pre_cross_attn_layernorm.weight"], c.rms_eps)
            q = F.linear(x, w[p + "cross_attn.q_proj.weight"]).view(T, c.n_heads, c.head_dim).transpose(0, 1)
            q = apply_rope(q, cos, sin)                       # PM-RoPE on cross queries, decoder positions
            a = attention(q, cross[l][0], cross[l][1], None, self.scaling, c.attn_softcap)
            a = F.linear(a, w[p + "cross_attn.o_proj.weight"])
            h = h + rmsnorm(a, w[p + "post_cross_attn_layernorm.weight"], c.rms_eps)

            x = rmsnorm(h, w[p + "pre_feedforward_layernorm.weight"], c.rms_eps)
            f = mlp(x, w[p + "mlp.gate_proj.weight"], w[p + "mlp.up_proj.weight"], w[p + "mlp.down_proj.weight"])
            h = h + rmsnorm(f, w[p + "post_feedforward_layernorm.weight"], c.rms_eps)
        return rmsnorm(h, w["backbone.model.decoder.norm.weight"], c.rms_eps)

    # ---- head (models/t5gemma.py:397-406,1058): Linear+b -> exact GELU -> Linear+b ----
    def head(self, hidden: torch.Tensor) -> torch.Tensor:
        w = self.w
        t = F.gelu(F.linear(hidden, w["predict_layer.0.0.weight"], w["predict_layer.0.0.bias"]))
        return F.linear(t, w["predict_layer.0.2.weight"], w["predict_layer.0.2.bias"])

    def embed_audio(self, ids: torch.Tensor) -> torch.Tensor:
        return self.w["audio_embedding.0.weight"][ids]

    # ---- teacher-forced logits over a given decoder input sequence ----
    def teacher_forced_logits(self, x_ids: torch.Tensor, dec_ids: torch.Tensor, est_total: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """dec_ids int64 [T] = BOS ++ tokens.  Returns (memory [S,h], logits [T, n_audio_tokens]),
        positions t/max(1,est_total-1)*scale as in the prefill (models/t5gemma.py:945-950)."""
        memory = self.encoder(x_ids)
        cross = self.cross_kv(memory)
        cache = [None] * self.cfg.n_dec_layers
        pos = self.decoder_prefill_positions(dec_ids.shape[0], est_total)
        hid = self.decoder(self.embed_audio(dec_ids), pos, cache, cross)
        return memory, self.head(hid)

    # ---- the generate loop (models/t5gemma.py:835-1129) ----
    @torch.no_grad()
    def inference_tts(self, x: torch.Tensor, x_lens: torch.Tensor, y: torch.Tensor, tgt_y_lens: torch.Tensor,
                      top_k=-100, top_p: float = 1.0, min_p: float = 0.0, temperature: float = 1.0,
                      prompt_frames: Optional[int] = None, uniforms: Optional[Sequence[float]] = None,
                      max_new_tokens: Optional[int] = None, return_logits: bool = False,
                      stop_repetition: int = 3, silence_tokens: Sequence[int] = (), logits_hook=None):
        """Same contract as the reference, bs=1.  `uniforms[i]` is the U[0,1) draw consumed by step i
        (the reference consumes torch.multinomial's stream instead; see sampler_oracle)."""
        c = self.cfg
        assert x.shape[0] == 1, "Current implementation only supports batch size 1."
        S = int(x_lens[0])
        memory = self.encoder(x[0, :S])
        cross = self.cross_kv(memory)
        yv = y.transpose(2, 1)[0, 0]                         # [Tp]
        y_len = yv.shape[0]
        prompt_frames = y_len if prompt_frames is None else prompt_frames
        target_total = None if tgt_y_lens is None else int(tgt_y_lens[0])
        cated = torch.cat([torch.tensor([c.empty_token], dtype=torch.long), yv.long()])
        current_length = cated.shape[0]
        prompt_offset = prompt_frames + 1
        if target_total is not None:
            est_total = target_total + 1
        else:       # models/t5gemma.py:925-933: no target -> progress_lookahead_secs (2.0 s) past the prompt
            est_total = int(current_length + int(c.encodec_sr) * 2.0)
        est_total = max(est_total, current_length)
        cache = [None] * c.n_dec_layers
        pos = self.decoder_prefill_positions(current_length, est_total)
        hid = self.decoder(self.embed_audio(cated), pos, cache, cross)
        last_hidden = hid[-1:]
        gen: List[int] = []
        all_logits = []
        cur_num_gen = 0
        prev_token, consec = -1, 0                       # models/t5gemma.py:967-968
        while True:
            logits = self.head(last_hidden)[0].clone()
            if logits_hook is not None:
                logits = torch.from_numpy(np.asarray(logits_hook(logits.numpy(), cur_num_gen), dtype=np.float32)).clone()
            if return_logits:
                all_logits.append(logits.clone())
            u = float(uniforms[cur_num_gen]) if uniforms is not None else 0.5
            kk = top_k[min(len(top_k) - 1, cur_num_gen)] if isinstance(top_k, (list, tuple)) else top_k
            token_id, det = sampler_oracle.sample_step(
                logits.numpy(), eos=c.eos, cur_num_gen=cur_num_gen, current_length=current_length,
                prompt_offset=prompt_offset, target_total=target_total, encodec_sr=c.encodec_sr,
                extra_cutoff=c.extra_cutoff, top_k=kk, top_p=top_p, min_p=min_p, temperature=temperature, u=u,
                x_len=S, text_guard_frames_per_token=c.text_guard_frames_per_token, return_detail=True,
                prev_token=prev_token, consec_silence_count=consec, stop_repetition=stop_repetition,
                silence_tokens=tuple(silence_tokens))
            prev_token, consec = det["prev_token"], det["consec_silence_count"]
            if max_new_tokens is not None and cur_num_gen + 1 >= max_new_tokens:
                token_id = c.eos
            gen.append(token_id)
            cur_num_gen += 1
            current_length += 1
            if token_id == c.eos:
                break
            p1 = self.decoder_step_position(current_length, est_total)
            hid = self.decoder(self.embed_audio(torch.tensor([token_id])), torch.tensor([p1], dtype=torch.float32),
                               cache, cross)
            last_hidden = hid
        gen_t = torch.tensor(gen, dtype=torch.long)[None, None, :]
        res = torch.cat([yv.long()[None, None, :], gen_t], dim=2)
        if return_logits:
            return res, gen_t, torch.stack(all_logits)
        return res, gen_t
