"""CPU oracle for the sampling step S (numpy, fp32, bit-exact definition).

TEST INFRASTRUCTURE ONLY (see oracle/t5gemma_voice_oracle.py header).

Restates, for one row of logits:
  * sample_helper's in-place eos edits and stop rules  -- /root/reference/models/t5gemma.py:971-1055
  * topk_sampling / top_k_top_p_filtering              -- /root/reference/models/utils.py:53-122
Parity status: PINNED -- survivor sets are checked against the reference's own
top_k_top_p_filtering on the committed fixtures (tests/golden/sampler_*.npz, incl.
SURVEY.md Appendix A's known-answer cases K1-K9).

What is (necessarily) re-defined: the reference draws with torch.multinomial from the global
torch RNG stream; the engine consumes one explicit uniform u in [0,1) per step and draws by
inverse CDF over the survivors in (value desc, index asc) order.  Both sample the same
distribution.  To make "bit-exact given identical logits and identical uniform draws"
checkable between this CPU code and the CUDA kernel, every floating-point operation below is
an individually rounded IEEE fp32 add/sub/mul/div (no fma, no library exp): det_exp() is a
fixed polynomial evaluated with those ops only, and all sums are sequential in a stated order.
The CUDA kernel (csrc/sampler.cu) performs the identical operation sequence with
__fadd_rn/__fmul_rn/__fdiv_rn.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

# ---- deterministic exp for x <= 0 ----------------------------------------------------------
_LOG2E = f32(1.4426950408889634)
_LN2_HI = f32(0.693359375)            # 9 significant bits: n*_LN2_HI exact for |n| < 2^15
_LN2_LO = f32(-2.12194440e-4)
# minimax-ish coefficients for e^r on |r| <= ln2/2 (Cephes expf)
_C = [f32(1.9875691500e-4), f32(1.3981999507e-3), f32(8.3334519073e-3),
      f32(4.1665795894e-2), f32(1.6666665459e-1), f32(5.0000001201e-1)]
_EXP_LO = f32(-86.0)


def det_exp(x):
    """exp(x) for x<=0 as a fixed sequence of fp32 ops (vectorised; every op rounds to fp32)."""
    x = np.asarray(x, dtype=f32)
    with np.errstate(over="ignore", invalid="ignore"):
        xc = np.maximum(x, _EXP_LO)
        n = np.rint(xc * _LOG2E).astype(f32)
        r = xc - n * _LN2_HI
        r = r - n * _LN2_LO
        p = _C[0]
        for c in _C[1:]:
            p = p * r + c                     # numpy: mul rounds, add rounds (no fma)
        r2 = r * r
        y = p * r2
        y = y + r
        y = y + f32(1.0)
        out = np.ldexp(y, n.astype(np.int32)).astype(f32)     # exact power-of-two scaling (normal range)
    return np.where(x < _EXP_LO, f32(0.0), out).astype(f32)


def _seq_sum(a):
    """Sequential left-to-right fp32 sum (np.cumsum accumulates in order; np.sum is pairwise)."""
    a = np.asarray(a, dtype=f32)
    if a.size == 0:
        return f32(0.0)
    return np.cumsum(a, dtype=f32)[-1]


CHUNK = 256  # full-vocab sums: per-chunk sequential, then sequential over chunk sums


def _chunked_sum(e):
    V = e.shape[0]
    nchunk = (V + CHUNK - 1) // CHUNK
    pad = np.zeros(nchunk * CHUNK, dtype=f32)
    pad[:V] = e
    csum = np.cumsum(pad.reshape(nchunk, CHUNK), axis=1, dtype=f32)[:, -1]
    return csum, _seq_sum(csum)


def order_desc(z, idx):
    """(value desc, index asc) ordering of the candidate indices."""
    idx = np.asarray(idx, dtype=np.int64)
    return idx[np.lexsort((idx, -z[idx].astype(np.float64)))]


def filter_survivors(z, top_k, top_p, min_p):
    """Indices kept by top_k_top_p_filtering (models/utils.py:53-111) on temperature-scaled z,
    returned in (value desc, index asc) order, plus whether any filtering happened."""
    V = z.shape[0]
    min_p_enabled = 0.0 < min_p < 1.0
    if min_p_enabled:
        m = z.max()
        e = det_exp(z - m)
        _, s = _chunked_sum(e)
        probs = e / s
        keep = ~(probs < f32(min_p))
        if keep.any():                       # "skip if everything would be removed"
            return order_desc(z, np.nonzero(keep)[0]), True
    cand = np.arange(V)
    filtered = False
    if isinstance(top_k, (int, np.integer)) and top_k > 0:
        k = min(max(int(top_k), 1), V)
        thr = np.partition(z, V - k)[V - k]          # k-th largest
        cand = np.nonzero(~(z < thr))[0]             # ties kept (`logits < threshold` removed)
        filtered = True
    if top_p < 1.0:
        order = order_desc(z, cand)
        zs = z[order]
        e = det_exp(zs - zs[0])
        s = _seq_sum(e)
        p = e / s
        cum = np.cumsum(p, dtype=f32)
        remove = np.zeros(order.shape[0], dtype=bool)
        remove[1:] = cum[:-1] > f32(top_p)           # shift right by one; first always kept
        return order[~remove], True
    if filtered:
        return order_desc(z, cand), True
    return cand, False


def draw(z, survivors, filtered, u):
    """Inverse-CDF draw with uniform u (fp32) over the survivors."""
    u = f32(u)
    if filtered:
        zs = z[survivors]
        e = det_exp(zs - zs[0])                      # survivors[0] holds the max
        s = _seq_sum(e)
        c = np.cumsum(e / s, dtype=f32)
        hit = np.nonzero(u < c)[0]
        j = int(hit[0]) if hit.size else survivors.shape[0] - 1
        return int(survivors[j])
    # unfiltered: full vocabulary in index order, chunked deterministic sums, target t = u*s
    m = z.max()
    e = det_exp(z - m)
    csum, s = _chunked_sum(e)
    t = u * s
    ccum = np.cumsum(csum, dtype=f32)
    hitc = np.nonzero(t < ccum)[0]
    if hitc.size == 0:
        return int(z.shape[0] - 1)
    cidx = int(hitc[0])
    base = ccum[cidx - 1] if cidx > 0 else f32(0.0)
    lo = cidx * CHUNK
    hi = min(lo + CHUNK, z.shape[0])
    run = base
    for i in range(lo, hi):
        run = f32(run + e[i])
        if t < run:
            return i
    return hi - 1


def sample_step(logits, *, eos, cur_num_gen, current_length, prompt_offset, target_total, encodec_sr=50,
                extra_cutoff=5.0, top_k=-100, top_p=1.0, min_p=0.0, temperature=1.0, u=0.5,
                x_len=0, text_guard_frames_per_token=0, return_detail=False,
                prev_token=-1, consec_silence_count=0, stop_repetition=3, silence_tokens=()):
    """One sample_helper call (models/t5gemma.py:971-1055) for text_input_type=="text",
    silence_tokens=[] (every shipped caller).  `logits` fp32 [V] is edited in place like the reference."""
    logits = np.asarray(logits)
    assert logits.dtype == np.float32
    effective_length = max(0, current_length - prompt_offset)
    if effective_length == 0:
        logits[eos] = f32(-1e9)
    if cur_num_gen <= int(encodec_sr) // 5:
        logits[eos] = f32(-10000.0)
    # silence-repetition penalty (models/t5gemma.py:999-1011)
    if stop_repetition > 0 and prev_token in silence_tokens and consec_silence_count > stop_repetition:
        fct = f32(consec_silence_count - (stop_repetition - 1))
        if logits[prev_token] < 0:
            logits[prev_token] = f32(logits[prev_token] * fct)
        else:
            logits[prev_token] = f32(logits[prev_token] / fct)
    amax = int(np.argmax(logits))
    z = logits / f32(temperature) if temperature != 1.0 else logits
    z = z.astype(f32, copy=False)
    survivors, filtered = filter_survivors(z, top_k, top_p, min_p)
    token_id = draw(z, survivors, filtered, u)
    sampled = token_id
    force = (token_id == eos) or (amax == eos)
    if text_guard_frames_per_token > 0:
        force = force or (effective_length > max(1, x_len) * text_guard_frames_per_token)
    budget = target_total is not None and cur_num_gen > (
        target_total - prompt_offset + int(encodec_sr) * extra_cutoff)
    if force or budget:
        token_id = eos
    if return_detail:
        # models/t5gemma.py:1050-1054
        if token_id in silence_tokens and token_id == prev_token:
            consec_silence_count += 1
        else:
            consec_silence_count = 0
        return token_id, dict(survivors=survivors, argmax=amax, sampled=sampled, prev_token=token_id,
                              consec_silence_count=consec_silence_count)
    return token_id
