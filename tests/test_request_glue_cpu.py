"""Host glue around the hot path vs the reference's inference_one_sample logic."""
from types import SimpleNamespace

import numpy as np

from oracle import fixtures
from t5gemma_tts_b200.request_glue import build_request, strip_sep_and_eos

CFG = SimpleNamespace(y_sep_token=65540, x_sep_token=255999, encodec_sr=50, add_eos_to_text=0, add_bos_to_text=0,
                      parallel_pattern=0)


def test_strip_matches_reference_fixture():
    z = np.load(f"{fixtures.GOLDEN}/glue_strip.npz")
    for i in range(6):
        got = strip_sep_and_eos(z[f"in_{i}"], 104, 103)
        assert np.array_equal(got, z[f"out_{i}"])
    assert strip_sep_and_eos(np.array([1, 2, 3]), None, None).tolist() == [1, 2, 3]


def test_build_request_follows_inference_one_sample():
    # voice prompt: y_sep appended, prompt_frames counts it, tgt_y_lens = prompt_frames + sr*sec (inference_tts_utils.py:229-286)
    r = build_request(CFG, [5, 6, 7], 10.0, prompt_codes=np.arange(150), prefix_text_ids=[1, 2])
    assert r.prompt_ids[-1] == 65540 and len(r.prompt_ids) == 151 and r.prompt_frames == 151
    assert r.text_ids.tolist() == [1, 2, 255999, 5, 6, 7]
    assert r.target_total == 151 + 500
    # no reference audio: no y_sep, no prompt
    r = build_request(CFG, [5, 6], 5.0)
    assert len(r.prompt_ids) == 0 and r.prompt_frames == 0 and r.target_total == 250 and r.text_ids.tolist() == [5, 6]
    # bos/eos text tokens and the parallel_pattern +2
    c2 = SimpleNamespace(**dict(vars(CFG), add_eos_to_text=9, add_bos_to_text=8, parallel_pattern=1))
    r = build_request(c2, [5], 1.0)
    assert r.text_ids.tolist() == [8, 5, 9] and r.target_total == 52
