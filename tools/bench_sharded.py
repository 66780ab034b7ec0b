"""BASELINE.json configs[3]: request-sharded throughput, N synthetic utterances (mixed 3-15 s) across the ranks of a
torchrun launch, one engine replica per GPU, LPT sharding, NCCL only for the final gather.

  python -m torch.distributed.run --nproc-per-node N tools/bench_sharded.py --utterances 2048 --slots 64
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from t5gemma_tts_b200 import EngineConfig, T5GemmaVoiceEngine, GenerationRequest  # noqa: E402
from t5gemma_tts_b200.random_init import iter_random_state_dict  # noqa: E402
from t5gemma_tts_b200.sharding import run_sharded  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--utterances", type=int, default=2048)
ap.add_argument("--slots", type=int, default=256)     # rows per engine: 64 -> 22.9 k, 256 -> 42.9 k tok/s/GPU
a = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
cfg = EngineConfig(max_slots=a.slots, max_text_len=128, max_dec_len=1024, max_prefill_tokens=8192)
eng = T5GemmaVoiceEngine(cfg, device=dev)
eng.load_state_dict(iter_random_state_dict(cfg, seed=0, device=dev))
rng = np.random.default_rng(2048)
reqs = []
for i in range(a.utterances):
    S = int(rng.integers(32, 129))
    tgt = int(50 * rng.uniform(3, 15))
    reqs.append(GenerationRequest(text_ids=rng.integers(2, 255000, S), prompt_ids=np.zeros(0, np.int64), target_total=tgt,
                                  prompt_frames=0, top_k=30, top_p=0.9, temperature=0.8))
torch.manual_seed(rank)
eng.generate(reqs[:2], chunk_steps=8)      # warm-up
costs = [r.target_total + 252 for r in reqs]
res, sec, total = run_sharded(lambda idx: eng.generate([reqs[i] for i in idx], chunk_steps=32), costs, device=dev)
if rank == 0:
    print(json.dumps({"workload": f"configs[3]: {a.utterances} utterances (3-15 s, S~U[32,128]), LPT-sharded over {world} GPU(s), {a.slots} rows/engine",
                      "n_gpus": world, "tokens": total, "seconds_max_rank": sec, "tokens_per_s": total / sec,
                      "real_time_factor": total / 50.0 / sec}))
if world > 1:
    dist.destroy_process_group()
