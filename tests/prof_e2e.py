"""Where the time of one configs[1] job goes outside the decode loop: host wall clock per stage (with a device synchronize after each) next to the
CUDA-event times of prefill / decode.    python tests/prof_e2e.py"""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chatttsplus_b200.commons.utils import InferCodeParams  # noqa: E402

dev = torch.device("cuda", 0)
cfg, pipe = bench.build_models(dev)
gpt = pipe.models_dict["gpt"]
gpt.record_timing = True
ids, mask, text_mask, spk_id = bench.synthetic_prompt(cfg, 32, seed=1234)
params = InferCodeParams(prompt="", spk_emb=None, temperature=0.3, top_P=0.7, top_K=20, repetition_penalty=1.05, max_new_token=512, min_new_token=512,
                         show_tqdm=False, ensure_non_empty=False)
ids_d, tm_d = ids.to(dev), text_mask.to(dev)


def job():
    for wavs in pipe.infer_ids(ids_d, mask, tm_d, params, spk_emb_ids=spk_id):
        pass
    return wavs


for _ in range(3):
    job()
torch.cuda.synchronize()
orig = pipe._decode_to_wavs
voc = [0.0]


def timed_voc(*a, **k):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = orig(*a, **k)
    torch.cuda.synchronize()
    voc[0] = time.perf_counter() - t0
    return r


pipe._decode_to_wavs = timed_voc
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    job()
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    print(f"job {tot*1e3:.2f} ms | prefill {gpt.timing['prefill_ms']:.2f} decode {gpt.timing['decode_ms']:.2f} (CUDA events) | vocoder call {voc[0]*1e3:.2f} ms (wall, synchronised) | "
          f"rest {tot*1e3 - gpt.timing['prefill_ms'] - gpt.timing['decode_ms'] - voc[0]*1e3:.2f} ms")
pipe._decode_to_wavs = orig
pr = cProfile.Profile()
pr.enable()
job()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
