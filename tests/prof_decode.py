"""Profiling driver: full-size GPT, B=32, 128-token prompt, a few graph-replayed decode steps (run under ncu).
    python tests/prof_decode.py [steps] [prompt_len]"""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from chatttsplus_b200 import synth  # noqa: E402
from chatttsplus_b200.gpt import GPT  # noqa: E402
from chatttsplus_b200.processors import gen_logits  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
L0 = int(sys.argv[2]) if len(sys.argv) > 2 else 128
cfg = synth.GPTConfig()
gpt = GPT(dict(hidden_size=768, intermediate_size=3072, num_attention_heads=12, num_hidden_layers=20), max_batch=32)
gpt.load_state_dict(synth.make_gpt_state(cfg, seed=1234))
gpt.to("cuda")
B = 32
g = torch.Generator().manual_seed(0)
ids = torch.randint(0, cfg.num_text_tokens, (B, L0, 1), generator=g).expand(-1, -1, 4).clone()
mask = torch.ones(B, L0, dtype=torch.long)
emb = gpt(ids.cuda(), mask.bool().cuda())
w, p = gen_logits(625, 0.7, 20, 1.05)
gpt.record_timing = True
for rep in range(2):
    list(gpt.generate(emb, ids.cuda(), torch.tensor([0.3] * 4), 625, mask, max_new_token=steps, min_new_token=steps, logits_warpers=w,
                      logits_processors=p, return_hidden=True, show_tqdm=False, ensure_non_empty=False))
    print("timing", gpt.timing, "us/step", 1e3 * gpt.timing["decode_ms"] / max(1, gpt.timing["decode_steps"]))
