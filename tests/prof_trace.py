"""Bring-up: in-graph timeline of one decode step (CTP_TRACE=1).  Every kernel of the step graph records globaltimer stamps
(CTA 0 enter / griddepcontrol.wait return / exit, latest exit over all CTAs); this prints them in launch order.
    CTP_TRACE=1 python tests/prof_trace.py [steps] [prompt_len]"""
import ctypes as C
import os
import sys

os.environ.setdefault("CTP_TRACE", "1")
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from chatttsplus_b200 import _lib, synth  # noqa: E402
from chatttsplus_b200.gpt import GPT  # noqa: E402
from chatttsplus_b200.processors import gen_logits  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 33
L0 = int(sys.argv[2]) if len(sys.argv) > 2 else 352
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
print("us/step", 1e3 * gpt.timing["decode_ms"] / max(1, gpt.timing["decode_steps"]))
lib = C.CDLL(_lib.lib()._name)
buf = (C.c_ulonglong * (8 * 256))()
lib.ctp_debug_trace.restype = C.c_int
lib.ctp_debug_trace.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]
n = lib.ctp_debug_trace(gpt._handle, buf, 256)
recs = [tuple(buf[8 * i + j] for j in range(8)) for i in range(n)]
t0 = recs[0][0]
print(f"{n} kernels of the last replayed step, launch order; times in us.  enter / wait_ret / exit0 / exit_last are relative to kernel 0's enter "
      f"(CTA 0 enters, its griddepcontrol.wait returns, CTA 0 leaves, the last CTA leaves); body = exit_last - wait_ret; handoff = wait_ret - previous "
      f"exit_last; s4..s7 = kernel-specific phase stamps relative to wait_ret (fused MLP kernel: s4 operand + row sums of all ranks landed, s5 D1 complete, s6 silu*up slice written, s7 D2 complete)")
prev_end = None
tot_body = tot_hand = 0.0
for i, r in enumerate(recs):
    a, b, c, d = r[:4]
    if not (a and b and d):
        print(f"{i:3d} (no record)")
        continue
    hand = (b - prev_end) / 1e3 if prev_end else 0.0
    extra = "  ".join(f"s{j} {(r[j] - b) / 1e3:6.2f}" for j in range(4, 8) if r[j] >= a)
    print(f"{i:3d} enter {(a - t0) / 1e3:8.2f} wait_ret {(b - t0) / 1e3:8.2f} exit0 {(c - t0) / 1e3:8.2f} exit_last {(d - t0) / 1e3:8.2f} | "
          f"body {(d - b) / 1e3:6.2f} handoff {hand:6.2f} | {extra}")
    tot_body += (d - b) / 1e3
    tot_hand += hand
    prev_end = d
print(f"sum of bodies {tot_body:.1f} us, sum of hand-offs {tot_hand:.1f} us, span (first enter -> last exit) {(recs[-1][3] - t0) / 1e3:.1f} us")

