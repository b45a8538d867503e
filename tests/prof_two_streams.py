"""Experiment: K independent GPT handles (batch 32/K each) on K streams / threads vs one handle with batch 32."""
import os, sys, threading, time
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from chatttsplus_b200 import synth
from chatttsplus_b200.gpt import GPT
from chatttsplus_b200.processors import gen_logits

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
L0, Btot = 128, 32
cfg = synth.GPTConfig()
sd = synth.make_gpt_state(cfg, seed=1234)
w, p = gen_logits(625, 0.7, 20, 1.05)
g = torch.Generator().manual_seed(0)
ids_all = torch.randint(0, cfg.num_text_tokens, (Btot, L0, 1), generator=g).expand(-1, -1, 4).clone()
handles = []
for k in range(K):
    gpt = GPT(dict(hidden_size=768, intermediate_size=3072, num_attention_heads=12, num_hidden_layers=20), max_batch=Btot // K)
    gpt.load_state_dict(sd); gpt.to("cuda")
    handles.append(gpt)

def run(k, stream, out):
    B = Btot // K
    ids = ids_all[k * B:(k + 1) * B].cuda()
    mask = torch.ones(B, L0, dtype=torch.long)
    with torch.cuda.stream(stream):
        emb = handles[k](ids, mask.bool().cuda())
        r = list(handles[k].generate(emb, ids, torch.tensor([0.3] * 4), 625, mask, max_new_token=steps, min_new_token=steps,
                                     logits_warpers=w, logits_processors=p, return_hidden=True, show_tqdm=False, ensure_non_empty=False))
    out[k] = r

streams = [torch.cuda.Stream() for _ in range(K)]
for rep in range(3):
    out = [None] * K
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ths = [threading.Thread(target=run, args=(k, streams[k], out)) for k in range(K)]
    for t in ths: t.start()
    for t in ths: t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"K={K} rep{rep}: {Btot * steps / dt:.0f} frames/s total, wall {dt*1e3:.1f} ms for {steps} steps ({1e6*dt/steps:.0f} us per step-set)")
