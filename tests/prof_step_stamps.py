"""Bring-up: per-phase work / barrier-wait times of the fused decode step (clock64 stamps around every grid barrier)."""
import ctypes, os, sys
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from chatttsplus_b200 import _lib, synth
from chatttsplus_b200.gpt import GPT
from chatttsplus_b200.processors import gen_logits
lib = _lib.lib()
lib.ctp_debug_step_stamps.argtypes = [ctypes.c_void_p]
L0 = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = synth.GPTConfig()
gpt = GPT(dict(hidden_size=768, intermediate_size=3072, num_attention_heads=12, num_hidden_layers=20), max_batch=32)
gpt.load_state_dict(synth.make_gpt_state(cfg, seed=1234)); gpt.to("cuda")
B = 32
g = torch.Generator().manual_seed(0)
ids = torch.randint(0, cfg.num_text_tokens, (B, L0, 1), generator=g).expand(-1, -1, 4).clone()
mask = torch.ones(B, L0, dtype=torch.long)
emb = gpt(ids.cuda(), mask.bool().cuda())
w, p = gen_logits(625, 0.7, 20, 1.05)
steps = 6
dbg = torch.zeros(148, 128, 8, dtype=torch.int64, device="cuda")
list(gpt.generate(emb, ids.cuda(), torch.tensor([0.3] * 4), 625, mask, max_new_token=steps, min_new_token=steps, logits_warpers=w,
                  logits_processors=p, return_hidden=True, show_tqdm=False, ensure_non_empty=False))
lib.ctp_debug_step_stamps(ctypes.c_void_p(dbg.data_ptr()))
list(gpt.generate(emb, ids.cuda(), torch.tensor([0.3] * 4), 625, mask, max_new_token=steps, min_new_token=steps, logits_warpers=w,
                  logits_processors=p, return_hidden=True, show_tqdm=False, ensure_non_empty=False))
torch.cuda.synchronize()
lib.ctp_debug_step_stamps(None)
d = dbg.cpu().double()  # last step's stamps
nb = 2 + 5 * 20
arr, lea = d[:, :nb, 0], d[:, :nb, 1]
work = arr[:, 1:] - lea[:, :-1]      # compute time of the phase ending at barrier e (per CTA)
wait = lea - arr
names = {0: "QKV", 1: "ATT", 2: "O", 3: "GU", 4: "DN"}
print("step total cycles (cta0):", lea[0, nb - 1] - arr[0, 0])
print("embed->bar0 wait mean", wait[:, 0].mean().item())
for k in range(5):
    idx = [1 + 5 * l + k for l in range(20)]
    wk = work[:, [i - 1 for i in idx]]
    wt = wait[:, idx]
    print(f"{names[k]:4s} work mean {wk.mean().item():8.0f} max-over-cta mean {wk.max(0).values.mean().item():8.0f} | wait mean {wt.mean().item():8.0f} min-over-cta mean {wt.min(0).values.mean().item():8.0f}")
i = nb - 1
for k, nm in [(0, "QKV"), (2, "O"), (3, "GU"), (4, "DN")]:
    idx = [1 + 5 * l + k for l in range(2, 20)]
    t_leave_prev = lea[:, [i - 1 for i in idx]]
    nd = d[:, idx, 2] - t_leave_prev
    ar = d[:, idx, 3] - t_leave_prev
    ep = d[:, idx, 4] - t_leave_prev
    m5 = d[:, idx, 5] - t_leave_prev; m6 = d[:, idx, 6] - t_leave_prev; m7 = d[:, idx, 7] - t_leave_prev
    act = d[:, idx, 3] > 0
    f = lambda t: (t * act).sum().item() / max(1, act.sum().item())
    print(f"{nm}: since phase start: norm_done {f(nd):8.0f} acc_ready {f(ar):8.0f} epi_done {f(ep):8.0f} | mma: a_ready_seen {f(m5):8.0f} slot0_full {f(m6):8.0f} all_issued {f(m7):8.0f} (active CTAs {act[:,0].sum().item()})")
print("HEAD work mean", work[:, i - 1].mean().item(), "max", work[:, i - 1].max().item(), "wait mean", wait[:, i].mean().item())
