"""Profiling driver: one prefill (B=32, L0=128, 20 layers) and one vocoder pass (32 utterances x 128 frames)."""
import os, sys
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from chatttsplus_b200 import synth
from chatttsplus_b200.gpt import GPT
from chatttsplus_b200.processors import gen_logits
from chatttsplus_b200.vocoder import DVAE, Vocos, VocoderEngine
cfg = synth.GPTConfig()
gpt = GPT(dict(hidden_size=768, intermediate_size=3072, num_attention_heads=12, num_hidden_layers=20), max_batch=32)
gpt.load_state_dict(synth.make_gpt_state(cfg, seed=1234)); gpt.to("cuda")
d = DVAE(decoder_config=dict(idim=384, odim=384, hidden=512, n_layer=12, bn_dim=128), dim=384); d.load_state_dict(synth.make_dvae_state(synth.DVAEConfig(), 1)); d.to("cuda")
v = Vocos(backbone_config=dict(input_channels=100, dim=512, intermediate_dim=1536, num_layers=8), head_config=dict(dim=512, n_fft=1024, hop_length=256, padding="center"))
v.load_state_dict(synth.make_vocos_state(synth.VocosConfig(), 2)); v.to("cuda")
eng = VocoderEngine(d, v)
B, L0 = 32, 128
g = torch.Generator().manual_seed(0)
ids = torch.randint(0, cfg.num_text_tokens, (B, L0, 1), generator=g).expand(-1, -1, 4).clone()
mask = torch.ones(B, L0, dtype=torch.long)
w, p = gen_logits(625, 0.7, 20, 1.05)
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 128
hid = [torch.randn(nf, 768, device="cuda") for _ in range(B)]
for rep in range(2):
    emb = gpt(ids.cuda(), mask.bool().cuda())
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    list(gpt.generate(emb, ids.cuda(), torch.tensor([0.3] * 4), 625, mask, max_new_token=1, min_new_token=1, logits_warpers=w,
                      logits_processors=p, return_hidden=True, show_tqdm=False, ensure_non_empty=False))
    e1.record()
    wavs, _ = eng.decode_batch(hid)
    e2.record()
    torch.cuda.synchronize()
    flop_v = 157.4e6 * B * nf
    print(f"prefill+1 sample {e0.elapsed_time(e1):.2f} ms ({1.55e12 / (e0.elapsed_time(e1) * 1e-3) / 1e12:.0f} TFLOP/s incl. attention); "
          f"vocoder {e1.elapsed_time(e2):.2f} ms for {B}x{nf} frames ({flop_v / (e1.elapsed_time(e2) * 1e-3) / 1e12:.0f} TFLOP/s)")
