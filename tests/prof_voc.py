"""Profiling driver: one vocoder pass over N utterances x F frames of hidden states (run under ncu).
    python tests/prof_voc.py [utterances] [frames]"""
import os, sys
import torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from chatttsplus_b200 import synth
from chatttsplus_b200.vocoder import DVAE, Vocos, VocoderEngine
d = DVAE(decoder_config=dict(idim=384, odim=384, hidden=512, n_layer=12, bn_dim=128), dim=384); d.load_state_dict(synth.make_dvae_state(synth.DVAEConfig(), 1)); d.to("cuda")
v = Vocos(backbone_config=dict(input_channels=100, dim=512, intermediate_dim=1536, num_layers=8), head_config=dict(dim=512, n_fft=1024, hop_length=256, padding="center"))
v.load_state_dict(synth.make_vocos_state(synth.VocosConfig(), 2)); v.to("cuda")
eng = VocoderEngine(d, v)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 512
hid = [torch.randn(nf, 768, device="cuda") for _ in range(B)]
for rep in range(5):
    e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
    e1.record()
    wavs, _ = eng.decode_batch(hid)
    e2.record()
    torch.cuda.synchronize()
    flop_v = 157.4e6 * B * nf
    print(f"vocoder {e1.elapsed_time(e2):.2f} ms for {B}x{nf} frames ({flop_v / (e1.elapsed_time(e2) * 1e-3) / 1e12:.0f} TFLOP/s)")
