"""Vocoder parity on the B200 (through the C ABI) against the fp32 CPU oracle.

Tolerance (north-star): waveform RMS error <= 1e-3 given identical vocoder input.  The kernels keep the residual
stream, LayerNorm, GELU, exp/sin/cos and the ISTFT in fp32; fp16 appears only as tensor-core GEMM operands.
"""
import pytest
import torch

from chatttsplus_b200 import synth
from oracle import ctp_oracle as O

pytestmark = pytest.mark.gpu


def _models(dcfg, vcfg, seed=1):
    from chatttsplus_b200.vocoder import DVAE, Vocos, VocoderEngine
    dsd = synth.make_dvae_state(dcfg, seed=seed)
    vsd = synth.make_vocos_state(vcfg, seed=seed + 1)
    kw = dict(decoder_config=dict(idim=dcfg.idim, odim=dcfg.odim, hidden=dcfg.hidden, n_layer=dcfg.n_layer, bn_dim=dcfg.bn_dim), dim=dcfg.dim)
    if dcfg.vq:
        kw["vq_config"] = dict(dim=dcfg.vq_dim, levels=list(dcfg.vq_levels), G=dcfg.vq_G, R=dcfg.vq_R)
    d = DVAE(**kw)
    d.load_state_dict(dsd)
    d.to("cuda")
    v = Vocos(backbone_config=dict(input_channels=100, dim=vcfg.dim, intermediate_dim=vcfg.intermediate_dim, num_layers=vcfg.num_layers),
              head_config=dict(dim=vcfg.dim, n_fft=1024, hop_length=256, padding="center"))
    v.load_state_dict(vsd)
    v.to("cuda")
    return d, v, VocoderEngine(d, v), dsd, vsd


def _rms(x):
    return float(x.double().pow(2).mean().sqrt())


def test_dvae_decode_varlen_batch_matches_oracle_per_utterance():
    dcfg = synth.DVAEConfig()
    d, v, eng, dsd, vsd = _models(dcfg, synth.VocosConfig(), seed=11)
    g = torch.Generator().manual_seed(3)
    lens = [5, 1, 17, 40]
    hid = [torch.randn(n, 768, generator=g) for n in lens]
    _, mels = eng.decode_batch([h.cuda() for h in hid], want_wav=False, want_mel=True)
    for h, m in zip(hid, mels):
        ref = O.dvae_decode(dsd, h.permute(1, 0)[None])[0].permute(1, 0)  # [2n, 100]
        err = _rms(m.cpu() - ref) / _rms(ref)
        print("dvae mel rel rms", err, "ref rms", _rms(ref))
        assert m.shape == ref.shape
        assert err < 3e-3


def test_dvae_call_interface_matches_reference_shape():
    dcfg = synth.DVAEConfig(n_layer=2)
    d, v, eng, dsd, vsd = _models(dcfg, synth.VocosConfig(num_layers=1), seed=12)
    x = torch.randn(2, 768, 9, generator=torch.Generator().manual_seed(1))
    mel = d(x.cuda())
    ref = O.dvae_decode(dsd, x, n_layer=2)
    assert mel.shape == (2, 100, 18)
    assert _rms(mel.cpu() - ref) / _rms(ref) < 3e-3


def test_codes_path_gfsq_embed_matches_oracle():
    dcfg = synth.DVAEConfig.codes_model()
    d, v, eng, dsd, vsd = _models(dcfg, synth.VocosConfig(num_layers=2), seed=13)
    g = torch.Generator().manual_seed(4)
    ids = [torch.randint(0, 625, (n, 4), generator=g) for n in (7, 23)]
    _, mels = eng.decode_batch([i.cuda() for i in ids], want_wav=False, want_mel=True)
    for i, m in zip(ids, mels):
        ref = O.dvae_decode(dsd, i.permute(1, 0)[None], vq=True)[0].permute(1, 0)
        err = _rms(m.cpu() - ref) / _rms(ref)
        print("codes mel rel rms", err)
        assert err < 3e-3


def test_vocos_decode_matches_oracle_waveform():
    vcfg = synth.VocosConfig()
    d, v, eng, dsd, vsd = _models(synth.DVAEConfig(n_layer=1), vcfg, seed=14)
    g = torch.Generator().manual_seed(5)
    mel = torch.randn(2, 100, 60, generator=g)
    wav = v.decode(mel.cuda())
    ref = O.vocos_decode(vsd, mel)
    assert wav.shape == ref.shape == (2, 256 * 59)
    e = _rms(wav.cpu() - ref)
    print("vocos wav rms err", e, "ref rms", _rms(ref), "peak", float(ref.abs().max()))
    assert e <= 1e-3, "waveform RMS error above the north-star tolerance (1e-3)"
    assert e / _rms(ref) <= 5e-3


def test_hidden_to_wav_end_to_end_batch():
    dcfg, vcfg = synth.DVAEConfig(), synth.VocosConfig()
    d, v, eng, dsd, vsd = _models(dcfg, vcfg, seed=15)
    g = torch.Generator().manual_seed(6)
    lens = [12, 3, 31]
    hid = [torch.randn(n, 768, generator=g) for n in lens]
    wavs, _ = eng.decode_batch([h.cuda() for h in hid])
    for h, w in zip(hid, wavs):
        ref = O.decode_to_wav(dsd, vsd, h)
        assert w.shape == ref.shape == (256 * (2 * h.shape[0] - 1),)
        e = _rms(w.cpu() - ref)
        print("e2e wav rms err", e, "ref rms", _rms(ref))
        # identical vocoder input (the hiddens); error budget covers DVAE + Vocos
        assert e <= 1e-3
        assert e / _rms(ref) <= 1e-2


def test_vocoder_grouping_when_workspace_is_small():
    """Utterances are processed in groups that fit the workspace; results do not depend on the grouping."""
    from chatttsplus_b200.vocoder import VocoderEngine
    dcfg, vcfg = synth.DVAEConfig(n_layer=2), synth.VocosConfig(num_layers=2)
    d, v, eng, dsd, vsd = _models(dcfg, vcfg, seed=16)
    g = torch.Generator().manual_seed(7)
    hid = [torch.randn(n, 768, generator=g).cuda() for n in (20, 20, 20, 9)]
    big, _ = eng.decode_batch(hid)
    small_eng = VocoderEngine(d, v, max_frames=96)
    small, _ = small_eng.decode_batch(hid)
    for a, b in zip(big, small):
        assert torch.allclose(a, b, atol=1e-6)
