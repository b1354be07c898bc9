"""Pins the encoder oracle against independent implementations available in this image:
the installed transformers WhisperFeatureExtractor / WhisperModel(config).encoder, and the reference's own
av_hubert/avhubert/resnet.py (imported by path when /root/reference is present; skipped on the GPU box)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import encoders as oe

transformers = pytest.importorskip("transformers")


def test_log_mel_matches_transformers_feature_extractor():
    fe = transformers.WhisperFeatureExtractor(feature_size=80)
    g = torch.Generator().manual_seed(0)
    audio = torch.randn(2, 16000 * 3 + 77, generator=g)
    want = fe(audio.numpy(), return_tensors="pt", sampling_rate=16000).input_features
    got = oe.log_mel(audio)
    assert got.shape == want.shape == (2, 80, 3000)
    assert torch.allclose(got, want, atol=2e-4, rtol=1e-4)
    assert np.allclose(oe.mel_filters(), fe.mel_filters, atol=1e-7)


def test_whisper_encoder_matches_transformers():
    cfg = transformers.WhisperConfig(d_model=64, encoder_layers=2, encoder_attention_heads=4, encoder_ffn_dim=128,
                                     num_mel_bins=80, max_source_positions=1500, decoder_layers=1,
                                     decoder_attention_heads=4, decoder_ffn_dim=128, vocab_size=100, pad_token_id=0,
                                     bos_token_id=1, eos_token_id=2, decoder_start_token_id=1)
    hf = transformers.WhisperModel(cfg).encoder.eval()
    mine = oe.WhisperEncoder(oe.WhisperCfg(64, 2, 4, 128))
    missing, unexpected = mine.load_state_dict(hf.state_dict(), strict=True)
    x = torch.randn(2, 80, 3000)
    with torch.no_grad():
        want = hf(x).last_hidden_state
        got = mine(x)
    assert torch.allclose(got, want, atol=1e-4, rtol=1e-4)
    assert torch.allclose(oe.sinusoids(1500, 64), hf.embed_positions.weight, atol=1e-6)


@pytest.mark.skipif(not os.path.exists("/root/reference/av_hubert/avhubert/resnet.py"), reason="reference not mounted")
def test_resencoder_matches_reference_file():
    spec = importlib.util.spec_from_file_location("ref_resnet", "/root/reference/av_hubert/avhubert/resnet.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(0)
    ref = mod.ResEncoder("prelu", None).eval()
    mine = oe.ResEncoder().eval()
    mine.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(1, 1, 6, 88, 88)
    with torch.no_grad():
        assert torch.allclose(mine(x), ref(x), atol=1e-5, rtol=1e-5)


def test_avhubert_lora_zero_is_identity_and_shapes():
    cfg = oe.AVHubertCfg(embed_dim=64, ffn=128, layers=2, heads=4, conv_pos=8, conv_pos_groups=4)
    m = oe.AVHubertVideo(cfg, widths=(8, 16, 16, 32)).eval()
    x = torch.randn(2, 1, 5, 88, 88)
    with torch.no_grad():
        y0 = m(x)
        for l in m.encoder.layers:
            torch.nn.init.normal_(l.self_attn.lora_up_Q.weight)       # up != 0 but down == 0 => no-op
        y1 = m(x)
    assert y0.shape == (2, 5, 64)
    assert torch.equal(y0, y1)
    keys = set(m.state_dict().keys())
    assert "encoder.pos_conv.0.weight_g" in keys and "encoder.pos_conv.0.weight_v" in keys
    assert "feature_extractor_video.resnet.frontend3D.0.weight" in keys
    assert "encoder.layers.0.self_attn.lora_down_Q.weight" in keys
