"""SURVEY row a16 pinned by REFERENCE EXECUTION: tests/golden/avhubert_golden.pt holds outputs of the reference's own
hubert.py (AVHubertModel / SubModel / extract_finetune :695-755), wav2vec2.py (TransformerEncoder :818-905, layer :916-1038 with
apply_lora), resnet.py and multihead_attention.forward_lora, run on CPU by tests/golden/make_avhubert_golden.py (sources loaded
by path, unmodified).  CPU: the oracle restatement reproduces them; GPU: the CUDA path reproduces them within bf16 tolerance."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = os.path.join(HERE, "golden", "avhubert_golden.pt")


def _weights(g):
    from make_avhubert_golden import named_weight
    return {n: named_weight(n, shp, g["seed"]) for n, shp in g["named_shapes"]}


def test_oracle_avhubert_matches_reference_execution():
    from oracle import encoders as oe
    g = torch.load(GOLD)
    cfg = oe.AVHubertCfg(embed_dim=g["E"], ffn=g["ffn"], layers=g["layers"], heads=g["heads"], conv_pos=g["conv_pos"],
                         conv_pos_groups=g["conv_pos_groups"])
    m = oe.AVHubertVideo(cfg).eval()
    w = _weights(g)
    sd = m.state_dict()
    floats = {n for n, p in sd.items() if p.is_floating_point()}
    assert floats == set(w), (sorted(floats - set(w))[:5], sorted(set(w) - floats)[:5])   # same key layout as hubert.py builds
    with torch.no_grad():
        for n in floats:
            sd[n].copy_(w[n])
        front = m.feature_extractor_video(g["video"])
        got = m(g["video"])
    assert torch.allclose(front, g["sub_model_out"], atol=1e-5, rtol=1e-5)
    assert torch.allclose(got, g["x"], atol=1e-5, rtol=1e-5), (got - g["x"]).abs().max().item()
    assert g["x"].abs().mean().item() > 0.1                       # a non-degenerate fixture


def test_golden_fixture_is_reproducible_from_the_reference():
    """When /root/reference is mounted (build container), re-run the generator's model and diff it with the committed file."""
    if not os.path.isdir("/root/reference/av_hubert"):
        pytest.skip("reference not mounted")
    import types
    import make_avhubert_golden as mk
    hub, _, _ = mk.import_reference_avhubert()
    g = torch.load(GOLD)
    model = hub.AVHubertModel(mk.small_cfg(E=g["E"]), types.SimpleNamespace(sample_rate=25), [None])
    mk.attach_lora(model, g["E"])
    model.eval()
    mk.fill(model, g["seed"], keep=lambda n: n.startswith(mk.VIDEO_KEYS))
    with torch.no_grad():
        x, _, layers = model.extract_finetune(source={"video": g["video"], "audio": None})
    assert torch.equal(x, g["x"])
    assert all(torch.equal(a.transpose(0, 1), b) for a, b in zip(layers, g["layer_outputs"]))


@pytest.mark.gpu
def test_cuda_avhubert_matches_reference_execution():
    """Product AV-HuBERT (tcgen05 front-end / trunk / encoder with the LoRA K-extension, bf16) vs the reference's fp32 CPU
    run with the same weights: max |a - b| <= 3e-2 * max |b|."""
    from omni_avsr_b200.encoders import AVHubertArch, AVHubertVideoEncoder
    g = torch.load(GOLD)
    arch = AVHubertArch(g["E"], g["ffn"], g["layers"], g["heads"], g["conv_pos"], g["conv_pos_groups"], (64, 128, 256, 512))
    enc = AVHubertVideoEncoder(arch, "cuda", None, use_lora=True)
    w = {n: t.cuda() for n, t in _weights(g).items()}
    missing, unexpected = enc.load_state_dict(w, strict=False)
    assert not [k for k in missing if "num_batches_tracked" not in k], missing[:8]
    assert not unexpected, unexpected[:8]
    with torch.no_grad():
        got, pad, _ = enc.extract_finetune({"video": g["video"].cuda().bfloat16(), "audio": None})
    assert pad is None
    want = g["x"].cuda()
    err = (got.float() - want).abs().max().item()
    assert err <= 3e-2 * want.abs().max().item(), (err, want.abs().max().item())
