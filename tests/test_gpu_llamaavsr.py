"""GPU parity of the Llama-AVSR / Llama-MTSK drop-in (omni_avsr_b200/modeling_LlamaAVSR.py) against
the CPU oracle (oracle/llama_avsr.py + the oracle encoders / LLM) on a small configuration with identical weights: the
Matryoshka mean loss over every (video rate, audio rate) pair computed in ONE packed LLM pass, projector gradients,
the single-rate inference embeddings and greedy ids, and the error behaviour the mirror adds.

Tolerances as in tests/test_gpu_model.py: loss |a-b| <= 5e-2, gradients max|a-b| <= 1e-1*max|b| and cosine >= 0.97,
inference embeddings max|a-b| <= 3e-2*max|b| (two bf16 encoders + projector in front of them)."""
import pytest
import torch

from oracle import llama_avsr as ola
from tests._small import small_llamaavsr_module

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a.float().cpu() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-9)


@pytest.fixture(scope="module")
def pair():
    from oracle.pairing import oracle_from_product
    mod = small_llamaavsr_module()
    return mod, oracle_from_product(mod)


def _batch(mod, B=2, seconds=2.0, L=12, seed=7):
    from omni_avsr_b200.synthetic import synthetic_batch, to_device
    cpu = synthetic_batch(B, mod.tokenizer, seconds=seconds, text_len=L, seed=seed)
    return cpu, to_device(cpu, "cuda")


def _oracle_sequences(mod, oracle, cpu, is_trainval, test_ratio=None):
    a = mod.args
    tok = mod.tokenizer
    v = tok.vocab
    marker = (v["<audio>"], v["</audio>"], v["<video>"], v["</video>"])
    prompt_ids = tok(a.prompt_audiovisual, return_tensors="pt").input_ids[:, 1:-1]
    max_len = max(cpu["lengths"])
    if is_trainval:
        al = [oracle.audio_proj[i](oracle.encode_audio(cpu["audio"], max_len, r)) for i, r in enumerate(a.downsample_ratio_audio)]
        vl = [oracle.video_proj[i](oracle.encode_video(cpu["video"], r)) for i, r in enumerate(a.downsample_ratio_video)]
        return ola.prepare_inputs(oracle.llm.model.embed_tokens, cpu["tokens"], cpu["labels"], al, vl, prompt_ids, marker,
                                  False, "audiovisual", True, True)
    rv, ra = test_ratio
    at = oracle.audio_proj[a.downsample_ratio_audio.index(ra)](oracle.encode_audio(cpu["audio"], max_len, ra))
    vt = oracle.video_proj[a.downsample_ratio_video.index(rv)](oracle.encode_video(cpu["video"], rv))
    return ola.prepare_inputs(oracle.llm.model.embed_tokens, cpu["tokens"][:, :1], None, at, vt, prompt_ids, marker, False,
                              "audiovisual", False, False)


def test_state_dict_has_no_prompt_buffers_and_reference_keys(pair):
    mod, _ = pair
    keys = set(mod.model.state_dict().keys())
    assert not any(k.startswith("prompt_") for k in keys)
    for k in ["audio_proj.0.0.weight", "audio_proj.1.2.bias", "video_proj.1.0.weight",
              "llm.model.layers.0.self_attn.lora_down_Q.weight", "llm.model.layers.1.self_attn.lora_up_V.weight"]:
        assert k in keys, k


def test_mtsk_loss_and_projector_grads(pair):
    mod, oracle = pair
    cpu, gpu = _batch(mod)
    oracle.zero_grad()
    seqs, labs = _oracle_sequences(mod, oracle, cpu, True)
    assert len(seqs) == 4
    want = ola.train_loss(oracle.llm, seqs, labs, True)
    want.backward()
    mod.zero_grad_flat()
    got = mod.model(gpu, is_trainval=True)
    got.backward()
    assert abs(got.item() - want.item()) <= 5e-2, (got.item(), want.item())
    m = mod.model
    for (g, w, name) in [(m.audio_proj[0][2].weight.grad, oracle.audio_proj[0][2].weight.grad, "audio_proj.0.2"),
                         (m.audio_proj[1][0].weight.grad, oracle.audio_proj[1][0].weight.grad, "audio_proj.1.0"),
                         (m.video_proj[1][2].bias.grad, oracle.video_proj[1][2].bias.grad, "video_proj.1.2.bias")]:
        assert w is not None and g is not None, name
        assert _rel(g, w) <= 1e-1, (name, _rel(g, w))
        cos = torch.nn.functional.cosine_similarity(g.float().cpu().flatten(), w.float().flatten(), dim=0).item()
        assert cos >= 0.97, (name, cos)
    # the packed pass saw exactly the reference's four sequence lengths, video rates outer / audio rates inner
    emb, lab = m.prepare_inputs(gpu, True)
    assert [e.shape[1] for e in emb] == [s.shape[1] for s in seqs]
    for l_gpu, l_cpu in zip(lab, labs):
        assert torch.equal(l_gpu.cpu(), l_cpu)


def test_inference_embeddings_and_greedy(pair):
    mod, oracle = pair
    cpu, gpu = _batch(mod, B=1)
    mod.on_test_epoch_start()
    with torch.no_grad():
        emb, _ = mod.model.prepare_inputs(gpu, False, test_ratio_matry=[5, 4])
        want, _ = _oracle_sequences(mod, oracle, cpu, False, test_ratio=[5, 4])
        assert emb.shape == want.shape
        assert _rel(emb, want) <= 3e-2, _rel(emb, want)
        ids = mod.test_step(gpu)
        v = mod.tokenizer.vocab
        ref, margins = oracle.llm.generate(want, 8, v["<|end_of_text|>"], v["<pad>"], modality=None, return_margins=True)
    n = min(ids.shape[1], ref.shape[1])
    differs = ids.cpu()[:, :n] != ref[:, :n]
    if differs.any():
        first = int(differs[0].float().argmax())
        assert float(margins[0][first]) < 0.05, (ids, ref, margins)


def test_error_behaviour():
    from omni_avsr_b200 import lightning_LlamaAVSR as pl_mod
    with pytest.raises(KeyError):                       # unknown rate at inference (:518-521)
        mod = small_llamaavsr_module()
        _, gpu = _batch(mod, B=1)
        mod.model.prepare_inputs(gpu, False, test_ratio_matry=[3, 4])
    with pytest.raises(NotImplementedError):            # Matryoshka layouts need a BOS token
        pl_mod.ModelModule_LLM(pl_mod.make_args(llm_model="Qwen/Qwen2.5-3B", is_matryoshka=True, modality="audio",
                                                downsample_ratio_audio=[4, 16]))
