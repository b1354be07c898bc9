"""Checkpoint compatibility: the drop-in modules load the state dicts of the models the reference
starts from -- built here from the INSTALLED transformers classes (random init: there are no weight files offline) -- and
reproduce those models' outputs; the fairseq AV-HuBERT checkpoint layout is exercised with a synthetic `{"model": ...}` file
that carries the unused audio / pre-training keys; checkpoint averaging follows utils/avg_checkpoints.py.

Tolerances: encoder features and logits max|a-b| <= 2e-2*max|b| (bf16 kernels vs the library model in bf16)."""
import os

import pytest
import torch

from omni_avsr_b200 import checkpoints as ck


def _rel(a, b):
    return (a.float().cpu() - b.float().cpu()).abs().max().item() / max(b.float().abs().max().item(), 1e-9)


# ------------------------------------------------------------------------------------------------ CPU
def test_average_checkpoints_matches_reference_semantics(tmp_path):
    g = torch.Generator().manual_seed(0)
    paths, sds = [], []
    for i in range(3):
        sd = {"model.a.weight": torch.randn(4, 3, generator=g), "model.b.bias": torch.randn(5, generator=g).bfloat16(),
              "model.steps": torch.tensor([10 + i, 21 + i]), "optimizer.junk": torch.zeros(1)}
        p = tmp_path / f"epoch={i}.ckpt"
        torch.save({"state_dict": sd, "epoch": i}, p)
        paths.append(str(p))
        sds.append(sd)
    avg = ck.average_checkpoints(paths)
    assert set(avg) == {"a.weight", "b.bias", "steps"}                       # `model.` stripped, other keys dropped
    assert torch.allclose(avg["a.weight"], sum(s["model.a.weight"] for s in sds) / 3)
    want_b = (sds[0]["model.b.bias"] + sds[1]["model.b.bias"] + sds[2]["model.b.bias"]) / 3     # bf16 running sum, as :22
    assert torch.equal(avg["b.bias"], want_b)
    assert torch.equal(avg["steps"], torch.tensor([(10 + 11 + 12) // 3, (21 + 22 + 23) // 3]))  # integer floor division

    class A:
        exp_dir, exp_name, max_epochs = str(tmp_path), "", 3
    out = ck.ensemble_original(A, num_average_epochs=3)
    assert os.path.basename(out) == "model_avg_3.pth"
    again = torch.load(out)
    assert torch.equal(again["a.weight"], avg["a.weight"])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_hf_whisper_state_dict_loads_and_matches():
    from transformers import WhisperConfig, WhisperModel
    from omni_avsr_b200.encoders import WhisperArch, WhisperEncoder
    torch.manual_seed(0)
    cfg = WhisperConfig(d_model=128, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=256, decoder_layers=1,
                        decoder_attention_heads=2, decoder_ffn_dim=64, num_mel_bins=80, max_source_positions=1500,
                        vocab_size=100, pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1)
    hf = WhisperModel(cfg).bfloat16().eval()
    enc = WhisperEncoder(WhisperArch(128, 2, 2, 256), "cuda")
    missing, ignored = ck.load_whisper_encoder(enc, hf.state_dict())
    assert not missing, missing
    assert all(k.startswith("decoder.") for k in ignored), [k for k in ignored if not k.startswith("decoder.")][:4]
    feats = torch.randn(2, 80, 3000, generator=torch.Generator().manual_seed(1)).bfloat16()
    with torch.no_grad():
        want = hf.encoder(feats).last_hidden_state
        got = enc(feats.cuda()).last_hidden_state
    assert _rel(got, want) <= 2e-2, _rel(got, want)
    # the `model.`-prefixed layout of WhisperForConditionalGeneration loads the same way
    enc2 = WhisperEncoder(WhisperArch(128, 2, 2, 256), "cuda")
    m2, _ = ck.load_whisper_encoder(enc2, {"model." + k: v for k, v in hf.state_dict().items()})
    assert not m2
    assert torch.equal(enc2.state_dict()["layers.1.fc2.weight"], enc.state_dict()["layers.1.fc2.weight"])


@pytest.mark.gpu
@pytest.mark.parametrize("family", ["llama", "qwen2"])
def test_hf_llm_state_dict_loads_and_matches(family):
    from transformers import LlamaConfig, LlamaForCausalLM, Qwen2Config, Qwen2ForCausalLM
    from omni_avsr_b200 import Llama_LoRA as pl
    from omni_avsr_b200 import Qwen_LoRA as pq
    torch.manual_seed(0)
    V0, V = 300, 305                                  # checkpoint vocabulary, vocabulary after add_special_tokens
    if family == "llama":
        hcfg = LlamaConfig(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                           num_key_value_heads=1, vocab_size=V0, rms_norm_eps=1e-5, rope_theta=500000.0, head_dim=64,
                           max_position_embeddings=16384, tie_word_embeddings=True,
                           rope_scaling=dict(rope_type="llama3", factor=32.0, low_freq_factor=1.0, high_freq_factor=4.0,
                                             original_max_position_embeddings=8192))
        hf = LlamaForCausalLM(hcfg)
        arch = pl.LLMArch("llama", 256, 512, 2, 4, 1, V, 1e-5, 500000.0, 64,
                          dict(factor=32.0, low_freq_factor=1.0, high_freq_factor=4.0, original_max_position_embeddings=8192),
                          False, True, max_position_embeddings=16384, inv_freq_dtype="bf16")
        model = pl.LlamaForCausalLM_lora(arch, pl.LoRA_config(4, 2, True, False, True, True))
    else:
        hcfg = Qwen2Config(hidden_size=512, intermediate_size=512, num_hidden_layers=2, num_attention_heads=8,
                           num_key_value_heads=1, vocab_size=V0, rms_norm_eps=1e-6, rope_theta=1000000.0,
                           max_position_embeddings=4096, tie_word_embeddings=True)
        hf = Qwen2ForCausalLM(hcfg)
        arch = pl.LLMArch("qwen2", 512, 512, 2, 8, 1, V, 1e-6, 1000000.0, 64, None, True, True, max_position_embeddings=4096,
                          inv_freq_dtype="fp32")
        model = pq.Qwen2ForCausalLM_lora(arch, pq.QwenLoRA_config(8, 2, IS_QWEN25_3B=True, IS_TASK_SPECIFIC=True, SHARED_LORA=True))
    with torch.no_grad():
        for n, p in hf.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
    hf.tie_weights()
    hf = hf.bfloat16().eval()
    missing, ignored = ck.load_llm(model, hf.state_dict())
    assert all("lora" in k for k in missing), [k for k in missing if "lora" not in k][:4]
    assert not ignored, ignored[:4]
    assert torch.equal(model.model.embed_tokens.weight[:V0].cpu(), hf.model.embed_tokens.weight)
    x_ids = torch.randint(0, V0, (2, 19), generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        want = hf(input_ids=x_ids).logits                                   # LoRA-down is zero-initialised: adapted == base
        got = model(inputs_embeds=model.model.embed_tokens(x_ids.cuda()), modality="audio").logits
    assert _rel(got[..., :V0], want) <= 2e-2, _rel(got[..., :V0], want)


@pytest.mark.gpu
def test_fairseq_avhubert_checkpoint_layout_loads_and_matches():
    from oracle import encoders as oe
    from omni_avsr_b200.encoders import AVHubertArch, AVHubertVideoEncoder
    torch.manual_seed(0)
    cfg = oe.AVHubertCfg(128, 256, 2, 2, 16, 4)
    ref = oe.AVHubertVideo(cfg, (16, 32, 32, 64)).eval()
    with torch.no_grad():
        for n, b in ref.named_buffers():
            if n.endswith("running_var"):
                b.uniform_(0.5, 1.5)
            elif n.endswith("running_mean"):
                b.normal_(0, 0.1)
    base = {k: v for k, v in ref.state_dict().items() if "lora_" not in k}
    ckpt = {"model": dict(base), "cfg": {"note": "synthetic"}, "args": None}
    # keys a real large_vox_iter5.pt also carries, none of them on the video-only extract_finetune path
    ckpt["model"].update({"mask_emb": torch.zeros(128), "label_embs_concat": torch.zeros(10, 16),
                          "final_proj.weight": torch.zeros(16, 128), "feature_extractor_audio.proj.weight": torch.zeros(128, 104)})
    enc = AVHubertVideoEncoder(AVHubertArch(128, 256, 2, 2, 16, 4, (16, 32, 32, 64)), "cuda", None, use_lora=True)
    missing, ignored = ck.load_avhubert(enc, ckpt)
    assert all("lora_" in k for k in missing), [k for k in missing if "lora_" not in k][:6]
    assert set(ignored) >= {"mask_emb", "label_embs_concat", "final_proj.weight", "feature_extractor_audio.proj.weight"}
    assert not [k for k in ignored if k.startswith(ck.AVHUBERT_VIDEO_PREFIXES)], ignored
    video = ((torch.rand(2, 1, 20, 88, 88, generator=torch.Generator().manual_seed(3)) - 0.421) / 0.165).bfloat16()
    with torch.no_grad():
        for layer in ref.encoder.layers:       # reference init of the adapters: down = 0 -> LoRA is the identity
            layer.self_attn.lora_down_Q.weight.zero_()
            layer.self_attn.lora_down_V.weight.zero_()
        want = ref.bfloat16()(video)
        got, _, _ = enc.extract_finetune({"video": video.cuda(), "audio": None})
    assert _rel(got, want) <= 3e-2, _rel(got, want)


@pytest.mark.gpu
def test_lightning_checkpoint_round_trip_through_averaging(tmp_path):
    from tests._small import small_module
    mod = small_module()
    paths = []
    for i in range(2):
        p = tmp_path / f"epoch={i}.ckpt"
        torch.save(ck.lightning_checkpoint(mod), p)
        paths.append(str(p))
    avg = ck.average_checkpoints(paths)
    sd = mod.model.state_dict()
    assert set(avg) == set(sd)
    mod.model.load_state_dict({k: v.cuda() for k, v in avg.items()})                 # lightning_OmniAVSR.py:148-150
    for k in ("audio_proj.0.0.weight", "llm.model.layers.0.self_attn.lora_up_Q.audio.weight", "prompt_audio"):
        assert torch.equal(mod.model.state_dict()[k].cpu(), avg[k])


# ---------------------------------------------------------------------------------------------------------------
# round-2 additions: real `model_avg_N.pth` files carry AV-HuBERT pre-training tensors, prompt buffers must
# follow the embedding table, the fused optimizer must not touch frozen tensors
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_reference_state_dict_with_unused_avhubert_tensors_loads_strictly():
    """The reference's `video_encoder` is the full fairseq AVHubertModel (remove_pretraining_modules is never called), so
    its state dict has mask_emb / label_embs_concat / final_proj / feature_extractor_audio; a strict load must accept it."""
    from tests._small import small_module
    mod = small_module()
    sd = {k: v.clone() for k, v in mod.model.state_dict().items()}
    sd.update({"video_encoder.mask_emb": torch.zeros(128), "video_encoder.label_embs_concat": torch.zeros(10, 16),
               "video_encoder.final_proj.weight": torch.zeros(16, 128), "video_encoder.final_proj.bias": torch.zeros(16),
               "video_encoder.feature_extractor_audio.proj.weight": torch.zeros(128, 104),
               "video_encoder.feature_extractor_audio.proj.bias": torch.zeros(128)})
    res = mod.model.load_state_dict(sd)                      # strict=True
    assert not res.missing_keys and not res.unexpected_keys
    sd["llm.model.layers.0.bogus.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        mod.model.load_state_dict(sd)


@pytest.mark.gpu
def test_prompt_buffers_follow_the_embedding_table():
    """The reference embeds the task prompts AFTER from_pretrained (modeling_OmniAVSR.py:218-221); loading base weights
    into the mirror must refresh the prompt buffers, otherwise they keep the random-init rows."""
    from omni_avsr_b200 import checkpoints
    from tests._small import small_module
    mod = small_module()
    m = mod.model
    g = torch.Generator().manual_seed(0)
    new = (torch.randn(m.llm.model.embed_tokens.weight.shape, generator=g) * 0.02).bfloat16()
    old_prompt = m.prompt_audio.clone()
    checkpoints.load_llm(m.llm, {"model.embed_tokens.weight": new}, owner=m)
    ids = m._prompt_ids["prompt_audio"].cpu()
    assert torch.equal(m.prompt_audio.cpu(), new[ids])
    assert not torch.equal(m.prompt_audio, old_prompt)
    ids = m._prompt_ids["prompt_audiovisual"].cpu()
    assert torch.equal(m.prompt_audiovisual.cpu(), new[ids])


@pytest.mark.gpu
def test_optimizer_leaves_frozen_tensors_alone_and_zero_grad_keeps_the_flat_views():
    """torch.optim.AdamW skips parameters without a gradient (weight decay included); the fused kernel therefore runs
    over the trainable spans of the flat buffer only.  zero_grad(set_to_none=True) must not detach the flat views."""
    from omni_avsr_b200.synthetic import synthetic_batch, to_device
    from tests._small import small_module
    mod = small_module(seed=2)
    m = mod.model
    m._unfreeze_PETF(["peft_llm"])                           # AV-HuBERT adapters frozen
    mod.args.weight_decay = 0.5
    mod.configure_optimizers()
    spans = m.flat.trainable_spans()
    covered = sum(b - a for a, b in spans)
    assert 0 < covered < m.flat.used
    avh = [p.detach().clone() for p in m.video_encoder.lora_parameters()]
    llm_before = m.llm.model.layers[0].self_attn.lora_up.detach().clone()
    gpu = to_device(synthetic_batch(2, mod.tokenizer, seconds=2.0, text_len=12, seed=7), "cuda")
    mod.zero_grad(set_to_none=True)
    att = m.llm.model.layers[0].self_attn
    assert att.lora_down.grad is not None and att.lora_down.grad.data_ptr() >= m.flat.grad.data_ptr()
    mod.train_step(gpu, rates=(4, 2), lr=0.1)
    for p, q in zip(m.video_encoder.lora_parameters(), avh):
        assert torch.equal(p.detach(), q)                    # no drift under lr * wd = 0.05
    assert not torch.equal(m.llm.model.layers[0].self_attn.lora_up.detach(), llm_before)
