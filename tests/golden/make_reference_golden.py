"""Generates tests/golden/reference_golden.pt by EXECUTING THE REFERENCE'S OWN SOURCES from /root/reference (read-only,
present in the build container only) on seeded inputs:

    python tests/golden/make_reference_golden.py

What runs is the unmodified reference code, imported through the name-only shims of `_ref_compat.py`:
  * `Omni_AVSR/Llama_LoRA.py`  LlamaForCausalLM_lora.forward (shared / task-specific / hybrid Omni-LoRA) and
    `prepare_inputs_for_generation` (driven by a minimal greedy loop, see `_greedy`);
  * `Omni_AVSR/Qwen_LoRA.py`   Qwen2ForCausalLM_lora.forward (hybrid, qkv bias);
  * `Omni_AVSR/modeling_OmniAVSR.py`  AVSR_LLMs.__init__ / forward / prepare_inputs (train + infer) / encode_audio /
    encode_video for avg-pooling and stack Matryoshka compression, Llama and Qwen layouts.  `from_pretrained` and the
    fairseq checkpoint loader are patched to return tiny random-init models (there are no checkpoints offline), the video
    encoder is a stand-in (its output is stored, the test feeds it back), `.cuda()` is patched to a no-op;
  * `Omni_AVSR/modeling_LlamaAVSR.py`  AVSR_LLMs (Llama-AVSR and Llama-MTSK: every rate / rate pair per step), same
    stand-ins;
  * `av_hubert/fairseq/fairseq/modules/multihead_attention.py`  MultiheadAttention.forward_lora;
  * `datamodule/transforms.py`  VideoTransform / AudioTransform (train and val pipelines, seeded RNGs).

Weights are NOT stored: they are regenerated from a seed by `golden_weights` (CPU mt19937 `randn`), a checksum guards
against RNG drift.  tests/test_reference_golden.py compares the oracle with this file on CPU and the CUDA path with it
on the GPU.  The fixture travels to the GPU box; /root/reference does not."""
import os
import random
import sys
import types

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(HERE, "reference_golden.pt")

LLAMA = dict(hidden_size=256, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=1,
             vocab_size=200, rms_norm_eps=1e-5, rope_theta=500000.0, max_position_embeddings=16384, head_dim=64,
             tie_word_embeddings=True,
             rope_scaling=dict(rope_type="llama3", factor=32.0, low_freq_factor=1.0, high_freq_factor=4.0,
                               original_max_position_embeddings=8192))
# Qwen2.5-3B-like GQA: Hkv*hd = H // 8 (the reference's IS_QWEN25_3B lora_up_V width, Qwen_LoRA.py:464-475)
QWEN = dict(hidden_size=512, intermediate_size=256, num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=1,
            vocab_size=200, rms_norm_eps=1e-6, rope_theta=1000000.0, max_position_embeddings=4096,
            tie_word_embeddings=True)


def golden_weights(named_shapes, seed, std_of=None):
    """Deterministic weights for a list of (name, shape): 1-D tensors whose name ends in 'norm.weight' or
    'layer_norm.weight' are 1 + 0.1*randn, biases 0.02*randn, LoRA matrices 0.05*randn, other matrices 0.02*randn
    (HF init would zero lora_down and make the adapters invisible)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in named_shapes:
        r = torch.randn(tuple(shape), generator=g)
        if len(shape) == 1 and ("norm" in name and name.endswith("weight")):
            out[name] = 1.0 + 0.1 * r
        elif "lora" in name:
            out[name] = 0.05 * r
        else:
            out[name] = 0.02 * r
    return out


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def load_golden_weights(module, seed, dtype=torch.bfloat16, skip=()):
    """Fill `module` (reference, oracle or product) from the seed; returns (named_shapes, checksum).  Tied lm_head is
    skipped by name so that the three implementations see the same list."""
    named = [(n, tuple(p.shape)) for n, p in module.state_dict().items()
             if n != "lm_head.weight" and not any(s in n for s in skip) and p.is_floating_point()]
    w = golden_weights(named, seed)
    with torch.no_grad():
        sd = module.state_dict()
        for n, _ in named:
            sd[n].copy_(w[n].to(dtype).to(sd[n].dtype))
    return named, checksum({n: w[n].to(dtype) for n, _ in named})


def _greedy(model, inputs_embeds, max_new_tokens, eos_id, pad_id, modality):
    """Minimal greedy driver with transformers-4.43.1 `_sample` semantics (unfinished-sequence tracking, pad after EOS,
    stop when all finished), calling the REFERENCE's prepare_inputs_for_generation (Llama_LoRA.py:400-444 /
    Qwen_LoRA.py:207-251) and forward with a DynamicCache each step.  The installed transformers-5.5 `generate` no longer
    passes `cache_position`, which the reference's hook requires."""
    from transformers.cache_utils import DynamicCache
    B, S0, _ = inputs_embeds.shape
    input_ids = torch.zeros(B, 0, dtype=torch.long)
    attention_mask = torch.ones(B, S0, dtype=torch.long)
    cache_position = torch.arange(S0)
    past = DynamicCache()
    unfinished = torch.ones(B, dtype=torch.long)
    margins = []
    for _ in range(max_new_tokens):
        mi = model.prepare_inputs_for_generation(input_ids, past_key_values=past, attention_mask=attention_mask,
                                                 inputs_embeds=inputs_embeds, cache_position=cache_position,
                                                 modality=modality)
        out = model(**mi, return_dict=True)
        logits = out.logits[:, -1, :].float()
        top2 = logits.topk(2, dim=-1).values
        margins.append(top2[:, 0] - top2[:, 1])
        nxt = logits.argmax(-1)
        nxt = nxt * unfinished + pad_id * (1 - unfinished)
        input_ids = torch.cat([input_ids, nxt[:, None]], dim=1)
        attention_mask = torch.cat([attention_mask, torch.ones(B, 1, dtype=torch.long)], dim=1)
        cache_position = cache_position[-1:] + 1
        unfinished = unfinished & (nxt != eos_id).long()
        if unfinished.max() == 0:
            break
    return input_ids, torch.stack(margins, 1)


def _patch_ref(ll, ql):
    import transformers.models.llama.modeling_llama as ml
    import transformers.models.qwen2.modeling_qwen2 as mq

    def ucm(self, attention_mask, input_tensor, cache_position, past_key_values, output_attentions):
        # 4.43.1 `_ignore_causal_mask_sdpa`: no mask / all-ones mask -> None (SDPA is_causal decides)
        assert attention_mask is None or bool((attention_mask == 1).all())
        return None
    ml.LlamaModel._update_causal_mask = ucm
    mq.Qwen2Model._update_causal_mask = ucm
    tied = {"lm_head.weight": "model.embed_tokens.weight"}      # 5.5 wants a dict, the reference declares a list
    ll.LlamaForCausalLM_lora._tied_weights_keys = tied
    ql.Qwen2ForCausalLM_lora._tied_weights_keys = tied


def build_llm_cases(ll, ql):
    from transformers import LlamaConfig, Qwen2Config
    cases = {}
    specs = [
        ("llama_S", "llama", ll.LoRA_config(4, 2, True, False, False, False)),
        ("llama_T", "llama", ll.LoRA_config(4, 2, True, False, True, False)),
        ("llama_ST", "llama", ll.LoRA_config(4, 2, True, False, True, True)),
        ("qwen_ST", "qwen2", ql.QwenLoRA_config(8, 2, False, False, True, False, False, False, True, True)),
    ]
    for i, (name, fam, lc) in enumerate(specs):
        if fam == "llama":
            cfg = LlamaConfig(**LLAMA)
            m = ll.LlamaForCausalLM_lora(cfg, lc)
        else:
            cfg = Qwen2Config(**QWEN)
            m = ql.Qwen2ForCausalLM_lora(cfg, lc)
        m.tie_weights()         # what `from_pretrained` (the reference's only constructor path, :203) does last
        m = m.bfloat16().eval()
        assert m.lm_head.weight is m.model.embed_tokens.weight
        named, csum = load_golden_weights(m, 100 + i)
        g = torch.Generator().manual_seed(200 + i)
        H = cfg.hidden_size
        x = (torch.randn(2, 21, H, generator=g) * 0.5).bfloat16()
        lab = torch.randint(0, cfg.vocab_size, (2, 21), generator=g)
        lab[:, :9] = -100
        lab[1, -2:] = -100
        c = dict(family=fam, lora=dict(vars(lc)), seed=100 + i, named_shapes=named, checksum=csum, x=x, labels=lab,
                 logits={}, loss={}, greedy={}, margins={})
        mods = ("audio", "video", "audiovisual") if lc.IS_TASK_SPECIFIC else (None,)
        with torch.no_grad():
            for t in mods:
                o = m(inputs_embeds=x, labels=lab, modality=t)
                c["logits"][t] = o.logits.clone()
                c["loss"][t] = o.loss.clone()
                ids, mg = _greedy(m, x[:, :13], 8, cfg.vocab_size - 1, cfg.vocab_size - 2, t)
                c["greedy"][t], c["margins"][t] = ids, mg
        cases[name] = c
    return cases


class _Tok:
    """Stand-in tokenizer: `tok(text, return_tensors='pt').input_ids`, `len(tok)`, `tok.vocab[name]`."""

    def __init__(self, base_vocab, is_qwen, prompts):
        self.is_qwen = is_qwen
        names = ["<pad>", "<audio>", "</audio>", "<video>", "</video>"]
        self.vocab = {n: base_vocab + i for i, n in enumerate(names)}
        self.vocab["<|begin_of_text|>"] = 1
        self.vocab["<|end_of_text|>"] = 2
        self.vocab["<|endoftext|>"] = 2
        self.n = base_vocab + len(names)
        self.prompts = prompts

    def __len__(self):
        return self.n

    def __call__(self, text, return_tensors="pt"):
        ids = list(self.prompts[text]) + [3]                     # reference strips the last id (and BOS for Llama)
        if not self.is_qwen:
            ids = [1] + ids
        return types.SimpleNamespace(input_ids=torch.tensor([ids]))


class _FakeVideoEncoder(nn.Module):
    """Stand-in for the fairseq AV-HuBERT model: [B,1,T,88,88] -> ([B,T,768], None, None); its OUTPUT is stored."""

    def __init__(self):
        super().__init__()
        self.lin = nn.Linear(16, 768)

    def extract_finetune(self, source, padding_mask=None, mask=False, ret_conv=False, output_layer=None):
        v = source["video"]
        assert v.dim() == 5 and v.shape[1] == 1 and source["audio"] is None
        B, _, T = v.shape[:3]
        f = torch.nn.functional.adaptive_avg_pool2d(v.reshape(B * T, 1, 88, 88).float(), 4).reshape(B, T, 16)
        return self.lin(f.to(self.lin.weight.dtype)), None, None


def build_omni_cases(ll, ql, mo):
    from transformers import LlamaConfig, Qwen2Config, WhisperConfig, WhisperFeatureExtractor, WhisperModel
    cases = {}
    prompts = {"PA": [11, 12, 13, 14, 15], "PV": [21, 22, 23, 24, 25, 26], "PAV": [31, 32, 33, 34, 35, 36, 37]}
    wcfg = WhisperConfig(d_model=64, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=128, decoder_layers=1,
                         decoder_attention_heads=2, decoder_ffn_dim=64, num_mel_bins=80, max_source_positions=1500,
                         vocab_size=100, pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1)
    torch.Tensor.cuda = lambda self, *a, **k: self               # encode_audio hard-codes .cuda() (:534)
    specs = [("llama_avg", "meta-llama/Llama-3.2-1B", "avg-pooling"), ("llama_stack", "meta-llama/Llama-3.2-1B", "stack"),
             ("qwen_avg", "Qwen/Qwen2.5-3B", "avg-pooling")]
    for i, (name, llm_name, mode) in enumerate(specs):
        is_qwen = "Qwen" in llm_name
        torch.manual_seed(300 + i)
        fake_video = _FakeVideoEncoder()

        def whisper_from_pretrained(_name):
            return WhisperModel(wcfg)
        mo.WhisperModel = types.SimpleNamespace(from_pretrained=whisper_from_pretrained)
        mo.AutoFeatureExtractor = types.SimpleNamespace(from_pretrained=lambda _n: WhisperFeatureExtractor())
        mo.fairseq = types.SimpleNamespace(checkpoint_utils=types.SimpleNamespace(
            load_model_ensemble_and_task=lambda paths: ([fake_video], None, None)))
        if is_qwen:
            lc = ql.QwenLoRA_config(8, 2, False, False, True, False, False, False, True, True)
            ql.Qwen2ForCausalLM_lora.from_pretrained = classmethod(
                lambda cls, _n, lcfg: cls(Qwen2Config(**QWEN), lcfg))
        else:
            lc = ll.LoRA_config(4, 2, True, False, True, True)
            ll.LlamaForCausalLM_lora.from_pretrained = classmethod(
                lambda cls, _n, lcfg: cls(LlamaConfig(**LLAMA), lcfg))
        tok = _Tok(200, is_qwen, prompts)
        weights = [1.0, 1.5, 0.5]
        model = mo.AVSR_LLMs(modality="audiovisual", pretrain_avhubert_enc_video="base_fake.pt", use_lora_avhubert=False,
                             llm_model=llm_name, hidden_size=(QWEN if is_qwen else LLAMA)["hidden_size"],
                             intermediate_size=96, tokenizer=tok, prompt_audio="PA", prompt_video="PV",
                             prompt_audiovisual="PAV", pad_id=tok.vocab["<pad>"], downsample_ratio_audio=[4, 16],
                             downsample_ratio_video=[2, 5], audio_encoder_name="openai/whisper-fake",
                             compression_mode=mode, unfrozen_modules=["peft_llm"], max_dec_tokens=6, num_beams=1,
                             PETF_LLM_name="lora", peft_config_llm=lc, remove_layernorm_from_projector=False,
                             matry_weights=weights, is_task_specific=True, is_matryoshka=True,
                             is_single_matry_projector=False)
        # prompt buffers were computed from the random-init embedding in __init__; re-derive them after loading the
        # seeded weights exactly as __init__ does (:218-225)
        model.llm.tie_weights()
        model = model.bfloat16().eval()
        assert model.llm.lm_head.weight is model.llm.model.embed_tokens.weight
        named_llm, csum_llm = load_golden_weights(model.llm, 400 + i)
        named_pa, csum_pa = load_golden_weights(model.audio_proj, 500 + i)
        named_pv, csum_pv = load_golden_weights(model.video_proj, 600 + i)
        start = 0 if is_qwen else 1
        with torch.no_grad():
            for key, text in (("prompt_audio", "PA"), ("prompt_video", "PV"), ("prompt_audiovisual", "PAV")):
                getattr(model, key).copy_(model.llm.model.embed_tokens(tok(text).input_ids[:, start:-1]))
        g = torch.Generator().manual_seed(700 + i)
        B, L, T = 2, 9, 23
        n_samples = 20000
        audio = torch.randn(B, n_samples, 1, generator=g)
        audio[1, 15000:] = 0                                                     # collate zero padding
        video = ((torch.rand(B, T, 1, 88, 88, generator=g) - 0.421) / 0.165).bfloat16()
        lengths = torch.tensor([n_samples, 15000])
        tokens = torch.randint(4, 200, (B, L), generator=g)
        if not is_qwen:
            tokens[:, 0] = 1
        tokens[0, -1] = 2
        tokens[1, -3] = 2
        tokens[1, -2:] = tok.vocab["<pad>"]
        labels = tokens.clone()
        labels[labels == tok.vocab["<pad>"]] = -100
        inputs = dict(audio=audio.bfloat16(), video=video, lengths=lengths, tokens=tokens, labels=labels)
        stored_inputs = {k: v for k, v in inputs.items() if k != "video"}      # video only feeds the stand-in encoder
        c = dict(llm_name=llm_name, mode=mode, lora=dict(vars(lc)), seeds=dict(llm=400 + i, pa=500 + i, pv=600 + i),
                 named=dict(llm=named_llm, pa=named_pa, pv=named_pv), checksum=dict(llm=csum_llm, pa=csum_pa, pv=csum_pv),
                 inputs=stored_inputs, prompts=prompts, vocab=dict(tok.vocab), n_vocab=len(tok), matry_weights=weights,
                 prompt_lens=(model.prompt_audio_len, model.prompt_video_len, model.prompt_audiovisual_len),
                 proj_repr=repr(model.audio_proj[0]), train={}, infer={})
        feats = {}
        h1 = model.audio_encoder.register_forward_hook(lambda m_, a, o: feats.__setitem__("audio_enc", o.last_hidden_state.clone()))
        orig_ef = fake_video.extract_finetune

        def ef(source, **kw):
            o = orig_ef(source, **kw)
            feats["video_enc"] = o[0].clone()
            return o
        fake_video.extract_finetune = ef
        orig_llm_forward = model.llm.forward
        with torch.no_grad():
            for ra, rv in ((4, 2), (16, 5)):
                calls = []

                def rec(*a, **kw):
                    calls.append(dict(inputs_embeds=kw["inputs_embeds"].clone(), labels=kw["labels"].clone(),
                                      modality=kw.get("modality")))
                    return orig_llm_forward(*a, **kw)
                model.llm.forward = rec
                random.seed(0)
                losses = model(inputs, is_trainval=True, test_ratio_matry_audio=ra, test_ratio_matry_video=rv)
                model.llm.forward = orig_llm_forward
                pi = model.prepare_inputs(inputs, True, test_ratio_matry_audio=ra, test_ratio_matry_video=rv)
                c["train"][(ra, rv)] = dict(
                    losses=[l.clone() for l in losses], llm_calls=calls, prepare_inputs={k: v.clone() for k, v in pi.items()},
                    audio_comp=model.encode_audio(inputs["audio"], max(inputs["lengths"]), is_trainval=True,
                                                  test_ratio_matry_audio=ra)[0].clone(),
                    video_comp=model.encode_video(inputs["video"], is_trainval=True, test_ratio_matry_video=rv)[0].clone())
            c["audio_enc"], c["video_enc"] = feats["audio_enc"][:, :100].clone(), feats["video_enc"]
            # inference branch: embeddings per task + greedy ids via the driver loop (B = 1 as in the reference's test loader)
            one = dict(audio=inputs["audio"][:1], video=inputs["video"][:1], lengths=lengths[:1],
                       tokens=(torch.zeros(1, 0, dtype=torch.long) if is_qwen else torch.tensor([[1]])), labels=None)
            for task in ("audio", "video", "audiovisual"):
                model.modality = task
                emb = model.prepare_inputs(one, False, test_ratio_matry_audio=4, test_ratio_matry_video=2)
                ids, mg = _greedy(model.llm, emb, 6, 2, 2 if is_qwen else tok.vocab["<pad>"], task)
                c["infer"][task] = dict(embeddings=emb.clone(), greedy=ids, margins=mg)
        h1.remove()
        cases[name] = c
    return cases


def build_llamaavsr_cases(ll, ql, mla):
    """Executes the reference's Omni_AVSR/modeling_LlamaAVSR.py AVSR_LLMs (Llama-AVSR / Llama-MTSK): real constructor,
    prepare_inputs (train + infer), forward, with the same stand-ins as build_omni_cases."""
    from transformers import LlamaConfig, Qwen2Config, WhisperConfig, WhisperFeatureExtractor, WhisperModel
    cases = {}
    prompts = {"P": [11, 12, 13, 14, 15, 16]}
    wcfg = WhisperConfig(d_model=64, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=128, decoder_layers=1,
                         decoder_attention_heads=2, decoder_ffn_dim=64, num_mel_bins=80, max_source_positions=1500,
                         vocab_size=100, pad_token_id=0, bos_token_id=1, eos_token_id=2, decoder_start_token_id=1)
    torch.Tensor.cuda = lambda self, *a, **k: self
    specs = [  # name, llm, modality, mode, matryoshka, remove_ln, rates_a, rates_v, test_ratio
        ("mtsk_av_avg", "meta-llama/Llama-3.2-1B", "audiovisual", "avg-pooling", True, False, [4, 16], [2, 5], [5, 4]),
        ("mtsk_audio_stack", "meta-llama/Llama-3.2-1B", "audio", "stack", True, True, [4, 16], [2, 5], 16),
        ("avsr_video_qwen", "Qwen/Qwen2.5-3B", "video", "avg-pooling", False, False, 4, 2, None),
        ("avsr_av_llama", "meta-llama/Llama-3.2-1B", "audiovisual", "avg-pooling", False, False, 4, 2, None),
    ]
    for i, (name, llm_name, modality, mode, matry, rm_ln, ra, rv, test_ratio) in enumerate(specs):
        is_qwen = "Qwen" in llm_name
        torch.manual_seed(1300 + i)
        fake_video = _FakeVideoEncoder()
        mla.WhisperModel = types.SimpleNamespace(from_pretrained=lambda _n: WhisperModel(wcfg))
        mla.AutoFeatureExtractor = types.SimpleNamespace(from_pretrained=lambda _n: WhisperFeatureExtractor())
        mla.fairseq = types.SimpleNamespace(checkpoint_utils=types.SimpleNamespace(
            load_model_ensemble_and_task=lambda paths: ([fake_video], None, None)))
        if is_qwen:
            lc = ql.QwenLoRA_config(8, 2, False, False, True, False)
            ql.Qwen2ForCausalLM_lora.from_pretrained = classmethod(lambda cls, _n, lcfg: cls(Qwen2Config(**QWEN), lcfg))
        else:
            lc = ll.LoRA_config(4, 2, True, False)
            ll.LlamaForCausalLM_lora.from_pretrained = classmethod(lambda cls, _n, lcfg: cls(LlamaConfig(**LLAMA), lcfg))
        tok = _Tok(200, is_qwen, prompts)
        model = mla.AVSR_LLMs(modality=modality, pretrain_avhubert_enc_video="base_fake.pt", use_lora_avhubert=False,
                              llm_model=llm_name, hidden_size=(QWEN if is_qwen else LLAMA)["hidden_size"],
                              intermediate_size=96, tokenizer=tok, prompt="P", pad_id=tok.vocab["<pad>"],
                              downsample_ratio_audio=ra, downsample_ratio_video=rv, audio_encoder_name="openai/whisper-fake",
                              compression_mode=mode, unfrozen_modules=["peft_llm"], max_dec_tokens=6, num_beams=1,
                              PETF_LLM_name="lora", peft_config_llm=lc, remove_layernorm_from_projector=rm_ln,
                              is_matryoshka=matry)
        model.llm.tie_weights()
        model = model.bfloat16().eval()
        assert model.llm.lm_head.weight is model.llm.model.embed_tokens.weight
        named_llm, csum_llm = load_golden_weights(model.llm, 1400 + i)
        named, csums = {"llm": named_llm}, {"llm": csum_llm}
        for key, seed in (("audio_proj", 1500 + i), ("video_proj", 1600 + i)):
            if hasattr(model, key):
                named[key], csums[key] = load_golden_weights(getattr(model, key), seed)
        g = torch.Generator().manual_seed(1700 + i)
        B, L, T, n_samples = 2, 9, 23, 20000
        audio = torch.randn(B, n_samples, 1, generator=g)
        audio[1, 15000:] = 0
        video = ((torch.rand(B, T, 1, 88, 88, generator=g) - 0.421) / 0.165).bfloat16()
        lengths = torch.tensor([n_samples, 15000])
        tokens = torch.randint(4, 200, (B, L), generator=g)
        if not is_qwen:
            tokens[:, 0] = 1
        tokens[0, -1] = 2
        tokens[1, -3] = 2
        tokens[1, -2:] = tok.vocab["<pad>"]
        labels = tokens.clone()
        labels[labels == tok.vocab["<pad>"]] = -100
        inputs = dict(audio=audio.bfloat16(), video=video, lengths=lengths, tokens=tokens, labels=labels)
        feats = {}
        hook = None
        if hasattr(model, "audio_encoder"):
            hook = model.audio_encoder.register_forward_hook(
                lambda m_, a, o: feats.__setitem__("audio_enc", o.last_hidden_state[:, :100].clone()))
        orig_ef = fake_video.extract_finetune

        def ef(source, **kw):
            o = orig_ef(source, **kw)
            feats["video_enc"] = o[0].clone()
            return o
        fake_video.extract_finetune = ef
        c = dict(llm_name=llm_name, modality=modality, mode=mode, is_matryoshka=matry, remove_layernorm=rm_ln,
                 rates_audio=ra, rates_video=rv, test_ratio=test_ratio, lora=dict(vars(lc)),
                 seeds=dict(llm=1400 + i, audio_proj=1500 + i, video_proj=1600 + i), named=named, checksum=csums,
                 inputs={k: v for k, v in inputs.items() if k != "video"}, prompt_ids=prompts["P"], vocab=dict(tok.vocab),
                 n_vocab=len(tok))
        with torch.no_grad():
            emb, lab = model.prepare_inputs(inputs, True)
            loss = model(inputs, is_trainval=True)
            as_list = lambda x: [t.clone() for t in x] if isinstance(x, list) else x.clone()
            c["train"] = dict(embeddings=as_list(emb), labels=as_list(lab), loss=loss.clone())
            c.update({k: v.clone() for k, v in feats.items()})       # encoder outputs of the TRAIN batch (B = 2)
            one = dict(audio=inputs["audio"][:1], video=inputs["video"][:1], lengths=lengths[:1],
                       tokens=(torch.zeros(1, 0, dtype=torch.long) if is_qwen else torch.tensor([[1]])), labels=None)
            e_inf, _ = model.prepare_inputs(one, False, test_ratio_matry=test_ratio)
            ids, mg = _greedy(model.llm, e_inf, 6, 2, 2 if is_qwen else tok.vocab["<pad>"], None)
            c["infer"] = dict(embeddings=e_inf.clone(), greedy=ids, margins=mg)
        if hook is not None:
            hook.remove()
        cases[name] = c
    return cases


def build_transform_cases():
    """Executes the reference's datamodule/transforms.py (VideoTransform / AudioTransform pipelines) on seeded inputs.
    AddNoise.__init__ loads babble_noise.wav (absent from the tree, and torchaudio.load needs torchcodec here): the module
    is built without calling it and given a synthetic noise waveform; its forward is the reference's."""
    import _ref_compat as rc
    tr = rc.load_by_path("ref_datamodule_transforms", "datamodule/transforms.py")
    g = torch.Generator().manual_seed(2100)
    video = torch.randint(0, 256, (8, 3, 96, 96), generator=g, dtype=torch.uint8)
    gray = torch.randint(0, 256, (16, 1, 90, 92), generator=g, dtype=torch.uint8)
    wave = torch.randn(24000, 1, generator=g) * 0.1
    noise = torch.randn(1, 40000, generator=g) * 0.3
    out = dict(video=video, gray=gray, wave=wave, noise=noise, cases={})

    def seeded(seed, fn):
        torch.manual_seed(seed)
        random.seed(seed)
        return fn()
    out["cases"]["video_train"] = dict(seed=11, out=seeded(11, lambda: tr.VideoTransform("train")(video)))
    out["cases"]["video_val"] = dict(seed=12, out=seeded(12, lambda: tr.VideoTransform("val")(video)))
    out["cases"]["gray_train"] = dict(seed=13, out=seeded(13, lambda: tr.VideoTransform("train")(gray)))

    def add_noise_module(snr_target=None):
        m = tr.AddNoise.__new__(tr.AddNoise)
        nn.Module.__init__(m)
        m.snr_levels = [snr_target] if snr_target else [-5, 0, 5, 10, 15, 20, 999999]      # transforms.py:66
        m.noise = noise
        return m
    ln = tr.FunctionalModule(lambda x: torch.nn.functional.layer_norm(x, x.shape, eps=1e-8))
    train_pipe = nn.Sequential(tr.AdaptiveTimeMask(6400, 16000), add_noise_module(), ln)       # :110-117
    out["cases"]["audio_train"] = dict(seed=21, out=seeded(21, lambda: train_pipe(wave)))
    out["cases"]["audio_train2"] = dict(seed=22, out=seeded(22, lambda: train_pipe(wave)))
    out["cases"]["audio_val"] = dict(seed=23, out=seeded(23, lambda: tr.AudioTransform("val")(wave)))
    snr_pipe = nn.Sequential(add_noise_module(5), ln)                                          # :119-128 with snr_target
    out["cases"]["audio_val_snr5"] = dict(seed=24, out=seeded(24, lambda: snr_pipe(wave)))
    return out


def build_mha_case(mha):
    torch.manual_seed(900)
    E, Hh, T, B = 128, 2, 11, 2
    att = mha.MultiheadAttention(E, Hh, dropout=0.0, self_attention=True).eval()
    att.rank, att.scaling_lora = 16, 2
    r = round(E / att.rank)
    att.lora_down_Q, att.lora_up_Q = nn.Linear(E, r, bias=False), nn.Linear(r, E, bias=False)
    att.lora_down_V, att.lora_up_V = nn.Linear(E, r, bias=False), nn.Linear(r, E, bias=False)
    named, csum = load_golden_weights(att, 901, dtype=torch.float32)
    g = torch.Generator().manual_seed(902)
    x = torch.randn(T, B, E, generator=g)
    pad = torch.zeros(B, T, dtype=torch.bool)
    pad[1, -3:] = True
    with torch.no_grad():
        y0, _ = att.forward_lora(x, x, x, key_padding_mask=None, need_weights=False)
        y1, _ = att.forward_lora(x, x, x, key_padding_mask=pad, need_weights=False)
    return dict(E=E, heads=Hh, rank=att.rank, scaling=att.scaling_lora, seed=901, named_shapes=named, checksum=csum, x=x,
                pad=pad, y_nomask=y0, y_mask=y1)


def main():
    import _ref_compat as rc
    assert rc.available(), "needs /root/reference (build container only)"
    ll, ql, mo, mha = rc.import_reference()
    _patch_ref(ll, ql)
    mla = rc.load_by_path("Omni_AVSR.modeling_LlamaAVSR", "Omni_AVSR/modeling_LlamaAVSR.py", "Omni_AVSR")
    out = dict(llm=build_llm_cases(ll, ql), omni=build_omni_cases(ll, ql, mo), llamaavsr=build_llamaavsr_cases(ll, ql, mla),
               mha=build_mha_case(mha), transforms=build_transform_cases(),
               meta=dict(torch=torch.__version__, note="outputs of /root/reference sources executed on CPU, bf16"))
    torch.save(out, OUT)
    print("written", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
