"""ORACLE-side test helper: build the CPU oracle that corresponds to a product `ModelModule_LLM` and load the
product's weights into it (state-dict keys are the reference's).  Test infrastructure only."""
from __future__ import annotations

import torch

from . import encoders as oe
from . import llm_lora as ol
from . import modeling as omod


def oracle_from_product(module, dtype=torch.bfloat16):
    """module: omni_avsr_b200.lightning_OmniAVSR.ModelModule_LLM -> oracle.modeling.AVSR_LLMs with the same weights."""
    m = module.model
    a = m.llm.config
    args = module.args
    tok = module.tokenizer
    is_qwen = "Qwen" in args.llm_model
    llm_cfg = ol.LLMConfig(a.family, a.hidden_size, a.intermediate_size, a.num_hidden_layers, a.num_attention_heads,
                           a.num_key_value_heads, a.vocab_size, a.rms_norm_eps, a.rope_theta, a.head_dim, a.rope_scaling,
                           a.attention_bias, a.tie_word_embeddings, inv_freq_dtype=a.inv_freq_dtype)
    lora_cfg = ol.make_lora_config(llm_cfg, args.llm_model, args.rank, args.alpha, args.is_task_specific,
                                   args.use_shared_lora_task_specific)
    wa = m.audio_encoder.config
    whisper_cfg = oe.WhisperCfg(wa.d_model, wa.encoder_layers, wa.encoder_attention_heads, wa.encoder_ffn_dim,
                                wa.num_mel_bins, wa.max_source_positions)
    va = m.video_encoder.arch
    avh_cfg = oe.AVHubertCfg(va.encoder_embed_dim, va.encoder_ffn_embed_dim, va.encoder_layers,
                             va.encoder_attention_heads, va.conv_pos, va.conv_pos_groups)
    start = 0 if is_qwen else 1
    prompts_ids = {k: tok(getattr(args, "prompt_" + k), return_tensors="pt").input_ids[:, start:-1]
                   for k in ("audio", "video", "audiovisual")}
    v = tok.vocab
    marker = (v["<audio>"], v["</audio>"], v["<video>"], v["</video>"])
    eos = v["<|endoftext|>"] if is_qwen else v["<|end_of_text|>"]
    pad = v["<|endoftext|>"] if is_qwen else v["<pad>"]
    rates_a = list(args.downsample_ratio_audio) if args.is_matryoshka else [args.downsample_ratio_audio]
    rates_v = list(args.downsample_ratio_video) if args.is_matryoshka else [args.downsample_ratio_video]
    oracle = omod.AVSR_LLMs(llm_cfg, lora_cfg, whisper_cfg, avh_cfg, args.intermediate_size, rates_a, rates_v,
                            args.compression_mode, prompts_ids, marker, is_qwen, args.matry_weights,
                            args.is_task_specific, va.resnet_widths, args.modality, args.max_dec_tokens, eos, pad,
                            single_projector=bool(args.is_matryoshka and args.is_single_matry_projector),
                            projector_layernorm=not args.no_layernorm_projector)
    sd = {k: t.detach().cpu() for k, t in m.state_dict().items()}
    missing, unexpected = oracle.load_state_dict(sd, strict=False)
    unexpected = [k for k in unexpected if not k.startswith("prompt_")]
    if missing or unexpected:
        raise RuntimeError(f"state-dict mismatch: missing={missing[:8]} unexpected={unexpected[:8]}")
    oracle = oracle.to(dtype).eval()
    # the product's prompt buffers must equal the oracle's embedded prompts
    for k, p in oracle.prompts().items():
        if not hasattr(m, "prompt_" + k):          # Llama-AVSR mirror: the prompt is embedded per call, no buffers
            continue
        if not torch.equal(getattr(m, "prompt_" + k).cpu().to(dtype), p.to(dtype)):
            raise RuntimeError("prompt buffer mismatch: " + k)
    return oracle
