"""Small Omni-AVSR configuration shared by the GPU model tests and __graft_entry__.smoke()."""
import torch


def small_module(llm="meta-llama/Llama-3.2-1B", task_specific=True, shared=True, compression="avg-pooling", seed=0,
                 lora_std=0.05, llm_over=None, **extra_args):
    from omni_avsr_b200 import lightning_OmniAVSR as pl_mod
    from omni_avsr_b200.encoders import AVHubertArch, WhisperArch
    torch.manual_seed(seed)
    is_qwen = "Qwen" in llm
    over = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=1,
                head_dim=64)
    if is_qwen:
        over = dict(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=8,
                    num_key_value_heads=1, head_dim=64)
    if llm_over:
        over.update(llm_over)
    args = pl_mod.make_args(llm_model=llm, rank=4 if not is_qwen else 8, alpha=2, intermediate_size=384,
                            is_task_specific=task_specific, use_shared_lora_task_specific=shared,
                            compression_mode=compression, max_dec_tokens=8, **extra_args)
    mk = dict(llm_overrides=over, hidden_size_override=over["hidden_size"],
              audio_arch=WhisperArch(128, 2, 2, 256), video_arch=AVHubertArch(128, 256, 2, 2, 16, 4, (16, 32, 32, 64)))
    mod = pl_mod.ModelModule_LLM(args, model_kwargs=mk)
    m = mod.model
    with torch.no_grad():
        for layer in m.llm.model.layers:
            layer.self_attn.reset_lora_parameters(down_std=lora_std)
        for layer in m.video_encoder.encoder.layers:
            layer.self_attn.lora_down_Q.weight.normal_(0, lora_std)
            layer.self_attn.lora_down_V.weight.normal_(0, lora_std)
    return mod


def small_llamaavsr_module(modality="audiovisual", is_matryoshka=True, compression="avg-pooling", seed=0, lora_std=0.05):
    """Small Llama-AVSR / Llama-MTSK configuration for the GPU tests."""
    from omni_avsr_b200 import lightning_LlamaAVSR as pl_mod
    from omni_avsr_b200.encoders import AVHubertArch, WhisperArch
    torch.manual_seed(seed)
    over = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=1,
                head_dim=64)
    args = pl_mod.make_args(modality=modality, llm_model="meta-llama/Llama-3.2-1B", rank=4, alpha=2, intermediate_size=384,
                            compression_mode=compression, max_dec_tokens=8, is_matryoshka=is_matryoshka,
                            downsample_ratio_audio=[4, 16] if is_matryoshka else 4,
                            downsample_ratio_video=[2, 5] if is_matryoshka else 2, no_layernorm_projector=True,
                            downsample_ratio_test_matry=[5, 4] if modality == "audiovisual" else 4)
    mk = dict(llm_overrides=over, hidden_size_override=256, audio_arch=WhisperArch(128, 2, 2, 256),
              video_arch=AVHubertArch(128, 256, 2, 2, 16, 4, (16, 32, 32, 64)))
    mod = pl_mod.ModelModule_LLM(args, model_kwargs=mk)
    m = mod.model
    with torch.no_grad():
        for layer in m.llm.model.layers:
            layer.self_attn.reset_lora_parameters(down_std=lora_std)
        if hasattr(m, "video_encoder"):
            for layer in m.video_encoder.encoder.layers:
                layer.self_attn.lora_down_Q.weight.normal_(0, lora_std)
                layer.self_attn.lora_down_V.weight.normal_(0, lora_std)
    return mod
