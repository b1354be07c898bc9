"""Checkpoint compatibility: load the weight files the reference starts from into the drop-in
modules, and average training checkpoints the way the reference does.

Reference behaviour mirrored (file:line in /root/reference):
  * `WhisperModel.from_pretrained(name).encoder`  (Omni_AVSR/modeling_OmniAVSR.py:59): HF state-dict keys
    `encoder.conv1.weight`, `encoder.layers.N.self_attn.q_proj.weight`, ... (optionally prefixed `model.`) -> our
    WhisperEncoder uses the same names below `encoder.`;
  * `LlamaForCausalLM_lora.from_pretrained(llm_model, cfg)` / Qwen twin (:203-208) followed by
    `resize_token_embeddings(len(tokenizer))` (:214): HF keys `model.embed_tokens.weight`, `model.layers.N...`,
    `lm_head.weight` (absent when tied); the checkpoint's V0 vocabulary rows go into the first V0 rows of the already
    resized embedding (the rows of the added special tokens keep their initialisation, as `resize_token_embeddings` does);
  * `fairseq.checkpoint_utils.load_model_ensemble_and_task([large_vox_iter5.pt])` (:123): a fairseq checkpoint
    `{"model": state_dict, "cfg": ...}`; the video-only `extract_finetune` path uses `feature_extractor_video.*`,
    `post_extract_proj.*`, `layer_norm.*`, `encoder.*`; everything else in the file (audio front-end, mask embedding,
    label embeddings, final projection) is not on this path and is reported back as ignored;
  * `utils/avg_checkpoints.py:14-45`: `average_checkpoints` / `ensemble_original` over Lightning checkpoints
    (`{"state_dict": {"model.<key>": tensor}}`), float tensors averaged, integer tensors floor-divided;
  * `ModelModule_LLM.__init__` loading the averaged `model_avg_N.pth` (lightning_OmniAVSR.py:148-150) = plain
    `AVSR_LLMs.load_state_dict`, which the drop-in already accepts (same key names).

Every loader returns `(missing, ignored)`: our keys that the file did not provide, and file keys that were not used.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Tuple

import torch

AVHUBERT_VIDEO_PREFIXES = ("feature_extractor_video.", "post_extract_proj.", "layer_norm.", "encoder.")


def _strip(sd: Dict[str, torch.Tensor], prefixes: Iterable[str]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        for p in prefixes:
            if k.startswith(p):
                out[k[len(p):]] = v
                break
    return out


def _load(module, sd) -> Tuple[List[str], List[str]]:
    own = module.state_dict()
    used = {k: v for k, v in sd.items() if k in own}
    bad = [k for k, v in used.items() if tuple(v.shape) != tuple(own[k].shape)]
    if bad:
        raise ValueError(f"shape mismatch for {bad[:4]}: file {tuple(used[bad[0]].shape)} vs model {tuple(own[bad[0]].shape)}")
    res = module.load_state_dict(used, strict=False)
    for m in module.modules():                # packed / transposed copies used by the kernels are rebuilt lazily
        if hasattr(m, "_wt"):
            m._wt = None
        if hasattr(m, "_head_t"):
            m._head_t = None
    return list(res.missing_keys), [k for k in sd if k not in own]


def load_whisper_encoder(encoder, state_dict) -> Tuple[List[str], List[str]]:
    """HF `WhisperModel` / `WhisperForConditionalGeneration` state dict (or just its encoder's) -> our WhisperEncoder."""
    if any(k.startswith("model.encoder.") for k in state_dict):
        sd = _strip(state_dict, ("model.encoder.",))
    elif any(k.startswith("encoder.") for k in state_dict):
        sd = _strip(state_dict, ("encoder.",))
    else:
        sd = dict(state_dict)
    missing, ignored = _load(encoder, sd)
    return missing, ignored


def load_llm(llm, state_dict, owner=None) -> Tuple[List[str], List[str]]:
    """HF `LlamaForCausalLM` / `Qwen2ForCausalLM` state dict -> our *_lora model.  The LoRA tensors are not in such a
    file and stay as initialised (reported in `missing`).  `owner`: the AVSR_LLMs that holds `llm`; its prompt buffers
    (embeddings of the task prompts, computed by the reference AFTER from_pretrained, modeling_OmniAVSR.py:218-221) are
    re-embedded from the freshly loaded table."""
    sd = dict(state_dict)
    emb = sd.get("model.embed_tokens.weight")
    own_rows = llm.model.embed_tokens.weight.shape[0]
    for key, target in (("model.embed_tokens.weight", llm.model.embed_tokens.weight),
                        ("lm_head.weight", llm.lm_head.weight)):
        w = sd.get(key)
        if w is not None and w.shape[0] != own_rows:
            if w.shape[0] > own_rows or w.shape[1] != target.shape[1]:
                raise ValueError(f"{key}: file {tuple(w.shape)} does not fit model {tuple(target.shape)}")
            with torch.no_grad():          # rows of the tokens added after from_pretrained keep their initialisation (:214)
                target.data[: w.shape[0]].copy_(w.to(target.dtype))
            del sd[key]
    if llm.config.tie_word_embeddings:
        sd.pop("lm_head.weight", None)     # tied: one storage, already filled through embed_tokens
    elif "lm_head.weight" not in state_dict and emb is not None:
        raise ValueError("untied architecture but the file has no lm_head.weight")
    missing, ignored = _load(llm, sd)
    missing = [k for k in missing if k not in ("lm_head.weight", "model.embed_tokens.weight")]
    if owner is not None and hasattr(owner, "refresh_prompts"):
        owner.refresh_prompts()
    return missing, ignored


def load_avhubert(video_encoder, checkpoint) -> Tuple[List[str], List[str]]:
    """fairseq AV-HuBERT checkpoint (`{"model": sd, ...}` or the bare state dict) -> our AVHubertVideoEncoder."""
    sd = checkpoint["model"] if isinstance(checkpoint, dict) and "model" in checkpoint and not torch.is_tensor(
        checkpoint["model"]) else checkpoint
    used = {k: v for k, v in sd.items() if k.startswith(AVHUBERT_VIDEO_PREFIXES)}
    missing, ignored = _load(video_encoder, used)
    ignored += [k for k in sd if not k.startswith(AVHUBERT_VIDEO_PREFIXES)]
    return missing, ignored


def _model_states(ckpt) -> Dict[str, torch.Tensor]:
    """The `model.`-prefixed entries of a Lightning checkpoint, prefix removed (what ModelModule_LLM.model holds)."""
    if isinstance(ckpt, (str, os.PathLike)):
        ckpt = torch.load(ckpt, map_location="cpu")
    return {k[len("model."):]: v for k, v in ckpt["state_dict"].items() if k.startswith("model.")}


def average_checkpoints(last):
    """Element-wise mean of the model tensors of several Lightning checkpoints, with the arithmetic of
    utils/avg_checkpoints.py:14-31: a running sum in each tensor's own dtype (bf16 sums round at every addition), then a true
    division for floating-point tensors and a floor division for integer ones.  `last`: paths or loaded checkpoint dicts."""
    count = len(last)
    total: Dict[str, torch.Tensor] = {}
    for n, item in enumerate(last):
        states = _model_states(item)
        if n == 0:
            total = {k: v.clone() for k, v in states.items()}
        else:
            for k in total:
                total[k].add_(states[k])
    for k, v in total.items():
        if v is None:
            continue
        if v.is_floating_point():
            v.div_(count)
        else:
            v.floor_divide_(count)
    return total


def ensemble_original(args, num_average_epochs=10):
    """utils/avg_checkpoints.py:34-45: average the last N epoch checkpoints into `model_avg_N.pth`."""
    last = [os.path.join(args.exp_dir, args.exp_name, f"epoch={n}.ckpt")
            for n in range(args.max_epochs - num_average_epochs, args.max_epochs)]
    model_path = os.path.join(args.exp_dir, args.exp_name, f"model_avg_{num_average_epochs}.pth")
    torch.save(average_checkpoints(last), model_path)
    return model_path


def lightning_checkpoint(module) -> dict:
    """What the reference's Trainer would save for our `ModelModule_LLM` (keys prefixed `model.`), so that checkpoints
    written by the drop-in average and reload exactly like the reference's."""
    return {"state_dict": {"model." + k: v.detach().cpu().clone() for k, v in module.model.state_dict().items()}}
