"""Golden vectors for SURVEY row a16: the reference's OWN AV-HuBERT video path executed on CPU --
    av_hubert/avhubert/hubert.py          AVHubertModel.__init__ / SubModel :318-333 / forward_features :539 / extract_finetune :695-755
    av_hubert/avhubert/resnet.py          ResEncoder :131-169
    fairseq/models/wav2vec/wav2vec2.py    TransformerEncoder :818-905, TransformerSentenceEncoderLayer :916-1038 (apply_lora)
    fairseq/modules/multihead_attention.py forward_lora :389-672, same_pad.py, layer_norm.py, gelu.py, transpose_last.py,
    fairseq/modules/transformer_sentence_encoder.py init_bert_params
all loaded BY PATH from /root/reference, unmodified.  What is stubbed is only what those files import but this path never
executes (fairseq's registry / dataclass / task / dictionary machinery, omegaconf.II) plus three one-line fairseq.utils
helpers (index_put / is_xla_tensor / buffered_arange, fairseq/utils.py:264,712,716 -- that file drags in the whole fairseq
package) and get_activation_fn("gelu") -> the reference's own modules/gelu.py.  LoRA is attached to the encoder layers the
way Omni_AVSR/modeling_OmniAVSR.py:128-142 does (rank 16, scaling 2) with non-zero down projections.

    python tests/golden/make_avhubert_golden.py      (build container only; writes tests/golden/avhubert_golden.pt)
"""
import importlib.util
import os
import sys
import types
import zlib
from copy import deepcopy

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
FS = os.path.join(REF, "av_hubert/fairseq/fairseq")
OUT = os.path.join(HERE, "avhubert_golden.pt")


def named_weight(name, shape, seed):
    """Deterministic, ORDER-INDEPENDENT tensor for a state-dict entry (reference, oracle and product regenerate the same
    weights from the name alone)."""
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ seed) & 0x7FFFFFFF)
    r = torch.randn(tuple(shape), generator=g)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "running_var":
        return 0.5 + torch.rand(tuple(shape), generator=g)
    if leaf == "running_mean":
        return 0.1 * r
    if leaf == "weight_g":
        return 0.5 + torch.rand(tuple(shape), generator=g)
    if len(shape) == 1 and leaf == "weight":                  # norm scales, PReLU slopes
        return (0.25 + 0.05 * r) if ("relu" in name or "frontend3D.2" in name) else (1.0 + 0.1 * r)
    if len(shape) == 1:
        return 0.02 * r
    if "lora" in name:
        return 0.05 * r
    fan_in = 1
    for d in shape[1:]:
        fan_in *= d
    return r * (1.0 / fan_in) ** 0.5


def fill(module, seed, keep=lambda n: True):
    names = []
    with torch.no_grad():
        for n, p in module.state_dict().items():
            if not p.is_floating_point() or not keep(n):
                continue
            p.copy_(named_weight(n, p.shape, seed))
            names.append((n, tuple(p.shape)))
    return names


def _load(name, path, **pre):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    for k, v in pre.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def import_reference_avhubert():
    sys.path.insert(0, HERE)
    import _ref_compat as rc
    rc.install_fairseq_stubs()
    # the reference's own small modules, by path
    gelu = _load("fairseq.modules.gelu", os.path.join(FS, "modules/gelu.py"))
    same_pad = _load("fairseq.modules.same_pad", os.path.join(FS, "modules/same_pad.py"))
    layer_norm = _load("fairseq.modules.layer_norm", os.path.join(FS, "modules/layer_norm.py"))
    transpose_last = _load("fairseq.modules.transpose_last", os.path.join(FS, "modules/transpose_last.py"))
    grad_mult = _load("fairseq.modules.grad_multiply", os.path.join(FS, "modules/grad_multiply.py"))
    mha = _load("fairseq.modules.multihead_attention", os.path.join(FS, "modules/multihead_attention.py"))

    def get_activation_fn(activation):
        assert activation == "gelu"                     # fairseq/utils.py:539-540
        return gelu.gelu

    def index_put(tensor, indices, value):              # fairseq/utils.py:716-725 (non-XLA branch)
        tensor[indices] = value
        return tensor

    utils = sys.modules["fairseq.utils"]
    utils.get_activation_fn = get_activation_fn
    utils.index_put = index_put
    utils.is_xla_tensor = lambda t: False
    utils.buffered_arange = lambda n: torch.arange(n)
    utils.get_available_activation_fns = lambda: ["relu", "gelu", "gelu_fast", "gelu_accurate", "tanh", "linear"]   # :556-564
    _stub("fairseq.data")
    _stub("fairseq.data.data_utils", compute_mask_indices=None)
    _stub("fairseq.data.dictionary", Dictionary=object)
    _stub("fairseq.dataclass", ChoiceEnum=lambda choices: str, FairseqDataclass=object)
    _stub("fairseq.models", BaseFairseqModel=nn.Module, register_model=lambda *a, **k: (lambda cls: cls))
    _stub("fairseq.modules", Fp32GroupNorm=None, Fp32LayerNorm=None, GradMultiply=grad_mult.GradMultiply,
          GumbelVectorQuantizer=None, LayerNorm=layer_norm.LayerNorm, MultiheadAttention=mha.MultiheadAttention,
          SamePad=same_pad.SamePad, TransposeLast=transpose_last.TransposeLast, PositionalEmbedding=None,
          TransformerSentenceEncoderLayer=None, FairseqDropout=sys.modules["fairseq.modules.fairseq_dropout"].FairseqDropout,
          LayerDropModuleList=None, gelu=gelu.gelu, gelu_accurate=gelu.gelu_accurate)
    _stub("fairseq.modules.multihead_attention_lora", MultiheadAttention_lora=None)
    tse = _load("fairseq.modules.transformer_sentence_encoder", os.path.join(FS, "modules/transformer_sentence_encoder.py"))
    _stub("fairseq.models.wav2vec")
    w2v = _load("fairseq.models.wav2vec.wav2vec2", os.path.join(FS, "models/wav2vec/wav2vec2.py"))
    _stub("omegaconf", II=lambda s: s)
    resnet = _load("av_hubert.avhubert.resnet", os.path.join(REF, "av_hubert/avhubert/resnet.py"))
    _stub("av_hubert.avhubert.hubert_pretraining", AVHubertPretrainingConfig=None, AVHubertPretrainingTask=None)
    _stub("av_hubert.avhubert.utils", compute_mask_indices=None)
    _stub("av_hubert.avhubert.decoder", TransformerDecoder=None)
    argv, sys.argv = sys.argv, sys.argv[:1]             # hubert.py:29 picks its absolute-import branch when len(argv) == 1
    try:
        hub = _load("av_hubert.avhubert.hubert", os.path.join(REF, "av_hubert/avhubert/hubert.py"))
    finally:
        sys.argv = argv
    return hub, w2v, tse


def small_cfg(E=128, ffn=256, layers=2, heads=2, conv_pos=16, groups=4):
    """The fields AVHubertModel.__init__ / TransformerEncoder read, with the values of conf/pretrain/large_vox_iter5.yaml:70-101
    scaled down (E / ffn / layers / heads / conv_pos); dropouts are inert in eval mode."""
    return types.SimpleNamespace(
        label_rate=25, sub_encoder_layers=0, resnet_relu_type="prelu", resnet_weights=None, audio_feat_dim=104,
        modality_dropout=0.5, audio_dropout=0.5, modality_fuse="concat", encoder_embed_dim=E, mask_prob_image=0.3,
        mask_prob_audio=0.8, mask_selection="static", mask_other=0, mask_length_image=5, mask_length_audio=10,
        no_mask_overlap=False, mask_min_space=1, mask_channel_prob=0.0, mask_channel_selection="static", mask_channel_other=0,
        mask_channel_length=10, no_mask_channel_overlap=False, mask_channel_min_space=1, dropout_input=0.1,
        dropout_features=0.1, feature_grad_mult=0.1, logit_temp=0.1, skip_masked=False, skip_nomask=False, sim_type="cosine",
        selection_type="same_seq", masking_type="input", final_dim=32, target_glu=False, untie_final_proj=True,
        dropout=0.1, attention_dropout=0.1, activation_dropout=0.0, encoder_layerdrop=0.05, encoder_layers=layers,
        encoder_ffn_embed_dim=ffn, encoder_attention_heads=heads, activation_fn="gelu", layer_norm_first=True, conv_pos=conv_pos,
        conv_pos_groups=groups)


def attach_lora(model, E, rank=16, scaling=2):
    """Omni_AVSR/modeling_OmniAVSR.py:128-137."""
    for layer in model.encoder.layers:
        layer.apply_lora = True
        att = layer.self_attn
        att.rank, att.scaling_lora = rank, scaling
        r = round(E / rank)
        att.lora_down_Q, att.lora_up_Q = nn.Linear(E, r, bias=False), nn.Linear(r, E, bias=False)
        att.lora_down_V, att.lora_up_V = nn.Linear(E, r, bias=False), nn.Linear(r, E, bias=False)


VIDEO_KEYS = ("feature_extractor_video.", "encoder.", "layer_norm.", "post_extract_proj.")


def main():
    hub, w2v, tse = import_reference_avhubert()
    torch.manual_seed(0)
    E = 128
    cfg = small_cfg(E=E)
    model = hub.AVHubertModel(cfg, types.SimpleNamespace(sample_rate=25), [None])
    attach_lora(model, E)
    model.eval()
    seed = 4242
    named = fill(model, seed, keep=lambda n: n.startswith(VIDEO_KEYS))
    g = torch.Generator().manual_seed(77)
    video = torch.randn(2, 1, 6, 40, 40, generator=g)          # [B, 1, T, H, W] (extract_finetune's layout, modeling_OmniAVSR.py:463)
    with torch.no_grad():
        x, pad, layers = model.extract_finetune(source={"video": video, "audio": None})
        front = model.feature_extractor_video(video)             # [B, E, T]: ResEncoder + proj (SubModel.forward)
    assert pad is None and len(layers) == cfg.encoder_layers
    out = dict(seed=seed, E=E, ffn=cfg.encoder_ffn_embed_dim, layers=cfg.encoder_layers, heads=cfg.encoder_attention_heads,
               conv_pos=cfg.conv_pos, conv_pos_groups=cfg.conv_pos_groups, rank=16, scaling=2, named_shapes=named, video=video,
               sub_model_out=front, x=x, layer_outputs=[t.transpose(0, 1).contiguous() for t in layers],
               note="outputs of /root/reference hubert.py / wav2vec2.py / resnet.py / multihead_attention.py executed on CPU, fp32, eval")
    torch.save(out, OUT)
    print("written", OUT, os.path.getsize(OUT), "x", tuple(x.shape), float(x.abs().mean()))


if __name__ == "__main__":
    main()
