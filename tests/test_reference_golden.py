"""Parity against OUTPUTS OF THE REFERENCE ITSELF.  tests/golden/reference_golden.pt was produced by executing the
unmodified sources under /root/reference (Llama_LoRA.py, Qwen_LoRA.py, modeling_OmniAVSR.py, fairseq
multihead_attention.py) on seeded inputs -- see tests/golden/make_reference_golden.py.  This file pins

  * the CPU oracle against those outputs (not gpu): splice / labels / compression bit-exact, LoRA-LLM logits, losses and
    greedy tokens, AV-HuBERT `forward_lora`;
  * the CUDA product path against the same outputs (gpu, through the C ABI): same quantities, with the tolerances of
    BASELINE.json's north_star (bit-exact indexing, logits <= 1e-2 of the logit range, token-for-token greedy).

Weights are regenerated from the seeds stored in the fixture (checksum-guarded), nothing under /root/reference is read."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.dirname(HERE))
from make_reference_golden import checksum, golden_weights  # noqa: E402

from oracle import encoders as oe  # noqa: E402
from oracle import llm_lora as ol  # noqa: E402
from oracle import matryoshka as om  # noqa: E402

GOLD = torch.load(os.path.join(HERE, "golden", "reference_golden.pt"), weights_only=False)
TASKS = ("audio", "video", "audiovisual")
LLAMA_CFG = dict(family="llama", hidden_size=256, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                 num_key_value_heads=1, rms_norm_eps=1e-5, rope_theta=500000.0, head_dim=64,
                 rope_scaling=dict(factor=32.0, low_freq_factor=1.0, high_freq_factor=4.0,
                                   original_max_position_embeddings=8192),
                 attention_bias=False, tie_word_embeddings=True, max_position_embeddings=16384, inv_freq_dtype="bf16")
QWEN_CFG = dict(family="qwen2", hidden_size=512, intermediate_size=256, num_hidden_layers=2, num_attention_heads=8,
                num_key_value_heads=1, rms_norm_eps=1e-6, rope_theta=1000000.0, head_dim=64, rope_scaling=None,
                attention_bias=True, tie_word_embeddings=True, max_position_embeddings=4096, inv_freq_dtype="fp32")


def bits_equal(a, b):
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    if a.dtype == torch.bfloat16:
        return torch.equal(a.view(torch.int16), b.view(torch.int16))
    return torch.equal(a, b)


def fixture_weights(named, seed, csum, dtype=torch.bfloat16):
    w = {n: t.to(dtype) for n, t in golden_weights(named, seed).items()}
    assert abs(checksum(w) - csum) <= 1e-6 * csum, "CPU RNG stream drifted: regenerate the fixture"
    return w


def load_named(module, weights):
    sd = module.state_dict()
    missing = [n for n in weights if n not in sd]
    assert not missing, missing
    with torch.no_grad():
        for n, t in weights.items():
            sd[n].copy_(t.to(sd[n].dtype))


def oracle_lora_cfg(family, lora):
    return (ol.QwenLoRA_config if family == "qwen2" else ol.LoRA_config)(**lora)


def oracle_llm(family, lora, vocab, named, seed, csum):
    base = dict(QWEN_CFG if family == "qwen2" else LLAMA_CFG)
    cfg = ol.LLMConfig(vocab_size=vocab, **base)
    m = ol.ForCausalLM_lora(cfg, oracle_lora_cfg(family, lora)).bfloat16().eval()
    load_named(m, fixture_weights(named, seed, csum))
    return m, cfg


def rel_err(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max())


# ------------------------------------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("name", ["llama_S", "llama_T", "llama_ST", "qwen_ST"])
def test_oracle_llm_matches_reference_outputs(name):
    c = GOLD["llm"][name]
    m, cfg = oracle_llm(c["family"], c["lora"], 200, c["named_shapes"], c["seed"], c["checksum"])
    with torch.no_grad():
        for t, ref_logits in c["logits"].items():
            o = m(inputs_embeds=c["x"], labels=c["labels"], modality=t)
            assert rel_err(o.logits, ref_logits) <= 1e-2, (name, t)            # same op order: expected ~1 bf16 ulp
            assert abs(float(o.loss) - float(c["loss"][t])) <= 2e-2
            ids = m.generate(c["x"][:, :13], 8, 199, 198, modality=t)
            ref_ids, mg = c["greedy"][t], c["margins"][t]
            assert ids.shape == ref_ids.shape
            differs = ids != ref_ids
            # a token may differ only where the reference's own top-2 margin is within bf16 noise
            if differs.any():
                first = differs.float().argmax(1)
                for b in range(ids.shape[0]):
                    if differs[b].any():
                        assert mg[b, first[b]] < 0.05, (name, t, b, ids[b], ref_ids[b])


def test_oracle_task_adapters_differ_in_reference_outputs():
    """Sanity of the fixture: the reference's three task adapters give different logits (the LoRA path is live)."""
    lg = GOLD["llm"]["llama_ST"]["logits"]
    assert rel_err(lg["audio"], lg["video"]) > 1e-2 and rel_err(lg["video"], lg["audiovisual"]) > 1e-2


@pytest.mark.parametrize("name", ["llama_avg", "llama_stack", "qwen_avg"])
def test_oracle_compress_splice_labels_bit_exact_vs_reference(name):
    c = GOLD["omni"][name]
    is_qwen = "Qwen" in c["llm_name"]
    n_tok = om.num_audio_tokens(max(c["inputs"]["lengths"]))
    assert n_tok == 62
    stack = c["mode"] == "stack"
    embed = torch.nn.Embedding(c["n_vocab"], (QWEN_CFG if is_qwen else LLAMA_CFG)["hidden_size"]).bfloat16()
    w_llm = fixture_weights(c["named"]["llm"], c["seeds"]["llm"], c["checksum"]["llm"])
    embed.weight.data.copy_(w_llm["model.embed_tokens.weight"])
    start = 0 if is_qwen else 1
    prompts = {t: embed(torch.tensor([c["prompts"][k]])) for t, k in zip(TASKS, ("PA", "PV", "PAV"))}
    v = c["vocab"]
    marker = (v["<audio>"], v["</audio>"], v["<video>"], v["</video>"])
    for (ra, rv), tr in c["train"].items():
        a = om.compress(c["audio_enc"][:, :n_tok], ra, c["mode"])
        vv = om.compress(c["video_enc"], rv, c["mode"])
        assert bits_equal(a, tr["audio_comp"]) and bits_equal(vv, tr["video_comp"])
        assert a.shape[1] == n_tok // ra and vv.shape[1] == 23 // rv
        # projector (reference quirk: nn.Linear(I, H, nn.LayerNorm(H)) -> no LayerNorm, bias on)
        D_a, D_v = 64 * (ra if stack else 1), 768 * (rv if stack else 1)
        H = embed.weight.shape[1]
        pa = om.make_projector(D_a, 96, H, False).bfloat16()
        pv = om.make_projector(D_v, 96, H, False).bfloat16()
        ia, iv = [4, 16].index(ra), [2, 5].index(rv)
        wa = fixture_weights(c["named"]["pa"], c["seeds"]["pa"], c["checksum"]["pa"])
        wv = fixture_weights(c["named"]["pv"], c["seeds"]["pv"], c["checksum"]["pv"])
        load_named(pa, {k[len(f"{ia}."):]: t for k, t in wa.items() if k.startswith(f"{ia}.")})
        load_named(pv, {k[len(f"{iv}."):]: t for k, t in wv.items() if k.startswith(f"{iv}.")})
        with torch.no_grad():
            ta, tv = pa(a), pv(vv)
        pi = tr["prepare_inputs"]
        assert rel_err(ta, pi["audio_tokens"][:, 1:-1]) <= 1e-2 and rel_err(tv, pi["video_tokens"][:, 1:-1]) <= 1e-2
        # splice + labels, fed with the reference's own projected tokens: bit-exact
        with torch.no_grad():
            seqs, labs = om.build_train_sequences(embed, c["inputs"]["tokens"], c["inputs"]["labels"],
                                                  pi["audio_tokens"][:, 1:-1], pi["video_tokens"][:, 1:-1], prompts, marker,
                                                  is_qwen)
        for call in tr["llm_calls"]:
            t = call["modality"]
            assert bits_equal(seqs[t], call["inputs_embeds"]), (name, ra, rv, t)
            assert torch.equal(labs[t], call["labels"])
            assert torch.equal(labs[t], pi[f"labels_{t}"])
    # inference layout: [bos, <audio> a </audio>, <video> v </video>, prompt]
    for t in TASKS:
        emb = c["infer"][t]["embeddings"]
        P = len(c["prompts"][{"audio": "PA", "video": "PV", "audiovisual": "PAV"}[t]])
        na, nv = n_tok // 4, 23 // 2
        pos = start
        a_tok = v_tok = None
        if t in ("audio", "audiovisual"):
            a_tok = emb[:, pos + 1: pos + 1 + na]
            pos += na + 2
        if t in ("video", "audiovisual"):
            v_tok = emb[:, pos + 1: pos + 1 + nv]
            pos += nv + 2
        assert emb.shape[1] == pos + P
        tokens = torch.zeros(1, 0, dtype=torch.long) if is_qwen else torch.tensor([[1]])
        with torch.no_grad():
            mine = om.build_infer_sequence(embed, tokens, a_tok, v_tok, prompts[t], marker, is_qwen)
        assert bits_equal(mine, emb), (name, t)


@pytest.mark.parametrize("name", ["llama_avg", "qwen_avg"])
def test_oracle_losses_and_decode_match_reference_model(name):
    c = GOLD["omni"][name]
    fam = "qwen2" if "Qwen" in c["llm_name"] else "llama"
    m, _ = oracle_llm(fam, c["lora"], c["n_vocab"], c["named"]["llm"], c["seeds"]["llm"], c["checksum"]["llm"])
    with torch.no_grad():
        for (ra, rv), tr in c["train"].items():
            for call, ref_loss, w in zip(tr["llm_calls"], tr["losses"], c["matry_weights"]):
                o = m(inputs_embeds=call["inputs_embeds"], labels=call["labels"], modality=call["modality"])
                assert abs(float(o.loss) * w - float(ref_loss)) <= 3e-2, (name, ra, rv, call["modality"])
        for t in TASKS:
            inf = c["infer"][t]
            ids = m.generate(inf["embeddings"], 6, 2, 2 if fam == "qwen2" else c["vocab"]["<pad>"], modality=t)
            n = min(ids.shape[1], inf["greedy"].shape[1])
            differs = ids[:, :n] != inf["greedy"][:, :n]
            if differs.any():
                first = int(differs[0].float().argmax())
                assert inf["margins"][0, first] < 0.05, (name, t, ids, inf["greedy"])


def test_oracle_whisper_truncation_rule_on_reference_features():
    """encode_audio keeps max(int(max_len/16000*50), 25) encoder rows (:537): the reference's compressed output has
    62 // rate tokens for max_len = 20000 samples."""
    c = GOLD["omni"]["llama_avg"]
    assert c["train"][(4, 2)]["audio_comp"].shape[1] == 15 and c["train"][(16, 5)]["audio_comp"].shape[1] == 3
    assert c["train"][(4, 2)]["video_comp"].shape[1] == 11 and c["train"][(16, 5)]["video_comp"].shape[1] == 4


def test_oracle_avhubert_forward_lora_matches_reference():
    c = GOLD["mha"]
    cfg = oe.AVHubertCfg(embed_dim=c["E"], heads=c["heads"], lora_rank_factor=c["rank"], lora_scaling=c["scaling"])
    att = oe.MHA_lora(cfg).eval()
    load_named(att, fixture_weights(c["named_shapes"], c["seed"], c["checksum"], dtype=torch.float32))
    with torch.no_grad():
        y = att(c["x"])
    assert torch.allclose(y, c["y_nomask"], atol=2e-5, rtol=1e-4)


# ------------------------------------------------------------------------------------------------ product (GPU)
def product_llm(family, lora, vocab, named, seed, csum):
    from omni_avsr_b200 import Llama_LoRA as pl
    from omni_avsr_b200 import Qwen_LoRA as pq
    c = dict(QWEN_CFG if family == "qwen2" else LLAMA_CFG)
    arch = pl.LLMArch(c["family"], c["hidden_size"], c["intermediate_size"], c["num_hidden_layers"],
                      c["num_attention_heads"], c["num_key_value_heads"], vocab, c["rms_norm_eps"], c["rope_theta"],
                      c["head_dim"], c["rope_scaling"], c["attention_bias"], c["tie_word_embeddings"],
                      max_position_embeddings=c["max_position_embeddings"], inv_freq_dtype=c["inv_freq_dtype"])
    if family == "qwen2":
        model = pq.Qwen2ForCausalLM_lora(arch, pq.QwenLoRA_config(**lora))
    else:
        model = pl.LlamaForCausalLM_lora(arch, pl.LoRA_config(**lora))
    w = fixture_weights(named, seed, csum)
    missing, unexpected = model.load_state_dict(w, strict=False)
    assert not unexpected, unexpected
    assert not [k for k in missing if k != "lm_head.weight"], missing
    return model, arch


def _greedy_ok(ids, ref_ids, margins, what):
    ids = ids.cpu()
    n = min(ids.shape[1], ref_ids.shape[1])
    differs = ids[:, :n] != ref_ids[:, :n]
    for b in range(ids.shape[0]):
        if differs[b].any():
            first = int(differs[b].float().argmax())
            # the continuation after a near-tie legitimately diverges; the tie itself must be within the logits tolerance
            assert margins[b, first] < 0.05, (what, b, ids[b], ref_ids[b], margins[b])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["llama_S", "llama_T", "llama_ST", "qwen_ST"])
def test_gpu_llm_matches_reference_outputs(name):
    """Omni-LoRA LLM on the tcgen05 kernels vs logits the reference's own LlamaForCausalLM_lora / Qwen2ForCausalLM_lora
    produced (bf16, CPU).  Tolerance: max|a-b| <= 1e-2 * max|b| (north_star), loss 2e-2, greedy token-for-token."""
    c = GOLD["llm"][name]
    model, arch = product_llm(c["family"], c["lora"], 200, c["named_shapes"], c["seed"], c["checksum"])
    x, lab = c["x"].cuda(), c["labels"].cuda()
    with torch.no_grad():
        for t, ref_logits in c["logits"].items():
            got = model(inputs_embeds=x, modality=t)
            assert rel_err(got.logits.cpu(), ref_logits) <= 1e-2, (name, t, rel_err(got.logits.cpu(), ref_logits))
            loss = model(inputs_embeds=x, labels=lab, modality=t).loss
            assert abs(float(loss) - float(c["loss"][t])) <= 2e-2, (name, t)
            ids = model.generate(inputs_embeds=x[:, :13].contiguous(), max_new_tokens=8, num_beams=1, eos_token_id=199,
                                 pad_token_id=198, modality=t)
            _greedy_ok(ids, c["greedy"][t], c["margins"][t], (name, t))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["llama_avg", "llama_stack", "qwen_avg"])
def test_gpu_compress_splice_labels_bit_exact_vs_reference(name):
    """CUDA Matryoshka compression + splice/label kernels vs what the reference's encode_audio / encode_video /
    prepare_inputs / forward produced: every byte of the three LLM input sequences and every label equal."""
    from omni_avsr_b200 import ops
    c = GOLD["omni"][name]
    is_qwen = "Qwen" in c["llm_name"]
    n_tok = 62
    w_llm = fixture_weights(c["named"]["llm"], c["seeds"]["llm"], c["checksum"]["llm"])
    embed_w = w_llm["model.embed_tokens.weight"].cuda()
    prompts = [embed_w[torch.tensor(c["prompts"][k], device="cuda")] for k in ("PA", "PV", "PAV")]
    v = c["vocab"]
    marker = (v["<audio>"], v["</audio>"], v["<video>"], v["</video>"])
    tokens, labels = c["inputs"]["tokens"].cuda(), c["inputs"]["labels"].cuda()
    for (ra, rv), tr in c["train"].items():
        a = ops.matryoshka_compress(c["audio_enc"].cuda(), n_tok, ra, c["mode"])
        vv = ops.matryoshka_compress(c["video_enc"].cuda(), 23, rv, c["mode"])
        assert bits_equal(a.cpu(), tr["audio_comp"]) and bits_equal(vv.cpu(), tr["video_comp"])
        pi = tr["prepare_inputs"]
        lay = ops.SpliceLayout(tokens=tokens, labels=labels, embed=embed_w,
                               audio_tok=pi["audio_tokens"][:, 1:-1].contiguous().cuda(),
                               video_tok=pi["video_tokens"][:, 1:-1].contiguous().cuda(), prompts=prompts,
                               marker_ids=marker, has_bos=not is_qwen)
        B, H = tokens.shape[0], embed_w.shape[1]
        outs = [torch.empty(B, s, H, device="cuda", dtype=torch.bfloat16) for s in lay.seq_len]
        outl = [torch.empty(B, s, device="cuda", dtype=torch.int64) for s in lay.seq_len]
        status = torch.zeros(1, device="cuda", dtype=torch.int32)
        ops.splice_prompt(lay, outs, outl, status)
        assert status.item() == 0
        for i, call in enumerate(tr["llm_calls"]):
            assert call["modality"] == TASKS[i]
            assert bits_equal(outs[i].cpu(), call["inputs_embeds"]), (name, ra, rv, TASKS[i])
            assert torch.equal(outl[i].cpu(), call["labels"])
    for i, t in enumerate(TASKS):
        emb = c["infer"][t]["embeddings"]
        na, nv = n_tok // 4, 23 // 2
        pos = 0 if is_qwen else 1
        a_tok = v_tok = None
        if t in ("audio", "audiovisual"):
            a_tok = emb[:, pos + 1: pos + 1 + na].contiguous().cuda()
            pos += na + 2
        if t in ("video", "audiovisual"):
            v_tok = emb[:, pos + 1: pos + 1 + nv].contiguous().cuda()
        toks = torch.zeros(1, 0, dtype=torch.long) if is_qwen else torch.tensor([[1]])
        lay = ops.SpliceLayout(tokens=toks.cuda(), labels=None, embed=embed_w, audio_tok=a_tok, video_tok=v_tok,
                               prompts=prompts, marker_ids=marker, has_bos=not is_qwen, task_mask=1 << i)
        outs = [None, None, None]
        outs[i] = torch.empty(1, lay.seq_len[i], embed_w.shape[1], device="cuda", dtype=torch.bfloat16)
        ops.splice_prompt(lay, outs, [None, None, None])
        assert bits_equal(outs[i].cpu(), emb), (name, t)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["llama_avg", "llama_stack", "qwen_avg"])
def test_gpu_fused_pool_project_splice_vs_reference(name):
    """The FUSED launch (omni_pool_project_splice: compression + projector MLP + splice + labels in one persistent kernel)
    against what the reference's encode_audio / encode_video / prepare_inputs / forward produced from the same encoder
    outputs and projector weights: compressed features, every marker / prompt / text row and every label bit-exact;
    the projected media rows (a bf16 matmul on the CPU in the reference, tcgen05 with fp32 accumulation here) within
    1e-2 of the value range."""
    from omni_avsr_b200 import ops
    c = GOLD["omni"][name]
    is_qwen = "Qwen" in c["llm_name"]
    n_tok = 62
    w_llm = fixture_weights(c["named"]["llm"], c["seeds"]["llm"], c["checksum"]["llm"])
    embed_w = w_llm["model.embed_tokens.weight"].cuda()
    prompts = [embed_w[torch.tensor(c["prompts"][k], device="cuda")] for k in ("PA", "PV", "PAV")]
    v = c["vocab"]
    marker = (v["<audio>"], v["</audio>"], v["<video>"], v["</video>"])
    tokens, labels = c["inputs"]["tokens"].cuda(), c["inputs"]["labels"].cuda()
    wa = fixture_weights(c["named"]["pa"], c["seeds"]["pa"], c["checksum"]["pa"])
    wv = fixture_weights(c["named"]["pv"], c["seeds"]["pv"], c["checksum"]["pv"])
    B, H = tokens.shape[0], embed_w.shape[1]
    xa, xv = c["audio_enc"].cuda(), c["video_enc"].cuda()
    bos = 0 if is_qwen else 1

    def proj(w, i):
        return [w[f"{i}.{k}"].cuda().contiguous() for k in ("0.weight", "0.bias", "2.weight", "2.bias")]

    for (ra, rv), tr in c["train"].items():
        ia, iv = [4, 16].index(ra), [2, 5].index(rv)
        na, nv = n_tok // ra, 23 // rv
        lay = ops.SpliceLayout(tokens=tokens, labels=labels, embed=embed_w, audio_tok=None, video_tok=None, prompts=prompts,
                               marker_ids=marker, has_bos=not is_qwen, n_audio=na, n_video=nv)
        outs = [torch.empty(B, s, H, device="cuda", dtype=torch.bfloat16) for s in lay.seq_len]
        outl = [torch.empty(B, s, device="cuda", dtype=torch.int64) for s in lay.seq_len]
        status = torch.zeros(1, device="cuda", dtype=torch.int32)
        res = ops.pool_project_splice(lay, outs, outl, ops.PoolProjectInput(xa, n_tok, ra, *proj(wa, ia)),
                                      ops.PoolProjectInput(xv, 23, rv, *proj(wv, iv)), c["mode"], status=status)
        assert status.item() == 0
        assert bits_equal(res["audio"][0].cpu().view(B, na, -1), tr["audio_comp"])
        assert bits_equal(res["video"][0].cpu().view(B, nv, -1), tr["video_comp"])
        for i, call in enumerate(tr["llm_calls"]):
            got, ref = outs[i].cpu(), call["inputs_embeds"]
            assert got.shape == ref.shape
            media = torch.zeros(ref.shape[1], dtype=torch.bool)
            pos = bos
            if i in (0, 2):
                media[pos + 1: pos + 1 + na] = True
                pos += na + 2
            if i in (1, 2):
                media[pos + 1: pos + 1 + nv] = True
            assert bits_equal(got[:, ~media].contiguous(), ref[:, ~media].contiguous()), (name, ra, rv, TASKS[i])
            assert rel_err(got[:, media], ref[:, media]) <= 1e-2, (name, ra, rv, TASKS[i])
            assert torch.equal(outl[i].cpu(), call["labels"])
    for i, t in enumerate(TASKS):
        emb = c["infer"][t]["embeddings"]
        na, nv = n_tok // 4, 23 // 2
        use_a, use_v = t in ("audio", "audiovisual"), t in ("video", "audiovisual")
        toks = torch.zeros(1, 0, dtype=torch.long) if is_qwen else torch.tensor([[1]])
        lay = ops.SpliceLayout(tokens=toks.cuda(), labels=None, embed=embed_w, audio_tok=None, video_tok=None,
                               prompts=prompts, marker_ids=marker, has_bos=not is_qwen, task_mask=1 << i,
                               n_audio=na if use_a else None, n_video=nv if use_v else None)
        outs = [None, None, None]
        outs[i] = torch.empty(1, lay.seq_len[i], H, device="cuda", dtype=torch.bfloat16)
        # the reference's inference example is clip 0 at rates (4, 2)
        ops.pool_project_splice(lay, outs, [None] * 3,
                                ops.PoolProjectInput(xa[:1].contiguous(), n_tok, 4, *proj(wa, 0)) if use_a else None,
                                ops.PoolProjectInput(xv[:1].contiguous(), 23, 2, *proj(wv, 0)) if use_v else None, c["mode"])
        got = outs[i].cpu()
        assert got.shape == emb.shape
        media = torch.zeros(emb.shape[1], dtype=torch.bool)
        pos = bos
        if use_a:
            media[pos + 1: pos + 1 + na] = True
            pos += na + 2
        if use_v:
            media[pos + 1: pos + 1 + nv] = True
        assert bits_equal(got[:, ~media].contiguous(), emb[:, ~media].contiguous()), (name, t)
        assert rel_err(got[:, media], emb[:, media]) <= 1e-2, (name, t)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["llama_avg", "qwen_avg"])
def test_gpu_losses_and_decode_match_reference_model(name):
    """The three task losses the reference's AVSR_LLMs.forward returned (matry_weights applied) and the greedy ids of
    its inference branch, reproduced by the CUDA LLM on the reference's own spliced sequences."""
    c = GOLD["omni"][name]
    fam = "qwen2" if "Qwen" in c["llm_name"] else "llama"
    model, _ = product_llm(fam, c["lora"], c["n_vocab"], c["named"]["llm"], c["seeds"]["llm"], c["checksum"]["llm"])
    with torch.no_grad():
        for (ra, rv), tr in c["train"].items():
            for call, ref_loss, w in zip(tr["llm_calls"], tr["losses"], c["matry_weights"]):
                o = model(inputs_embeds=call["inputs_embeds"].cuda(), labels=call["labels"].cuda(),
                          modality=call["modality"])
                assert abs(float(o.loss) * w - float(ref_loss)) <= 5e-2, (name, ra, rv, call["modality"])
        for t in TASKS:
            inf = c["infer"][t]
            ids = model.generate(inputs_embeds=inf["embeddings"].cuda(), max_new_tokens=6, num_beams=1, eos_token_id=2,
                                 pad_token_id=2 if fam == "qwen2" else c["vocab"]["<pad>"], modality=t)
            _greedy_ok(ids, inf["greedy"], inf["margins"], (name, t))


@pytest.mark.gpu
def test_gpu_avhubert_lora_attention_matches_reference():
    """AV-HuBERT self-attention with LoRA (fairseq MultiheadAttention.forward_lora, executed from the reference tree)
    vs the LoRA-fused q|k|v GEMM + tcgen05 attention of the product encoder layer.  bf16 path vs fp32 reference:
    max|a-b| <= 2e-2 * max|b|."""
    from omni_avsr_b200 import encoders as pe
    from omni_avsr_b200 import ops
    c = GOLD["mha"]
    E, T, B = c["E"], c["x"].shape[0], c["x"].shape[1]
    arch = pe.AVHubertArch(E, 256, 1, c["heads"], 16, 4, (16, 32, 32, 64))
    att = pe.AVHAttention_lora(arch, "cuda", None, True, 0)
    assert att.rank == c["rank"] and att.scaling_lora == c["scaling"]
    w = fixture_weights(c["named_shapes"], c["seed"], c["checksum"], dtype=torch.float32)
    missing, unexpected = att.load_state_dict({k: t.bfloat16() for k, t in w.items()}, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    h = c["x"].transpose(0, 1).reshape(B * T, E).contiguous().cuda().bfloat16()       # [T,B,E] -> batch-major rows
    with torch.no_grad():
        Tm = ops.gemm(h, att.lora_down.data, n=att.plan.t_cols, alpha=att.plan.scaling, b_row_table=att.plan.brow_fwd,
                      block_n=64)
        qkv = ops.gemm(h, att.qkv_weight, bias=att.qkv_bias, ext=(Tm, att.lora_up.data, att.plan.ext_fwd),
                       block_n=att.plan.block_n, pair_aligned=True)
        o = att.sdpa(qkv, B, T)
        y = ops.gemm(o, att.out_proj.weight.data, bias=att.out_proj.bias.data, block_n=64)
    y = y.float().cpu().view(B, T, E).transpose(0, 1)
    assert rel_err(y, c["y_nomask"]) <= 2e-2, rel_err(y, c["y_nomask"])


# ------------------------------------------------------------------------------------------------ Llama-AVSR / Llama-MTSK
# (Omni_AVSR/modeling_LlamaAVSR.py executed from the reference tree; fixtures under GOLD["llamaavsr"])
from oracle import llama_avsr as ola  # noqa: E402

LA_CASES = ["mtsk_av_avg", "mtsk_audio_stack", "avsr_video_qwen", "avsr_av_llama"]


def _la_media_slices(c, emb, is_trainval, ra, rv):
    """Positions of the projected media tokens inside a reference sequence [bos?, <a> A </a>, <v> V </v>, prompt, ...]."""
    is_qwen = "Qwen" in c["llm_name"]
    stack = c["mode"] == "stack"
    pos = 0 if is_qwen else 1
    a_tok = v_tok = None
    if c["modality"] in ("audio", "audiovisual"):
        na = 62 // ra
        a_tok = emb[:, pos + 1: pos + 1 + na].contiguous()
        pos += na + 2
    if c["modality"] in ("video", "audiovisual"):
        nv = 23 // rv
        v_tok = emb[:, pos + 1: pos + 1 + nv].contiguous()
        pos += nv + 2
    return a_tok, v_tok


def _la_rate_grid(c):
    if not c["is_matryoshka"]:
        return [(c["rates_audio"], c["rates_video"])]
    if c["modality"] == "audiovisual":
        return [(a, v) for v in c["rates_video"] for a in c["rates_audio"]]        # video outer, audio inner (:312-324)
    if c["modality"] == "audio":
        return [(a, None) for a in c["rates_audio"]]
    return [(None, v) for v in c["rates_video"]]


def _la_embed(c):
    H = (QWEN_CFG if "Qwen" in c["llm_name"] else LLAMA_CFG)["hidden_size"]
    w = fixture_weights(c["named"]["llm"], c["seeds"]["llm"], c["checksum"]["llm"])
    embed = torch.nn.Embedding(c["n_vocab"], H).bfloat16()
    embed.weight.data.copy_(w["model.embed_tokens.weight"])
    return embed


@pytest.mark.parametrize("name", LA_CASES)
def test_oracle_llamaavsr_sequences_labels_bit_exact_vs_reference(name):
    c = GOLD["llamaavsr"][name]
    is_qwen = "Qwen" in c["llm_name"]
    embed = _la_embed(c)
    v = c["vocab"]
    marker = (v["<audio>"], v["</audio>"], v["<video>"], v["</video>"])
    prompt_ids = torch.tensor([c["prompt_ids"]])
    tr = c["train"]
    seqs = tr["embeddings"] if c["is_matryoshka"] else [tr["embeddings"]]
    labs = tr["labels"] if c["is_matryoshka"] else [tr["labels"]]
    grid = _la_rate_grid(c)
    assert len(seqs) == len(grid)
    # compression: the reference's media rows are projector(compress(encoder output)) -- check the compressed lengths and
    # (avg-pooling, LN-free stack) the projector output against the oracle projector on the oracle compression
    a_list, v_list = [], []
    for (ra, rv), e in zip(grid, seqs):
        a_tok, v_tok = _la_media_slices(c, e, True, ra or 1, rv or 1)
        a_list.append(a_tok)
        v_list.append(v_tok)
    for which, enc_key, rates, n_in in (("audio_proj", "audio_enc", c["rates_audio"], 62), ("video_proj", "video_enc", c["rates_video"], 23)):
        if which not in c["named"]:
            continue
        wts = fixture_weights(c["named"][which], c["seeds"][which], c["checksum"][which])
        rl = rates if c["is_matryoshka"] else [rates]
        for i, r in enumerate(rl):
            comp = om.compress(c[enc_key][:, :n_in], r, c["mode"])
            ln = not c["remove_layernorm"]
            proj = om.make_projector(comp.shape[-1], 96, embed.weight.shape[1], ln).bfloat16()
            pre = f"{i}." if c["is_matryoshka"] else ""
            load_named(proj, {k[len(pre):]: t for k, t in wts.items() if k.startswith(pre)})
            with torch.no_grad():
                want = proj(comp)
            k = next(j for j, (ra, rv) in enumerate(grid) if (ra if which == "audio_proj" else rv) == r)
            got = (a_list if which == "audio_proj" else v_list)[k]
            assert rel_err(want, got) <= 1e-2, (name, which, r)
    # splice + labels from the reference's own projected tokens: every byte equal
    if c["is_matryoshka"]:
        ua = [a_list[k] for k in range(len(c["rates_audio"]))] if c["modality"] != "video" else None
        if c["modality"] == "audiovisual":
            uv = [v_list[k * len(c["rates_audio"])] for k in range(len(c["rates_video"]))]
        else:
            uv = v_list if c["modality"] == "video" else None
        with torch.no_grad():
            mine, mlab = ola.prepare_inputs(embed, c["inputs"]["tokens"], c["inputs"]["labels"], ua, uv, prompt_ids, marker,
                                            is_qwen, c["modality"], True, True)
    else:
        with torch.no_grad():
            s, l = ola.prepare_inputs(embed, c["inputs"]["tokens"], c["inputs"]["labels"], a_list[0], v_list[0], prompt_ids,
                                      marker, is_qwen, c["modality"], False, True)
        mine, mlab = [s], [l]
    for a, b, la, lb in zip(mine, seqs, mlab, labs):
        assert bits_equal(a, b), name
        assert torch.equal(la, lb), name
    # inference layout
    inf = c["infer"]["embeddings"]
    tr_ = c["test_ratio"]
    if c["is_matryoshka"]:
        ra, rv = (tr_[1], tr_[0]) if c["modality"] == "audiovisual" else (tr_, tr_)
    else:
        ra, rv = c["rates_audio"], c["rates_video"]
    a_tok, v_tok = _la_media_slices(c, inf, False, ra, rv)
    toks = torch.zeros(1, 0, dtype=torch.long) if is_qwen else torch.tensor([[1]])
    with torch.no_grad():
        s, _ = ola.prepare_inputs(embed, toks, None, a_tok, v_tok, prompt_ids, marker, is_qwen, c["modality"], False, False)
    assert bits_equal(s, inf), name


@pytest.mark.parametrize("name", LA_CASES)
def test_oracle_llamaavsr_loss_and_decode_vs_reference(name):
    c = GOLD["llamaavsr"][name]
    fam = "qwen2" if "Qwen" in c["llm_name"] else "llama"
    m, _ = oracle_llm(fam, c["lora"], c["n_vocab"], c["named"]["llm"], c["seeds"]["llm"], c["checksum"]["llm"])
    with torch.no_grad():
        loss = ola.train_loss(m, c["train"]["embeddings"], c["train"]["labels"], c["is_matryoshka"])
        assert abs(float(loss) - float(c["train"]["loss"])) <= 2e-2, name
        inf = c["infer"]
        ids = m.generate(inf["embeddings"], 6, 2, 2 if fam == "qwen2" else c["vocab"]["<pad>"], modality=None)
        n = min(ids.shape[1], inf["greedy"].shape[1])
        differs = ids[:, :n] != inf["greedy"][:, :n]
        if differs.any():
            assert inf["margins"][0, int(differs[0].float().argmax())] < 0.05, (name, ids, inf["greedy"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", LA_CASES)
def test_gpu_llamaavsr_splice_loss_decode_vs_reference(name):
    """CUDA path of the Llama-AVSR / Llama-MTSK step against the reference's outputs: splice + labels of every Matryoshka
    sequence bit-exact, ONE packed LLM pass over all sequences -> mean loss within 5e-2, greedy ids token-for-token."""
    from omni_avsr_b200 import ops
    from omni_avsr_b200.Llama_LoRA import PackedRows, pack_segments
    c = GOLD["llamaavsr"][name]
    is_qwen = "Qwen" in c["llm_name"]
    fam = "qwen2" if is_qwen else "llama"
    model, _ = product_llm(fam, c["lora"], c["n_vocab"], c["named"]["llm"], c["seeds"]["llm"], c["checksum"]["llm"])
    embed_w = model.model.embed_tokens.weight.data
    v = c["vocab"]
    marker = (v["<audio>"], v["</audio>"], v["<video>"], v["</video>"])
    prompt = embed_w[torch.tensor(c["prompt_ids"], device="cuda")].contiguous()
    t = TASKS.index(c["modality"])
    tr = c["train"]
    seqs = tr["embeddings"] if c["is_matryoshka"] else [tr["embeddings"]]
    labs = tr["labels"] if c["is_matryoshka"] else [tr["labels"]]
    tokens, labels = c["inputs"]["tokens"].cuda(), c["inputs"]["labels"].cuda()
    mine, mlab = [], []
    for (ra, rv), e in zip(_la_rate_grid(c), seqs):
        a_tok, v_tok = _la_media_slices(c, e, True, ra or 1, rv or 1)
        lay = ops.SpliceLayout(tokens=tokens, labels=labels, embed=embed_w,
                               audio_tok=None if a_tok is None else a_tok.cuda(),
                               video_tok=None if v_tok is None else v_tok.cuda(), prompts=[prompt] * 3, marker_ids=marker,
                               has_bos=not is_qwen, task_mask=1 << t)
        outs, outl = [None] * 3, [None] * 3
        outs[t] = torch.empty(2, lay.seq_len[t], embed_w.shape[1], device="cuda", dtype=torch.bfloat16)
        outl[t] = torch.empty(2, lay.seq_len[t], device="cuda", dtype=torch.int64)
        ops.splice_prompt(lay, outs, outl)
        mine.append(outs[t])
        mlab.append(outl[t])
    for a, b, la, lb in zip(mine, seqs, mlab, labs):
        assert bits_equal(a.cpu(), b), name
        assert torch.equal(la.cpu(), lb), name
    with torch.no_grad():
        rows = PackedRows.get([(0, 2, s.shape[1]) for s in mine], "cuda")
        hid = model.model.forward_packed(pack_segments(mine, rows), rows)
        losses = model.loss_from_hidden(hid, [(b, s, off) for (_, b, s, off) in rows.segments], mlab, [1.0] * len(mine))
        loss = sum(float(l) for l in losses) / len(mine)
        assert abs(loss - float(tr["loss"])) <= 5e-2, (name, loss, float(tr["loss"]))
        inf = c["infer"]
        ids = model.generate(inputs_embeds=inf["embeddings"].cuda(), max_new_tokens=6, num_beams=1, eos_token_id=2,
                             pad_token_id=2 if is_qwen else v["<pad>"])
        _greedy_ok(ids, inf["greedy"], inf["margins"], name)


# ------------------------------------------------------------------------------------------------ input pipeline
# (datamodule/transforms.py executed from the reference tree; fixtures under GOLD["transforms"])
def _seed(s):
    import random
    torch.manual_seed(s)
    random.seed(s)


def test_oracle_transforms_match_reference_bit_for_bit():
    from oracle import transforms as otr
    t = GOLD["transforms"]
    c = t["cases"]
    _seed(c["video_train"]["seed"])
    assert torch.equal(otr.video_transform(t["video"], "train"), c["video_train"]["out"])
    _seed(c["video_val"]["seed"])
    assert torch.equal(otr.video_transform(t["video"], "val"), c["video_val"]["out"])
    _seed(c["gray_train"]["seed"])
    assert torch.equal(otr.video_transform(t["gray"], "train"), c["gray_train"]["out"])
    for name in ("audio_train", "audio_train2"):
        _seed(c[name]["seed"])
        assert torch.equal(otr.audio_transform(t["wave"], "train", noise=t["noise"]), c[name]["out"])
    _seed(c["audio_val"]["seed"])
    assert torch.equal(otr.audio_transform(t["wave"], "val"), c["audio_val"]["out"])
    _seed(c["audio_val_snr5"]["seed"])
    assert torch.equal(otr.audio_transform(t["wave"], "val", noise=t["noise"], snr_target=5), c["audio_val_snr5"]["out"])
    # the fixture exercises the masks: some but not all frames / samples are zeroed before normalisation
    masked = (c["gray_train"]["out"] == (0.0 - 0.421) / 0.165).flatten(1).all(1)
    assert 0 < int(masked.sum()) < masked.numel()


@pytest.mark.gpu
def test_gpu_video_transform_bit_exact_vs_reference():
    """CUDA VideoTransform (x/255, crop, grayscale, time mask, normalise in one kernel) against the reference's torchvision
    pipeline: every fp32 value identical; same seeds -> same RandomCrop offsets and mask spans."""
    from omni_avsr_b200 import transforms as ptr
    t = GOLD["transforms"]
    c = t["cases"]
    for name, clip, subset in (("video_train", "video", "train"), ("video_val", "video", "val"), ("gray_train", "gray", "train")):
        _seed(c[name]["seed"])
        got = ptr.VideoTransform(subset)(t[clip])
        assert got.shape == c[name]["out"].shape and got.dtype == torch.float32
        assert torch.equal(got.cpu(), c[name]["out"]), name
    _seed(c["video_train"]["seed"])
    got16 = ptr.VideoTransform("train", out_dtype=torch.bfloat16)(t["video"])
    assert torch.equal(got16.cpu(), c["video_train"]["out"].bfloat16())


@pytest.mark.gpu
def test_gpu_audio_transform_vs_reference():
    """CUDA AudioTransform (time mask, add_noise at the sampled SNR, utterance layer-norm) against the reference's
    torchaudio pipeline.  The reductions run in fp64 on the device and in fp32 on the CPU: max|a-b| <= 1e-4 on unit-variance
    output; samples inside a mask span of the noise-free pipeline are the same constant."""
    from omni_avsr_b200 import transforms as ptr
    t = GOLD["transforms"]
    c = t["cases"]
    for name, kw in (("audio_train", dict(subset="train", noise=t["noise"])), ("audio_train2", dict(subset="train", noise=t["noise"])),
                     ("audio_val", dict(subset="val")), ("audio_val_snr5", dict(subset="val", snr_target=5, noise=t["noise"]))):
        tr = ptr.AudioTransform(**kw)
        _seed(c[name]["seed"])
        got = tr(t["wave"])
        want = c[name]["out"]
        assert got.shape == want.shape
        assert (got.cpu() - want).abs().max().item() <= 1e-4, (name, (got.cpu() - want).abs().max().item())


def test_product_span_sampler_draws_like_the_reference():
    """Host logic of the on-device transforms (no GPU needed): with the same seeds the product's mask-span sampler consumes the
    RNGs exactly like the reference's AdaptiveTimeMask (executed through the oracle restatement, itself pinned bit for bit
    above) and yields the same zeroed ranges -- checked by applying both to an all-ones signal."""
    import random
    from oracle import transforms as otr
    from omni_avsr_b200.transforms import adaptive_time_mask_spans
    for seed, length, window, stride in [(1, 400, 10, 25), (2, 256000, 6400, 16000), (3, 7, 10, 25), (4, 24000, 6400, 16000),
                                         (5, 30, 10, 25), (6, 100001, 6400, 16000)]:
        _seed(seed)
        want = otr.adaptive_time_mask(torch.ones(length), window, stride)
        after_ref = (torch.rand(1).item(), random.random())
        _seed(seed)
        spans = adaptive_time_mask_spans(length, window, stride)
        after_prod = (torch.rand(1).item(), random.random())
        got = torch.ones(length)
        for a, b in spans:
            assert 0 <= a < b <= length
            got[a:b] = 0
        assert torch.equal(got, want), (seed, spans)
        assert after_ref == after_prod, "RNG streams diverged"
