"""Decode-step microbenchmark: LLM only (no encoders), greedy, CUDA-graph replay.
   python tools/decode_step.py [--llm meta-llama/Llama-3.2-1B] [--batch 64] [--prefill 413] [--steps 32] [--profile]
Prints ms per decode step (CUDA events around the graph replays), the HBM roofline of the step
(weights streamed once + KV cache read) and, with --profile, the per-kernel CUPTI breakdown of the replays."""
import argparse
import json
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def decode_step_bytes(a, B, ctx_len, vocab):
    """Algorithmic HBM bytes of one decode step: every bf16 weight once + the K/V rows of the context."""
    H, I, L = a.hidden_size, a.intermediate_size, a.num_hidden_layers
    per_layer = (a.q_dim + 2 * a.kv_dim) * H + a.q_dim * H + 2 * I * H + I * H
    weights = 2 * (L * per_layer + vocab * H)
    kv = 2 * L * a.kv_dim * ctx_len * B * 2
    return weights, kv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--llm", default="meta-llama/Llama-3.2-1B")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--prefill", type=int, default=413)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    from omni_avsr_b200 import Llama_LoRA as pl
    from omni_avsr_b200 import Qwen_LoRA as pq
    from omni_avsr_b200 import decode as dec
    torch.manual_seed(0)
    is_qwen = "Qwen" in args.llm
    arch = pl.arch_from_name(args.llm)
    lc = (pq.QwenLoRA_config(32, 4, IS_QWEN25_3B=True, IS_TASK_SPECIFIC=True, SHARED_LORA=True) if is_qwen
          else pl.LoRA_config(32, 4, True, False, True, True))
    llm = (pq.Qwen2ForCausalLM_lora if is_qwen else pl.LlamaForCausalLM_lora)(arch, lc)
    vocab = 151669 if is_qwen else 128261
    llm.resize_token_embeddings(vocab)
    for layer in llm.model.layers:
        layer.self_attn.reset_lora_parameters(down_std=0.02)
    B, S0, n = args.batch, args.prefill, args.steps
    x = (torch.randn(B, S0, arch.hidden_size, device="cuda") * 0.02).bfloat16()
    with torch.no_grad():
        for _ in range(2):      # graph capture + warm-up
            dec.greedy_generate(llm, x, n, eos_token_id=-1, pad_token_id=0, modality="audiovisual", trim=False)
        step = dec._get_step(llm, B, (S0 + n + 127) // 128 * 128, n, x.device)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

        def replays():
            step.cache.len_idx.fill_(S0)
            step.rows.pos.fill_(S0)
            step.step_idx.zero_()
            step.cache.graph_mode = True
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.zero_()
            s.record()
            step.run(n)
            e.record()
            torch.cuda.synchronize()
            step.cache.graph_mode = False
            return s.elapsed_time(e) / n
        ts = sorted(replays() for _ in range(5))
        ms = ts[len(ts) // 2]
        w, kv = decode_step_bytes(arch, B, S0 + n // 2, vocab)
        peaks = {"hbm_gbs": 6543.7}
        try:
            peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
        except Exception:
            pass
        bound_ms = (w + kv) / (peaks["hbm_gbs"] * 1e9) * 1e3
        print(json.dumps({"llm": args.llm, "batch": B, "prefill": S0, "steps": n, "ms_per_step": round(ms, 4),
                          "weight_MB": round(w / 1e6, 1), "kv_MB": round(kv / 1e6, 1), "hbm_bound_ms": round(bound_ms, 4),
                          "frac_of_hbm_roofline": round(bound_ms / ms, 3)}), flush=True)
        if args.profile:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                replays()
            agg = defaultdict(lambda: [0.0, 0])
            for ev in prof.events():
                if ev.device_type == torch.autograd.DeviceType.CUDA:
                    agg[ev.name][0] += ev.device_time / 1e3
                    agg[ev.name][1] += 1
            evs = sorted((ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA),
                         key=lambda ev: ev.time_range.start)
            # the launches of one decoder layer in the middle of the second step, in order, with the gaps between them
            per_step = len(evs) // n
            seq = evs[per_step + per_step // 2: per_step + per_step // 2 + 14]
            for a_, b_ in zip(seq, seq[1:] + [None]):
                gap = (b_.time_range.start - a_.time_range.end) if b_ is not None else 0.0
                print(f"   {a_.device_time:7.1f} us  (+{gap:5.1f} us gap)  {a_.name[:90]}")
            rows = sorted(((v[0], v[1], k) for k, v in agg.items()), reverse=True)
            total = sum(r[0] for r in rows)
            print(f"kernel time {total / n:.4f} ms/step, {sum(r[1] for r in rows) // n} launches/step")
            for t, c, name in rows[:30]:
                print(f"{t / n * 1e3:9.1f} us/step {100 * t / total:5.1f}% x{c // n:4d}/step  avg {t / c * 1e3:6.1f} us  {name[:100]}")


if __name__ == "__main__":
    main()
