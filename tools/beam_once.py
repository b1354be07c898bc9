"""Beam-search microbenchmark, LLM only (no encoders): CUDA-graph replays of the beam step (ranking + scorer + forward of the
B*K rows) with CUDA events, and the per-kernel CUPTI breakdown of the replays.
   python tools/beam_once.py [--llm meta-llama/Llama-3.2-1B] [--batch 8] [--beams 15] [--prefill 413] [--steps 32]"""
import argparse
import json
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--llm", default="meta-llama/Llama-3.2-1B")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--beams", type=int, default=15)
    ap.add_argument("--prefill", type=int, default=413)
    ap.add_argument("--steps", type=int, default=32)
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
    llm.resize_token_embeddings(151669 if is_qwen else 128261)
    for layer in llm.model.layers:
        layer.self_attn.reset_lora_parameters(down_std=0.02)
    B, K, S0, n = args.batch, args.beams, args.prefill, args.steps
    x = (torch.randn(B, S0, arch.hidden_size, device="cuda") * 0.02).bfloat16()
    with torch.no_grad():
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            dec.beam_generate(llm, x, n, K, eos_token_id=-1, pad_token_id=0, modality="audiovisual")
        torch.cuda.synchronize()
        s.record()
        dec.beam_generate(llm, x, n, K, eos_token_id=-1, pad_token_id=0, modality="audiovisual")
        e.record()
        torch.cuda.synchronize()
        total_ms = s.elapsed_time(e)
        step = dec._get_beam_step(llm, B, K, (S0 + n + 127) // 128 * 128, n, x.device)

        def replays():
            step.start(0, S0, step.h_last.clone(), -1, 0)
            s.record()
            step.run(n)
            e.record()
            torch.cuda.synchronize()
            step.finish()
            return s.elapsed_time(e) / n
        ms = sorted(replays() for _ in range(5))[2]
        print(json.dumps({"llm": args.llm, "utterances": B, "beams": K, "rows_per_step": B * K, "prefill": S0, "steps": n,
                          "ms_per_beam_step": round(ms, 4), "ms_generate_total": round(total_ms, 2)}), flush=True)
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            replays()
        agg = defaultdict(lambda: [0.0, 0])
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                agg[ev.name][0] += ev.device_time / 1e3
                agg[ev.name][1] += 1
        rows = sorted(((v[0], v[1], k) for k, v in agg.items()), reverse=True)
        total = sum(r[0] for r in rows)
        print(f"kernel time {total / n:.4f} ms/step, {sum(r[1] for r in rows) // n} launches/step")
        for t, c, name in rows[:14]:
            print(f"{t / n * 1e3:9.1f} us/step {100 * t / total:5.1f}% x{c // n:4d}/step  avg {t / c * 1e3:6.1f} us  {name[:100]}")


if __name__ == "__main__":
    main()
