#!/usr/bin/env python
"""Omni-AVSR hot-path benchmark (BASELINE.json metric: utterances/sec, train step, 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--mode train|decode] [--impl ours|reference]

Workload (config.workload): BASELINE config 2 -- AVSR train step, Whisper-medium + AV-HuBERT-Large + Llama-3.2-1B,
hybrid (task-specific + shared) Omni-LoRA, audio rates {4,16} x video rates {2,5}, bf16, synthetic 16 s clips,
random-init weights.  One "step" = forward of the three task sequences (ASR, VSR, AVSR) of B utterances + backward +
gradient all-reduce (N>1) + global-norm clip + AdamW; step k uses rate pair k mod 4 so K steps sweep the grid.
Data parallel over utterances: weak scaling (per-GPU batch fixed), one NCCL all-reduce of the flat trainable-gradient
buffer per step.

`value`  : utterances/s with the batch already resident in HBM.
`e2e`    : the same step through the public API (ModelModule_LLM.train_step) with every step's batch copied from pinned
           host memory inside the timed region (DevicePrefetcher: the copy of batch k+1 overlaps step k on a copy stream)
           and the loss read back to the host every step.
`--impl reference`: the reference's own PyTorch CPU path (the oracle restatement, since the reference cannot be
           imported here -- see DESIGN.md) on the host cores, bounded sample (batch 1 per step).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

RATE_GRID = [(4, 2), (4, 5), (16, 2), (16, 5)]
WORKLOAD = ("Omni-AVSR AVSR train step: Whisper-medium + AV-HuBERT-Large + Llama-3.2-1B, audio rates {4,16} x video "
            "rates {2,5} (step k uses pair k mod 4), hybrid Omni-LoRA (task-specific + shared, r=64), bf16, 3 tasks/utterance")


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu capture
    (profiles/gemm_traffic.json, written by tools/ncu_traffic.py from an ncu pass over one B=32 train step).  The capture is
    tied to the source digest it was taken from: `stale` says whether the library running now is a different build."""
    path = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    try:
        t = json.load(open(path))
        from omni_avsr_b200 import build as b
        dig = b._digest(sorted(b.CSRC.glob("*.cu")) + sorted(b.CSRC.glob("*.cuh")) + sorted((b.ROOT / "include").glob("*.h")))
        return t["plain"]["mean_bytes_per_launch"], {
            "file": "profiles/gemm_traffic.json", "launches_captured": t["plain"]["launches"],
            "stale": dig != t.get("source_digest"), "how": t.get("how")}
    except Exception as e:          # no capture committed: say so instead of quoting a number from another build
        return None, {"file": "profiles/gemm_traffic.json", "error": repr(e)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


MTSK_WORKLOAD = ("Llama-MTSK AVSR train step (SURVEY 8f rank 2): Whisper-medium + AV-HuBERT-Large + Llama-3.2-1B, shared LoRA "
                 "(r=64), audiovisual, ALL audio rates {4,16} x video rates {2,5} in every step (4 sequences/utterance, "
                 "one packed LLM pass), bf16")


def build_module(args, device, llm=None, task_specific=True, shared=True):
    if getattr(args, "workload", "omni") == "mtsk":
        from omni_avsr_b200 import lightning_LlamaAVSR as L
        margs = L.make_args(modality="audiovisual", is_matryoshka=True, downsample_ratio_audio=[4, 16],
                            downsample_ratio_video=[2, 5], num_beams=1, max_dec_tokens=32, no_layernorm_projector=True,
                            downsample_ratio_test_matry=[2, 4])
    else:
        from omni_avsr_b200 import lightning_OmniAVSR as L
        margs = L.make_args(num_beams=1, max_dec_tokens=32, is_task_specific=task_specific,
                            use_shared_lora_task_specific=shared)
    llm = llm or getattr(args, "llm", None)
    if llm:                                  # BASELINE configs 4 / 5 (Qwen2.5-3B, Llama-3.1-8B): extra lines, not the judged one
        margs.llm_model = llm
    torch.manual_seed(0)
    mod = L.ModelModule_LLM(margs, device=device)
    with torch.no_grad():   # non-degenerate adapters (LoRA down AND up ~ N(0, 0.02))
        for layer in mod.model.llm.model.layers:
            layer.self_attn.reset_lora_parameters(down_std=0.02)
        for layer in mod.model.video_encoder.encoder.layers:
            layer.self_attn.lora_down_Q.weight.normal_(0, 0.02)
            layer.self_attn.lora_down_V.weight.normal_(0, 0.02)
    mod.configure_optimizers()
    return mod



SETTINGS = [("audio", 4, None), ("audio", 16, None), ("video", None, 2), ("video", None, 5),
            ("audiovisual", 4, 2), ("audiovisual", 4, 5), ("audiovisual", 16, 2), ("audiovisual", 16, 5)]


def decode_step_bytes(a, B, ctx_len, vocab):
    """Algorithmic HBM bytes of one decode step: every bf16 weight once + the K/V rows of the context."""
    H, I, L = a.hidden_size, a.intermediate_size, a.num_hidden_layers
    per_layer = (a.q_dim + 2 * a.kv_dim) * H + a.q_dim * H + 2 * I * H + I * H
    weights = 2 * (L * per_layer + vocab * H)
    kv = 2 * L * a.kv_dim * ctx_len * B * 2
    return weights, kv


def measure_decode(mod, Bd, device, rank, world, barrier, llm_name):
    """Greedy decode sweep over the 8 (task, rate) settings of eval_OmniAVSR.py:310-337 (second half of BASELINE's metric)
    + the HBM roofline of the decode STEP (graph replays of the audiovisual (4, 2) setting: weights once + KV cache)."""
    import torch.distributed as dist
    from omni_avsr_b200 import decode as dec
    from omni_avsr_b200.synthetic import synthetic_batch, to_device
    dhost = synthetic_batch(Bd, mod.tokenizer, seconds=16.0, text_len=48, seed=4321 + rank, pin=True)
    dres = to_device(dhost, device)
    dres["tokens"] = dres["tokens"][:, :1].contiguous()

    def sweep(which):
        for task, ra, rv in which:
            mod.args.modality = task
            mod.args.downsample_ratio_test_matry_audio, mod.args.downsample_ratio_test_matry_video = ra, rv
            mod.on_test_epoch_start()
            mod.model.decode_no_trim = True           # timing protocol: always 32 new tokens
            mod.test_step(dres)
    with torch.no_grad():
        sweep(SETTINGS)                               # untimed: CUDA-graph capture per cache bucket, allocator warm-up
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        sweep(SETTINGS)
        e.record()
        barrier()
        t = torch.tensor([s.elapsed_time(e)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # ---- decode-step roofline: replay the captured step of the audiovisual (4, 2) setting
        sweep([SETTINGS[4]])
        llm = mod.model.llm
        steps = getattr(llm, "_graphed_steps", {})
        roof = None
        if steps:
            (B_, max_len, n_new), step = max(steps.items(), key=lambda kv: kv[0][1])
            S0 = step.cache.len - n_new if step.cache.len > n_new else step.cache.len
            S0 = 413 if mod.model._has_bos else 412
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
            times = []
            for _ in range(3):
                step.cache.len_idx.fill_(S0)
                step.rows.pos.fill_(S0)
                step.step_idx.zero_()
                step.cache.graph_mode = True
                flush.zero_()
                s.record()
                step.run(n_new)
                e.record()
                torch.cuda.synchronize()
                step.cache.graph_mode = False
                times.append(s.elapsed_time(e) / n_new)
            ms_step = sorted(times)[1]
            w, kv = decode_step_bytes(llm.config, B_, S0 + n_new // 2, llm.config.vocab_size)
            peaks = load_peaks()
            bound = (w + kv) / (peaks["hbm_gbs"] * 1e9) * 1e3
            roof = {"bound": "hbm", "bytes_per_step": w + kv, "weight_bytes": w, "kv_bytes": kv, "ms_per_step": round(ms_step, 4),
                    "achieved": round((w + kv) / ms_step / 1e6, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": round(bound / ms_step, 3), "peak_source": peaks["_source"],
                    "how": "CUDA events around 32 CUDA-graph replays of the decode step (B=%d, context %d+), L2 flushed before; "
                           "bytes = every bf16 weight once + K/V rows of the context" % (B_, S0)}
        # ---- the reference's evaluation default (eval_OmniAVSR.py:216-226): beam search, num_beams = 15, on the audiovisual
        # (4, 2) setting; ranking, scorer bookkeeping and KV-cache indirection on the device, the step in one CUDA graph
        beam = None
        try:
            Bb, Kb = min(8, Bd), 15
            small = {k: (v[:Bb].contiguous() if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == Bd else v)
                     for k, v in dres.items()}
            mod.model.num_beams = Kb
            task, ra, rv = SETTINGS[4]
            mod.args.modality = task
            mod.args.downsample_ratio_test_matry_audio, mod.args.downsample_ratio_test_matry_video = ra, rv
            mod.on_test_epoch_start()
            mod.test_step(small)                      # untimed: graph capture
            torch.cuda.synchronize()
            tb = []
            for _ in range(3):
                s.record()
                mod.test_step(small)
                e.record()
                torch.cuda.synchronize()
                tb.append(s.elapsed_time(e))
            ms_b = sorted(tb)[1]
            beam = {"num_beams": Kb, "batch_per_gpu": Bb, "rows_per_step": Bb * Kb, "ms_per_batch": round(ms_b, 2),
                    "value": round(Bb / (ms_b * 1e-3), 2), "unit": "utterances/s (per GPU)",
                    "includes": "encoders + compression + projector + splice + prefill (once per utterance) + up to 32 beam "
                                "steps (lm_head, omni_beam_topk_rows, omni_beam_select, forward of the B*K rows)"}
        except Exception as ex:                       # reported, never fatal for the headline line
            beam = {"error": repr(ex)[:200]}
        finally:
            mod.model.num_beams = 1
    return {"metric": "utterances/sec (greedy decode, 32 new tokens, sweep over 8 task x rate settings)",
            "value": round(Bd * world * len(SETTINGS) / (t.item() * 1e-3), 2), "unit": "utterances/s",
            "batch_per_gpu": Bd, "ms_per_sweep": round(t.item(), 2), "llm": llm_name,
            "includes": "encoders + compression + projector + splice + prefill + 32 decode steps", "roofline": roof,
            "beam_search": beam}


def measure_train(mod, B, device, rank, world, barrier, steps, warmup, seed=1234, micro=None):
    """utterances/s of the train step with the batch resident in HBM (max over ranks).  micro: split the per-GPU batch into
    micro-batches of this size whose gradients accumulate in the flat buffer before ONE optimizer step (single GPU only:
    the same step as one big batch, used where the big batch does not fit)."""
    import torch.distributed as dist
    from omni_avsr_b200.synthetic import synthetic_batch, to_device
    host = synthetic_batch(B, mod.tokenizer, seconds=16.0, text_len=48, seed=seed + rank, pin=True)
    res = to_device(host, device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    if micro and micro < B:
        parts = [{k: (v[i: i + micro].contiguous() if torch.is_tensor(v) else v) for k, v in res.items()}
                 for i in range(0, B, micro)]

        def one_step(k):
            mod.zero_grad_flat()
            for i, p_ in enumerate(parts):
                mod.no_sync = i < len(parts) - 1          # gradients accumulate locally; the last micro-batch reduces
                # loss * W / sum(B) divides the mean loss by the batch size: weight (b / B) ** 2 reproduces the one-batch step
                (mod.training_step(p_, 0, RATE_GRID[k % 4]) * (p_["tokens"].shape[0] / B) ** 2).backward()
            mod.no_sync = False
            mod.optimizer_step(1e-4)
    else:
        def one_step(k):
            mod.train_step(res, rates=RATE_GRID[k % 4], lr=1e-4)
    for k in range(4 + warmup):
        one_step(k)
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for k in range(steps):
        flush.zero_()
        one_step(k)
    e.record()
    barrier()
    t = torch.tensor([s.elapsed_time(e)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return B * world * steps / (t.item() * 1e-3), t.item() / steps


def measure_ragged(mod, device, rank, world, barrier, max_frames=12800, n_utts=384, passes=1):
    """Train-step throughput on VARIABLE-length utterances batched the reference's way (datamodule/data_module.py:82-144:
    length buckets + frame budget, collate_LLM padding), instead of the fixed 16 s clips of the headline: a synthetic
    LRS3-like length distribution (log-normal, median 4 s, clipped to [1 s, 16 s]) -> SyntheticLengthDataset ->
    CustomBucketDataset(max_frames) -> collate_LLM -> ModelModule_LLM.train_step.  The budget is the headline batch's media
    volume (32 x 400 frames); the README recipe's 1500 frames is a 24 GB-GPU figure."""
    import math
    import torch.distributed as dist
    from omni_avsr_b200 import data_module as dm
    from omni_avsr_b200.synthetic import to_device
    g = torch.Generator().manual_seed(7 + rank)
    secs = torch.exp(torch.randn(n_utts, generator=g) * 0.6 + math.log(4.0)).clamp(1.0, 16.0)
    frames = (secs * 25).round().int().tolist()
    data = dm.SyntheticLengthDataset(frames, seed=rank)
    ds = dm.CustomBucketDataset(data, data.input_lengths, max_frames, num_buckets=50)
    host = []
    for i in range(len(ds)):
        b = dm.collate_LLM(ds[i], mod.tokenizer, "audiovisual", is_trainval=True)
        b["audio"], b["video"] = b["audio"].to(torch.bfloat16), b["video"].to(torch.bfloat16)
        host.append({k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in b.items()})
    n_b = torch.tensor([len(host)], device=device)
    if world > 1:                                   # every rank must take the same number of optimizer steps
        dist.all_reduce(n_b, op=dist.ReduceOp.MIN)
    host = host[: int(n_b.item())]
    dev = [to_device(b, device) for b in host]
    torch.cuda.synchronize()
    for k, b in enumerate(dev):                     # untimed: every batch shape once (allocator, row-layout caches)
        mod.train_step(b, rates=RATE_GRID[k % 4], lr=1e-4)
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(passes):
        for k, b in enumerate(dev):
            mod.train_step(b, rates=RATE_GRID[k % 4], lr=1e-4)
    e.record()
    barrier()
    t = torch.tensor([s.elapsed_time(e)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    utts = sum(b["tokens"].shape[0] for b in host) * passes
    real_frames = sum(int(b["lengths"].sum()) // 640 for b in host) * passes
    padded_frames = sum(b["video"].shape[0] * b["video"].shape[1] for b in host) * passes
    tot = torch.tensor([utts, real_frames, padded_frames], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot)
    utts, real_frames, padded_frames = (float(x) for x in tot.tolist())
    sec = t.item() * 1e-3
    return {"value": round(utts / sec, 2), "unit": "utterances/s", "media_seconds_per_s": round(real_frames / 25 / sec, 1),
            "headline_equivalent": "the fixed 16 s headline processes value x 16 media seconds per second",
            "batches_per_gpu": len(host), "utterances_per_batch": [b["tokens"].shape[0] for b in host],
            "max_frames": max_frames, "padding_fraction": round(1 - real_frames / padded_frames, 3),
            "ms_total": round(t.item(), 1), "n_gpus": world,
            "lengths": "log-normal, median 4 s, sigma 0.6, clipped to [1, 16] s; 50 buckets (data_module.py:101-140)"}


def extra_configs(args, device, rank, world, barrier):
    """BASELINE configs 3 / 4 / 5, the strong-scaling point and the ragged (bucketed) workload as extra objects of the ONE
    JSON line (bounded: a few steps each).  None of them is the judged headline; each names its configuration."""
    import gc
    out = {}

    def run(name, build_kw, fn):
        """Build a module, measure, and drop the module (with its CUDA graphs and caches) before the next configuration."""
        m = None
        torch.cuda.reset_peak_memory_stats()
        before = torch.cuda.memory_allocated()
        try:
            m = build_module(args, device, **build_kw)
            out[name] = fn(m)
        except Exception as ex:      # noqa: BLE001  (reported in the line, never fatal for the headline)
            out[name] = {"error": repr(ex)[:300]}
        finally:
            m = None
            gc.collect()
            torch.cuda.empty_cache()
            out[name]["hbm_gb"] = {"peak": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                                   "live_before": round(before / 2 ** 30, 1),
                                   "live_after": round(torch.cuda.memory_allocated() / 2 ** 30, 1)}

    # variable-length utterances through the reference's frame-budget bucketing
    run("ragged_bucketed_train", {}, lambda m: measure_ragged(m, device, rank, world, barrier))

    def config3(m):     # joint 3-task training, TASK-SPECIFIC LoRA (no shared adapter), data parallel at this N
        v, ms = measure_train(m, args.batch, device, rank, world, barrier, steps=4, warmup=1)
        return {"value": round(v, 2), "unit": "utterances/s", "ms_per_step": round(ms, 2), "per_gpu_batch": args.batch,
                "n_gpus": world, "scaling": "weak", "lora": "task-specific (IS_TASK_SPECIFIC, no shared adapter), Llama-3.2-1B"}
    run("config3_task_specific_lora", dict(task_specific=True, shared=False), config3)

    def strong(m):      # strong scaling: global batch fixed at 256 utterances
        per = 256 // world
        v, ms = measure_train(m, per, device, rank, world, barrier, steps=3, warmup=0, micro=64)
        return {"value": round(v, 2), "unit": "utterances/s", "ms_per_step": round(ms, 2), "per_gpu_batch": per,
                "n_gpus": world, "scaling": "strong",
                "note": "one optimizer step per 256 utterances; per-GPU batches above 64 run as micro-batches of 64 accumulating "
                        "into the flat gradient buffer (one all-reduce per optimizer step)"}
    run("strong_scaling_global_batch_256", {}, strong)

    # config 4: elastic greedy decode sweep, batch 64, Qwen2.5-3B backbone
    run("config4_qwen25_3b_decode_B64", dict(llm="Qwen/Qwen2.5-3B"),
        lambda m: measure_decode(m, 64, device, rank, world, barrier, "Qwen/Qwen2.5-3B"))

    def config5(m):     # Llama-3.1-8B backbone, hybrid Omni-LoRA, full rate grid
        v, ms = measure_train(m, 16, device, rank, world, barrier, steps=4, warmup=0)
        return {"value": round(v, 2), "unit": "utterances/s", "ms_per_step": round(ms, 2), "per_gpu_batch": 16,
                "n_gpus": world, "scaling": "weak"}
    run("config5_llama31_8b_train", dict(llm="meta-llama/Meta-Llama-3.1-8B"), config5)
    return out


def fused_stage_roofline(B, device, peaks):
    """north_star subsystem (1) as ONE launch (omni_pool_project_splice): live CUDA-event timing at the bench batch, rates
    (4, 2), against max(bytes / HBM peak, flops / tensor peak) with the algorithmic figures of SURVEY 8(d)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from omni_avsr_b200 import ops
    from tools.bench_kernels import fused_bytes_flops, fused_setup
    c = fused_setup(B, 4, 2)
    lay = ops.SpliceLayout(tokens=c["tokens"], labels=c["tokens"], embed=c["embed"], audio_tok=None, video_tok=None,
                           prompts=c["prompts"], marker_ids=c["marker"], has_bos=True, n_audio=c["na"], n_video=c["nv"])
    outs = [torch.empty(B, sl, c["H"], device=device, dtype=torch.bfloat16) for sl in lay.seq_len]
    outl = [torch.empty(B, sl, device=device, dtype=torch.int64) for sl in lay.seq_len]
    a_in = ops.PoolProjectInput(c["xa"], 800, 4, *c["pa"])
    v_in = ops.PoolProjectInput(c["xv"], 400, 2, *c["pv"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    ts = []
    for i in range(8):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.pool_project_splice(lay, outs, outl, a_in, v_in, "avg-pooling")
        e.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    byts, wbytes, flops = fused_bytes_flops(c)
    t_hbm = (byts + wbytes) / (peaks["hbm_gbs"] * 1e9) * 1e3
    t_tc = flops / (peaks["bf16_tflops_sustained"] * 1e12) * 1e3
    return {"kernel": "omni::pool_project_splice_kernel (compression + projector MLP + splice + labels, one launch)",
            "batch": B, "rates": [4, 2], "us": round(ms * 1e3, 1), "MB_per_utt": round(byts / B / 1e6, 3),
            "GFLOP_per_utt": round(flops / B / 1e9, 3), "projector_tflops": round(flops / ms / 1e9, 1),
            "frac_tensor_sustained": round(flops / ms / 1e9 / peaks["bf16_tflops_sustained"], 3),
            "GBs": round((byts + wbytes) / ms / 1e6, 1), "frac_hbm": round((byts + wbytes) / ms / 1e6 / peaks["hbm_gbs"], 3),
            "bound": "tensor" if t_tc > t_hbm else "hbm", "bound_us": round(1e3 * max(t_hbm, t_tc), 1),
            "frac_of_bound": round(max(t_hbm, t_tc) / ms, 3),
            "how": "median of 6 launches, CUDA events, 256 MiB written between launches; the HBM stage (7.15 MB/utt) runs "
                   "under the tensor stage (5.03 GFLOP/utt): the roofline is max(bytes / HBM, flops / tensor)"}


def run_ours(args):
    import torch.distributed as dist
    from omni_avsr_b200 import ops
    from omni_avsr_b200.synthetic import host_bytes, synthetic_batch, to_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sm_100a); there is no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    mod = build_module(args, device)
    B = args.batch
    host = synthetic_batch(B, mod.tokenizer, seconds=16.0, text_len=48, seed=1234 + rank, pin=True)
    resident = to_device(host, device)
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(k):
        return mod.train_step(resident, rates=RATE_GRID[k % 4], lr=1e-4)

    from omni_avsr_b200.synthetic import DevicePrefetcher
    feed = DevicePrefetcher(lambda k: host, device)         # H2D of every step's inputs from pinned memory, inside the timed
                                                            # region, double-buffered: batch k+1 is copied under step k

    def step_e2e(k):
        loss = mod.train_step(feed.next(), rates=RATE_GRID[k % 4], lr=1e-4)
        return float(loss.item())                           # D2H read of the step's result

    def timed(fn, K):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.LAUNCHES
        s.record()
        for k in range(K):
            l2_flush.zero_()                                 # flush L2 between timed iterations
            fn(k)
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)         # max over ranks
        return t.item(), ops.LAUNCHES - l0

    for k in range(4):                  # prime every rate pair once (allocator, cuDNN plans, layout caches): untimed
        step_resident(k)
    for k in range(args.warmup):
        step_resident(k)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, launches = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    for k in range(min(args.warmup, 2)):
        step_e2e(k)
    ms_e2e, _ = timed(step_e2e, args.steps)

    # ---- roofline of the dominant kernel (tcgen05 GEMM): one instrumented step, CUDA events per launch ---------
    roof = None
    if True:   # every rank runs the instrumented step (it contains the step's collectives); rank 0 reports
        recs = []
        orig = ops.gemm

        def traced(a, b, **kw):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = orig(a, b, **kw)
            e.record()
            n = kw.get("n") or b.shape[0]
            k_ext = 0
            if kw.get("ext") is not None:
                k_ext = kw["ext"][2].shape[-2] * 64
            recs.append((s, e, 2.0 * a.shape[0] * n * (a.shape[1] + k_ext), (a.shape[0], n, a.shape[1] + k_ext),
                         kw.get("act") in ("swiglu64", "gelu_keep", "swiglu_bwd64", "gelu_bwd", "prelu_ring")))
            return out
        ops.gemm = traced
        orig_conv = ops.conv_frames

        def traced_conv(x, spec, prelu=None):
            # table-driven convolutions of the ResNet trunk (K-extension-only GEMM launches): executed MACs of the tables
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            out = orig_conv(x, spec, prelu=prelu)
            e.record()
            recs.append((s, e, 2.0 * x.N * spec.macs_per_frame, (x.N, spec.N, spec.n_ext * 64), True))
            return out
        ops.conv_frames = traced_conv
        import omni_avsr_b200.autograd_ops as ag
        try:
            step_resident(0)
            torch.cuda.synchronize()
        finally:
            ops.gemm = orig
            ops.conv_frames = orig_conv
        # the dominant kernel = the plain-epilogue tcgen05 GEMM; the launches whose epilogue also runs SwiGLU / GELU and
        # writes a second output (gemm_bf16_tn_2cta<5,1|2>) are reported next to it: their time contains that extra work
        fused = [r for r in recs if r[4]]
        recs = [r[:4] for r in recs if not r[4]]
        tot_ms = sum(s.elapsed_time(e) for s, e, _, _ in recs)
        tot_fl = sum(f for _, _, f, _ in recs)
        fused_ms = sum(s.elapsed_time(e) for s, e, _, _, _ in fused)
        fused_fl = sum(f for _, _, f, _, _ in fused)
        if rank == 0 and os.path.isdir(os.path.join(ROOT, "gpurun_out")):
            by_shape = {}
            for s_, e_, f_, shp in recs:
                d = by_shape.setdefault(shp, [0, 0.0, 0.0])
                d[0] += 1
                d[1] += s_.elapsed_time(e_)
                d[2] += f_
            rows_ = sorted(({"M": k[0], "N": k[1], "K": k[2], "launches": v[0], "ms": round(v[1], 3),
                             "tflops": round(v[2] / (v[1] * 1e-3) / 1e12, 1)} for k, v in by_shape.items()),
                           key=lambda r: -r["ms"])
            json.dump(rows_, open(os.path.join(ROOT, "gpurun_out", "gemm_shapes.json"), "w"), indent=0)
        peaks = load_peaks()
        ach = tot_fl / (tot_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "omni::gemm_bf16_tn_{2cta,cluster,persistent} (tcgen05)", "achieved": round(ach, 1),
                "peak": peaks["bf16_tflops_sustained"], "peak_source": peaks["_source"] + " (sustained: timed inside a long step)",
                "unit": "TFLOP/s", "frac": round(ach / peaks["bf16_tflops_sustained"], 3), "traffic": load_traffic()[0],
                "traffic_source": load_traffic()[1],
                "launches": len(recs), "gemm_ms_per_step": round(tot_ms, 2),
                "fused_epilogue_gemms": {"launches": len(fused), "ms_per_step": round(fused_ms, 2),
                                         "achieved_gemm_flops_only": round(fused_fl / max(fused_ms, 1e-9) / 1e9, 1),
                                         "note": "SwiGLU (LLM gate_up) / GELU-keep (AV-HuBERT fc1) / SwiGLU backward (down_proj "
                                                 "dgrad) / BasicBlock tail (trunk convolutions) computed in the epilogue; "
                                                 "replaces separate elementwise kernels"},
                "how": "algorithmic 2*M*N*(K+K_ext) per launch / CUDA-event duration per launch, summed over one step"}

    # ---- greedy decode (second half of the BASELINE metric): elastic sweep over the 8 (task, rate) settings -------
    decode = None
    if not args.no_decode and args.workload == "omni":
        decode = measure_decode(mod, args.decode_batch, device, rank, world, barrier, args.llm or "meta-llama/Llama-3.2-1B")

    # ---- parity at benchmark scale (rank 0): the CPU oracle with THIS module's weights, one utterance, rates (4, 2) ----
    parity = None
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "omni":
        parity, cpu_base = parity_and_cpu_baseline(mod, device)

    extras = None
    if args.extras and args.workload == "omni" and not args.llm:
        del mod, resident, feed
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        extras = extra_configs(args, device, rank, world, barrier)
        mod = None

    hbm = None
    if rank == 0 and args.workload == "omni":
        hbm = hbm_kernel_rooflines(B, device, load_peaks())
        hbm["fused_pool_project_splice"] = fused_stage_roofline(B, device, load_peaks())
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    utt = B * world * args.steps
    value = utt / (ms_total * 1e-3)
    e2e = utt / (ms_e2e * 1e-3)
    line = {
        "metric": "utterances/sec (train step)", "value": round(value, 3), "unit": "utterances/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": (WORKLOAD if args.workload == "omni" else MTSK_WORKLOAD) +
                   (f" [LLM replaced by {args.llm}]" if args.llm else ""), "per_gpu_batch": B, "global_batch": B * world, "clip_seconds": 16,
                   "text_tokens": 48, "parallelism": f"dp{world}", "l2": "256 MiB buffer written between timed steps",
                   "optimizer": "NCCL all-reduce of the flat gradient buffer (LLM adapters' range launched from an autograd hook "
                                "under the rest of the backward), then one clip(10) + AdamW kernel", "random_init": True},
        "e2e": {"value": round(e2e, 3), "unit": "utterances/s", "h2d_bytes_per_step": host_bytes(host),
                "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3),
                "h2d": "every step copies its batch from pinned host memory inside the timed region; double-buffered "
                       "(DevicePrefetcher: batch k+1 on a copy stream under step k)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "hbm_kernels": hbm if args.workload == "omni" else None,
        "decode": decode,
        "parity": parity,
        "extra_configs": extras,
    }
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def hbm_kernel_rooflines(B, device, peaks):
    """Live CUDA-event timings (L2 flushed) of the HBM-bound kernels of north_star subsystem (1) at the bench batch:
    Matryoshka compression (avg-pool, audio rate 4 / video rate 2 from the untruncated encoder outputs) and the splice /
    label kernel (rates (4, 2): 200 audio + 200 video tokens).  Algorithmic bytes as in SURVEY.md 8(d)."""
    from omni_avsr_b200 import ops
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def timeit(fn, iters=9):
        fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        return sorted(ts)[len(ts) // 2]
    out = {"peak_gbs": peaks["hbm_gbs"], "peak_source": peaks["_source"], "batch": B,
           "how": "median of 9 launches, CUDA events, 256 MiB written between launches; bytes = algorithmic read + write",
           "note": "compress_* / splice_* are the STANDALONE C-ABI kernels (Llama-AVSR 'stack' compression, API parity); the "
                   "Omni-AVSR train / decode path runs fused_pool_project_splice instead (one launch, see below)"}
    for name, T, n_tok, rate in (("compress_avg_audio_r4", 1500, 800, 4), ("compress_avg_video_r2", 400, 400, 2)):
        x = torch.randn(B, T, 1024, device=device).bfloat16()
        ms = timeit(lambda: ops.matryoshka_compress(x, n_tok, rate, "avg-pooling"))
        n = n_tok // rate
        byts = B * n * rate * 1024 * 2 + B * n * 1024 * 2
        out[name] = {"us": round(ms * 1e3, 1), "GBs": round(byts / ms / 1e6, 1), "frac": round(byts / ms / 1e6 / peaks["hbm_gbs"], 3)}
    H, V, L, n_a, n_v = 2048, 128261, 48, 200, 200
    embed = torch.randn(V, H, device=device).bfloat16()
    tokens = torch.randint(0, V, (B, L), device=device)
    a = torch.randn(B, n_a, H, device=device).bfloat16()
    v = torch.randn(B, n_v, H, device=device).bfloat16()
    prompts = [torch.randn(pl, H, device=device).bfloat16() for pl in (6, 6, 8)]
    lay = ops.SpliceLayout(tokens=tokens, labels=tokens, embed=embed, audio_tok=a, video_tok=v, prompts=prompts,
                           marker_ids=(V - 4, V - 3, V - 2, V - 1), has_bos=True)
    outs = [torch.empty(B, sl, H, device=device, dtype=torch.bfloat16) for sl in lay.seq_len]
    outl = [torch.empty(B, sl, device=device, dtype=torch.int64) for sl in lay.seq_len]
    ms = timeit(lambda: ops.splice_prompt(lay, outs, outl))
    rows = B * sum(lay.seq_len)
    byts = rows * H * 2 + rows * 8 + (B * (n_a + n_v) + rows - 2 * B * (n_a + n_v)) * H * 2
    out["splice_train_3tasks"] = {"us": round(ms * 1e3, 1), "GBs": round(byts / ms / 1e6, 1),
                                  "frac": round(byts / ms / 1e6 / peaks["hbm_gbs"], 3), "rows": rows}
    return out


def parity_and_cpu_baseline(mod, device):
    """The CPU oracle built from THIS module's weights (oracle/pairing.py), one utterance, rates (4, 2): (a) |loss_gpu -
    loss_oracle| per task at the benchmark's full geometry, (b) the timed CPU train step (the reported baseline)."""
    from omni_avsr_b200.synthetic import synthetic_batch, to_device
    from oracle.modeling import training_step
    from oracle.pairing import oracle_from_product
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oracle = oracle_from_product(mod)
    for p_ in oracle.parameters():
        p_.requires_grad_(False)
    for n_, p_ in oracle.named_parameters():
        if "lora_" in n_ or n_.startswith("audio_proj") or n_.startswith("video_proj"):
            p_.requires_grad_(True)
    cpu = synthetic_batch(1, mod.tokenizer, seconds=16.0, text_len=48, seed=99)
    gpu = to_device(cpu, device)
    with torch.no_grad():
        mod.training_step(gpu, 0, rates=(4, 2))
        got = [float(x) for x in mod.last_losses]
    opt = torch.optim.AdamW([p_ for p_ in oracle.parameters() if p_.requires_grad], lr=1e-4, weight_decay=0.1, betas=(0.9, 0.98))
    t0 = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    loss, parts = training_step(oracle, cpu, 4, 2)
    loss.backward()
    torch.nn.utils.clip_grad_norm_([p_ for p_ in oracle.parameters() if p_.requires_grad], 10.0)
    opt.step()
    dt = time.perf_counter() - t0
    want = [float(x) for x in parts]
    parity = {"what": "three task losses (matry_weights applied) of one 16 s utterance at rates (4, 2), full config-2 geometry: CUDA "
                      "path vs the CPU oracle holding the same weights",
              "loss_gpu": [round(x, 4) for x in got], "loss_oracle": [round(x, 4) for x in want],
              "max_abs_diff": round(max(abs(a - b) for a, b in zip(got, want)), 4), "tolerance": 5e-2}
    base = {"value": round(1.0 / dt, 4), "unit": "utterances/s", "cores": cores, "kind": "port",
            "sample": f"1 train step of batch 1 (16 s clip, same architecture / weights / config, bf16, torch {torch.__version__} "
                      f"CPU, {cores} threads), no warm-up", "ms_per_step": round(1e3 * dt, 1)}
    # informational: the SAME oracle in PyTorch eager on this GPU (what a user of the reference gets by moving it to the
    # B200 unchanged: library kernels, three LLM passes, full-vocabulary fp32 logits) -- context for the speed-up, not a target
    eager = None
    try:
        Be = 8
        oracle_g = oracle.to(device)
        cpu_e = synthetic_batch(Be, mod.tokenizer, seconds=16.0, text_len=48, seed=77)
        gpu_e = {k: (v.to(device) if torch.is_tensor(v) and k != "lengths" else v) for k, v in cpu_e.items()}
        opt_g = torch.optim.AdamW([p_ for p_ in oracle_g.parameters() if p_.requires_grad], lr=1e-4, weight_decay=0.1,
                                  betas=(0.9, 0.98))
        ts = []
        for k in range(4):
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            opt_g.zero_grad(set_to_none=True)
            l_, _ = training_step(oracle_g, gpu_e, *RATE_GRID[k % 4])
            l_.backward()
            torch.nn.utils.clip_grad_norm_([p_ for p_ in oracle_g.parameters() if p_.requires_grad], 10.0)
            opt_g.step()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t1)
        dt_e = sum(ts[1:]) / len(ts[1:])
        eager = {"value": round(Be / dt_e, 2), "unit": "utterances/s", "batch": Be, "ms_per_step": round(1e3 * dt_e, 1),
                 "what": "oracle restatement of the reference's train step in PyTorch eager (bf16, cuBLAS / cuDNN / ATen) on the "
                         "same B200, same weights; 3 timed steps after 1 warm-up"}
    except Exception as ex:      # noqa: BLE001
        eager = {"error": repr(ex)[:300]}
    base["gpu_eager_baseline"] = eager
    del oracle
    torch.cuda.empty_cache()
    return parity, base


def build_oracle():
    """The reference's CPU path = the oracle restatement at the benchmark's architecture (random init)."""
    from oracle import encoders as oe
    from oracle import llm_lora as ol
    from oracle import modeling as omod
    torch.manual_seed(0)
    cfg = ol.llama_3_2_1b()
    lc = ol.LoRA_config(32, 4, True, False, True, True)
    prompts = {k: torch.randint(0, 1000, (1, n)) for k, n in (("audio", 6), ("video", 6), ("audiovisual", 8))}
    m = omod.AVSR_LLMs(cfg, lc, oe.WHISPER["openai/whisper-medium.en"], oe.AVHubertCfg(), 2048, [4, 16], [2, 5],
                       "avg-pooling", prompts, (128257, 128258, 128259, 128260), False, [1.0, 1.5, 1.0], True,
                       eos_id=128001, pad_id=128256).bfloat16().eval()
    for p in m.parameters():
        p.requires_grad_(False)
    for n, p in m.named_parameters():
        if "lora_" in n or n.startswith("audio_proj") or n.startswith("video_proj"):
            p.requires_grad_(True)
    return m, omod


def cpu_baseline(steps=1, warmup=0, B=1):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m, omod = build_oracle()
    g = torch.Generator().manual_seed(1234)
    batch = {"tokens": torch.randint(0, 128000, (B, 48), generator=g),
             "audio": torch.nn.functional.layer_norm(torch.randn(B, 256000, generator=g), (256000,)).unsqueeze(-1).bfloat16(),
             "lengths": torch.full((B,), 256000, dtype=torch.int64),
             "video": ((torch.rand(B, 400, 1, 88, 88, generator=g) - 0.421) / 0.165).bfloat16()}
    batch["tokens"][:, 0], batch["tokens"][:, -1] = 128000, 128001
    batch["labels"] = batch["tokens"].clone()
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-4, weight_decay=0.1, betas=(0.9, 0.98))
    times = []
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss, _ = omod.training_step(m, batch, *RATE_GRID[k % 4])
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in m.parameters() if p.requires_grad], 10.0)
        opt.step()
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
    tot = sum(times)
    return {"value": round(B * len(times) / tot, 4), "unit": "utterances/s", "cores": cores, "kind": "port",
            "sample": f"{len(times)} train step(s) of batch {B} (16 s clip, same architecture/config, bf16, "
                      f"torch {torch.__version__} CPU, {cores} threads) after {warmup} warm-up",
            "ms_per_step": round(1e3 * tot / len(times), 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(steps=args.steps, warmup=args.warmup, B=1)
    line = {"impl": "reference", "metric": "utterances/sec (train step)", "value": cb["value"], "unit": "utterances/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": 1, "note": "reference CPU path (oracle port; the reference "
                       "itself cannot be imported in this image, see DESIGN.md), bounded sample: batch 1 per step"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--decode-batch", type=int, default=64)
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="skip BASELINE configs 3 / 4 / 5 and the strong-scaling point (extra objects of the JSON line)")
    ap.add_argument("--llm", default=None, help="other backbone of the reference's table, e.g. Qwen/Qwen2.5-3B (BASELINE config "
                    "4) or meta-llama/Meta-Llama-3.1-8B (config 5); default = the judged Llama-3.2-1B line")
    ap.add_argument("--workload", default="omni", choices=["omni", "mtsk"],
                    help="omni = BASELINE config 2 (default, the judged line); mtsk = Llama-MTSK train step (extra line)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
