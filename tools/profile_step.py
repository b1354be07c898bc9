"""Kernel-time breakdown of one train step with torch.profiler (CUPTI timeline, no replay):
   python tools/profile_step.py [--batch 8] [--out gpurun_out/step_profile.json]
Prints the top kernels by total device time and their share of the step."""
import argparse
import json
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--decode", action="store_true", help="profile one greedy decode instead of a train step")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "step_profile.json"))
    args = ap.parse_args()
    from omni_avsr_b200.synthetic import synthetic_batch, to_device
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    mod = bench.build_module(args, dev)
    batch = to_device(synthetic_batch(args.batch, mod.tokenizer, seed=1234), dev)
    from torch.profiler import ProfilerActivity, profile
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.decode:
        # one greedy decode (audiovisual, rates 4/2, 32 new tokens) of `batch` utterances
        batch["tokens"] = batch["tokens"][:, :1].contiguous()
        mod.args.modality = "audiovisual"
        mod.args.downsample_ratio_test_matry_audio, mod.args.downsample_ratio_test_matry_video = 4, 2
        mod.on_test_epoch_start()
        mod.model.decode_no_trim = True

        def run():
            with torch.no_grad():
                mod.test_step(batch)
        for _ in range(2):
            run()
    else:
        def run():
            mod.train_step(batch, rates=(4, 2), lr=1e-4)
        for k in range(5):
            mod.train_step(batch, rates=bench.RATE_GRID[k % 4], lr=1e-4)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        s.record()
        run()
        e.record()
        torch.cuda.synchronize()
    step_ms = s.elapsed_time(e)
    agg = defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            agg[ev.name][0] += ev.device_time / 1e3 if hasattr(ev, "device_time") else ev.cuda_time / 1e3
            agg[ev.name][1] += 1
    rows = sorted(((v[0], v[1], k) for k, v in agg.items()), reverse=True)
    total = sum(r[0] for r in rows)
    out = {"step_ms_events": step_ms, "sum_kernel_ms": total, "batch": args.batch,
           "kernels": [{"name": n[:160], "ms": round(ms, 3), "launches": c, "share_of_kernel_time": round(ms / total, 4)}
                       for ms, c, n in rows[:60]]}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(f"step {step_ms:.2f} ms, kernel time {total:.2f} ms, {sum(r[1] for r in rows)} launches")
    for ms, c, n in rows[:45]:
        print(f"{ms:9.3f} ms {100 * ms / total:5.1f}% x{c:5d}  {n[:110]}")


if __name__ == "__main__":
    main()
