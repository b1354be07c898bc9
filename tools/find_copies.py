"""Where do the big ATen strided copies of a train step come from?  python tools/find_copies.py [--batch 32]
Prints the Python stack of every `aten::copy_` / `aten::contiguous` call that moves more than 32 MB."""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
args = ap.parse_args()
from omni_avsr_b200.synthetic import synthetic_batch, to_device  # noqa: E402
dev = torch.device("cuda", 0)
mod = bench.build_module(args, dev)
batch = to_device(synthetic_batch(args.batch, mod.tokenizer, seed=1234), dev)
for _ in range(2):
    mod.train_step(batch, rates=(4, 2), lr=1e-4)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    mod.train_step(batch, rates=(4, 2), lr=1e-4)
    torch.cuda.synchronize()
seen = {}
for ev in prof.events():
    if ev.name in ("aten::copy_", "aten::contiguous", "aten::clone") and ev.device_time_total > 100:
        st = [s for s in (ev.stack or []) if "omni_avsr_b200" in s or "bench.py" in s][:3]
        key = (ev.name, tuple(st), str(ev.input_shapes)[:80])
        seen[key] = seen.get(key, 0) + ev.device_time_total
for k, v in sorted(seen.items(), key=lambda kv: -kv[1])[:12]:
    print(f"{v / 1e3:8.3f} ms  {k[0]}  {k[2]}")
    for s in k[1]:
        print("      ", s)
