"""Runs warm-up steps, then ONE train step inside cudaProfilerStart/Stop so that
   ncu --profile-from-start off ...  python tools/one_step.py [--batch 8]
profiles exactly one step (launch list) or the first N launches of a kernel (--set full -k regex:... -c 3)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--warm", type=int, default=2)
    args = ap.parse_args()
    from omni_avsr_b200.synthetic import synthetic_batch, to_device
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    mod = bench.build_module(args, dev)
    batch = to_device(synthetic_batch(args.batch, mod.tokenizer, seed=1234), dev)
    for _ in range(args.warm):
        mod.train_step(batch, rates=(4, 2), lr=1e-4)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    mod.train_step(batch, rates=(4, 2), lr=1e-4)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("one step done")


if __name__ == "__main__":
    main()
