"""One greedy decode (audiovisual, rates 4/2) of --batch utterances inside cudaProfilerStart/Stop, for
   ncu --profile-from-start off ... python tools/decode_once.py   (set OMNI_DECODE_NO_GRAPH=1 for eager steps)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    from omni_avsr_b200.synthetic import synthetic_batch, to_device
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    mod = bench.build_module(args, dev)
    batch = to_device(synthetic_batch(args.batch, mod.tokenizer, seed=1234), dev)
    batch["tokens"] = batch["tokens"][:, :1].contiguous()
    mod.args.modality = "audiovisual"
    mod.args.downsample_ratio_test_matry_audio, mod.args.downsample_ratio_test_matry_video = 4, 2
    mod.on_test_epoch_start()
    mod.model.decode_no_trim = True
    with torch.no_grad():
        mod.test_step(batch)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        mod.test_step(batch)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("decode done")


if __name__ == "__main__":
    main()
