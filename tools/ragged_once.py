"""Ragged (bucketed, frame-budget) train-step throughput alone: python tools/ragged_once.py [max_frames] [n_utts]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class A:
    workload, llm, batch = "omni", None, 32


if __name__ == "__main__":
    dev = torch.device("cuda:0")
    mod = bench.build_module(A(), dev)
    mf = int(sys.argv[1]) if len(sys.argv) > 1 else 12800
    nu = int(sys.argv[2]) if len(sys.argv) > 2 else 384
    print(json.dumps(bench.measure_ragged(mod, dev, 0, 1, lambda: torch.cuda.synchronize(), max_frames=mf, n_utts=nu)))
