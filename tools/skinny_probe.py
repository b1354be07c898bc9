"""Weight-streaming GEMMs of the decode step (64 tokens) in isolation, for the split-K factor given by OMNI_SKINNY_SPLIT
(unset = the dispatcher's choice): CUDA events, median of 15, L2 flushed between launches (the events add ~5 us of launch
latency to every figure: compare columns, not absolute values).
   python tools/skinny_probe.py [llama1b|qwen3b]"""
import json
import os
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from omni_avsr_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
which = sys.argv[1] if len(sys.argv) > 1 else "llama1b"
H, I, QKV = (2048, 8192, 3072) if which == "llama1b" else (2048, 11008, 2560)
shapes = {"qkv": (QKV, H), "o_proj": (H, H), "down": (H, I), "lora_down": (256, H), "gate_up": (2 * I, H)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=15):
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2] * 1e3


res = {"model": which, "split": os.environ.get("OMNI_SKINNY_SPLIT", "auto")}
for name, (N, K) in shapes.items():
    x = (torch.randn(64, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    if name == "gate_up":
        act = torch.empty(64, N // 2, device="cuda", dtype=torch.bfloat16)
        res[name] = round(timeit(lambda: ops.gemm(x, W, act="swiglu64", out2=act, skinny=True)), 1)
    else:
        res[name] = round(timeit(lambda: ops.gemm(x, W, skinny=True)), 1)
print(json.dumps(res))
