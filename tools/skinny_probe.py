"""Decode-step qkv GEMM (64 tokens, N=3072, K=2048) on the weight-streaming kernel, in isolation: with / without the LoRA
K-extension, for the split-K factor given by OMNI_SKINNY_SPLIT (unset = the dispatcher's choice).  CUDA events, L2 flushed."""
import json
import os
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from omni_avsr_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
M, N, K = 64, 3072, 2048
x = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
T = (torch.randn(M, 256, device="cuda", generator=g) * 0.5).bfloat16()
up = (torch.randn(N, 256, device="cuda", generator=g) * 0.05).bfloat16()
nt = N // 128
tab = torch.full((1, nt, 4, 4), -1, dtype=torch.int32)
for t in range(nt):
    if t < 16 or t >= 20:          # q and v tiles carry two adapters (task + shared), one 64-column block each
        tab[0, t, 0] = torch.tensor([0, t * 128, 0, 0])
        tab[0, t, 1] = torch.tensor([64, t * 128, 64, 0])
tab = tab.cuda()
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


res = {"split": os.environ.get("OMNI_SKINNY_SPLIT", "auto")}
res["plain_us"] = round(timeit(lambda: ops.gemm(x, W, skinny=True)), 2)
res["ext_us"] = round(timeit(lambda: ops.gemm(x, W, ext=(T, up, tab), block_n=128, skinny=True)), 2)
res["o_proj_us"] = round(timeit(lambda: ops.gemm(x, W[:2048], skinny=True)), 2)
print(json.dumps(res))
