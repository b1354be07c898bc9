"""K = 1024 encoder GEMMs with the GELU epilogues (Whisper fc1: act="gelu"; AV-HuBERT fc1: "gelu_keep") and the plain launch of
the same shape: CUDA events, median of 9, L2 flushed."""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from omni_avsr_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=9):
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


for M, N, K in ((48000, 4096, 1024), (12800, 4096, 1024)):
    x = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = (torch.randn(N, device="cuda", generator=g) * 0.1).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * M * N * K
    r = {"M": M, "N": N, "K": K}
    for name, fn in (("plain_bias", lambda: ops.gemm(x, W, bias=b, out=out, block_n=256)),
                     ("gelu", lambda: ops.gemm(x, W, bias=b, out=out, act="gelu", block_n=256)),
                     ("gelu_keep", lambda: ops.gemm(x, W, bias=b, out=out, out2=out2, act="gelu_keep", block_n=256))):
        ms = timeit(fn)
        r[name] = {"ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}
    print(json.dumps(r))
