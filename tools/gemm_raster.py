"""gate_up GEMM (M=31232, N=16384, K=2048, SwiGLU epilogue) and a plain GEMM of the same shape: CUDA-event time per launch
(L2 flushed between launches).  Run twice (OMNI_GEMM_NO_NGROUP=1 and unset) to compare the N-grouped raster with N-fastest;
under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_bf16_tn_2cta` the same
script gives the DRAM traffic per launch."""
import json
import os
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from omni_avsr_b200 import ops  # noqa: E402

M, N, K = 31232, 16384, 2048
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
gu = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
act = torch.empty(M, N // 2, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=7):
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


fused = timeit(lambda: ops.gemm(x, W, out=gu, out2=act, act="swiglu64", block_n=256))
plain = timeit(lambda: ops.gemm(x, W, out=gu, block_n=256))
fl = 2.0 * M * N * K
print(json.dumps({"raster": "n_fastest" if os.environ.get("OMNI_GEMM_NO_NGROUP") else "n_grouped", "M": M, "N": N, "K": K,
                  "swiglu64_ms": round(fused, 4), "swiglu64_tflops": round(fl / fused / 1e9, 1), "plain_ms": round(plain, 4),
                  "plain_tflops": round(fl / plain / 1e9, 1)}))
