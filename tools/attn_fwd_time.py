"""Times the attention forward alone on the path shapes (CUDA events, L2 flushed): python tools/attn_fwd_time.py"""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from omni_avsr_b200 import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, B, S, nh, nkv, hd, causal in [("whisper-m", 16, 1500, 16, 16, 64, False), ("avhubert-l", 16, 400, 16, 16, 64, False),
                                         ("llama1b-avsr", 16, 460, 32, 8, 64, True)]:
    M = B * S
    qkv = torch.randn(M, (nh + 2 * nkv) * hd, device="cuda").bfloat16()
    out = torch.empty(M, nh * hd, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(nh, M, device="cuda")
    ts = []
    for it in range(13):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.attention_fwd(qkv, out, [(0, B, S, 0)], nh, nkv, hd, causal, lse=lse)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts[3:])[5]
    fl = 4.0 * B * nh * S * S * hd * (0.5 if causal else 1.0)
    print(json.dumps({"shape": name, "fwd_ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}), flush=True)
