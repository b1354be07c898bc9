"""Single-token attention kernel alone at the decode step's geometry (for ncu captures and A/B timing):
   python tools/decode_attn_once.py [B nh nkv hd max_len pos]        (OMNI_DA_FFMA=1: the FFMA formulation)
   ncu --set full -k regex:decode_attn -c 2 --launch-skip 4 ... python tools/decode_attn_once.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omni_avsr_b200 import ops  # noqa: E402

if __name__ == "__main__":
    B, nh, nkv, hd, max_len, pos = (int(a) for a in sys.argv[1:7]) if len(sys.argv) >= 7 else (64, 32, 8, 64, 512, 445)
    L = 16                                                     # one cache per layer: the step never re-reads a warm cache
    g = torch.Generator(device="cuda").manual_seed(0)
    kc = [torch.randn(B, nkv, max_len, hd, device="cuda", generator=g).bfloat16() for _ in range(L)]
    vc = [torch.randn(B, nkv, max_len, hd, device="cuda", generator=g).bfloat16() for _ in range(L)]
    qkv = (torch.randn(B, (nh + 2 * nkv) * hd, device="cuda", generator=g)).bfloat16()
    ang = torch.rand(max_len, hd // 2, device="cuda", generator=g) * 6.28
    emb = torch.cat([ang, ang], dim=-1)
    cos_t, sin_t = emb.cos().bfloat16().contiguous(), emb.sin().bfloat16().contiguous()
    out = torch.empty(B, nh * hd, device="cuda", dtype=torch.bfloat16)
    len_idx = torch.tensor([pos], device="cuda", dtype=torch.int64)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for i in range(4):
        ops.decode_attention(qkv, kc[i], vc[i], len_idx, out, B, nh, nkv, hd, rope=(cos_t, sin_t))
    times = []
    for rep in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(L):
            ops.decode_attention(qkv, kc[i], vc[i], len_idx, out, B, nh, nkv, hd, rope=(cos_t, sin_t))
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e) * 1e3 / L)
    kv_bytes = 2 * B * nkv * (pos + 1) * hd * 2
    us = sorted(times)[2]
    print({"us_per_launch": round(us, 2), "kv_MB": round(kv_bytes / 1e6, 1), "GBs": round(kv_bytes / us / 1e3, 1),
           "variant": "ffma" if os.environ.get("OMNI_DA_FFMA") == "1" else "mma", "geometry": (B, nh, nkv, hd, max_len, pos)})
