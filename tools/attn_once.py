"""One forward + backward attention launch per path shape (for ncu captures):
   ncu --set full -k regex:attn_ ... python tools/attn_once.py [whisper|avh|llama]"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from omni_avsr_b200 import ops  # noqa: E402

SHAPES = {"whisper": (16, 1500, 16, 16, 64, False), "avh": (16, 400, 16, 16, 64, False),
          "llama": (16, 460, 32, 8, 64, True), "llama8b": (16, 460, 32, 8, 128, True)}
name = sys.argv[1] if len(sys.argv) > 1 else "whisper"
B, S, nh, nkv, hd, causal = SHAPES[name]
M = B * S
qkv = torch.randn(M, (nh + 2 * nkv) * hd, device="cuda").bfloat16()
out = torch.empty(M, nh * hd, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(nh, M, device="cuda")
dout = torch.randn(M, nh * hd, device="cuda").bfloat16()
dqkv = torch.empty_like(qkv)
seg = [(0, B, S, 0)]
for _ in range(2):
    ops.attention_fwd(qkv, out, seg, nh, nkv, hd, causal, lse=lse)
    ops.attention_bwd(qkv, out, dout, lse, dqkv, seg, nh, nkv, hd, causal)
torch.cuda.synchronize()
print("done", name)
