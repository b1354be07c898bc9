"""One fused compression -> projector -> splice launch at BASELINE config-2 geometry (for ncu captures)."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from omni_avsr_b200 import ops  # noqa: E402
from tools.bench_kernels import fused_setup  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ra, rv = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4, 2)
c = fused_setup(B, ra, rv)
lay = ops.SpliceLayout(tokens=c["tokens"], labels=c["tokens"], embed=c["embed"], audio_tok=None, video_tok=None,
                       prompts=c["prompts"], marker_ids=c["marker"], has_bos=True, n_audio=c["na"], n_video=c["nv"])
outs = [torch.empty(B, s, c["H"], device="cuda", dtype=torch.bfloat16) for s in lay.seq_len]
outl = [torch.empty(B, s, device="cuda", dtype=torch.int64) for s in lay.seq_len]
a_in = ops.PoolProjectInput(c["xa"], 800, ra, *c["pa"])
v_in = ops.PoolProjectInput(c["xv"], 400, rv, *c["pv"])
for _ in range(3):
    ops.pool_project_splice(lay, outs, outl, a_in, v_in, "avg-pooling")
torch.cuda.synchronize()
