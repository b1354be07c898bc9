"""LoRA phase-1 (down-projection) GEMM shapes of the train step on the tile widths the dispatcher offers:
   python tools/lora_phase1_probe.py
Measured (B200, L2 flushed, CUDA events): M=31232 N=256 K=2048: 59.4 / 51.2 / 43.0 us at block_n 64 / 128 / 256 (HBM floor
22 us); N=128: 43.0 / 38.9; the K <= 1024 shapes sit at 15 - 18 us whatever the tile (fixed cost).  Moving phase 1 to the
256-wide tiles would save ~0.3 ms of a 189 ms step: not done (the per-task row tables are built for 64-column blocks)."""
import sys, torch
sys.path.insert(0, "/root/repo")
from omni_avsr_b200 import ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, n=9):
    fn(); ts=[]
    for _ in range(n):
        flush.zero_()
        s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e)*1e3)
    return sorted(ts)[n//2]
for (M,N,K) in [(31232,256,2048),(31232,128,2048),(31232,128,512),(12800,128,1024),(12800,64,1024)]:
    x=torch.randn(M,K,device="cuda").bfloat16(); w=torch.randn(N,K,device="cuda").bfloat16()
    res={}
    for bn in (64,128,256):
        if bn > N and bn != 64: continue
        try:
            res[bn]=round(t(lambda: ops.gemm(x,w,block_n=bn)),1)
        except Exception as ex:
            res[bn]=repr(ex)[:60]
    print((M,N,K), res, "HBM floor us", round((M*K*2+M*N*2)/6.5e6,1))
