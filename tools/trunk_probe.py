import os, sys, time, torch
sys.path.insert(0, "/root/repo")
from omni_avsr_b200.encoders import _ResEncoder
torch.manual_seed(0)
enc = _ResEncoder((64, 128, 256, 512)).cuda().bfloat16().eval()
with torch.no_grad():
    for m in enc.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            m.running_var.uniform_(0.5, 1.5); m.running_mean.normal_(0, 0.1)
x = torch.randn(32, 1, 400, 88, 88, device="cuda").bfloat16()
with torch.no_grad():
    y = enc(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); y = enc(x); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
print("L1 frames" if os.environ.get("OMNI_TRUNK_FRAMES_L1") else "L1 ring", "ms", sorted(ts)[2], float(y.float().abs().mean()), float(y.float().std()))
torch.save(y.cpu(), "/tmp/y_%s.pt" % ("f" if os.environ.get("OMNI_TRUNK_FRAMES_L1") else "r"))
