import sys, torch
sys.path.insert(0, "/root/repo")
from tests._small import small_module
from oracle.pairing import oracle_from_product
from oracle.modeling import training_step
from omni_avsr_b200.synthetic import synthetic_batch
mod = small_module()
o = oracle_from_product(mod).cuda()
cpu = synthetic_batch(2, mod.tokenizer, seconds=2.0, text_len=12, seed=7)
gpu = {k: (v.cuda() if torch.is_tensor(v) and k != "lengths" else v) for k, v in cpu.items()}
l, parts = training_step(o, gpu, 4, 2)
l.backward()
print("eager ok", float(l))
