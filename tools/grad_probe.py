"""Where does the AV-HuBERT adapter-gradient error of the CUDA path come from?  Small model, same weights in the product
(GPU, bf16), the bf16 oracle and the fp32 oracle (CPU).  Compares, against the fp32 oracle: d(loss)/d(video encoder output)
-- everything downstream of the encoder: LLM, splice, projector, compression backward -- and the LoRA gradients of both
encoder layers -- the encoder's own backward."""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from oracle.modeling import training_step  # noqa: E402
from oracle.pairing import oracle_from_product  # noqa: E402
from omni_avsr_b200.synthetic import synthetic_batch, to_device  # noqa: E402
from tests._small import small_module  # noqa: E402


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def cos(a, b):
    return torch.nn.functional.cosine_similarity(a.float().cpu().flatten(), b.float().cpu().flatten(), dim=0).item()


def hook_output(module, store):
    orig = module.forward

    def fwd(*a, **k):
        y = orig(*a, **k)
        t = y[0] if isinstance(y, tuple) else y
        if t.requires_grad:
            t.retain_grad()
        store.append(t)
        return y
    module.forward = fwd


mod = small_module()
o16 = oracle_from_product(mod)
o32 = oracle_from_product(mod, dtype=torch.float32)
cpu = synthetic_batch(2, mod.tokenizer, seconds=2.0, text_len=12, seed=7)
gpu = to_device(cpu, "cuda")
cpu32 = {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in cpu.items()}
ra, rv = 4, 2
s16, s32, sp = [], [], []
hook_output(o16.video_encoder, s16)
hook_output(o32.video_encoder, s32)
ef = mod.model.video_encoder.extract_finetune


def ef_hook(*a, **k):
    y = ef(*a, **k)
    if y[0].requires_grad:
        y[0].retain_grad()
    sp.append(y[0])
    return y


mod.model.video_encoder.extract_finetune = ef_hook
l32, _ = training_step(o32, cpu32, ra, rv)
l32.backward()
l16, _ = training_step(o16, cpu, ra, rv)
l16.backward()
mod.zero_grad_flat()
loss = mod.training_step(gpu, 0, rates=(ra, rv))
loss.backward()
torch.cuda.synchronize()
print("loss", loss.item(), l16.item(), l32.item())
g32, g16, gp = s32[-1].grad, s16[-1].grad, sp[-1].grad
print("features   : prod vs fp32", rel(sp[-1], s32[-1]), " bf16 oracle vs fp32", rel(s16[-1], s32[-1]))
print("d(features): prod vs fp32", rel(gp, g32), cos(gp, g32), " bf16 oracle vs fp32", rel(g16, g32), cos(g16, g32),
      " |g| max", g32.abs().max().item())
for li in (1, 0):
    vatt = mod.model.video_encoder.encoder.layers[li].self_attn
    r_ = round(128 / 16)
    for got, key in ((vatt.lora_up.grad[:128, :r_], "lora_up_Q"), (vatt.lora_down.grad[:r_], "lora_down_Q"),
                     (vatt.lora_up.grad[128:, :r_], "lora_up_V"), (vatt.lora_down.grad[vatt.lora_down.shape[0] // 2:][:r_], "lora_down_V")):
        w16 = getattr(o16.video_encoder.encoder.layers[li].self_attn, key).weight.grad
        w32 = getattr(o32.video_encoder.encoder.layers[li].self_attn, key).weight.grad
        print(f"layer {li} {key:12s}: prod vs fp32 {rel(got, w32):.3f} cos {cos(got, w32):.4f} | bf16 oracle vs fp32 {rel(w16, w32):.3f} "
              f"cos {cos(w16, w32):.4f} | max {w32.abs().max().item():.2e}")
