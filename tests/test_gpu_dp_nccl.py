"""Two ranks on two GPUs (NCCL): one data-parallel train step == the single-process step on the concatenated batch.

Semantics under test (reference: DDPStrategy train_OmniAVSR.py:46-49 averages gradients; lightning_OmniAVSR.py:171-173 scales
the loss by W / sum(B)): with equal per-rank batches b the 2-rank gradient is (1 / W) sum_r grad(loss_r * W / (W b)) and the
1-rank gradient on the W b utterances is grad(loss_all / (W b)) with loss_all = mean_r loss_r (equal label counts), i.e.
grad_2rank == W * grad_1rank -- the reference's own scaling rule, kept.  Skipped on boxes with one GPU."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _slice(batch, idx):
    return {k: (v[idx].contiguous() if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == 4 else v) for k, v in batch.items()}


def _worker(rank, world, port, q):
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    from omni_avsr_b200.synthetic import synthetic_batch, to_device
    from tests._small import small_module
    mod = small_module(seed=0)
    cpu = synthetic_batch(4, mod.tokenizer, seconds=2.0, text_len=12, seed=7)
    used = mod.model.flat.used

    def step(batch):
        mod.zero_grad_flat()
        loss = mod.training_step(to_device(batch, "cuda"), 0, rates=(4, 2))
        loss.backward()
        red = getattr(mod, "_reducer", None)
        factor = 1.0
        if red is not None:
            factor = red.finish()
            mod._reducer = None
        torch.cuda.synchronize()
        return mod.model.flat.grad[:used].float().clone() * factor, loss.item()

    single = None
    if rank == 0:
        single, _ = step(cpu)                     # world size 1: no process group yet
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    mine = list(range(rank, 4, world))
    g2, loss2 = step(_slice(cpu, mine))
    # gradient accumulation under data parallelism (Lightning accumulate_grad_batches): the rank's two utterances as two
    # micro-batches, the first with no_sync (no collective, gradients accumulate in the flat buffer), ONE reduction after the
    # second -- the same gradient as the one-batch step
    mod.zero_grad_flat()
    for i, u in enumerate(mine):
        mod.no_sync = i < len(mine) - 1
        # the reference's rule divides the MEAN loss by the batch size (loss * W / sum B): a micro-batch of b of the rank's
        # B utterances therefore carries the weight (b / B) ** 2
        (mod.training_step(to_device(_slice(cpu, [u]), "cuda"), 0, rates=(4, 2)) * (1.0 / len(mine)) ** 2).backward()
        assert (getattr(mod, "_reducer", None) is None) == mod.no_sync
    mod.no_sync = False
    factor = mod._reducer.finish()
    mod._reducer = None
    torch.cuda.synchronize()
    g3 = mod.model.flat.grad[:used].float().clone() * factor
    if rank == 0:
        want = single * world
        err = (g2 - want).abs().max().item() / want.abs().max().item()
        err_acc = (g3 - want).abs().max().item() / want.abs().max().item()
        q.put((max(err, err_acc), float(want.abs().max()), loss2))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_step_equals_single_rank_step_on_the_concatenated_batch():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, scale, _ = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert scale > 0
    assert err <= 3e-2, err       # two bf16 backward chains over different batch splits
