"""CPU, world_size 2, gloo: the data-parallel host logic (utterance sharding, W/sum(B) loss scale, flat-gradient
all-reduce + 1/W factor) reproduces single-process results."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from omni_avsr_b200 import dp
    g = torch.Generator().manual_seed(0)
    per_utt_grads = torch.randn(6, 37, generator=g)          # gradient contribution of each of 6 utterances
    mine = dp.shard_utterances(6, rank, world)
    local_b = len(mine)
    # each rank: mean over its utterances (what a per-rank mean loss produces), then reference scaling W / sum(B)
    scale = dp.loss_scale(local_b)
    flat = per_utt_grads[mine].sum(0) * scale
    flat2 = flat.clone()
    factor = dp.allreduce_flat_grad(flat)
    # the asynchronous forms used by ModelModule_LLM: gather started early, gradient buffer reduced in two pieces
    ls = dp.LossScale(local_b)
    assert abs(float(ls.value()) - float(scale)) < 1e-9
    red = dp.GradReducer(flat2, split=10)
    red.hook(None)                       # tail [10, 37) goes out first (autograd hook on the LLM input)
    assert red.finish() == factor
    assert torch.equal(flat, flat2)
    red = dp.GradReducer(flat2.clone(), split=10)      # hook never fired: one reduction of the whole buffer
    assert red.finish() == factor
    q.put((rank, mine, float(scale), (flat * factor).tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3, 5]
    assert abs(res[0][2] - 2 / 6) < 1e-7
    g = torch.Generator().manual_seed(0)
    per_utt_grads = torch.randn(6, 37, generator=g)
    # DDP semantics: average over ranks of (sum over local utterances * W/sum(B))  ==  mean over all utterances
    want = per_utt_grads.sum(0) * (2 / 6) / 2
    for r in res:
        assert torch.allclose(torch.tensor(r[3]), want, atol=1e-6)
