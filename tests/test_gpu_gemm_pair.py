"""GPU: the K-extended (Omni-LoRA) GEMM on the CTA-pair (cta_group::2) kernel -- tasks constant over pairs of 128-row
tiles (segments aligned to 256 rows), enough tiles to engage the pair kernel; same fp32 reference as test_gpu_gemm.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pair_aligned", [True, False])
def test_lora_extension_pair_kernel(pair_aligned):
    from omni_avsr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(17)
    H, Hq, Hkv, r, s = 1024, 2048, 512, 64, 0.125
    tiles = [0, 0] * 6 + [2, 2] * 5 + [1, 1] * 7          # 36 tiles = 18 pairs x 12 n-tiles = 216 pair tiles
    M = 128 * len(tiles) - 40                              # ragged tail inside the last pair
    x = torch.randn(M, H, device="cuda", generator=g).bfloat16()
    Wqkv = (torch.randn(Hq + 2 * Hkv, H, device="cuda", generator=g) * 0.05).bfloat16()
    Aq = (torch.randn(4, r, H, device="cuda", generator=g) * 0.05).bfloat16()
    Av = (torch.randn(4, r, H, device="cuda", generator=g) * 0.05).bfloat16()
    Bq = (torch.randn(4, Hq, r, device="cuda", generator=g) * 0.05).bfloat16()
    Bv = (torch.randn(4, Hkv, r, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(Hq + 2 * Hkv, device="cuda", generator=g).bfloat16()
    tile_group = torch.tensor(tiles, device="cuda", dtype=torch.int32)
    A_pack = torch.cat([Aq.reshape(4 * r, H), Av.reshape(4 * r, H)], dim=0).contiguous()
    brow = torch.empty(3, 4, dtype=torch.int32)
    for grp in range(3):
        brow[grp] = torch.tensor([grp * r, 3 * r, 4 * r + grp * r, 4 * r + 3 * r])
    T = ops.gemm(x, A_pack, n=4 * r, alpha=s, tile_group=tile_group, b_row_table=brow.cuda(), block_n=64)
    BN = 256
    n_tiles = (Hq + 2 * Hkv) // BN
    B_pack = torch.cat([Bq.reshape(4 * Hq, r), Bv.reshape(4 * Hkv, r)], dim=0).contiguous()
    ext = torch.full((3, n_tiles, 2, 4), -1, dtype=torch.int32)
    for grp in range(3):
        for nt in range(n_tiles):
            n0 = nt * BN
            if n0 < Hq:
                ext[grp, nt, 0] = torch.tensor([0, grp * Hq + n0, 0, 0])
                ext[grp, nt, 1] = torch.tensor([r, 3 * Hq + n0, 0, 0])
            elif n0 >= Hq + Hkv:
                nv = n0 - Hq - Hkv
                ext[grp, nt, 0] = torch.tensor([2 * r, 4 * Hq + grp * Hkv + nv, 0, 0])
                ext[grp, nt, 1] = torch.tensor([3 * r, 4 * Hq + 3 * Hkv + nv, 0, 0])
    qkv = ops.gemm(x, Wqkv, bias=bias, tile_group=tile_group, ext=(T, B_pack, ext.cuda().contiguous()), block_n=BN,
                   pair_aligned=pair_aligned)
    xf = x.float()
    ref = xf @ Wqkv.float().t() + bias.float()
    for i, grp in enumerate(tiles):
        rows = slice(i * 128, min((i + 1) * 128, M))
        tq = (s * (xf[rows] @ Aq[grp].float().t())).bfloat16().float()
        tqs = (s * (xf[rows] @ Aq[3].float().t())).bfloat16().float()
        tv = (s * (xf[rows] @ Av[grp].float().t())).bfloat16().float()
        tvs = (s * (xf[rows] @ Av[3].float().t())).bfloat16().float()
        ref[rows, :Hq] += tq @ Bq[grp].float().t() + tqs @ Bq[3].float().t()
        ref[rows, Hq + Hkv:] += tv @ Bv[grp].float().t() + tvs @ Bv[3].float().t()
    err = (qkv.float() - ref).abs().max().item()
    assert err <= 5e-3 * ref.abs().max().item(), err
