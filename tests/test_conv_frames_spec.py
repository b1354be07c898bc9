"""CPU check of the table-driven convolution's host logic (ops.ConvFramesSpec): replaying the K-extension table and the
filter-pattern matrix with plain torch matmuls must reproduce F.conv2d (resnet.py:35-74 geometries: 3x3 pad 1 stride 1 / 2,
1x1 stride 2; ring-padded or plain input frames; 1, 2 or 4 output pixels per 256-wide N tile)."""
import pytest
import torch


def _replay(spec, x):
    """out[frame, N] from the table exactly as the GEMM walks it: per N tile, sum over entries A2[:, col:+64] @ B2[row:+256, col:+64]^T"""
    N = x.shape[0]
    if spec.in_ring:
        xr = torch.zeros(N, spec.Hin + 2, spec.Win + 2, spec.Ci)
        xr[:, 1:-1, 1:-1] = x.permute(0, 2, 3, 1)
    else:
        xr = x.permute(0, 2, 3, 1)
    a2 = xr.reshape(N, -1)
    b2 = spec.b2.float()
    out = torch.zeros(N, spec.N)
    tab = spec.table[0]
    for t in range(tab.shape[0]):
        for j in range(tab.shape[1]):
            ac, br, bc, _ = tab[t, j].tolist()
            if br < 0:
                continue
            out[:, t * 256:(t + 1) * 256] += a2[:, ac:ac + 64] @ b2[br:br + 256, bc:bc + 64].T
    P = spec.Hout * spec.Wout
    return out.view(N, spec.PA, spec.Co)[:, :P].reshape(N, spec.Hout, spec.Wout, spec.Co).permute(0, 3, 1, 2)


@pytest.mark.parametrize("H,W,Ci,Co,k,stride,ring", [(22, 22, 64, 128, 3, 2, True), (11, 11, 128, 128, 3, 1, False),
                                                     (11, 11, 128, 256, 3, 2, False), (6, 6, 64, 256, 3, 1, False),
                                                     (6, 6, 64, 512, 3, 2, False), (3, 3, 64, 512, 3, 1, False),
                                                     (22, 22, 64, 128, 1, 2, True), (7, 5, 64, 64, 3, 1, False),
                                                     (5, 7, 128, 64, 1, 2, True)])
def test_conv_frames_table_reproduces_conv2d(H, W, Ci, Co, k, stride, ring):
    from omni_avsr_b200.ops import ConvFramesSpec
    g = torch.Generator().manual_seed(H * 7 + Co)
    w = (torch.randn(Co, Ci, k, k, generator=g) / (k * Ci ** 0.5)).bfloat16()
    x = torch.randn(3, Ci, H, W, generator=g).bfloat16().float()
    spec = ConvFramesSpec(w, H, W, stride, ring)
    want = torch.nn.functional.conv2d(x, w.float(), stride=stride, padding=1 if k == 3 else 0)
    got = _replay(spec, x)
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 1e-4 * want.abs().max().item()
    # no entry multiplies the zero padding, every entry addresses a 64-column block inside the operands
    tab = spec.table[0]
    valid = tab[..., 1] >= 0
    assert (tab[..., 0][valid] % 64 == 0).all() and (tab[..., 2][valid] + 64 <= spec.b2.shape[1]).all()
    assert spec.N % 256 == 0 and spec.PA * Co == spec.N
