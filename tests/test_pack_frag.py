"""Host-side weight packing of the one-launch decoder stage (no GPU needed)."""
import torch


def test_pack_frag_layout_cpu_side():
    """Fragment order of the packed weights: lane (g, t) of n-tile j, k-step s holds w[8j+g, 16s+2t:+2] and [.. +8 ..]."""
    from focalformer3d_b200 import ops
    w = torch.randn(30, 48, generator=torch.Generator().manual_seed(0))
    f = ops.pack_frag(w)
    assert tuple(f.shape) == (4, 3, 32, 4) and f.dtype == torch.int32
    wp = torch.zeros(32, 48)
    wp[:30] = w
    hi, lo = ops.split_f16(wp)
    h16 = f.view(torch.float16).view(4, 3, 32, 4, 2)
    for j, s, g, t in ((0, 0, 0, 0), (1, 2, 3, 1), (3, 1, 7, 3), (2, 0, 5, 2)):
        r, c = 8 * j + g, 16 * s + 2 * t
        assert torch.equal(h16[j, s, g * 4 + t, 0], hi[r, c:c + 2]) and torch.equal(h16[j, s, g * 4 + t, 1], hi[r, c + 8:c + 10])
        assert torch.equal(h16[j, s, g * 4 + t, 2], lo[r, c:c + 2]) and torch.equal(h16[j, s, g * 4 + t, 3], lo[r, c + 8:c + 10])
    uh, ul = ops.unpack_frag(f, 30, 48)
    assert torch.equal(uh, hi[:30].float()) and torch.equal(ul, lo[:30].float())
