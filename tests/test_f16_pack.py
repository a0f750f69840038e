"""Host side of the fp16 hi/lo GEMM operand format (ff3d_tcgemm_f16): the weight images round-trip, keep fp32-grade
precision and follow the same swizzle rule as the TF32 images."""
import pytest
import torch


@pytest.mark.parametrize("taps,cin,cout", [(27, 128, 128), (9, 64, 256), (27, 16, 32), (49, 8, 64), (1, 32, 16), (3, 832, 64)])
def test_f16_weight_images_round_trip(taps, cin, cout):
    from focalformer3d_b200.ops import tc_weight_images_f16, unpack_images_f16
    from focalformer3d_b200.lib import lib
    f16_stages = lib.ff3d_tcgemm_f16_stages
    g = torch.Generator().manual_seed(taps * 1000 + cin)
    w = torch.randn(taps, cin, cout, generator=g) / (taps * cin) ** 0.5
    imgs, bn = tc_weight_images_f16(w)
    assert imgs.dtype == torch.float16 and tuple(imgs.shape) == (cout // bn, f16_stages(cin, taps), 2, bn, 64)
    hi, lo = unpack_images_f16(imgs, taps, cin, cout)
    rec = hi.double() + lo.double() / 2048.0
    # 11 + 11 significand bits: |w - (hi + lo/2048)| <= 2^-22 |w| (+ the fp16 subnormal floor of the scaled lo part)
    assert ((rec - w.double()).abs() <= w.double().abs() * 2.0 ** -21 + 2.0 ** -36).all()
    assert torch.equal(hi, w.half().float())


def test_f16_and_tf32_images_share_the_swizzle_rule():
    """chunk j of row n sits at chunk j ^ (n % 8) in both formats (8 halves vs 4 floats per 16-byte chunk)."""
    from focalformer3d_b200.ops import tc_weight_images_f16
    w = torch.zeros(1, 64, 16)
    w[0, 8 * 3 + 2, 5] = 1.0                                  # K = 26 -> chunk 3, half 2 of output row 5
    imgs, bn = tc_weight_images_f16(w)
    row = imgs[0, 0, 0, 5]
    pos = int(row.nonzero()[0])
    assert pos == ((3 ^ (5 & 7)) * 8 + 2)
    assert imgs[0, 0, 1].abs().sum() == 0                     # 1.0 is exact in fp16: empty lo image


def test_f16_split_matches_the_numerics_study():
    from focalformer3d_b200.ops import split_f16
    g = torch.Generator().manual_seed(0)
    a = torch.randn(64, 512, generator=g) * 3.0
    b = torch.randn(512, 32, generator=g) / 512 ** 0.5
    ah, al = split_f16(a)
    bh, bl = split_f16(b)
    d = lambda x, y: x.double() @ y.double()
    got = d(ah, bh) + (d(ah, bl) + d(al, bh)) / 2048.0
    ref = a.double() @ b.double()
    scale = a.double().abs() @ b.double().abs()
    assert ((got - ref).abs() / scale).max().item() < 5e-7
