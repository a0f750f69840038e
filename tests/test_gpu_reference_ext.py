"""The reference's OWN CUDA extensions (locatt_ops, bev_pool), compiled from the sources in place under /root/reference
into oracle/_ref/ by oracle/build_ref.py, as the checker on the GPU box: they pin the oracle's restatements
(oracle/bev.py local_similar / local_weighting, oracle/camera.py voxel pooling) and the CUDA product kernels
(ff3d_local_attention, ff3d_lss_splat) to the reference's real kernels.  Skipped when oracle/_ref was not built."""
import math
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ext(name):
    from oracle.build_ref import load
    m = load(name)
    if m is None:
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py needs /root/reference)")
    return m


def test_local_attention_against_the_reference_kernels():
    from focalformer3d_b200 import ops
    from oracle.bev import local_similar, local_weighting
    ext = _ext("ref_localattention")
    B, C, H, W, K = 2, 128, 13, 17, 9
    g = torch.Generator().manual_seed(21)
    q, k, v = (torch.randn(B, C, H, W, generator=g) * s for s in (1.3, 1.3, 1.0))
    sim = ext.similar_forward(q.cuda(), k.cuda(), K, K)                     # [B, H, W, K*K]   (kernels.cuh cc2k)
    assert tuple(sim.shape) == (B, H, W, K * K)
    assert (sim.cpu() - local_similar(q, k, K, K)).abs().max().item() < 1e-4
    wgt = torch.softmax(sim / math.sqrt(C), -1).contiguous()                # encoder_utils.py:161
    out = ext.weighting_forward(v.cuda(), wgt, K, K)                        # [B, C, H, W]     (kernels.cuh ck2c_ori)
    assert (out.cpu() - local_weighting(v, wgt.cpu(), K, K)).abs().max().item() < 1e-5
    ours = torch.empty(B, H, W, C, device="cuda")
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().cuda()
    ops.local_attention(nhwc(q), nhwc(k), nhwc(v), ours, K)
    assert (ours.permute(0, 3, 1, 2) - out).abs().max().item() < 2e-5


def test_lift_splat_against_the_reference_bev_pool():
    from focalformer3d_b200 import ops
    from focalformer3d_b200.synth import synth_cameras
    from oracle.camera import LiftSplatShoot, lidar2img_to_rots_trans
    ext = _ext("ref_bev_pool")
    B, N, (H, W) = 2, 6, (64, 96)
    lss = LiftSplatShoot(img_scale=(H, W), pc_range=[-7.2, -7.2, -5.0, 7.2, 7.2, 3.0], grid=0.6, inputC=256, outputC=128,
                         camC=64, downsample=4)
    D, fH, fW = lss.D, lss.fH, lss.fW
    g = torch.Generator().manual_seed(4)
    logits = torch.randn(B * N, D + 64, fH, fW, generator=g) * 2.0
    rt = [lidar2img_to_rots_trans(synth_cameras(N, (H, W), seed=70 + b)) for b in range(B)]
    rots, trans = torch.stack([r for r, _ in rt]), torch.stack([t for _, t in rt])
    depth = logits[:, :D].softmax(1)
    feat = (depth.unsqueeze(1) * logits[:, D:].unsqueeze(2)).view(B, N, 64, D, fH, fW).permute(0, 1, 3, 4, 5, 2)
    geom = lss.get_geometry(rots, trans)
    # ---- the reference's bev_pool path (lss.py:286-322 + bev_pool_op.py:78-97) on its own CUDA kernel
    nx = lss.nx.tolist()
    idx = lss.voxel_indices(geom).view(-1, 3)
    batch_ix = torch.arange(B).repeat_interleave(idx.shape[0] // B)[:, None]
    coords = torch.cat([idx, batch_ix], 1)
    kept = (idx[:, 0] >= 0) & (idx[:, 0] < nx[0]) & (idx[:, 1] >= 0) & (idx[:, 1] < nx[1]) & (idx[:, 2] >= 0) & (idx[:, 2] < nx[2])
    x = feat.reshape(-1, 64)[kept].cuda()
    coords = coords[kept].cuda()
    Dz, Hx, Wy = nx[2], nx[0], nx[1]
    ranks = coords[:, 0] * (Wy * Dz * B) + coords[:, 1] * (Dz * B) + coords[:, 2] * B + coords[:, 3]
    order = ranks.argsort()
    x, coords, ranks = x[order].contiguous(), coords[order], ranks[order]
    first = torch.ones(x.shape[0], dtype=torch.bool, device="cuda")
    first[1:] = ranks[1:] != ranks[:-1]
    starts = torch.where(first)[0].int()
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = x.shape[0] - starts[-1]
    pooled = ext.bev_pool_forward(x, coords.int().contiguous(), lengths, starts, B, Dz, Hx, Wy)    # [B, Z, X, Y, C]
    ref = pooled.permute(0, 3, 2, 1, 4).reshape(B, Wy, Hx, Dz * 64).cpu()                          # ours: [B, y, x, z*64 + c]
    # ---- the oracle restatement and the fused CUDA kernel
    vox = lss.voxel_pooling(geom, feat)                                                             # [B, C, Z, X, Y]
    orc = vox.permute(0, 4, 3, 2, 1).reshape(B, Wy, Hx, Dz * 64)
    assert torch.equal(orc != 0, ref != 0)
    assert (orc - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())
    dn = torch.zeros(B * N, fH, fW, 128)
    dn[..., :64] = logits[:, D:].permute(0, 2, 3, 1)
    dn[..., 64:64 + D] = logits[:, :D].permute(0, 2, 3, 1)
    bev = torch.empty((B, Wy, Hx, Dz * 64), device="cuda")
    ops.lss_splat(dn.cuda(), lss.frustum.data.cuda(), rots.reshape(-1, 9).cuda(), trans.reshape(-1, 3).cuda(), bev, N, D,
                  (lss.bx - lss.dx / 2.0).tolist(), lss.dx.tolist())
    got = bev.cpu()
    assert torch.equal(got != 0, ref != 0), "voxel indices differ from the reference's bev_pool path"
    assert (got - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())
