"""Oracle self-checks by independent formulations: the [upstream] restatements are unpinned by any
reference fixture, so each is cross-checked against a second, structurally different implementation."""
import numpy as np
import torch

from oracle import sparse, voxelize
from oracle.transformer import ms_deform_attn_core, ms_deform_attn_loop, MultiheadAttention


def _rand_sparse(B=2, shape=(9, 12, 12), n=150, C=6, seed=0):
    g = torch.Generator().manual_seed(seed)
    D, H, W = shape
    lin = torch.randperm(B * D * H * W, generator=g)[:n]
    idx = torch.stack([lin // (D * H * W), (lin // (H * W)) % D, (lin // W) % H, lin % W], 1)
    return sparse.SparseTensor(torch.randn(n, C, generator=g), idx, shape, B)


def _dense_eq(a, b, tol=1e-4):
    da, db = a.dense(), b.dense()
    assert da.shape == db.shape
    assert (da - db).abs().max().item() < tol


def test_subm_conv_matches_masked_dense_conv3d():
    x = _rand_sparse()
    conv = sparse.SpConv3d(6, 5, 3, 1, 1, subm=True)
    torch.nn.init.normal_(conv.weight)
    with torch.no_grad():
        y = conv(x)
        ref = sparse.dense_reference_conv(x, conv.weight, 3, 1, 1, subm=True)
    _dense_eq(y, ref)


def test_strided_sparse_conv_matches_dense_conv3d():
    for k, s, p in [((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (0, 1, 1)), ((3, 1, 1), (2, 1, 1), (0, 0, 0))]:
        x = _rand_sparse(seed=3)
        conv = sparse.SpConv3d(6, 4, k, s, p, subm=False)
        torch.nn.init.normal_(conv.weight)
        with torch.no_grad():
            y = conv(x)
            ref = sparse.dense_reference_conv(x, conv.weight, k, s, p, subm=False)
        assert y.indices.shape[0] == ref.indices.shape[0]
        _dense_eq(y, ref)


def test_spatial_shapes_of_the_shipped_encoder():
    # z 41 -> 21 -> 11 -> 5 -> 2 ; xy 1440 -> 720 -> 360 -> 180
    s = (41, 1440, 1440)
    s = sparse.out_shape(s, (3, 3, 3), (2, 2, 2), (1, 1, 1)); assert s == (21, 720, 720)
    s = sparse.out_shape(s, (3, 3, 3), (2, 2, 2), (1, 1, 1)); assert s == (11, 360, 360)
    s = sparse.out_shape(s, (3, 3, 3), (2, 2, 2), (0, 1, 1)); assert s == (5, 180, 180)
    s = sparse.out_shape(s, (3, 1, 1), (2, 1, 1), (0, 0, 0)); assert s == (2, 180, 180)


def test_hard_voxelize_matches_point_loop():
    rng = np.random.default_rng(0)
    pts = rng.uniform(-2.2, 2.2, (3000, 5)).astype(np.float32)
    pts[:, 2] = rng.uniform(-1.2, 1.2, 3000)
    args = ([0.25, 0.25, 0.5], [-2, -2, -1, 2, 2, 1], 3, 180)
    a = voxelize.hard_voxelize(pts, *args)
    b = voxelize.hard_voxelize_loop(pts, *args)
    for u, v in zip(a, b):
        assert u.shape == v.shape and np.array_equal(u, v)
    assert a[1].shape[0] == 180 and a[2].max() == 3          # both caps exercised


def test_hard_voxelize_edges():
    empty = voxelize.hard_voxelize(np.zeros((0, 5), np.float32), [0.5, 0.5, 0.5], [0, 0, 0, 2, 2, 2], 4, 10)
    assert empty[0].shape == (0, 4, 5) and empty[1].shape == (0, 3)
    # a point exactly on the upper bound is outside ([min, max) bins); on the lower bound it is inside
    p = np.array([[2.0, 1.0, 1.0, 0, 0], [0.0, 0.0, 0.0, 1, 1]], np.float32)
    v, c, n = voxelize.hard_voxelize(p, [0.5, 0.5, 0.5], [0, 0, 0, 2, 2, 2], 4, 10)
    assert c.tolist() == [[0, 0, 0]] and n.tolist() == [1]


def test_msda_grid_sample_matches_explicit_bilinear():
    g = torch.Generator().manual_seed(1)
    shapes = [(12, 10), (6, 5), (3, 3)]
    B, Nq, h, d, P = 2, 17, 4, 8, 4
    value = torch.randn(B, sum(a * b for a, b in shapes), h, d, generator=g)
    loc = torch.rand(B, Nq, h, len(shapes), P, 2, generator=g) * 1.4 - 0.2     # includes out-of-range samples
    aw = torch.rand(B, Nq, h, len(shapes), P, generator=g)
    a = ms_deform_attn_core(value, shapes, loc, aw)
    b = ms_deform_attn_loop(value, shapes, loc, aw)
    assert (a - b).abs().max().item() < 1e-5


def test_mha_wrapper_adds_pos_to_q_and_k_only():
    torch.manual_seed(0)
    m = MultiheadAttention(32, 4).eval()
    x, pos = torch.randn(9, 2, 32), torch.randn(9, 2, 32)
    with torch.no_grad():
        out = m(x, x, x, None, query_pos=pos, key_pos=pos)
        w, bq = m.attn.in_proj_weight, m.attn.in_proj_bias
        q = (x + pos) @ w[:32].t() + bq[:32]
        k = (x + pos) @ w[32:64].t() + bq[32:64]
        v = x @ w[64:].t() + bq[64:]
        def heads(t):
            return t.view(9, 2, 4, 8).permute(1, 2, 0, 3)
        a = torch.softmax(heads(q) @ heads(k).transpose(-1, -2) / 8 ** 0.5, -1) @ heads(v)
        ref = x + m.attn.out_proj(a.permute(2, 0, 1, 3).reshape(9, 2, 32))
    assert (out - ref).abs().max().item() < 1e-5


def test_rotated_iou_and_nms_restatements():
    """[upstream] iou3d boxes_iou_bev / nms_gpu and circle_nms restatements (oracle/head.py) against independent
    formulations: a raster estimate of the rotated IoU, and brute-force properties of the greedy NMS outputs."""
    import math
    import torch
    from oracle.head import boxes_iou_bev, xywhr2xyxyr, circle_nms, nms_rotated_bev
    a = torch.tensor([[0., 0., 2., 1., 0.3], [0.5, 0.2, 1.5, 1., -1.0], [5, 5, 1, 1, 0.]])
    b = torch.tensor([[0.2, 0.1, 2., 1., 0.9], [0., 0., 2., 1., 0.3], [0., 0., 1., 0.5, 0.3]])
    iou = boxes_iou_bev(xywhr2xyxyr(a), xywhr2xyxyr(b))
    assert abs(iou[0, 1].item() - 1.0) < 1e-6 and abs(iou[0, 2].item() - 0.25) < 1e-6 and iou[2].abs().max().item() == 0

    def inside(p, bx):
        c, s = math.cos(bx[4]), math.sin(bx[4])
        dx, dy = p[0] - bx[0], p[1] - bx[1]
        return abs(dx * c + dy * s) <= bx[2] / 2 and abs(-dx * s + dy * c) <= bx[3] / 2
    N = 160
    for i, j in ((0, 0), (1, 0), (1, 1), (1, 2)):
        ia = ib = ii = 0
        for yi in range(N):
            for xi in range(N):
                p = (-2 + 4.5 * xi / N, -2 + 4.5 * yi / N)
                A, B = inside(p, a[i].tolist()), inside(p, b[j].tolist())
                ia += A; ib += B; ii += A and B
        assert abs(ii / max(ia + ib - ii, 1) - iou[i, j].item()) < 0.03
    g = torch.Generator().manual_seed(0)
    xy, sc = torch.rand(80, 2, generator=g) * 3, torch.rand(80, generator=g)
    keep = circle_nms(torch.cat([xy, sc[:, None]], 1).numpy(), 0.175)
    kept = torch.tensor(keep)
    d2 = ((xy[kept][:, None] - xy[kept][None]) ** 2).sum(-1) + torch.eye(len(keep)) * 9
    assert bool((d2 > 0.175).all())                                   # no two kept boxes within the radius
    for i in set(range(80)) - set(keep):                              # every dropped box has a better kept box in range
        assert bool(((((xy[kept] - xy[i]) ** 2).sum(-1) <= 0.175) & (sc[kept] >= sc[i])).any())
    boxes = torch.cat([xy, torch.rand(80, 2, generator=g) + 0.5, torch.rand(80, 1, generator=g) * 3], 1)
    k = nms_rotated_bev(xywhr2xyxyr(boxes), sc, 0.1)
    m = boxes_iou_bev(xywhr2xyxyr(boxes[k]), xywhr2xyxyr(boxes[k]))
    assert bool((m - torch.eye(len(k)) <= 0.1 + 1e-6).all()) and bool((sc[k][:-1] >= sc[k][1:]).all())
