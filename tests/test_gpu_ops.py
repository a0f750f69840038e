"""Per-operator parity: every C-ABI entry point against the CPU oracle / plain torch fp32 on seeded inputs.
Tolerances: indices, counts, coordinates, labels bit-exact; fp32 values 1e-4 abs on O(1) data unless noted."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from focalformer3d_b200 import ops as _ops
    return _ops


def _pack_lin(w):
    from focalformer3d_b200.model import pack_linear
    return pack_linear(w, None, "cuda")


def _pack_conv(w):
    from focalformer3d_b200.model import pack_conv2d
    return pack_conv2d(w, None, "cuda")


# ------------------------------------------------------------------------------------------------ voxelize
@pytest.mark.parametrize("max_voxels,max_points", [(5000, 10), (150, 3)])
def test_voxelize_matches_oracle(ops, max_voxels, max_points):
    from oracle.voxelize import hard_voxelize
    rng = np.random.default_rng(0)
    vs, rg = [0.25, 0.25, 0.5], [-4.0, -4.0, -1.0, 4.0, 4.0, 1.0]
    samples = []
    for n in (4000, 0, 2500):                      # includes an empty sample
        p = rng.uniform(-4.4, 4.4, (n, 5)).astype(np.float32)
        p[:, 2] = rng.uniform(-1.2, 1.2, n)
        samples.append(p)
    offs = np.cumsum([0] + [len(s) for s in samples]).tolist()
    allp = torch.from_numpy(np.concatenate(samples)).cuda()
    out = ops.voxelize(allp, offs, vs, rg, max_points, max_voxels, mean_ld=8, want_voxels=True)
    n_dev = out["n_dev"].cpu().numpy()
    base = 0
    for b, p in enumerate(samples):
        v, c, n = hard_voxelize(p, vs, rg, max_points, max_voxels)
        M = v.shape[0]
        assert n_dev[1 + b] == M
        sl = slice(base, base + M)
        assert np.array_equal(out["coors"][sl].cpu().numpy(), np.concatenate([np.full((M, 1), b, np.int32), c], 1))
        assert np.array_equal(out["num_points"][sl].cpu().numpy(), n)
        assert np.array_equal(out["voxels"][sl].cpu().numpy(), v)                      # bit-exact copies
        mean = v[:, :, :5].sum(1) / np.maximum(n, 1)[:, None].astype(np.float32)
        np.testing.assert_allclose(out["mean"][sl, :5].cpu().numpy(), mean, rtol=1e-6, atol=1e-6)
        assert (out["mean"][sl, 5:] == 0).all()
        base += M
    assert n_dev[0] == base


def test_hard_vfe_matches_oracle(ops):
    from oracle.voxelize import HardVFE
    from focalformer3d_b200.model import bn_scale_shift
    rng = np.random.default_rng(3)
    p = rng.uniform(-3.9, 3.9, (5000, 5)).astype(np.float32)
    p[:, 2] = rng.uniform(-0.9, 0.9, 5000)
    vox = ops.voxelize(torch.from_numpy(p).cuda(), [0, 5000], [0.25, 0.25, 0.5], [-4.0, -4.0, -1.0, 4.0, 4.0, 1.0], 5, 4000,
                       want_voxels=True)
    n = int(vox["n_dev"][0].item())
    torch.manual_seed(0)
    vfe = HardVFE(5, (64,)).eval()
    for prm in vfe.parameters():
        torch.nn.init.normal_(prm, std=0.5)
    vfe.vfe_layers[0].norm.running_mean.normal_(0, 0.3)
    vfe.vfe_layers[0].norm.running_var.uniform_(0.5, 1.5)
    with torch.no_grad():
        want = vfe(vox["voxels"][:n].cpu(), vox["num_points"][:n].cpu())
    sd = {k: v.detach() for k, v in vfe.state_dict().items()}
    s, b = bn_scale_shift(sd, "vfe_layers.0.norm", 1e-3)
    w = (sd["vfe_layers.0.linear.weight"].double() * s.view(-1, 1)).t().contiguous().float().cuda()
    got = ops.vfe_hard(vox, w, b.float().cuda(), 64, 5, 5)[:n].cpu()
    assert (vox["num_points"][:n] < 5).any() and (vox["num_points"][:n] == 5).any()   # padded and full voxels
    assert (got - want).abs().max().item() < 1e-4


# ------------------------------------------------------------------------------------------------ implicit GEMM
@pytest.mark.parametrize("M,K,N,act", [(300, 128, 128, 1), (257, 256, 20, 0), (1000, 1024, 128, 0), (64, 8, 16, 2),
                                       (130, 384, 288, 0)])
def test_linear(ops, M, K, N, act):
    g = torch.Generator().manual_seed(M + N)
    x, x2 = torch.randn(M, K, generator=g), torch.randn(M, K, generator=g)
    w, b, r = torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    ref = (x + x2).double() @ w.double().t() + b.double() + r.double()
    ref = {0: ref, 1: ref.relu(), 2: ref.clamp(0, 6)}[act].float()
    bias = torch.cat([b, torch.zeros((-N) % 4)]).cuda()
    y = ops.linear(x.cuda(), _pack_lin(w), bias, act=act, res=r.cuda(), x2=x2.cuda(), cout=N)
    assert (y.cpu() - ref).abs().max().item() < 2e-4
    # act before the residual add (query_feat += roi_feat)
    y2 = ops.linear(x.cuda(), _pack_lin(w), bias, act=1, res=r.cuda(), cout=N, res_after_act=True)
    ref2 = (x.double() @ w.double().t() + b.double()).relu() + r.double()
    assert (y2.cpu() - ref2.float()).abs().max().item() < 2e-4


@pytest.mark.parametrize("cin,cout,k,stride,H,W", [(16, 24, 3, 1, 13, 11), (32, 64, 3, 2, 12, 12), (32, 64, 3, 2, 11, 9),
                                                   (64, 10, 3, 1, 9, 9), (256, 128, 1, 1, 7, 5), (128, 128, 3, 1, 20, 20),
                                                   (32, 128, 3, 1, 140, 141), (32, 256, 3, 2, 200, 200)])
def test_conv2d_nhwc(ops, cin, cout, k, stride, H, W):
    g = torch.Generator().manual_seed(cin + cout + k)
    B = 2
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=k // 2).relu().float()
    Ho, Wo = ref.shape[2:]
    wide = torch.full((B, Ho, Wo, cout + 8), 7.0, device="cuda")             # channel-slice output (concat-free cat)
    ops.conv2d(x.permute(0, 2, 3, 1).contiguous().cuda(), _pack_conv(w), b.cuda(), wide[..., 4:4 + cout], k,
               stride=stride, act=1)
    assert (wide[..., 4:4 + cout].permute(0, 3, 1, 2).cpu() - ref).abs().max().item() < 2e-4
    assert (wide[..., :4] == 7).all() and (wide[..., 4 + cout:] == 7).all()    # neighbours untouched


def test_conv2d_batch_strided_views_and_residual(ops):
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 2, 16, 8, 8
    tokens = H * W + (H // 2) * (W // 2)
    ms = torch.randn(B, tokens, C, generator=g).cuda()
    w = torch.randn(C, C, 3, 3, generator=g) / 12
    lv0 = ms[:, :H * W].view(B, H, W, C)
    lv1 = ms[:, H * W:].view(B, H // 2, W // 2, C)
    ref = F.conv2d(lv0.permute(0, 3, 1, 2).cpu().double(), w.double(), stride=2, padding=1).float()
    ops.conv2d(lv0, _pack_conv(w), None, lv1, 3, stride=2)
    assert (lv1.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() < 2e-4
    x = torch.randn(B, H, W, C, generator=g).cuda()
    out = torch.empty_like(x)
    ops.conv2d(x, _pack_conv(w), None, out, 3, res=x)
    ref = F.conv2d(x.permute(0, 3, 1, 2).cpu().double(), w.double(), padding=1).float() + x.permute(0, 3, 1, 2).cpu()
    assert (out.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() < 2e-4


def test_transposed_conv_lattice(ops):
    g = torch.Generator().manual_seed(9)
    B, ci, co, H, W = 2, 32, 24, 6, 5
    x = torch.randn(B, ci, H, W, generator=g)
    w = torch.randn(ci, co, 2, 2, generator=g) / ci ** 0.5
    ref = F.conv_transpose2d(x.double(), w.double(), stride=2).float()
    out = torch.empty((B, 2 * H, 2 * W, co), device="cuda")
    from focalformer3d_b200.model import pack_taps
    for dy in range(2):
        for dx in range(2):
            ops.conv2d(x.permute(0, 2, 3, 1).contiguous().cuda(), pack_taps(w[:, :, dy, dx].unsqueeze(0), "cuda"), None,
                       out, 1, up=(2, dy, dx))
    assert (out.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() < 2e-4


def test_dwconv_and_layernorm(ops):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 64, 9, 7, generator=g) * 3
    w, b = torch.randn(64, 1, 3, 3, generator=g), torch.randn(64, generator=g)
    ref = F.conv2d(x, w, b, padding=1, groups=64).clamp(0, 6)
    out = torch.empty((2, 9, 7, 64), device="cuda")
    ops.dwconv3x3(x.permute(0, 2, 3, 1).contiguous().cuda(), w.reshape(64, 9).t().contiguous().cuda(), b.cuda(), out, act=2)
    assert (out.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() < 1e-5
    y = torch.randn(77, 128, generator=g) * 2 + 1
    gm, bt = torch.randn(128, generator=g), torch.randn(128, generator=g)
    o = ops.layernorm(y.cuda(), gm.cuda(), bt.cuda())
    assert (o.cpu() - F.layer_norm(y, (128,), gm, bt)).abs().max().item() < 1e-5


# ------------------------------------------------------------------------------------------------ sparse conv
def _rand_level(B, shape, n, C, seed):
    g = torch.Generator().manual_seed(seed)
    D, H, W = shape
    lin = torch.randperm(B * D * H * W, generator=g)[:n]
    idx = torch.stack([lin // (D * H * W), (lin // (H * W)) % D, (lin // W) % H, lin % W], 1).int()
    return idx, torch.randn(n, C, generator=g)


def _to_dense(feat, coors, n, shape, B):
    D, H, W = shape
    out = torch.zeros(B, D, H, W, feat.shape[1])
    c = coors[:n].long()
    out[c[:, 0], c[:, 1], c[:, 2], c[:, 3]] = feat[:n]
    return out


@pytest.mark.parametrize("sort_rows", [False, True])
@pytest.mark.parametrize("k,s,p", [(None, None, None), ((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (0, 1, 1)),
                                   ((3, 1, 1), (2, 1, 1), (0, 0, 0))])
def test_sparse_conv_matches_oracle(ops, k, s, p, sort_rows):
    """SubM and strided sparse convs vs the oracle; ``sort_rows`` stores the input level in tap-mask order first (what the
    sparse encoder does), so both homogeneous (mask-sorted, sparse tile masks) and random (full tile masks) tiles run."""
    from oracle import sparse as osp
    from focalformer3d_b200.model import pack_taps
    B, shape, n, Ci, Co = 2, (9, 20, 20), 900, 16, 32
    idx, feat = _rand_level(B, shape, n, Ci, 11)
    cap = n + 50                                                   # capacity > active count
    coors = torch.zeros((cap, 4), dtype=torch.int32)
    coors[:n] = idx
    x = torch.zeros((cap, Ci))
    x[:n] = feat
    n_dev = torch.tensor([n], dtype=torch.int32).cuda()
    lvl = ops.SparseLevel(coors.cuda(), n_dev, cap, B, shape)
    lvl.build_hash()
    subm = k is None
    kk = (3, 3, 3) if subm else k
    wt = torch.randn(*kk, Ci, Co, generator=torch.Generator().manual_seed(3)) / 8
    bias, res = torch.randn(Co), torch.randn(cap, Co)
    oconv = osp.SpConv3d(Ci, Co, kk, 1 if subm else s, 1 if subm else p, subm=subm)
    oconv.weight.data.copy_(wt)
    with torch.no_grad():
        ref = oconv(osp.SparseTensor(feat, idx, shape, B))
    wpk = pack_taps(wt.reshape(-1, Ci, Co), "cuda")
    xg, resg = x.cuda(), res.cuda()
    if sort_rows:
        perm = lvl.sort_by_mask()
        pl = perm[:n].long()
        assert torch.equal(pl.sort().values.cpu(), torch.arange(n))                  # a permutation of the active rows
        assert torch.equal(lvl.coors[:n].cpu(), coors[pl.cpu()])
        xg = ops.gather_rows(xg, perm, n_dev, Ci)
        resg = ops.gather_rows(resg, perm, n_dev, Co)
    if subm:
        out = torch.zeros((cap, Co), device="cuda")
        rb = lvl.subm_map()
        ops.sparse_conv(xg, rb, lvl.n_dev, wpk, bias.cuda(), out, act=1, res=resg)
        got = _to_dense(out.cpu(), lvl.coors.cpu(), n, shape, B)
        r = torch.relu(ref.features + bias + res[:n])               # oracle rows are in input order for SubM
        want = _to_dense(r, idx, n, shape, B)
    else:
        overflow = torch.zeros(1, dtype=torch.int32, device="cuda")
        nl, rb = lvl.downsample(k, s, p, 4 * cap, overflow, ldy=Co)
        out = torch.zeros((nl.cap, Co), device="cuda")
        ops.sparse_conv(xg, rb, nl.n_dev, wpk, bias.cuda(), out, act=0)
        no = int(nl.n_dev.item())
        assert int(overflow.item()) == 0 and no == ref.indices.shape[0]          # same output-site set size
        assert nl.shape == tuple(ref.spatial_shape)
        got = _to_dense(out.cpu(), nl.coors.cpu(), no, nl.shape, B)
        want = _to_dense(ref.features + bias, ref.indices, no, ref.spatial_shape, B)
        occ_g = _to_dense(torch.ones(nl.cap, 1), nl.coors.cpu(), no, nl.shape, B)
        occ_w = _to_dense(torch.ones(no, 1), ref.indices, no, ref.spatial_shape, B)
        assert torch.equal(occ_g, occ_w)                                         # identical active sites
    # the tile masks cover every (row, tap) pair of the rulebook and nothing outside the kernel volume
    no = int((nl if not subm else lvl).n_dev.item())
    nbr = rb.nbr[:, :no].cpu()
    tm = rb.tile_mask[:(no + 127) // 128].cpu().long() & 0xFFFFFFFF          # entries past the last tile are never read
    for t in range(nbr.shape[0]):
        rows = (nbr[t] >= 0).nonzero().flatten()
        assert bool(((tm[rows // 128] >> t) & 1).all())
    assert int(tm.max()) < (1 << nbr.shape[0])
    assert (got - want).abs().max().item() < 2e-4


def test_mask_sort_makes_tiles_homogeneous(ops):
    """A clustered (LiDAR-like: thin horizontal sheets) level: after sort_by_mask the per-tile OR masks must select far
    fewer (tile, tap) pairs than the 27 of unsorted tiles, and the SubM conv result must not depend on the order."""
    from oracle import sparse as osp
    from focalformer3d_b200.model import pack_taps
    g = torch.Generator().manual_seed(5)
    B, shape, Ci, Co = 2, (11, 120, 120), 32, 32
    cells = []
    for b in range(B):
        yx = torch.randint(0, 120, (9000, 2), generator=g)
        z = torch.randint(3, 5, (9000, 1), generator=g)                   # two z slabs -> kz = 0 / 2 taps mostly absent
        cells.append(torch.cat([torch.full((9000, 1), b), z, yx], 1))
    idx = torch.unique(torch.cat(cells), dim=0).int()
    idx = idx[torch.randperm(idx.shape[0], generator=g)]
    n = idx.shape[0]
    feat = torch.randn(n, Ci, generator=g)
    n_dev = torch.tensor([n], dtype=torch.int32).cuda()
    wt = torch.randn(3, 3, 3, Ci, Co, generator=g) / 16
    wpk = pack_taps(wt.reshape(-1, Ci, Co), "cuda")
    outs, pairs = [], []
    for sort_rows in (False, True):
        lvl = ops.SparseLevel(idx.cuda(), n_dev, n, B, shape)
        lvl.build_hash()
        x = feat.cuda()
        if sort_rows:
            x = ops.gather_rows(x, lvl.sort_by_mask(), n_dev, Ci)
        rb = lvl.subm_map()
        out = torch.zeros((n, Co), device="cuda")
        ops.sparse_conv(x, rb, n_dev, wpk, None, out, act=0)
        outs.append(_to_dense(out.cpu(), lvl.coors.cpu(), n, shape, B))
        tm = rb.tile_mask.cpu().long() & 0xFFFFFFFF
        pairs.append(sum(bin(int(v)).count("1") for v in tm[:(n + 127) // 128]) * 128)
        valid = int((rb.nbr[:, :n] >= 0).sum().item())
    assert torch.equal(outs[0], outs[1])             # bit-identical: skipped taps only ever added exact zeros
    assert pairs[1] < 0.6 * pairs[0], (pairs, valid)
    oconv = osp.SpConv3d(Ci, Co, 3, 1, 1, subm=True)
    oconv.weight.data.copy_(wt)
    with torch.no_grad():
        ref = oconv(osp.SparseTensor(feat, idx, shape, B))
    assert (outs[1] - _to_dense(ref.features, idx, n, shape, B)).abs().max().item() < 2e-4


@pytest.mark.parametrize("n,bits", [(1, 27), (31, 3), (4096, 9), (4097, 18), (100000, 27), (300000, 27)])
def test_radix_sort_pairs_matches_stable_sort(ops, n, bits):
    """ff3d_sort_pairs == torch.sort(stable=True) on the low `bits` key bits, with a device-side count < capacity."""
    import ctypes as C
    from focalformer3d_b200.lib import lib, check
    g = torch.Generator().manual_seed(n)
    cap = n + 77
    keys = torch.randint(0, 1 << bits, (cap,), generator=g, dtype=torch.int64)
    if n > 1000:
        keys[: n // 2] = keys[: n // 2] & 0x1C7                                   # heavy duplicates: stability matters
    k32 = keys.to(torch.int32).cuda()
    n_dev = torch.tensor([n], dtype=torch.int32).cuda()
    ws_bytes = lib.ff3d_sort_workspace_bytes(cap)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    ko = torch.full((cap,), -1, dtype=torch.int32, device="cuda")
    vo = torch.full((cap,), -1, dtype=torch.int32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(lib.ff3d_sort_pairs(C.c_void_p(k32.data_ptr()), None, C.c_void_p(n_dev.data_ptr()), cap, bits,
                              C.c_void_p(ko.data_ptr()), C.c_void_p(vo.data_ptr()), C.c_void_p(ws.data_ptr()), ws_bytes, st),
          "ff3d_sort_pairs")
    ref_k, ref_v = torch.sort(keys[:n], stable=True)
    assert torch.equal(ko[:n].cpu().long(), ref_k) and torch.equal(vo[:n].cpu().long(), ref_v)
    assert bool((vo[n:] == -1).all())                                             # nothing written past the count


def test_sparse_conv_many_tiles_256_row_path(ops):
    """Enough rows that the BN=128 kernel takes its 256-rows-per-CTA (MT=2) persistent path, SubM 32 -> 128."""
    from oracle import sparse as osp
    from focalformer3d_b200.model import pack_taps
    B, shape, n, Ci, Co = 2, (9, 90, 90), 42000, 32, 128
    idx, feat = _rand_level(B, shape, n, Ci, 21)
    n_dev = torch.tensor([n], dtype=torch.int32).cuda()
    wt = torch.randn(3, 3, 3, Ci, Co, generator=torch.Generator().manual_seed(5)) / 12
    oconv = osp.SpConv3d(Ci, Co, 3, 1, 1, subm=True)
    oconv.weight.data.copy_(wt)
    with torch.no_grad():
        ref = oconv(osp.SparseTensor(feat, idx, shape, B))
    want = _to_dense(ref.features, idx, n, shape, B)
    for sort_rows in (False, True):
        lvl = ops.SparseLevel(idx.cuda(), n_dev, n, B, shape)
        lvl.build_hash()
        x = feat.cuda()
        if sort_rows:
            x = ops.gather_rows(x, lvl.sort_by_mask(), n_dev, Ci)
        out = torch.zeros((n, Co), device="cuda")
        ops.sparse_conv(x, lvl.subm_map(), lvl.n_dev, pack_taps(wt.reshape(-1, Ci, Co), "cuda"), None, out, act=0)
        assert (_to_dense(out.cpu(), lvl.coors.cpu(), n, shape, B) - want).abs().max().item() < 1e-4


def test_sparse_to_bev_scatter(ops):
    """conv_out's row map (y_mode 2 of ff3d_sp_nbr_build): NHWC BEV element offsets == SparseConvTensor.dense() + view."""
    from focalformer3d_b200.model import pack_taps
    B, shape, n, C = 2, (5, 6, 5), 60, 8
    idx, feat = _rand_level(B, shape, n, C, 4)
    n_dev = torch.tensor([n], dtype=torch.int32).cuda()
    lvl = ops.SparseLevel(idx.cuda(), n_dev, n, B, shape)
    lvl.build_hash()
    overflow = torch.zeros(1, dtype=torch.int32, device="cuda")
    D2 = (shape[0] - 3) // 2 + 1
    bev = torch.zeros((B, shape[1], shape[2], D2 * C), device="cuda")
    nl, rb = lvl.downsample((3, 1, 1), (2, 1, 1), (0, 0, 0), n, overflow, ldy=D2 * C, sort_level=False,
                            bev=(shape[1], shape[2], C))
    w = torch.zeros(3, 1, 1, C, C)
    w[0, 0, 0] = torch.eye(C)                                      # out(o) = in(2*o_z) (tap 0 only): a pure scatter
    ops.sparse_conv(feat.cuda(), rb, nl.n_dev, pack_taps(w.reshape(-1, C, C), "cuda"), None, bev, act=0, cout=C)
    dense = _to_dense(feat, idx, n, shape, B)                      # [B,D,H,W,C]
    want = dense[:, 0:2 * D2:2].permute(0, 2, 3, 1, 4).reshape(B, shape[1], shape[2], D2 * C)
    assert torch.equal(bev.cpu(), want)


def test_down_build_overflow_flag(ops):
    B, shape, n = 1, (5, 8, 8), 200
    idx, _ = _rand_level(B, shape, n, 4, 8)
    lvl = ops.SparseLevel(idx.cuda(), torch.tensor([n], dtype=torch.int32).cuda(), n, B, shape)
    lvl.build_hash()
    overflow = torch.zeros(1, dtype=torch.int32, device="cuda")
    nl, _ = lvl.downsample((3, 3, 3), (2, 2, 2), (1, 1, 1), 16, overflow, ldy=4)
    assert int(overflow.item()) == 1 and int(nl.n_dev.item()) == 16


# ------------------------------------------------------------------------------------------------ HIP stage
@pytest.mark.parametrize("dataset,C,exempt", [("nuScenes", 10, (8, 9)), ("Waymo", 3, (1, 2))])
def test_hip_stage_matches_reference_logic(ops, dataset, C, exempt):
    from oracle.head import canonical_topk
    g = torch.Generator().manual_seed(7)
    B, H, W, Cf, k = 2, 20, 18, 32, 30
    logits = torch.randn(B, C, H, W, generator=g) * 2
    acc = (torch.rand(B, C, H, W, generator=g) > 0.2).float()
    feat = torch.randn(B, Cf, H, W, generator=g)
    cls_w, cls_b = torch.randn(Cf, C, generator=g), torch.randn(Cf, generator=g)
    # --- reference logic (focal_decoder.py:662-782)
    heat = logits.sigmoid() * acc
    lm = torch.zeros_like(heat)
    lm[:, :, 1:-1, 1:-1] = F.max_pool2d(heat, 3, 1, 0)
    for c in range(exempt[0], exempt[1] + 1):
        lm[:, c] = heat[:, c]
    nms = (heat * (heat == lm)).view(B, C, -1)
    top = canonical_topk(nms.view(B, -1), k)
    cls, pos = top // (H * W), top % (H * W)
    qf = feat.view(B, Cf, -1).gather(2, pos[:, None].expand(-1, Cf, -1)) + (cls_w[:, cls].permute(1, 0, 2) + cls_b[None, :, None])
    qs = nms.gather(2, pos[:, None].expand(-1, C, -1))
    sel = torch.zeros(B, C * H * W).scatter_(1, top, 1.0).view(B, C, H, W)
    selk = F.max_pool2d(sel, 3, 1, 1)
    selk[:, exempt[0]:exempt[1] + 1] = sel[:, exempt[0]:exempt[1] + 1]
    acc_ref = acc * (1 - selk)
    # --- kernel
    lg = torch.zeros(B, H, W, 12)
    lg[..., :C] = logits.permute(0, 2, 3, 1)
    accd, nmsd = acc.clone().cuda(), torch.empty(B, C, H, W, device="cuda")
    nq = 2 * k
    topd = torch.empty((B, k), dtype=torch.int32, device="cuda")
    qfd, qpd = torch.zeros(B * nq, Cf, device="cuda"), torch.zeros(B * nq, 2, device="cuda")
    qsd, qld = torch.zeros(B * nq, C, device="cuda"), torch.zeros(B * nq, dtype=torch.int32, device="cuda")
    ops.hip_stage(lg.cuda(), accd, nmsd, feat.permute(0, 2, 3, 1).contiguous().cuda(), cls_w.t().contiguous().cuda(),
                  cls_b.cuda(), k, 3, exempt, k, nq, topd, qfd, qpd, qsd, qld)
    assert (nmsd.cpu().view(B, C, -1) - nms).abs().max().item() < 1e-6
    assert torch.equal(topd.cpu().long(), top)                                    # bit-exact, canonical order
    sl = slice(k, 2 * k)
    assert torch.equal(qld.view(B, nq)[:, sl].cpu().long(), cls)
    assert (qfd.view(B, nq, Cf)[:, sl].cpu() - qf.transpose(1, 2)).abs().max().item() < 1e-5
    assert (qsd.view(B, nq, C)[:, sl].cpu() - qs.transpose(1, 2)).abs().max().item() < 1e-6
    px = torch.stack([(pos % W).float() + 0.5, (pos // W).float() + 0.5], -1)
    assert torch.equal(qpd.view(B, nq, 2)[:, sl].cpu(), px)
    assert torch.equal(accd.cpu(), acc_ref)


def test_hip_stage_degenerate_fewer_candidates_than_k(ops):
    B, C, H, W, Cf, k = 1, 3, 6, 6, 8, 20
    logits = torch.full((B, H, W, 4), -3.0)
    acc = torch.zeros(B, C, H, W)
    acc[0, 0, 2, 2] = acc[0, 2, 4, 1] = 1.0                                       # only two positive candidates
    nmsd = torch.empty(B, C, H, W, device="cuda")
    topd = torch.empty((B, k), dtype=torch.int32, device="cuda")
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device="cuda")
    ops.hip_stage(logits.cuda(), acc.cuda(), nmsd, z(B, H, W, Cf), z(C, Cf), z(Cf), k, 3, (1, 2), 0, k, topd,
                  z(B * k, Cf), z(B * k, 2), z(B * k, C), z(B * k, dt=torch.int32))
    t = topd.cpu()[0].tolist()
    assert set(t[:2]) == {0 * 36 + 14, 2 * 36 + 25} and len(set(t)) == k
    zeros = [i for i in range(C * H * W) if i not in (14, 97)]
    assert t[2:] == zeros[:k - 2]                                                 # ties at 0 -> lowest flat index


# ------------------------------------------------------------------------------------------------ decoder pieces
def test_sine_embed(ops):
    from oracle.head import gen_sineembed_for_position
    g = torch.Generator().manual_seed(0)
    pos = torch.rand(2, 50, 2, generator=g) * 200 - 10
    ref = gen_sineembed_for_position(pos / torch.tensor([180.0, 180.0]))
    dim_t = 10000 ** (2 * (torch.arange(128, dtype=torch.float32) // 2) / 128)
    out = ops.sine_embed(pos.view(-1, 2).cuda(), 180.0, 180.0, dim_t.cuda())
    assert (out.cpu().view(2, 50, 256) - ref).abs().max().item() < 2e-5


def test_mha_core(ops):
    g = torch.Generator().manual_seed(0)
    B, Nq, h, d = 2, 77, 8, 16
    q, k, v = (torch.randn(B * Nq, h * d, generator=g) for _ in range(3))
    qk = torch.cat([q, k], 1).cuda()
    out = torch.empty(B * Nq, h * d, device="cuda")
    ops.mha_core(qk[:, :h * d], qk[:, h * d:], v.cuda(), out, B, Nq, h, d)
    sp = lambda t: t.view(B, Nq, h, d).transpose(1, 2)
    ref = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(B * Nq, h * d)
    assert (out.cpu() - ref).abs().max().item() < 1e-5


def test_msda_matches_oracle(ops):
    from oracle.transformer import ms_deform_attn_core
    g = torch.Generator().manual_seed(3)
    shapes = [(12, 10), (6, 5), (3, 2)]
    B, Nq, h, d, P = 2, 40, 8, 16, 4
    T = sum(a * b for a, b in shapes)
    value = torch.randn(B, T, 3 * h * d, generator=g)                           # 3 layers' projected values side by side
    pos = torch.rand(B, Nq, 2, generator=g) * torch.tensor([10.0, 12.0]) * 1.2 - 1.0
    offs = torch.randn(B, Nq, h, 3, P, 2, generator=g) * 2
    attw = torch.randn(B, Nq, h, 3 * P, generator=g)
    ref_pts = pos / torch.tensor([10.0, 12.0])
    norm = torch.tensor([[w, hh] for hh, w in shapes], dtype=torch.float32)
    loc = ref_pts[:, :, None, None, None, :] + offs / norm[None, None, None, :, None, :]
    col = h * d                                                                  # use the middle layer's columns
    want = ms_deform_attn_core(value[:, :, col:2 * col].reshape(B, T, h, d), shapes, loc,
                               attw.softmax(-1).view(B, Nq, h, 3, P))
    geom = ops.LevelGeom(shapes)
    oa = torch.cat([offs.reshape(B * Nq, -1), attw.reshape(B * Nq, -1)], 1).cuda()
    n_off = h * 3 * P * 2
    out = torch.empty(B * Nq, h * d, device="cuda")
    ops.msda(value.cuda(), col, geom, P, pos.view(-1, 2).cuda(), 10.0, 12.0, oa[:, :n_off], oa[:, n_off:], out, B, Nq, h, d)
    assert (out.cpu().view(B, Nq, -1) - want).abs().max().item() < 2e-5


def test_roi_sample_matches_reference_logic(ops):
    from oracle.head import FocalDecoder, TransFusionBBoxCoder, rotation_3d_in_axis_z
    g = torch.Generator().manual_seed(5)
    shapes = [(16, 16), (8, 8), (4, 4)]
    B, Nq, C, gs = 2, 9, 8, 7
    T = sum(a * b for a, b in shapes)
    ms = torch.randn(B, T, C, generator=g)
    coder = TransFusionBBoxCoder(pc_range=[-54.0, -54.0], out_size_factor=8, voxel_size=[0.075, 0.075])
    qb = torch.randn(B, 10, Nq, generator=g)
    qb[:, :2] = torch.rand(B, 2, Nq, generator=g) * 180
    qb[:, 3:6] = qb[:, 3:6] * 0.8 + 1.0
    # reference logic, focal_decoder.py:890-919
    std = coder.decode_box(qb[:, 6:8].clone(), qb[:, 3:6].clone() * 1.2, qb[:, 0:2].clone(), qb[:, 2:3].clone(), qb[:, 8:].clone())
    std = std.reshape(B * Nq, -1)
    gp = FocalDecoder.get_dense_grid_points(std, B * Nq, gs)
    gp = rotation_3d_in_axis_z(gp, std[:, 6]) + std[:, None, :2]
    gp = gp.view(B, Nq, gs * gs, 2)
    pcr = torch.tensor([-54, -54, -5.0, 54, 54, 3.0])
    gp = ((gp - pcr[:2]) / (pcr[3:5] - pcr[:2]) * 2 - 1).clip(-2, 2)
    feats, s0 = [], 0
    for hh, ww in shapes:
        feats.append(ms[:, s0:s0 + hh * ww].view(B, hh, ww, C).permute(0, 3, 1, 2))
        s0 += hh * ww
    rf = torch.cat([F.grid_sample(f, gp, mode="bilinear", align_corners=False) for f in feats], 1)   # [B, 3C, Nq, 49]
    want = rf.view(B, 3, C, Nq, gs * gs).permute(0, 3, 1, 4, 2).reshape(B * Nq, -1)                  # (lvl, pt, c)
    out = torch.empty(B * Nq, 3 * gs * gs * C, device="cuda")
    prev = qb.permute(0, 2, 1).reshape(B * Nq, 10).contiguous().cuda()
    ops.roi_sample(prev, ms.cuda(), ops.LevelGeom(shapes), C, gs, 1.2, (0.6, 0.6), (-54.0, -54.0), (-54.0, -54.0, 54.0, 54.0),
                   out, B, Nq)
    assert (out.cpu() - want).abs().max().item() < 1e-4


def test_head_update_and_box_decode(ops):
    from oracle.head import TransFusionBBoxCoder
    g = torch.Generator().manual_seed(1)
    B, Nq, C = 2, 50, 10
    pred = torch.randn(B * Nq, 20, generator=g)
    qpos = torch.rand(B * Nq, 2, generator=g) * 180
    prev = torch.randn(B * Nq, 20, generator=g)
    p, qp = pred.clone().cuda(), qpos.clone().cuda()
    ops.head_update(p, qp, prev.cuda())
    want = pred.clone()
    want[:, :2] += qpos
    want[:, 3:5] += prev[:, 3:5]
    want[:, 6:8] += prev[:, 6:8]
    assert torch.allclose(p.cpu(), want) and torch.allclose(qp.cpu(), want[:, :2])
    # decode (focal_decoder.py:1313-1321 + transfusion_bbox_coder.py:71-158)
    coder = TransFusionBBoxCoder(pc_range=[-54.0, -54.0], out_size_factor=8, voxel_size=[0.075, 0.075],
                                 post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], score_threshold=0.0, code_size=10)
    qs = torch.rand(B * Nq, C, generator=g)
    ql = torch.randint(0, C, (B * Nq,), generator=g)
    qs[3, ql[3]] = 0.0                                                            # zero score -> label 0 by argmax
    cm = lambda a, b: want[:, a:b].view(B, Nq, b - a).transpose(1, 2).clone()
    score = cm(10, 20).sigmoid() * qs.view(B, Nq, C).transpose(1, 2) * F.one_hot(ql.view(B, Nq), C).permute(0, 2, 1)
    ref = coder.decode(score, cm(6, 8), cm(3, 6), cm(0, 2), cm(2, 3), cm(8, 10), filter=False)
    boxes = torch.empty(B * Nq, 9, device="cuda")
    scores = torch.empty(B * Nq, device="cuda")
    labels = torch.empty(B * Nq, dtype=torch.int32, device="cuda")
    keep = torch.empty(B * Nq, dtype=torch.uint8, device="cuda")
    ops.box_decode(p, 10, True, qs.cuda(), ql.int().cuda(), C, (0.6, 0.6), (-54.0, -54.0),
                   [-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], boxes, scores, labels, keep)
    rb = torch.stack([r["bboxes"] for r in ref]).view(B * Nq, 9)
    assert (boxes.cpu() - rb).abs().max().item() < 1e-4
    assert (scores.cpu() - torch.stack([r["scores"] for r in ref]).view(-1)).abs().max().item() < 1e-6
    assert torch.equal(labels.cpu().long(), torch.stack([r["labels"] for r in ref]).view(-1))
    pr = torch.tensor([-61.2, -61.2, -10.0, 61.2, 61.2, 10.0])
    assert torch.equal(keep.cpu().bool(), ((rb[:, :3] >= pr[:3]) & (rb[:, :3] <= pr[3:])).all(1))


# ------------------------------------------------------------------------------------------------ tcgen05 path
@pytest.mark.parametrize("M,K,N", [(128, 32, 16), (1000, 128, 128), (300, 1024, 128), (2400, 128, 384), (129, 64, 64),
                                   (4096, 512, 256), (60000, 288, 128), (80000, 64, 32), (40000, 96, 64)])
def test_tcgemm_3xtf32_accuracy_vs_fp64(ops, M, K, N):
    """The tensor-core kernel must be fp32-grade (3xTF32), not TF32-grade: error vs fp64 ~1e-6 of the row norm,
    and indistinguishable from the SIMT fp32 kernel."""
    g = torch.Generator().manual_seed(M + K + N)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    ref = x.double() @ w.double().t() + b.double()
    pw = _pack_lin(w)
    assert pw.img is not None
    old = ops.USE_TC
    try:
        ops.USE_TC = True
        y_tc = ops.linear(x.cuda(), pw, b.cuda()).cpu()
        ops.USE_TC = False
        y_simt = ops.linear(x.cuda(), pw, b.cuda()).cpu()
    finally:
        ops.USE_TC = old
    e_tc = (y_tc.double() - ref).abs().max().item()
    e_simt = (y_simt.double() - ref).abs().max().item()
    print(f"M={M} K={K} N={N}: tcgen05 3xTF32 err {e_tc:.2e}, SIMT fp32 err {e_simt:.2e}")
    assert e_tc < 3e-5 and e_tc < 8 * e_simt, f"tcgen05 3xTF32 max abs err {e_tc} (SIMT fp32: {e_simt})"
    assert e_simt < 2e-5
    # a single TF32 pass would be ~1e-3 here: the split accumulation is what buys the parity bar
    tf32 = lambda t: ((t.contiguous().view(torch.int32) + 0x1000) & -8192).view(torch.float32)
    e_1x = ((tf32(x).double() @ tf32(w).double().t() + b.double()) - ref).abs().max().item()
    assert e_tc < 0.2 * e_1x


@pytest.mark.parametrize("kind", ["f16", "tf32"])
def test_tcgemm_long_k_error_budget(ops, kind):
    """Where the split-precision GEMM's error comes from at long K (K = 8192, all-positive products: a biased
    accumulation shows as a systematic offset).  Reports, against fp64: the kernel on the full K, the same kernel on 8
    K-chunks summed in fp32 on the host (an accumulator that is rounded 8x less often), and the SIMT fp32 kernel."""
    from focalformer3d_b200.ops import PackedW
    M, K, N = 512, 8192, 128
    g = torch.Generator().manual_seed(1)
    x = torch.rand(M, K, generator=g) + 0.5
    w = (torch.rand(N, K, generator=g) + 0.5) / K
    ref = x.double() @ w.double().t()

    def pack(wt):
        return PackedW(torch.nn.functional.pad(wt.t().unsqueeze(0), (0, 0)).contiguous(), "cuda", kind=kind)
    y_full = ops.linear(x.cuda(), pack(w)).cpu().double()
    y_chunks = torch.zeros(M, N)
    for c in range(8):
        sl = slice(c * 1024, (c + 1) * 1024)
        y_chunks += ops.linear(x[:, sl].contiguous().cuda(), pack(w[:, sl])).cpu()
    old = ops.USE_TC
    try:
        ops.USE_TC = False
        y_simt = ops.linear(x.cuda(), pack(w)).cpu().double()
    finally:
        ops.USE_TC = old
    rel = lambda y: ((y.double() - ref) / ref).abs().max().item()
    bias = lambda y: ((y.double() - ref) / ref).mean().item()
    msg = (f"long-K budget [{kind}] K={K}: full rel {rel(y_full):.2e} (mean {bias(y_full):+.2e}) | 8 chunks rel "
           f"{rel(y_chunks):.2e} (mean {bias(y_chunks):+.2e}) | SIMT fp32 rel {rel(y_simt):.2e} (mean {bias(y_simt):+.2e})")
    if kind == "f16" and ops.tma_enabled():
        # the TMA-fed kernel accumulates in chunks of 8 stages (512 K-values) summed in fp32 registers: the bias must be gone
        xs = ops.split_rows(x.cuda())
        xq = ops.unsplit_rows(xs, M).cpu().double()
        ref_q = xq @ w.double().t()
        y_tma = ops.linear(xs, pack(w)).cpu().double()
        r_tma, b_tma = ((y_tma - ref_q) / ref_q).abs().max().item(), ((y_tma - ref_q) / ref_q).mean().item()
        msg += f" | TMA kernel (chunked) rel {r_tma:.2e} (mean {b_tma:+.2e})"
        assert r_tma < 2e-5 and abs(b_tma) < 1e-5
    print(msg)
    assert rel(y_full) < 2e-4


def test_tcgemm_conv_small_cin_taps_share_a_kstep(ops):
    """cin=16 (two taps per 128-byte K-step) and cin=8 (four) against fp64 conv."""
    for cin, cout in ((16, 16), (8, 32), (16, 64)):
        g = torch.Generator().manual_seed(cin * cout)
        B, H, W = 2, 11, 9
        x = torch.randn(B, cin, H, W, generator=g)
        w = torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
        ref = F.conv2d(x.double(), w.double(), padding=1).float()
        out = torch.empty((B, H, W, cout), device="cuda")
        pw = _pack_conv(w)
        assert pw.img is not None
        ops.conv2d(x.permute(0, 2, 3, 1).contiguous().cuda(), pw, None, out, 3)
        assert (out.permute(0, 3, 1, 2).cpu() - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("shape", [(2, 13, 17, 128), (1, 24, 24, 128), (1, 5, 4, 256)])
def test_local_attention(shape):
    """ff3d_local_attention vs the oracle's restatement of locatt_ops (similar -> softmax -> weighting), incl. the
    zero-scored out-of-map neighbours that take part in the soft-max."""
    from focalformer3d_b200 import ops
    from oracle.bev import local_similar, local_weighting
    import math
    B, H, W, C = shape
    g = torch.Generator().manual_seed(H * 100 + W)
    q, k, v = (torch.randn(B, C, H, W, generator=g) * s for s in (1.3, 1.3, 1.0))
    w = torch.softmax(local_similar(q, k, 9, 9) / math.sqrt(C), -1)
    ref = local_weighting(v, w, 9, 9)
    assert w.max().item() > 0.3                      # peaked soft-max: the test is sensitive to the scores
    # operands as channel slices of a wider buffer (how the encoder feeds them)
    buf = torch.zeros(B, H, W, 3 * C + 4, device="cuda")
    buf[..., :C], buf[..., C:2 * C], buf[..., 2 * C:3 * C] = (t.permute(0, 2, 3, 1).cuda() for t in (q, k, v))
    out = torch.empty(B, H, W, 2 * C, device="cuda")
    ops.local_attention(buf[..., :C], buf[..., C:2 * C], buf[..., 2 * C:3 * C], out[..., C:], 9)
    err = (out[..., C:].permute(0, 3, 1, 2).cpu() - ref).abs().max().item()
    assert err < 2e-5, err


def test_class_select():
    """class-aware regression heads (focal_decoder.py:940-943): keep the regression set of each query's own class."""
    from focalformer3d_b200 import ops
    g = torch.Generator().manual_seed(9)
    rows, nc, gk = 37, 3, [2, 1, 3, 2]
    full = torch.randn(rows, 28, generator=g)                      # 3 * 8 = 24 regression columns + 3 class logits (+ pad)
    lab = torch.randint(0, nc, (rows,), generator=g).int()
    out = ops.class_select(full.cuda(), lab.cuda(), gk, nc, nc, 12).cpu()
    c_full, c_out = 0, 0
    for k in gk:
        blk = full[:, c_full:c_full + nc * k].view(rows, nc, k)
        ref = blk[torch.arange(rows), lab.long()]
        assert torch.equal(out[:, c_out:c_out + k], ref)
        c_full += nc * k
        c_out += k
    assert torch.equal(out[:, c_out:c_out + nc], full[:, c_full:c_full + nc])


# ------------------------------------------------------------------------------------------------ TMA-fed GEMM (split activations)
def _split_ref(x):
    """host restatement of the split format: (hi, lo) halves with x ~= hi + lo / 2048"""
    hi = x.clamp(-65504, 65504).half()
    lo = ((x - hi.float()) * 2048.0).clamp(-65504, 65504).half()
    return hi, lo


def test_split_rows_round_trip(ops):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1000, 192, generator=g) * torch.logspace(-6, 3, 192)[None]
    xs = ops.split_rows(x.cuda(), zero_row=True)
    assert xs.zero_row == 1000
    hi, lo = _split_ref(x)
    assert torch.equal(xs.t[:1000, :192].cpu(), hi) and torch.equal(xs.t[:1000, 192:].cpu(), lo)
    back = ops.unsplit_rows(xs, 1000).cpu()
    assert ((back - x).abs() <= x.abs() * 2.0 ** -21 + 2.0 ** -36).all()
    assert int(ops.gemm_flag().item()) == 0
    ops.split_rows(torch.full((8, 64), 1.0e6, device="cuda"))            # beyond the fp16 range: the flag must trip
    assert int(ops.gemm_flag().item()) == 1
    ops.gemm_flag().zero_()


@pytest.mark.parametrize("M,K,N", [(300, 128, 128), (4096, 1024, 256), (129, 64, 64), (70000, 256, 128), (2400, 18816, 512)])
def test_tma_linear_vs_fp64(ops, M, K, N):
    """ff3d_tmagemm ROWS mode (2-D TMA tile loads of the split A rows): fp32-grade accuracy, fp32 + split outputs,
    fp32 and split residuals."""
    if not ops.tma_enabled():
        pytest.skip("TMA path disabled (FF3D_GEMM=tf32 / FF3D_TMA=0)")
    g = torch.Generator().manual_seed(M + K + N)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g)
    pw = _pack_lin(w)
    assert ops.tma_ok(pw, K, N)
    xs = ops.split_rows(x.cuda())
    xq = ops.unsplit_rows(xs, M).cpu().double()                          # the 22-bit values the kernel really multiplies
    ref = xq @ w.double().t() + b.double()
    ys = ops.Split.empty((M,), N, "cuda")
    y = ops.linear(xs, pw, b.cuda(), act=1, res=res.cuda(), out_s=ys).cpu()
    want = torch.relu(ref + res.double())
    tol = 6e-5 * max(1.0, (K / 1024) ** 0.5)
    assert (y.double() - want).abs().max().item() < tol
    assert (ops.unsplit_rows(ys, M).cpu() - y).abs().max().item() <= (y.abs() * 2.0 ** -21 + 2.0 ** -30).max().item()
    # split residual, split-only output, activation after the residual add disabled
    rs = ops.split_rows(res.cuda())
    rq = ops.unsplit_rows(rs, M).cpu().double()
    ys2 = ops.Split.empty((M,), N, "cuda")
    assert ops.linear(xs, pw, b.cuda(), act=1, res=rs, res_after_act=True, out_s=ys2, want_out=False) is None
    want2 = torch.relu(ref) + rq
    assert (ops.unsplit_rows(ys2, M).cpu().double() - want2).abs().max().item() < tol + 1e-5
    assert int(ops.gemm_flag().item()) == 0


@pytest.mark.parametrize("cin,cout,k,H,W", [(64, 128, 3, 13, 11), (128, 128, 3, 32, 32), (256, 64, 1, 20, 17), (128, 256, 3, 45, 45),
                                            (512, 128, 3, 24, 24), (64, 64, 7, 9, 30)])
def test_tma_conv2d_vs_fp64(ops, cin, cout, k, H, W):
    """ff3d_tmagemm CONV2D mode: 4-D TMA box loads, zero padding = out-of-bounds fill, patches that overhang the map."""
    if not ops.tma_enabled():
        pytest.skip("TMA path disabled")
    g = torch.Generator().manual_seed(cin + cout + k + H)
    B = 2
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / (k * k * cin) ** 0.5
    b = torch.randn(cout, generator=g)
    res = torch.randn(B, cout, H, W, generator=g)
    pw = _pack_conv(w)
    assert ops.tma_ok(pw, cin, cout)
    xs = ops.split_rows(x.permute(0, 2, 3, 1).contiguous().cuda())
    xq = ops.unsplit_rows(xs, B * H * W).cpu().view(B, H, W, cin).permute(0, 3, 1, 2).double()
    ref = F.conv2d(xq, w.double(), b.double(), padding=k // 2) + res.double()
    out = torch.empty((B, H, W, cout), device="cuda")
    ys = ops.Split.empty((B, H, W), cout, "cuda")
    ops.conv2d(xs, pw, b.cuda(), out, k, act=2, res=res.permute(0, 2, 3, 1).contiguous().cuda(), out_s=ys)
    want = ref.clamp(0, 6)
    got = out.permute(0, 3, 1, 2).cpu().double()
    assert (got - want).abs().max().item() < 3e-5
    back = ops.unsplit_rows(ys, B * H * W).cpu().view(B, H, W, cout).permute(0, 3, 1, 2)
    assert (back.double() - got).abs().max().item() < 1e-5
    assert int(ops.gemm_flag().item()) == 0


@pytest.mark.parametrize("cin,cout,H,W", [(128, 256, 30, 37), (64, 128, 17, 16)])
def test_tma_conv2d_stride2_vs_fp64(ops, cin, cout, H, W):
    """ff3d_tmagemm CONV2D with stride 2: tensor map with a traversal stride of 2 on W / H (every second pixel of the box),
    odd and even map sizes, boxes that start at -1 and overhang the map."""
    if not ops.tma_enabled():
        pytest.skip("TMA path disabled")
    g = torch.Generator().manual_seed(cin + H)
    B = 2
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
    b = torch.randn(cout, generator=g)
    pw = _pack_conv(w)
    xs = ops.split_rows(x.permute(0, 2, 3, 1).contiguous().cuda())
    xq = ops.unsplit_rows(xs, B * H * W).cpu().view(B, H, W, cin).permute(0, 3, 1, 2).double()
    ref = F.conv2d(xq, w.double(), b.double(), stride=2, padding=1).clamp(min=0)
    Ho, Wo = ref.shape[2], ref.shape[3]
    out = torch.empty((B, Ho, Wo, cout), device="cuda")
    ys = ops.Split.empty((B, Ho, Wo), cout, "cuda")
    ops.conv2d(xs, pw, b.cuda(), out, 3, stride=2, act=1, out_s=ys)
    got = out.permute(0, 3, 1, 2).cpu().double()
    assert (got - ref).abs().max().item() < 3e-5
    back = ops.unsplit_rows(ys, B * Ho * Wo).cpu().view(B, Ho, Wo, cout).permute(0, 3, 1, 2)
    assert (back.double() - got).abs().max().item() < 1e-5
    assert int(ops.gemm_flag().item()) == 0


def test_tma_transposed_conv_lattice(ops):
    """ConvTranspose2d with kernel == stride as one 1x1 GEMM per output lattice position on the TMA-fed kernel (split input,
    fp32 + split outputs on the (oy*u + dy, ox*u + dx) lattice)."""
    if not ops.tma_enabled():
        pytest.skip("TMA path disabled")
    g = torch.Generator().manual_seed(5)
    B, H, W, cin, cout, u = 2, 13, 11, 128, 128, 2
    x = torch.randn(B, cin, H, W, generator=g)
    wt = torch.randn(cin, cout, u, u, generator=g) / cin ** 0.5                    # ConvTranspose2d layout
    xs = ops.split_rows(x.permute(0, 2, 3, 1).contiguous().cuda())
    xq = ops.unsplit_rows(xs, B * H * W).cpu().view(B, H, W, cin).permute(0, 3, 1, 2).double()
    ref = F.conv_transpose2d(xq, wt.double(), stride=u)
    out = torch.zeros((B, H * u, W * u, cout), device="cuda")
    ys = ops.Split.empty((B, H * u, W * u), cout, "cuda")
    for dy in range(u):
        for dx in range(u):
            pw = _pack_conv(wt[:, :, dy, dx].t().reshape(cout, cin, 1, 1).contiguous())
            ops.conv2d(xs, pw, None, out, 1, up=(u, dy, dx), out_s=ys)
    got = out.permute(0, 3, 1, 2).cpu().double()
    assert (got - ref).abs().max().item() < 3e-5
    back = ops.unsplit_rows(ys, B * H * u * W * u).cpu().view(B, H * u, W * u, cout).permute(0, 3, 1, 2)
    assert (back.double() - got).abs().max().item() < 1e-5


def test_tma_conv2d_channel_slices(ops):
    """concat-free views: the conv reads channels [64, 192) of a 256-channel split buffer and writes its fp32 and split
    outputs into channel slices of wider buffers."""
    if not ops.tma_enabled():
        pytest.skip("TMA path disabled")
    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 18, 21
    x = torch.randn(B, H, W, 256, generator=g)
    w = torch.randn(128, 128, 3, 3, generator=g) / (9 * 128) ** 0.5
    pw = _pack_conv(w)
    xs = ops.split_rows(x.cuda()).slice(64, 192)
    xq = ops.unsplit_rows(xs, B * H * W).cpu().view(B, H, W, 128).permute(0, 3, 1, 2).double()
    ref = F.conv2d(xq, w.double(), padding=1)
    out = torch.zeros((B, H, W, 384), device="cuda")
    out_s = ops.Split.empty((B, H, W), 384, "cuda")
    out_s.t.zero_()
    ops.conv2d(xs, pw, None, out[..., 128:256], 3, out_s=out_s.slice(256, 384))
    assert (out[..., 128:256].permute(0, 3, 1, 2).cpu().double() - ref).abs().max().item() < 3e-5
    assert bool((out[..., :128] == 0).all()) and bool((out[..., 256:] == 0).all())
    back = ops.unsplit_rows(out_s.slice(256, 384), B * H * W).cpu().view(B, H, W, 128).permute(0, 3, 1, 2)
    assert (back.double() - ref).abs().max().item() < 1e-4
    assert bool((out_s.t[..., :256] == 0).all()) and bool((out_s.t[..., 384:640] == 0).all())


@pytest.mark.parametrize("Ci,Co", [(64, 64), (64, 128), (128, 128), (8, 16), (16, 16), (16, 32), (32, 32), (32, 64), (64, 16)])
@pytest.mark.parametrize("subm", [True, False])
def test_tma_sparse_conv_matches_oracle(ops, Ci, Co, subm):
    """ff3d_tmagemm SPARSE mode: tile::gather4 of the rulebook rows (absent neighbours -> the all-zero row), per-tile tap
    skipping, row-mapped / split outputs and split residuals -- mask-sorted clustered level vs the oracle."""
    if not ops.tma_enabled():
        pytest.skip("TMA path disabled")
    from oracle import sparse as osp
    from focalformer3d_b200.model import pack_taps
    g = torch.Generator().manual_seed(Ci + Co)
    B, shape = 2, (9, 60, 60)
    cells = []
    for b in range(B):
        yx = torch.randint(0, 60, (2500, 2), generator=g)
        z = torch.randint(2, 5, (2500, 1), generator=g)
        cells.append(torch.cat([torch.full((2500, 1), b), z, yx], 1))
    idx = torch.unique(torch.cat(cells), dim=0).int()
    idx = idx[torch.randperm(idx.shape[0], generator=g)]
    n = idx.shape[0]
    cap = n + 37
    coors = torch.zeros((cap, 4), dtype=torch.int32)
    coors[:n] = idx
    feat = torch.randn(n, Ci, generator=g)
    x = torch.zeros((cap, Ci))
    x[:n] = feat
    n_dev = torch.tensor([n], dtype=torch.int32).cuda()
    lvl = ops.SparseLevel(coors.cuda(), n_dev, cap, B, shape)
    lvl.build_hash()
    perm = lvl.sort_by_mask()
    xg = ops.gather_rows(x.cuda(), perm, n_dev, Ci)
    xs = ops.split_rows(xg, n_dev=n_dev, zero_row=True)
    xq = ops.unsplit_rows(xs, cap, n_dev=n_dev).cpu()[:n]
    sorted_idx = lvl.coors[:n].cpu()
    k, s, p = ((3, 3, 3), (1, 1, 1), (1, 1, 1)) if subm else ((3, 3, 3), (2, 2, 2), (1, 1, 1))
    wt = torch.randn(*k, Ci, Co, generator=g) / (27 * Ci) ** 0.5
    bias = torch.randn(Co, generator=g)
    oconv = osp.SpConv3d(Ci, Co, k, s[0], p[0], subm=subm)
    oconv.weight.data.copy_(wt)
    with torch.no_grad():
        ref = oconv(osp.SparseTensor(xq, sorted_idx, shape, B))
    wpk = pack_taps(wt.reshape(-1, Ci, Co), "cuda")
    assert ops.tma_ok(wpk, Ci, Co, sparse=True)
    if subm:
        rb = lvl.subm_map()
        ys = ops.Split.empty((cap,), Co, "cuda", zero_row=True)
        if Ci == Co:                                                    # residual block: res = the split input itself
            ops.sparse_conv(xs, rb, n_dev, wpk, bias.cuda(), None, act=1, res=xs, out_s=ys)
            want = torch.relu(ref.features + bias + xq)
        else:
            ops.sparse_conv(xs, rb, n_dev, wpk, bias.cuda(), None, act=1, out_s=ys)
            want = torch.relu(ref.features + bias)
        got = ops.unsplit_rows(ys, cap, n_dev=n_dev).cpu()[:n]
        assert (got - want).abs().max().item() < 1e-4                   # oracle rows are in (sorted) input order for SubM
    else:
        overflow = torch.zeros(1, dtype=torch.int32, device="cuda")
        nl, rb = lvl.downsample(k, s, p, 4 * cap, overflow, ldy=Co)
        no = int(nl.n_dev.item())
        assert no == ref.indices.shape[0] and rb.y_row is not None
        out = torch.zeros((nl.cap, Co), device="cuda")
        ys = ops.Split.empty((nl.cap,), Co, "cuda", zero_row=True)
        ops.sparse_conv(xs, rb, nl.n_dev, wpk, bias.cuda(), out, act=0, out_s=ys)
        got = _to_dense(out.cpu(), nl.coors.cpu(), no, nl.shape, B)
        want = _to_dense(ref.features + bias, ref.indices, no, ref.spatial_shape, B)
        assert (got - want).abs().max().item() < 1e-4
        back = ops.unsplit_rows(ys, nl.cap, n_dev=nl.n_dev).cpu()
        assert (back[:no] - out.cpu()[:no]).abs().max().item() < 1e-5
    assert int(ops.gemm_flag().item()) == 0


# ------------------------------------------------------------------------------------------------ output side (NMS, TTA merge)
def _rand_boxes(n, seed, spread=12.0):
    g = torch.Generator().manual_seed(seed)
    centres = torch.rand(n // 4 + 1, 2, generator=g) * spread - spread / 2
    xy = centres[torch.randint(0, centres.shape[0], (n,), generator=g)] + torch.randn(n, 2, generator=g) * 0.6
    z = torch.randn(n, 1, generator=g) * 0.2
    dims = torch.rand(n, 3, generator=g) * torch.tensor([3.0, 1.5, 1.0]) + torch.tensor([0.8, 0.5, 0.8])
    yaw = (torch.rand(n, 1, generator=g) - 0.5) * 6.2
    vel = torch.randn(n, 2, generator=g)
    return torch.cat([xy, z, dims, yaw, vel], 1)


def test_boxes_iou_bev_matches_oracle(ops):
    from oracle.head import boxes_iou_bev, xywhr2xyxyr
    a, b = _rand_boxes(40, 1), _rand_boxes(50, 2)
    b[:5] = a[:5]                                            # identical boxes: IoU 1
    b[5, :2] = a[5, :2]; b[5, 3:6] = a[5, 3:6] * 0.5; b[5, 6] = a[5, 6]      # contained, same yaw: IoU 0.25
    got = ops.boxes_iou_bev(a.cuda(), b.cuda()).cpu()
    want = boxes_iou_bev(xywhr2xyxyr(a[:, [0, 1, 3, 4, 6]]), xywhr2xyxyr(b[:, [0, 1, 3, 4, 6]]))
    assert (got - want).abs().max().item() < 1e-4
    assert (got[:5].diagonal() - 1).abs().max().item() < 1e-5 and abs(got[5, 5].item() - 0.25) < 1e-5


@pytest.mark.parametrize("nms_type,dataset", [("circle", "nuScenes"), ("rotate", "nuScenes"), ("circle", "Waymo"), ("rotate", "Waymo")])
def test_nms_tasks_match_oracle(ops, nms_type, dataset):
    """ff3d_nms_tasks vs the reference's per-task loop (focal_decoder.py:1352-1385) on [upstream] circle_nms / nms_gpu."""
    from oracle.head import circle_nms, nms_rotated_bev, xywhr2xyxyr
    B, nq = 2, 300
    nc = 10 if dataset == "nuScenes" else 3
    tasks = [(list(range(8)), -1.0), ([8], 0.175), ([9], 0.175)] if dataset == "nuScenes" else [([0], 0.7), ([1], 0.7), ([2], 0.7)]
    g = torch.Generator().manual_seed(7)
    boxes = torch.stack([_rand_boxes(nq, 10 + b, spread=6.0) for b in range(B)])
    scores = torch.rand(B, nq, generator=g)
    labels = torch.randint(0, nc, (B, nq), generator=g)
    if dataset == "nuScenes":
        labels[:, :200] = torch.randint(8, 10, (B, 200), generator=g)       # plenty of pedestrians / cones close together
    keep_in = torch.rand(B, nq, generator=g) > 0.1
    pre, post = 150, 40
    got = ops.nms_tasks(boxes.cuda(), scores.cuda(), labels.int().cuda(), keep_in.to(torch.uint8).cuda(), tasks, nms_type, pre, post).cpu().bool()
    for b in range(B):
        want = torch.zeros(nq, dtype=torch.bool)
        for idx, radius in tasks:
            tm = keep_in[b] & torch.isin(labels[b], torch.tensor(idx))
            rows = tm.nonzero().flatten()
            if radius <= 0:
                want[rows] = True
            elif nms_type == "circle":
                dets = torch.cat([boxes[b][rows][:, :2], scores[b][rows][:, None]], 1).numpy()
                want[rows[torch.tensor(circle_nms(dets, radius), dtype=torch.long)]] = True
            else:
                k = nms_rotated_bev(xywhr2xyxyr(boxes[b][rows][:, [0, 1, 3, 4, 6]]), scores[b][rows], radius, pre, post)
                want[rows[k]] = True
        assert torch.equal(got[b], want), f"{nms_type}/{dataset} scene {b}: {(got[b] != want).sum().item()} keep flags differ"
    assert bool((got <= keep_in).all())


def test_merge_aug_bboxes_matches_oracle(ops):
    """TTA merge (merge_augs.py:14-184): map back flips / scale, per-class rotated NMS, box voting."""
    from oracle.head import merge_aug_bboxes_3d
    base = _rand_boxes(60, 3, spread=20.0)
    g = torch.Generator().manual_seed(5)
    augs, metas = [], []
    for hf, vf, sf in ((False, False, 1.0), (True, False, 1.0), (False, True, 0.95), (True, True, 1.05)):
        b = base.clone() + torch.randn(60, 9, generator=g) * 0.03
        b[:, :6] *= sf; b[:, 7:] *= sf                       # forward augmentation: scale, then flips
        if vf:
            b[:, 0] = -b[:, 0]; b[:, 7] = -b[:, 7]; b[:, 6] = -b[:, 6]
        if hf:
            b[:, 1] = -b[:, 1]; b[:, 8] = -b[:, 8]; b[:, 6] = -b[:, 6] + torch.pi
        augs.append(dict(boxes_3d=b, scores_3d=torch.rand(60, generator=g), labels_3d=torch.arange(60) % 3))
        metas.append(dict(pcd_scale_factor=sf, pcd_horizontal_flip=hf, pcd_vertical_flip=vf))
    want = merge_aug_bboxes_3d(augs, metas)
    got = ops.merge_aug_bboxes_3d(augs, metas)
    assert got["boxes_3d"].shape == want["boxes_3d"].shape
    assert torch.equal(got["labels_3d"].cpu().long(), want["labels_3d"])
    assert (got["scores_3d"].cpu() - want["scores_3d"]).abs().max().item() < 1e-6
    d = (got["boxes_3d"].cpu() - want["boxes_3d"]).abs()
    d[:, 6] = torch.minimum(d[:, 6], (2 * torch.pi - d[:, 6]).abs())
    assert d.max().item() < 1e-3
