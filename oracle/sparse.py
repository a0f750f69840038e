"""Oracle: sparse 3-D convolution + SparseEncoder.  TEST INFRASTRUCTURE (see oracle/__init__.py).

[upstream] mmdet3d v0.17.1 ``ops/spconv`` (spconv-v1 fork) and ``models/middle_encoders/sparse_encoder.py``
as configured at ``projects/configs/focalformer3d/FocalFormer3D_L.py:198-206`` and called at
``projects/mmdet3d_plugin/models/detectors/focalformer3d.py:168``.

Algorithm class is the reference's: build a rulebook (per kernel offset a list of (in_row, out_row)
pairs), then per offset gather -> mm -> scatter-add.
"""
import torch
from torch import nn


def _to3(v):
    return tuple(v) if isinstance(v, (list, tuple)) else (v, v, v)


class SparseTensor:
    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features = features            # [N, C] fp32
        self.indices = indices.long()       # [N, 4] (b, z, y, x)
        self.spatial_shape = tuple(int(s) for s in spatial_shape)  # (D, H, W)
        self.batch_size = int(batch_size)

    def dense(self):
        """SparseConvTensor.dense(): [B, C, D, H, W], inactive sites exactly 0."""
        D, H, W = self.spatial_shape
        out = self.features.new_zeros(self.batch_size, D, H, W, self.features.shape[1])
        i = self.indices
        out[i[:, 0], i[:, 1], i[:, 2], i[:, 3]] = self.features
        return out.permute(0, 4, 1, 2, 3).contiguous()


def _lin(idx, shape):
    D, H, W = shape
    return ((idx[:, 0] * D + idx[:, 1]) * H + idx[:, 2]) * W + idx[:, 3]


def _lookup(sorted_keys, order, q):
    pos = torch.searchsorted(sorted_keys, q).clamp(max=max(sorted_keys.numel() - 1, 0))
    hit = sorted_keys[pos] == q if sorted_keys.numel() else torch.zeros_like(q, dtype=torch.bool)
    return order[pos], hit


def out_shape(shape, k, s, p):
    return tuple((shape[i] + 2 * p[i] - k[i]) // s[i] + 1 for i in range(3))


def build_rulebook(indices, spatial_shape, ksize, stride, padding, subm):
    """Returns (out_indices [No,4], out_spatial_shape, pairs: list over K offsets of (in_rows, out_rows)).

    spconv-v1 getValidOutPos: an input at i touches output o with tap k iff o*s = i + p - k
    (cross-correlation, same orientation as torch conv3d).  SubM: outputs == inputs, padding k//2.
    """
    k, s, p = _to3(ksize), _to3(stride), _to3(padding)
    idx = indices.long()
    dev = idx.device
    T = lambda v: torch.tensor(v, device=dev)
    if subm:
        oshape = tuple(spatial_shape)
        out_idx = idx
    else:
        oshape = out_shape(spatial_shape, k, s, p)
        cands = []
        for kz in range(k[0]):
            for ky in range(k[1]):
                for kx in range(k[2]):
                    v = idx[:, 1:] + T([p[0] - kz, p[1] - ky, p[2] - kx])
                    st = T(s)
                    ok = (v % st == 0).all(1)
                    o = torch.div(v, st, rounding_mode="floor")
                    ok &= (o >= 0).all(1) & (o < T(oshape)).all(1)
                    cands.append(torch.cat([idx[ok, :1], o[ok]], 1))
        allc = torch.cat(cands, 0)
        keys = torch.unique(_lin(allc, oshape))          # sorted; the order of output rows is
        W, H, D = oshape[2], oshape[1], oshape[0]        # implementation-defined in spconv
        out_idx = torch.stack([keys // (D * H * W), (keys // (H * W)) % D, (keys // W) % H, keys % W], 1)
    in_keys = _lin(idx, spatial_shape)
    sk, order = torch.sort(in_keys)
    pairs = []
    No = out_idx.shape[0]
    orow = torch.arange(No, device=dev)
    for kz in range(k[0]):
        for ky in range(k[1]):
            for kx in range(k[2]):
                ic = out_idx[:, 1:] * T(s) - T(p) + T([kz, ky, kx])
                ok = (ic >= 0).all(1) & (ic < T(spatial_shape)).all(1)
                q = _lin(torch.cat([out_idx[:, :1], ic], 1), spatial_shape)
                rows, hit = _lookup(sk, order, q.clamp(min=0))
                ok &= hit
                pairs.append((rows[ok], orow[ok]))
    return out_idx, oshape, pairs


def indice_conv(features, weight, pairs, n_out):
    """spconv-v1 indice_conv: for each offset, gather rows, torch.mm with W[k], scatter-add."""
    kvol = weight.shape[0] * weight.shape[1] * weight.shape[2]
    w = weight.reshape(kvol, weight.shape[3], weight.shape[4])
    out = features.new_zeros(n_out, w.shape[2])
    for kk, (ir, orr) in enumerate(pairs):
        if ir.numel():
            out.index_add_(0, orr, features[ir] @ w[kk])
    return out


class SpConv3d(nn.Module):
    """SubMConv3d / SparseConv3d, bias=False.  Weight layout [kD, kH, kW, Cin, Cout] (spconv v1)."""

    def __init__(self, cin, cout, ksize, stride=1, padding=0, subm=False):
        super().__init__()
        self.k, self.s, self.p, self.subm = _to3(ksize), _to3(stride), _to3(padding), subm
        self.weight = nn.Parameter(torch.zeros(*self.k, cin, cout))

    def forward(self, x):
        out_idx, oshape, pairs = build_rulebook(x.indices, x.spatial_shape, self.k, self.s, self.p, self.subm)
        f = indice_conv(x.features, self.weight, pairs, out_idx.shape[0])
        return SparseTensor(f, out_idx, oshape, x.batch_size)


class SparseSeq(nn.Sequential):
    """SparseSequential: dense modules act on .features."""

    def forward(self, x):
        for m in self:
            if isinstance(m, (SpConv3d, SparseBasicBlock, SparseSeq)):
                x = m(x)
            else:
                x = SparseTensor(m(x.features), x.indices, x.spatial_shape, x.batch_size)
        return x


def _bn1d(c):
    return nn.BatchNorm1d(c, eps=1e-3, momentum=0.01)


class SparseBasicBlock(nn.Module):
    """[upstream] mmdet3d SparseBasicBlock: SubM-BN-ReLU-SubM-BN-(+identity)-ReLU; no indice_key."""

    def __init__(self, c):
        super().__init__()
        self.conv1 = SpConv3d(c, c, 3, 1, 1, subm=True)
        self.bn1 = _bn1d(c)
        self.conv2 = SpConv3d(c, c, 3, 1, 1, subm=True)
        self.bn2 = _bn1d(c)

    def forward(self, x):
        identity = x.features
        out = self.conv1(x)
        out.features = torch.relu(self.bn1(out.features))
        out = self.conv2(out)
        out.features = self.bn2(out.features)
        out.features = torch.relu(out.features + identity)
        return out


class SparseEncoder(nn.Module):
    """[upstream] mmdet3d v0.17.1 SparseEncoder, block_type='basicblock', order (conv, norm, act)."""

    def __init__(self, in_channels, sparse_shape, output_channels=128, base_channels=16,
                 encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
                 encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, (0, 1, 1)), (0, 0)),
                 block_type="basicblock", order=("conv", "norm", "act"), **kw):
        super().__init__()
        assert block_type == "basicblock" and tuple(order) == ("conv", "norm", "act")
        self.sparse_shape = tuple(sparse_shape)
        self.conv_input = SparseSeq(SpConv3d(in_channels, base_channels, 3, 1, 1, subm=True),
                                    _bn1d(base_channels), nn.ReLU())
        self.encoder_layers = SparseSeq()
        cin = base_channels
        for i, blocks in enumerate(encoder_channels):
            mods = []
            for j, cout in enumerate(blocks):
                pad = encoder_paddings[i][j]
                if j == len(blocks) - 1 and i != len(encoder_channels) - 1:
                    mods.append(SparseSeq(SpConv3d(cin, cout, 3, 2, pad, subm=False), _bn1d(cout), nn.ReLU()))
                else:
                    mods.append(SparseBasicBlock(cout))
                cin = cout
            self.encoder_layers.add_module(f"encoder_layer{i + 1}", SparseSeq(*mods))
        self.conv_out = SparseSeq(SpConv3d(cin, output_channels, (3, 1, 1), (2, 1, 1), 0, subm=False),
                                  _bn1d(output_channels), nn.ReLU())

    def forward(self, voxel_features, coors, batch_size):
        x = SparseTensor(voxel_features, coors.int(), self.sparse_shape, batch_size)
        x = self.conv_input(x)
        x = self.encoder_layers(x)
        out = self.conv_out(x)
        d = out.dense()
        N, C, D, H, W = d.shape
        return d.view(N, C * D, H, W)


def dense_reference_conv(x: SparseTensor, weight, ksize, stride, padding, subm):
    """Independent formulation (self-check): dense F.conv3d, then keep only the active output sites."""
    import torch.nn.functional as F
    k, s, p = _to3(ksize), _to3(stride), _to3(padding)
    d = x.dense()
    w = weight.permute(4, 3, 0, 1, 2).contiguous()
    y = F.conv3d(d, w, stride=s, padding=p)
    if subm:
        act = x.indices
    else:
        occ = (x.dense().abs().sum(1, keepdim=True) * 0)
        occ[x.indices[:, 0], 0, x.indices[:, 1], x.indices[:, 2], x.indices[:, 3]] = 1.0
        reach = F.conv3d(occ, torch.ones(1, 1, *k), stride=s, padding=p)
        act = (reach[:, 0] > 0).nonzero()
    f = y[act[:, 0], :, act[:, 1], act[:, 2], act[:, 3]]
    return SparseTensor(f, act, y.shape[2:], x.batch_size)
