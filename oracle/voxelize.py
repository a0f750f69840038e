"""Oracle: hard voxelisation + voxel feature encoders.  TEST INFRASTRUCTURE (see oracle/__init__.py).

[upstream] mmdet3d v0.17.1 ``ops/voxel`` (hard_voxelize, deterministic=True) as called from
``projects/mmdet3d_plugin/models/detectors/focalformer3d.py:189-209`` (per-sample loop, batch
index padded in front of the (z, y, x) coordinate) and ``HardSimpleVFE`` / ``HardVFE`` called
at ``focalformer3d.py:166``.
"""
import numpy as np
import torch


def grid_size_of(voxel_size, pc_range):
    """mmdet3d Voxelization.__init__: grid = round((range[3:] - range[:3]) / voxel_size) (x, y, z)."""
    r = np.asarray(pc_range, dtype=np.float32)
    v = np.asarray(voxel_size, dtype=np.float32)
    return np.round((r[3:] - r[:3]) / v).astype(np.int64)


def hard_voxelize(points, voxel_size, pc_range, max_points, max_voxels):
    """One sample.  points [N, F] float32 (x, y, z, ...).

    Returns (voxels [M, max_points, F] f32, coors [M, 3] i32 as (z, y, x), num_points [M] i32).

    Semantics (mmdet3d v0.17.1 voxelization): c = floor((p - range_min) / voxel) in
    fp32; points outside the grid are dropped; voxel ids are assigned in order of first appearance
    in the input; a new voxel is only created while #voxels < max_voxels; each voxel keeps the
    first ``max_points`` points in input order; unused slots are zero.
    """
    pts = np.ascontiguousarray(points, dtype=np.float32)
    N, F = pts.shape
    r = np.asarray(pc_range, dtype=np.float32)
    v = np.asarray(voxel_size, dtype=np.float32)
    g = grid_size_of(voxel_size, pc_range)
    c = np.floor((pts[:, :3] - r[None, :3]) / v[None, :]).astype(np.int64)  # (cx, cy, cz)
    valid = np.all((c >= 0) & (c < g[None, :]), axis=1)
    vidx = np.nonzero(valid)[0]
    cv = c[vidx]
    key = (cv[:, 2] * g[1] + cv[:, 1]) * g[0] + cv[:, 0]
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # unique-id -> first-seen rank
    rank_of_uniq = np.empty_like(order)
    rank_of_uniq[order] = np.arange(order.size)
    vid = rank_of_uniq[inv]                           # voxel id per valid point
    keep = vid < max_voxels
    vidx, vid, cv = vidx[keep], vid[keep], cv[keep]
    M = int(min(uniq.size, max_voxels))
    # rank of each point inside its voxel, in input order
    srt = np.argsort(vid, kind="stable")
    vs = vid[srt]
    start = np.r_[0, np.nonzero(np.diff(vs))[0] + 1] if vs.size else np.zeros(0, np.int64)
    seg_start = np.repeat(start, np.diff(np.r_[start, vs.size])) if vs.size else start
    within = np.arange(vs.size) - seg_start
    sel = within < max_points
    voxels = np.zeros((M, max_points, F), np.float32)
    voxels[vs[sel], within[sel]] = pts[vidx[srt][sel]]
    num_points = np.bincount(vs[sel], minlength=M).astype(np.int32)
    coors = np.zeros((M, 3), np.int32)
    firstpt = srt[start] if vs.size else start
    coors[vs[start]] = cv[firstpt][:, ::-1]           # stored (z, y, x)
    return voxels, coors, num_points


def hard_voxelize_loop(points, voxel_size, pc_range, max_points, max_voxels):
    """Independent formulation of hard_voxelize: the literal point-by-point loop (small inputs only)."""
    pts = np.asarray(points, dtype=np.float32)
    r = np.asarray(pc_range, dtype=np.float32)
    v = np.asarray(voxel_size, dtype=np.float32)
    g = grid_size_of(voxel_size, pc_range)
    table, voxels, coors, nump = {}, [], [], []
    for p in pts:
        c = np.floor((p[:3] - r[:3]) / v).astype(np.int64)
        if np.any(c < 0) or np.any(c >= g):
            continue
        k = (int(c[2]), int(c[1]), int(c[0]))
        if k not in table:
            if len(voxels) >= max_voxels:
                continue
            table[k] = len(voxels)
            voxels.append(np.zeros((max_points, pts.shape[1]), np.float32))
            coors.append(k)
            nump.append(0)
        i = table[k]
        if nump[i] < max_points:
            voxels[i][nump[i]] = p
            nump[i] += 1
    M = len(voxels)
    return (np.stack(voxels) if M else np.zeros((0, max_points, pts.shape[1]), np.float32),
            np.asarray(coors, np.int32).reshape(M, 3), np.asarray(nump, np.int32))


def voxelize_batch(points_list, voxel_size, pc_range, max_points, max_voxels):
    """focalformer3d.py:189-209: per-sample hard voxelisation, concatenate, pad batch index in front."""
    vox, coo, num = [], [], []
    for b, pts in enumerate(points_list):
        p = pts.detach().cpu().numpy() if torch.is_tensor(pts) else pts
        v, c, n = hard_voxelize(p, voxel_size, pc_range, max_points, max_voxels)
        vox.append(v)
        num.append(n)
        coo.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c], axis=1))
    return (torch.from_numpy(np.concatenate(vox)), torch.from_numpy(np.concatenate(num)),
            torch.from_numpy(np.concatenate(coo)))


def dynamic_voxelize_mean(points_list, voxel_size, pc_range):
    """[upstream] mmdet3d v0.17.1 dynamic voxelisation (Voxelization(max_num_points=-1): one (z, y, x) per point, -1 when
    outside the grid) + DynamicSimpleVFE (DynamicScatter, mean of ALL points of a voxel) as used at
    focalformer3d.py:159-163,213-238.  Returns (features [M, F] float32, coors [M, 4] (b, z, y, x)); voxels are listed in
    key order -- the order is implementation-defined in the reference and irrelevant after SparseConvTensor.dense()."""
    feats, coors = [], []
    r = np.asarray(pc_range, dtype=np.float32)
    v = np.asarray(voxel_size, dtype=np.float32)
    g = grid_size_of(voxel_size, pc_range)
    for b, pts in enumerate(points_list):
        p = np.ascontiguousarray(pts.detach().cpu().numpy() if torch.is_tensor(pts) else pts, dtype=np.float32)
        c = np.floor((p[:, :3] - r[None, :3]) / v[None, :]).astype(np.int64)
        ok = np.all((c >= 0) & (c < g[None, :]), axis=1)
        p, c = p[ok], c[ok]
        key = (c[:, 2] * g[1] + c[:, 1]) * g[0] + c[:, 0]
        uniq, inv = np.unique(key, return_inverse=True)
        sums = np.zeros((uniq.size, p.shape[1]), np.float64)
        np.add.at(sums, inv, p.astype(np.float64))
        cnt = np.bincount(inv, minlength=uniq.size).astype(np.float64)
        feats.append((sums / cnt[:, None]).astype(np.float32))
        z, rem = uniq // (g[1] * g[0]), uniq % (g[1] * g[0])
        coors.append(np.stack([np.full_like(uniq, b), z, rem // g[0], rem % g[0]], 1).astype(np.int32))
    return torch.from_numpy(np.concatenate(feats)), torch.from_numpy(np.concatenate(coors))


class HardSimpleVFE(torch.nn.Module):
    """[upstream] mmdet3d v0.17.1 HardSimpleVFE: mean of the (<= max_points) points of a voxel."""

    def __init__(self, num_features=4):
        super().__init__()
        self.num_features = num_features

    def forward(self, features, num_points, coors=None):
        s = features[:, :, :self.num_features].sum(dim=1, keepdim=False)
        return (s / num_points.type_as(features).view(-1, 1)).contiguous()


class _VFELayer(torch.nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear = torch.nn.Linear(cin, cout, bias=False)
        self.norm = torch.nn.BatchNorm1d(cout, eps=1e-3, momentum=0.01)


class HardVFE(torch.nn.Module):
    """[upstream] mmdet3d v0.17.1 HardVFE as configured at FocalFormer3D_Waymo_L.py:141-152
    (no distance / cluster-centre / voxel-centre features, one VFELayer, max-pool output)."""

    def __init__(self, in_channels=5, feat_channels=(64,)):
        super().__init__()
        assert len(feat_channels) == 1
        self.vfe_layers = torch.nn.ModuleList([_VFELayer(in_channels, feat_channels[0])])

    def forward(self, features, num_points, coors=None):
        M, P, _ = features.shape
        mask = (torch.arange(P, device=features.device)[None, :] < num_points[:, None]).type_as(features)
        x = features * mask[..., None]
        l = self.vfe_layers[0]
        y = l.linear(x)                                        # [M, P, C]
        y = l.norm(y.permute(0, 2, 1).contiguous()).permute(0, 2, 1)
        y = torch.relu(y)
        return y.max(dim=1)[0]
