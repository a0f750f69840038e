"""Oracle: camera branch (ResNet-50 + FPN image features, Lift-Splat-Shoot camera->BEV).  TEST INFRASTRUCTURE.

Reference: ``projects/mmdet3d_plugin/models/detectors/focalformer3d.py:133-153`` (extract_img_feat),
``projects/mmdet3d_plugin/models/necks/lss.py:126-383`` (CamEncode, LiftSplatShoot) and the camera path of
``models/necks/focal_encoder.py:171-197``.  [upstream] mmdet 2.14 ResNet (depth 50, style 'pytorch', norm_eval;
structurally torchvision's resnet50) and FPN (lateral 1x1 + nearest top-down + 3x3; only level 0 is
consumed, focalformer3d.py:186).
"""
import torch
from torch import nn
import torch.nn.functional as F
import torchvision


class ResNet50(nn.Module):
    """[upstream] mmdet ResNet-50, out_indices (0,1,2,3).  torchvision module/key names."""

    def __init__(self, **kw):
        super().__init__()
        r = torchvision.models.resnet50(weights=None)
        self.conv1, self.bn1, self.relu, self.maxpool = r.conv1, r.bn1, r.relu, r.maxpool
        self.layer1, self.layer2, self.layer3, self.layer4 = r.layer1, r.layer2, r.layer3, r.layer4

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        outs = []
        for l in (self.layer1, self.layer2, self.layer3, self.layer4):
            x = l(x)
            outs.append(x)
        return tuple(outs)


class _Conv(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=k // 2)

    def forward(self, x):
        return self.conv(x)


class FPN(nn.Module):
    """[upstream] mmdet 2.14 FPN (no norm, no activation, add_extra_convs=False)."""

    def __init__(self, in_channels, out_channels, num_outs, **kw):
        super().__init__()
        self.num_outs = num_outs
        self.lateral_convs = nn.ModuleList([_Conv(c, out_channels, 1) for c in in_channels])
        self.fpn_convs = nn.ModuleList([_Conv(out_channels, out_channels, 3) for _ in in_channels])

    def forward(self, inputs):
        lat = [l(x) for l, x in zip(self.lateral_convs, inputs)]
        for i in range(len(lat) - 1, 0, -1):
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
        outs = [c(l) for c, l in zip(self.fpn_convs, lat)]
        for _ in range(self.num_outs - len(outs)):
            outs.append(F.max_pool2d(outs[-1], 1, stride=2))
        return tuple(outs)


def gen_dx_bx(xbound, ybound, zbound):                                        # lss.py:77-82
    dx = torch.Tensor([row[2] for row in [xbound, ybound, zbound]])
    bx = torch.Tensor([row[0] + row[2] / 2.0 for row in [xbound, ybound, zbound]])
    nx = torch.LongTensor([(row[1] - row[0]) / row[2] for row in [xbound, ybound, zbound]])
    return dx, bx, nx


class CamEncode(nn.Module):                                                   # lss.py:126-146
    def __init__(self, D, C, inputC):
        super().__init__()
        self.D, self.C = D, C
        self.depthnet = nn.Conv2d(inputC, D + C, kernel_size=1, padding=0)

    def forward(self, x):
        x = self.depthnet(x)
        depth = x[:, :self.D].softmax(dim=1)
        return depth.unsqueeze(1) * x[:, self.D:self.D + self.C].unsqueeze(2), depth


class LiftSplatShoot(nn.Module):                                              # lss.py:149-383
    def __init__(self, img_scale=(900, 1600), camera_depth_range=(4.0, 45.0, 1.0), pc_range=(-50, -50, -5, 50, 50, 3),
                 downsample=4, grid=3, inputC=256, outputC=128, camC=64, **kw):
        super().__init__()
        self.dx, self.bx, self.nx = gen_dx_bx([pc_range[0], pc_range[3], grid], [pc_range[1], pc_range[4], grid],
                                              [pc_range[2], pc_range[5], grid])
        self.fH, self.fW = img_scale[0] // downsample, img_scale[1] // downsample
        self.camC = camC
        ogfH, ogfW = img_scale
        ds = torch.arange(*camera_depth_range, dtype=torch.float).view(-1, 1, 1).expand(-1, self.fH, self.fW)
        D = ds.shape[0]
        xs = torch.linspace(0, ogfW - 1, self.fW, dtype=torch.float).view(1, 1, self.fW).expand(D, self.fH, self.fW)
        ys = torch.linspace(0, ogfH - 1, self.fH, dtype=torch.float).view(1, self.fH, 1).expand(D, self.fH, self.fW)
        self.frustum = nn.Parameter(torch.stack((xs, ys, ds), -1), requires_grad=False)      # :215-226
        self.D = D
        self.camencode = CamEncode(D, camC, inputC)
        cz = int(camC * ((pc_range[5] - pc_range[2]) // grid))
        self.bevencode = nn.Sequential(
            nn.Conv2d(cz, cz, 3, padding=1, bias=False), nn.BatchNorm2d(cz), nn.ReLU(inplace=True),
            nn.Conv2d(cz, 512, 3, padding=1, bias=False), nn.BatchNorm2d(512), nn.ReLU(inplace=True),
            nn.Conv2d(512, 512, 3, padding=1, bias=False), nn.BatchNorm2d(512), nn.ReLU(inplace=True),
            nn.Conv2d(512, outputC, 3, padding=1, bias=False), nn.BatchNorm2d(outputC), nn.ReLU(inplace=True))

    def get_geometry(self, rots, trans):                                      # :228-271 (no image / lidar aug at test)
        """(u d, v d, d) -> R p + t.  The reference does the 3x3 product with a batched matmul whose summation order is
        the BLAS library's; the oracle fixes it (left to right, every product and sum rounded to fp32, no FMA) so that
        the truncated voxel indices downstream are a well-defined function of the inputs."""
        B, N, _ = trans.shape
        fr = self.frustum.view(1, 1, *self.frustum.shape)                     # [1,1,D,fH,fW,3]
        pd = fr[..., 2]
        px, py = fr[..., 0] * pd, fr[..., 1] * pd
        R = rots.view(B, N, 1, 1, 1, 3, 3)
        t = trans.view(B, N, 1, 1, 1, 3)
        out = [((R[..., i, 0] * px + R[..., i, 1] * py) + R[..., i, 2] * pd) + t[..., i] for i in range(3)]
        return torch.stack(out, -1)

    def get_geometry_matmul(self, rots, trans):
        """The literal reference formulation (self-check for get_geometry; equal up to fp32 rounding)."""
        B, N, _ = trans.shape
        points = self.frustum.repeat(B, N, 1, 1, 1, 1).unsqueeze(-1)
        points = torch.cat((points[:, :, :, :, :, :2] * points[:, :, :, :, :, 2:3], points[:, :, :, :, :, 2:3]), 5)
        points = rots.view(B, N, 1, 1, 1, 3, 3).matmul(points).squeeze(-1)
        return points + trans.view(B, N, 1, 1, 1, 3)

    def voxel_indices(self, geom):
        """((geom - (bx - dx/2)) / dx).long() -- truncation towards zero, as the reference (:335)."""
        return ((geom - (self.bx - self.dx / 2.0).to(geom)) / self.dx.to(geom)).long()

    def voxel_pooling(self, geom, x):
        """:324-362.  The reference sums the points of a voxel with a sort + global fp32 cumsum + difference
        ('cumsum trick'); the oracle states the intended per-voxel sum directly with index_add_ (the cumsum trick loses
        ~1e-7 x running-total per voxel)."""
        B, N, D, H, W, C = x.shape
        g = self.voxel_indices(geom).view(B, -1, 3)
        nx = self.nx.tolist()
        final = x.new_zeros(B, C, nx[2], nx[0], nx[1])
        for b in range(B):
            gi = g[b]
            kept = (gi[:, 0] >= 0) & (gi[:, 0] < nx[0]) & (gi[:, 1] >= 0) & (gi[:, 1] < nx[1]) & (gi[:, 2] >= 0) & (gi[:, 2] < nx[2])
            gi = gi[kept]
            lin = (gi[:, 2] * nx[0] + gi[:, 0]) * nx[1] + gi[:, 1]
            acc = x.new_zeros(nx[2] * nx[0] * nx[1], C)
            acc.index_add_(0, lin, x[b].reshape(-1, C)[kept])
            final[b] = acc.view(nx[2], nx[0], nx[1], C).permute(3, 0, 1, 2)
        return final

    def forward(self, x, rots, trans):
        B, N, C, H, W = x.shape
        geom = self.get_geometry(rots, trans)
        feat, depth = self.camencode(x.view(B * N, C, H, W))
        self.debug = dict(depth=depth, geom_idx=self.voxel_indices(geom))
        feat = feat.view(B, N, self.camC, self.D, H, W).permute(0, 1, 3, 4, 5, 2)
        vox = self.voxel_pooling(geom, feat)                                  # [B, C, Z, X, Y]
        Bv, Cv, Zv, Xv, Yv = vox.shape
        bev = vox.reshape(Bv, Cv * Zv, Xv, Yv).permute(0, 1, 3, 2)           # s2c :373-377
        self.debug["pooled"] = bev
        return self.bevencode(bev), depth


def lidar2img_to_rots_trans(lidar2img):
    """focal_encoder.py:178-194: per camera inverse of lidar2img -> (rot 3x3, trans 3)."""
    inv = torch.inverse(torch.as_tensor(lidar2img, dtype=torch.float32))
    return inv[..., :3, :3].contiguous(), inv[..., :3, 3].contiguous()


class CameraFocalEncoder(nn.Module):
    """focal_encoder.py:90-222 with input_img=True, cam_lss=True, input_pts=False, no fusion layers (C_R50)."""

    def __init__(self, hidden_channel=128, pc_range=None, img_scale=None, num_layers=None, input_pts=True,
                 multistage_heatmap=None, cam_lss=False, **kw):
        super().__init__()
        assert cam_lss and not input_pts and not num_layers and not multistage_heatmap
        self.cam_lss = LiftSplatShoot(grid=0.6, inputC=256, outputC=hidden_channel, camC=64, pc_range=pc_range,
                                      img_scale=img_scale, downsample=4)

    def forward(self, img_feats, pts_feats, img_metas):
        B = len(img_metas)
        rots, trans = zip(*[lidar2img_to_rots_trans(m["lidar2img"]) for m in img_metas])
        rots, trans = torch.stack(rots).to(img_feats), torch.stack(trans).to(img_feats)
        bev, _ = self.cam_lss(img_feats.view(B, -1, *img_feats.shape[-3:]), rots, trans)
        return None, [bev, bev]                                               # :196-197
