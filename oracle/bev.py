"""Oracle: SECOND backbone, SECONDFPN neck, FocalEncoder neck.  TEST INFRASTRUCTURE (see oracle/__init__.py).

[upstream] mmdet3d v0.17.1 SECOND / SECONDFPN (cfg ``FocalFormer3D_L.py:207-222``, called at
``focalformer3d.py:169-171``); in-tree ``models/necks/focal_encoder.py:15-222`` and
``models/utils/encoder_utils.py:10-33``.
"""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F
import torchvision.models.mobilenetv2 as mobilenetv2


class SECOND(nn.Module):
    def __init__(self, in_channels=128, out_channels=(128, 128, 256), layer_nums=(3, 5, 5),
                 layer_strides=(2, 2, 2), norm_cfg=None, conv_cfg=None, **kw):
        super().__init__()
        eps = (norm_cfg or {}).get("eps", 1e-3)
        mom = (norm_cfg or {}).get("momentum", 0.01)
        in_filters = [in_channels, *out_channels[:-1]]
        blocks = []
        for i, layer_num in enumerate(layer_nums):
            block = [nn.Conv2d(in_filters[i], out_channels[i], 3, stride=layer_strides[i], padding=1, bias=False),
                     nn.BatchNorm2d(out_channels[i], eps=eps, momentum=mom), nn.ReLU(inplace=True)]
            for _ in range(layer_num):
                block += [nn.Conv2d(out_channels[i], out_channels[i], 3, padding=1, bias=False),
                          nn.BatchNorm2d(out_channels[i], eps=eps, momentum=mom), nn.ReLU(inplace=True)]
            blocks.append(nn.Sequential(*block))
        self.blocks = nn.ModuleList(blocks)

    def forward(self, x):
        outs = []
        for b in self.blocks:
            x = b(x)
            outs.append(x)
        return tuple(outs)


class SECONDFPN(nn.Module):
    def __init__(self, in_channels=(128, 128, 256), out_channels=(256, 256, 256), upsample_strides=(1, 2, 4),
                 norm_cfg=None, upsample_cfg=None, conv_cfg=None, use_conv_for_no_stride=False, **kw):
        super().__init__()
        eps = (norm_cfg or {}).get("eps", 1e-3)
        mom = (norm_cfg or {}).get("momentum", 0.01)
        deblocks = []
        for i, oc in enumerate(out_channels):
            stride = upsample_strides[i]
            if stride > 1 or (stride == 1 and not use_conv_for_no_stride):
                up = nn.ConvTranspose2d(in_channels[i], oc, kernel_size=stride, stride=stride, bias=False)
            else:
                stride = int(round(1 / stride))
                up = nn.Conv2d(in_channels[i], oc, kernel_size=stride, stride=stride, bias=False)
            deblocks.append(nn.Sequential(up, nn.BatchNorm2d(oc, eps=eps, momentum=mom), nn.ReLU(inplace=True)))
        self.deblocks = nn.ModuleList(deblocks)

    def forward(self, x):
        ups = [d(x[i]) for i, d in enumerate(self.deblocks)]
        return [torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]]


class ConvBNReLU(nn.Module):
    """encoder_utils.py:10-33."""

    def __init__(self, cin, cout, kernel_size=3, stride=1, dilation=1, groups=1, norm_layer=nn.BatchNorm2d,
                 activation_layer=nn.ReLU, bias="auto"):
        super().__init__()
        padding = dilation * (kernel_size - 1) // 2
        self.use_norm = norm_layer is not None
        self.use_activation = activation_layer is not None
        if bias == "auto":
            bias = not self.use_norm
        self.conv = nn.Conv2d(cin, cout, kernel_size, stride, padding, dilation=dilation, groups=groups, bias=bias)
        if self.use_norm:
            self.bn = norm_layer(cout)
        if self.use_activation:
            self.activation = activation_layer(inplace=True)

    def forward(self, x):
        x = self.conv(x)
        if self.use_norm:
            x = self.bn(x)
        if self.use_activation:
            x = self.activation(x)
        return x


def local_similar(x_ori, x_loc, kH, kW):
    """locatt_ops similar_forward (kernels.cuh:4-42 `cc2k`, float in / double accumulate): y[b,h,w,k] =
    sum_c x_ori[b,c,h,w] * x_loc[b,c,h+dy,w+dx], 0 for neighbours outside the map; k = (dy+rH)*kW + (dx+rW)."""
    B, C, H, W = x_loc.shape
    nb = F.unfold(x_loc.double(), (kH, kW), padding=(kH // 2, kW // 2)).view(B, C, kH * kW, H, W)
    return (x_ori.double().unsqueeze(2) * nb).sum(1).permute(0, 2, 3, 1).float().contiguous()


def local_weighting(x_loc, weight, kH, kW):
    """locatt_ops weighting_forward (kernels.cuh:44-80 `ck2c_ori`): y[b,c,h,w] = sum_k x_loc[b,c,h+dy,w+dx] * weight[b,h,w,k]."""
    B, C, H, W = x_loc.shape
    nb = F.unfold(x_loc.double(), (kH, kW), padding=(kH // 2, kW // 2)).view(B, C, kH * kW, H, W)
    return (nb * weight.double().permute(0, 3, 1, 2).unsqueeze(1)).sum(2).float()


class LocalContextAttentionBlock(nn.Module):
    """encoder_utils.py:109-163: 9x9 local attention with 1x1 ConvBNReLU projections."""

    def __init__(self, cin, cout, kernel_size):
        super().__init__()
        self.kernel_size = kernel_size
        mk = lambda: ConvBNReLU(cin, cout, kernel_size=1, norm_layer=nn.BatchNorm2d, activation_layer=nn.ReLU)
        mk2 = lambda: ConvBNReLU(cout, cout, kernel_size=1, norm_layer=nn.BatchNorm2d, activation_layer=nn.ReLU)
        self.query_project = nn.Sequential(mk(), mk2())
        self.key_project = nn.Sequential(mk(), mk2())
        self.value_project = mk()

    def forward(self, target_feats, source_feats):
        import math
        query, key, value = self.query_project(target_feats), self.key_project(source_feats), self.value_project(source_feats)
        k = self.kernel_size
        weight = local_similar(query, key, k, k)                                           # :160
        weight = F.softmax(weight / math.sqrt(key.size(1)), -1)                            # :161 (zeros of the padding take part)
        self.debug = dict(query=query, key=key, value=value, weight=weight)
        return local_weighting(value, weight, k, k)                                        # :162


class I2P(nn.Module):
    """encoder_utils.py:184-262: image -> BEV projection of the 'proj' fusion variant (FocalFormer3D_LC_Proj).  Every BEV
    cell owns `max_points_height` sample points along z (cell centres of a W x H x P grid over the HARD-CODED range
    +-54 m, [-5, 3] m, :208); each point is projected into every camera with `lidar2img`, the image feature is sampled
    bilinearly (grid_sample, zeros outside), averaged over the cameras that see the point, and one single-head attention
    (query = the cell's LiDAR feature, keys = values = its P sampled features, unseen points masked) yields the cell's
    decorated feature; cells no camera sees stay zero."""

    def __init__(self, pts_channels, img_channels, dropout, max_points_height=5):
        super().__init__()
        self.pts_channels, self.img_channels, self.max_points_height = pts_channels, img_channels, max_points_height
        self.learnedAlign = nn.MultiheadAttention(pts_channels, 1, dropout=dropout, kdim=img_channels, vdim=img_channels,
                                                  batch_first=True)

    def sample_points(self, ref):
        """[P*H*W, 3] sample points, height-major (index = (p*H + h)*W + w), :209-212 + create_3D_grid :174-182."""
        P = self.max_points_height
        W, H = ref.shape[-1], ref.shape[-2]
        pz, py, px = torch.meshgrid(torch.linspace(0, P - 1, P), torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W),
                                    indexing="ij")
        cell = torch.stack([px + 0.5, py + 0.5, pz + 0.5], 0).view(3, -1).t().to(ref)
        rng = ref.new_tensor([-54.0, -54.0, -5.0, 54.0, 54.0, 3.0])
        return cell / ref.new_tensor([W, H, P]) * (rng[3:] - rng[:3]) + rng[:3]

    def forward(self, lidar_feat, img_feat, img_metas):
        """lidar_feat [B, C, H, W]; img_feat [B, N, C, fH, fW]; returns [B, C, H, W]."""
        B, C, H, W = lidar_feat.shape
        P = self.max_points_height
        out = torch.zeros_like(lidar_feat)
        pts = self.sample_points(lidar_feat)
        homo = torch.cat([pts, torch.ones_like(pts[:, :1])], 1)                             # [V, 4]
        for b in range(B):
            l2i = lidar_feat.new_tensor(np.asarray(img_metas[b]["lidar2img"]))              # [N, 4, 4]
            cam = torch.matmul(l2i.unsqueeze(1), homo[None, :, :, None]).squeeze(-1)        # [N, V, 4]   :224
            eps = 1e-5
            seen = cam[..., 2:3] > eps
            uv = cam[..., 0:2] / torch.maximum(cam[..., 2:3], torch.ones_like(cam[..., 2:3]) * eps)
            ih, iw = img_metas[b]["input_shape"]
            uv = torch.stack([uv[..., 0] / iw, uv[..., 1] / ih], -1)
            uv = (uv - 0.5) * 2                                                              # :236
            seen = seen & (uv[..., 0:1] > -1.0) & (uv[..., 0:1] < 1.0) & (uv[..., 1:2] > -1.0) & (uv[..., 1:2] < 1.0)
            samp = F.grid_sample(img_feat[b], uv.unsqueeze(-2), align_corners=False).squeeze(-1)   # [N, C, V]   :243 (torch default)
            N = samp.shape[0]
            seen = seen.view(N, 1, P, H, W)
            samp = samp.view(N, self.img_channels, P, H, W)
            mean = (samp * seen).sum(0) / (seen.sum(0) + 1e-10)                              # over cameras :249
            kv = mean.flatten(2, 3).transpose(0, 2)                                         # [H*W, P, C]
            vis = (seen[:, 0].sum(0) > 0).view(P, H * W).t()                                 # [H*W, P]
            q = lidar_feat[b].flatten(1, 2).t().unsqueeze(1)                                 # [H*W, 1, C]
            valid = vis.sum(1) > 0
            att = lidar_feat.new_zeros(H * W, 1, self.pts_channels)
            att[valid] = self.learnedAlign(q[valid], kv[valid], kv[valid], attn_mask=(~vis[valid]).unsqueeze(1))[0]
            out[b] = att.squeeze(1).t().view(self.pts_channels, H, W)
        return out


class FocalEncoderLayer(nn.Module):
    """focal_encoder.py:15-87: 'bevfusionmb2' (LiDAR-only, iterbev_wo_img) and 'bevfusion' (LiDAR + camera BEV with
    iter_bev_cam and no I2P projection: cam_lss supplies the image BEV feature) branches."""

    def __init__(self, hidden_channel, iterbev="bevfusionmb2", iterbev_wo_img=True, iter_bev_cam=None, need_projbev=True,
                 layer_id=None, max_points_height=5, **kw):
        super().__init__()
        self.iterbev, self.iterbev_wo_img = iterbev, iterbev_wo_img
        self.project = bool(need_projbev and not iterbev_wo_img and iter_bev_cam and layer_id == 0)   # :28-31
        hc = hidden_channel
        if iterbev == "bevfusionmb2":
            assert iterbev_wo_img, "oracle covers the LiDAR-only mb2 branch"
            IR = mobilenetv2.InvertedResidual
            self.P_IML = IR(hc, hc, stride=1, expand_ratio=2, norm_layer=nn.BatchNorm2d)
            self.P_out_proj = IR(2 * hc, hc, stride=1, expand_ratio=1, norm_layer=nn.BatchNorm2d)
            self.P_integration = IR(2 * hc, hc, stride=1, expand_ratio=1, norm_layer=nn.BatchNorm2d)
        else:
            assert iterbev == "bevfusion" and (iterbev_wo_img or iter_bev_cam), \
                "oracle covers 'bevfusion' with iter_bev_cam (camera BEV from cam_lss, or the layer-0 I2P projection)"
            if self.project:
                self.I2P_block = I2P(hc, hc, 0.1, max_points_height=max_points_height)        # :31
            self.P_IML = LocalContextAttentionBlock(hc, hc, 9)                              # :40
            self.P_out_proj = ConvBNReLU(2 * hc, hc, kernel_size=1, norm_layer=nn.BatchNorm2d, activation_layer=None)
            self.P_integration = ConvBNReLU(2 * hc, hc, kernel_size=1, norm_layer=nn.BatchNorm2d, activation_layer=None)
        self.iterimg_conv = None
        if not iterbev_wo_img:
            import torchvision.models.resnet as resnet
            self.iterimg_conv = nn.Sequential(resnet.BasicBlock(hc, hc, norm_layer=nn.BatchNorm2d))   # :50-52

    def forward(self, img_feat, lidar_feat, img_metas=None, extra_args=None):
        I2P_feat = lidar_feat if self.iterbev_wo_img else img_feat     # :58-70 (iter_bev_cam)
        if self.project:                                               # :62-64: layer 0 lifts the image-plane feature
            B = lidar_feat.shape[0]
            I2P_feat = self.I2P_block(lidar_feat, img_feat.view(B, -1, *img_feat.shape[1:]), img_metas)
            img_feat = I2P_feat
        if self.iterbev == "bevfusion":
            P2P_feat = self.P_IML(lidar_feat, lidar_feat)              # :72
        else:
            P2P_feat = self.P_IML(lidar_feat)                          # :76
        P_Aug_feat = self.P_out_proj(torch.cat((I2P_feat, P2P_feat), dim=1))       # :73,77
        new_lidar_feat = self.P_integration(torch.cat((P_Aug_feat, lidar_feat), dim=1))  # :74,78
        new_img_feat = self.iterimg_conv(img_feat) if self.iterimg_conv is not None else None   # :82-85
        return new_img_feat, new_lidar_feat


class FocalEncoder(nn.Module):
    """focal_encoder.py:90-222: LiDAR-only (input_img=False) and LiDAR + camera (cam_lss) paths."""

    def __init__(self, num_layers=2, in_channels_img=64, in_channels_pts=384, hidden_channel=128, bn_momentum=0.1,
                 bias="auto", iterbev="bevfusion", max_points_height=5, multistage_heatmap=False, input_img=True,
                 input_pts=True, iterbev_wo_img=False, extra_feat=False, iter_bev_cam=False, cam_lss=False, pc_range=None,
                 img_scale=None, **kw):
        super().__init__()
        assert input_pts and cam_lss != "proj"
        self.use_lss = bool(input_img and cam_lss)
        self.iterbev_wo_img = iterbev_wo_img
        self.multistage_heatmap = multistage_heatmap
        self.input_img = input_img
        self.shared_conv_pts = nn.Conv2d(in_channels_pts, hidden_channel, 3, padding=1, bias=bool(bias))  # :120
        if self.use_lss:
            from .camera import LiftSplatShoot
            self.cam_lss = LiftSplatShoot(grid=0.6, inputC=256, outputC=hidden_channel, camC=64, pc_range=pc_range,
                                          img_scale=img_scale, downsample=4)                 # :129-131
        elif input_img:
            self.shared_conv_img = nn.Conv2d(in_channels_img, hidden_channel, 3, padding=1, bias=bool(bias))   # :134-141
        self.num_layers = num_layers if num_layers else 0
        self.fusion_blocks = nn.ModuleList(
            [FocalEncoderLayer(hidden_channel, iterbev=iterbev, iterbev_wo_img=iterbev_wo_img, iter_bev_cam=iter_bev_cam,
                               need_projbev=not cam_lss, layer_id=i, max_points_height=max_points_height)
             for i in range(self.num_layers)])
        self.extra_feat = extra_feat
        if extra_feat:
            self.extra_output = ConvBNReLU(hidden_channel, hidden_channel, 3, norm_layer=nn.BatchNorm2d, activation_layer=None)

    def forward(self, img_feats, pts_feats, img_metas=None):
        new_img_feat = None
        if self.input_img and not self.use_lss:
            new_img_feat = self.shared_conv_img(img_feats)             # :199 (image-plane feature; I2P lifts it in layer 0)
        elif self.input_img:                                           # :173-197
            from .camera import lidar2img_to_rots_trans
            B = len(img_metas)
            rots, trans = zip(*[lidar2img_to_rots_trans(m["lidar2img"]) for m in img_metas])
            rots, trans = torch.stack(rots).to(img_feats), torch.stack(trans).to(img_feats)
            new_img_feat, _ = self.cam_lss(img_feats.view(B, -1, *img_feats.shape[-3:]), rots, trans)
            self.debug = dict(img_bev=new_img_feat)
        new_pts_feat = self.shared_conv_pts(pts_feats)                 # :204
        pts_feat_conv = new_pts_feat.clone()                           # :207
        if self.input_img or self.iterbev_wo_img:
            multistage = []
            for i in range(self.num_layers):
                new_img_feat, new_pts_feat = self.fusion_blocks[i](new_img_feat, new_pts_feat, img_metas, {})
                if self.multistage_heatmap:
                    multistage.append(new_pts_feat)
            if self.multistage_heatmap:
                new_pts_feat = multistage
                if self.extra_feat:
                    new_pts_feat.append(self.extra_output(new_pts_feat[-1]))   # :218-219
            return new_img_feat, [pts_feat_conv, new_pts_feat]
        return None, [new_pts_feat, None]
