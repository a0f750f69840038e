"""Oracle: SECOND backbone, SECONDFPN neck, FocalEncoder neck.  TEST INFRASTRUCTURE (see oracle/__init__.py).

[upstream] mmdet3d v0.17.1 SECOND / SECONDFPN (cfg ``FocalFormer3D_L.py:207-222``, called at
``focalformer3d.py:169-171``); in-tree ``models/necks/focal_encoder.py:15-222`` and
``models/utils/encoder_utils.py:10-33``.
"""
import torch
from torch import nn
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


class FocalEncoderLayer(nn.Module):
    """focal_encoder.py:15-87, LiDAR-only 'bevfusionmb2' branch (iterbev_wo_img=True)."""

    def __init__(self, hidden_channel, iterbev="bevfusionmb2", iterbev_wo_img=True, **kw):
        super().__init__()
        assert iterbev == "bevfusionmb2" and iterbev_wo_img, "oracle covers the LiDAR-only mb2 branch"
        IR = mobilenetv2.InvertedResidual
        self.P_IML = IR(hidden_channel, hidden_channel, stride=1, expand_ratio=2, norm_layer=nn.BatchNorm2d)
        self.P_out_proj = IR(2 * hidden_channel, hidden_channel, stride=1, expand_ratio=1, norm_layer=nn.BatchNorm2d)
        self.P_integration = IR(2 * hidden_channel, hidden_channel, stride=1, expand_ratio=1, norm_layer=nn.BatchNorm2d)

    def forward(self, img_feat, lidar_feat, img_metas=None, extra_args=None):
        I2P_feat = lidar_feat                                          # :70 (iterbev_wo_img)
        P2P_feat = self.P_IML(lidar_feat)                              # :76
        P_Aug_feat = self.P_out_proj(torch.cat((I2P_feat, P2P_feat), dim=1))       # :77
        new_lidar_feat = self.P_integration(torch.cat((P_Aug_feat, lidar_feat), dim=1))  # :78
        return None, new_lidar_feat


class FocalEncoder(nn.Module):
    """focal_encoder.py:90-222, input_img=False path."""

    def __init__(self, num_layers=2, in_channels_img=64, in_channels_pts=384, hidden_channel=128, bn_momentum=0.1,
                 bias="auto", iterbev="bevfusion", max_points_height=5, multistage_heatmap=False, input_img=True,
                 input_pts=True, iterbev_wo_img=False, extra_feat=False, **kw):
        super().__init__()
        assert not input_img and input_pts
        self.iterbev_wo_img = iterbev_wo_img
        self.multistage_heatmap = multistage_heatmap
        self.input_img = input_img
        self.shared_conv_pts = nn.Conv2d(in_channels_pts, hidden_channel, 3, padding=1, bias=bool(bias))  # :120
        self.num_layers = num_layers if num_layers else 0
        self.fusion_blocks = nn.ModuleList(
            [FocalEncoderLayer(hidden_channel, iterbev=iterbev, iterbev_wo_img=iterbev_wo_img) for _ in range(self.num_layers)])
        self.extra_feat = extra_feat
        if extra_feat:
            self.extra_output = ConvBNReLU(hidden_channel, hidden_channel, 3, norm_layer=nn.BatchNorm2d, activation_layer=None)

    def forward(self, img_feats, pts_feats, img_metas=None):
        new_img_feat = None
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
