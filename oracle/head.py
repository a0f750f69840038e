"""Oracle: FocalDecoder head (eval path), prediction FFN, TransFusionBBoxCoder.  TEST INFRASTRUCTURE.

In-tree reference: ``projects/mmdet3d_plugin/models/dense_heads/focal_decoder.py`` (ctor :118-335,
forward :522-992 eval branches only, get_bboxes :1313-1413, get_dense_grid_points :1655-1664),
``models/utils/utils.py:16-66``, ``models/utils/decoder_utils.py:495-578``,
``core/bbox/coders/transfusion_bbox_coder.py:54-158``.  [upstream] pieces: mmcv ConvModule
(conv(bias = not norm) + BN + ReLU) and mmdet3d v0.17.1 ``rotation_3d_in_axis``.
"""
import copy
import math
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from .transformer import DeformableDetrTransformerDecoder


class ConvModule(nn.Module):
    """[upstream] mmcv ConvModule with a norm: conv(no bias) -> BN -> ReLU.  Attribute names conv / bn."""

    def __init__(self, cin, cout, kernel_size, stride=1, padding=0, conv="2d"):
        super().__init__()
        C = nn.Conv2d if conv == "2d" else nn.Conv1d
        B = nn.BatchNorm2d if conv == "2d" else nn.BatchNorm1d
        self.conv = C(cin, cout, kernel_size, stride=stride, padding=padding, bias=False)
        self.bn = B(cout)

    def forward(self, x):
        return torch.relu(self.bn(self.conv(x)))


class MLP(nn.Module):
    """utils.py:16-28."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = F.relu(layer(x)) if i < self.num_layers - 1 else layer(x)
        return x


def gen_sineembed_for_position(pos_tensor):
    """utils.py:40-66 (2-D branch): 128 sin/cos features per coordinate, T=10000, scale 2*pi, (y, x) order."""
    scale = 2 * math.pi
    dim_t = torch.arange(128, dtype=torch.float32, device=pos_tensor.device)
    dim_t = 10000 ** (2 * (dim_t // 2) / 128)
    x_embed = pos_tensor[:, :, 0] * scale
    y_embed = pos_tensor[:, :, 1] * scale
    pos_x = x_embed[:, :, None] / dim_t
    pos_y = y_embed[:, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, 0::2].sin(), pos_x[:, :, 1::2].cos()), dim=3).flatten(2)
    pos_y = torch.stack((pos_y[:, :, 0::2].sin(), pos_y[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((pos_y, pos_x), dim=2)


class PredFFN(nn.Module):
    """decoder_utils.py:495-578: per-query Conv1d heads (ConvModule(128->64) + Conv1d(64->k, bias))."""

    def __init__(self, in_channels, heads, head_conv=64):
        super().__init__()
        self.heads = heads
        for head, (classes, num_conv) in heads.items():
            layers, c = [], in_channels
            for _ in range(num_conv - 1):
                layers.append(ConvModule(c, head_conv, 1, conv="1d"))
                c = head_conv
            layers.append(nn.Conv1d(head_conv, classes, 1, bias=True))
            setattr(self, head, nn.Sequential(*layers))

    def forward(self, x):
        return {h: getattr(self, h)(x) for h in self.heads}


class TransFusionBBoxCoder:
    """transfusion_bbox_coder.py:8-158 (decode side)."""

    def __init__(self, pc_range, out_size_factor, voxel_size, post_center_range=None, score_threshold=None,
                 code_size=8, **kw):
        self.pc_range, self.out_size_factor, self.voxel_size = pc_range, out_size_factor, voxel_size
        self.post_center_range, self.score_threshold, self.code_size = post_center_range, score_threshold, code_size

    def decode_box(self, rot, dim, center, height, vel):            # :54-69 (in-place on the clones passed in)
        center[:, 0, :] = center[:, 0, :] * self.out_size_factor * self.voxel_size[0] + self.pc_range[0]
        center[:, 1, :] = center[:, 1, :] * self.out_size_factor * self.voxel_size[1] + self.pc_range[1]
        dim[:, 0, :] = dim[:, 0, :].exp()
        dim[:, 1, :] = dim[:, 1, :].exp()
        dim[:, 2, :] = dim[:, 2, :].exp()
        height = height - dim[:, 2:3, :] * 0.5
        rot = torch.atan2(rot[:, 0:1, :], rot[:, 1:2, :])
        parts = [center, height, dim, rot] + ([] if vel is None else [vel])
        return torch.cat(parts, dim=1).permute(0, 2, 1)

    def decode(self, heatmap, rot, dim, center, height, vel, filter=False):     # :71-158
        final_preds = heatmap.max(1, keepdims=False).indices
        final_scores = heatmap.max(1, keepdims=False).values
        final_box_preds = self.decode_box(rot, dim, center, height, vel)
        if not filter:
            return [dict(bboxes=final_box_preds[i], scores=final_scores[i], labels=final_preds[i])
                    for i in range(heatmap.shape[0])]
        if self.score_threshold is not None:
            thresh_mask = final_scores > self.score_threshold
        assert self.post_center_range is not None
        pcr = torch.tensor(self.post_center_range, device=heatmap.device)
        mask = (final_box_preds[..., :3] >= pcr[:3]).all(2)
        mask &= (final_box_preds[..., :3] <= pcr[3:]).all(2)
        out = []
        for i in range(heatmap.shape[0]):
            cmask = mask[i, :]
            if self.score_threshold:                                 # 0.0 is falsy -> not applied (:140)
                cmask = cmask & thresh_mask[i]
            out.append(dict(bboxes=final_box_preds[i, cmask], scores=final_scores[i, cmask],
                            labels=final_preds[i, cmask], keep=cmask))
        return out


def rotation_3d_in_axis_z(points, angles):
    """[upstream] mmdet3d v0.17.1 rotation_3d_in_axis(points, angles, axis=2) restricted to (x, y):
    rot_mat_T = [[cos, -sin], [sin, cos]]; out = einsum('aij,jka->aik') -> x' = x cos + y sin, y' = -x sin + y cos."""
    s, c = torch.sin(angles)[:, None], torch.cos(angles)[:, None]
    x, y = points[..., 0], points[..., 1]
    return torch.stack([x * c + y * s, -x * s + y * c], dim=-1)


def canonical_topk(flat, k):
    """torch.topk(sorted=False) leaves order (and the choice among ties) implementation-defined
    (focal_decoder.py:688).  The oracle and the product both use the canonical choice: descending value,
    ties broken towards the LOWER flat index, output in that order."""
    B, n = flat.shape
    idx = torch.arange(n, device=flat.device)[None].expand(B, -1)
    # stable descending sort == lexicographic (value desc, index asc)
    order = torch.argsort(flat, dim=-1, descending=True, stable=True)
    return order[:, :k]


class FocalDecoder(nn.Module):
    def __init__(self, num_proposals=128, hidden_channel=128, hidden_channel_roi=512, num_classes=4,
                 num_decoder_layers=1, num_heads=8, initialize_by_heatmap=False, nms_kernel_size=1,
                 common_heads=dict(), num_heatmap_convs=2, bias="auto", train_cfg=None, test_cfg=None,
                 bbox_coder=None, multiscale=False, multistage_heatmap=False, reuse_first_heatmap=False,
                 extra_feat=False, heatmap_box=False, bevpos=False, input_img=True, iterbev_wo_img=False,
                 mask_heatmap_mode="poscls", roi_feats=0, roi_dropout_rate=0.0, roi_expand_ratio=1.0,
                 roi_based_reg=False, classaware_reg=False, boxpos=None, decoder_cfg=None,
                 loss_cls=dict(type='GaussianFocalLoss', reduction='mean'), **unused):
        super().__init__()
        assert initialize_by_heatmap and not heatmap_box and boxpos is None
        self.classaware_reg = classaware_reg
        # focal_decoder.py:164-166: a background class is appended unless loss_cls.use_sigmoid; every shipped config
        # sets use_sigmoid=True (the softmax variant would break class_encoding's channel count in the reference)
        assert loss_cls.get('use_sigmoid', False), 'only the use_sigmoid=True classification head is restated'
        self.num_classes = num_classes
        self.num_proposals_ori = self.num_proposals = num_proposals
        self.num_decoder_layers = num_decoder_layers
        self.nms_kernel_size = nms_kernel_size
        self.test_cfg = test_cfg
        self.multiscale, self.extra_feat, self.bevpos = multiscale, extra_feat, bevpos
        self.multistage_heatmap = multistage_heatmap
        self.reuse_first_heatmap = reuse_first_heatmap
        if reuse_first_heatmap:
            self.multistage_heatmap += 1                                         # :138-139
        self.input_img, self.iterbev_wo_img = input_img, iterbev_wo_img
        self.mask_heatmap_mode = mask_heatmap_mode
        hc = hidden_channel
        if multiscale:
            self.dconv = ConvModule(hc, hc, 3, stride=2, padding=1)              # :150-162
            self.dconv2 = ConvModule(hc, hc, 3, stride=2, padding=1)
        self.bbox_coder = TransFusionBBoxCoder(**{k: v for k, v in bbox_coder.items() if k != "type"})
        self.roi_feats, self.roi_based_reg = roi_feats, roi_based_reg
        self.roi_expand_ratio = [roi_expand_ratio] * num_decoder_layers if isinstance(roi_expand_ratio, float) \
            else roi_expand_ratio
        if roi_feats:                                                            # :186-200
            fc, pre = [], roi_feats ** 2 * hc * (3 if multiscale else 1)
            for i in range(3):
                chl = hidden_channel_roi if i < 2 else hc
                fc += [nn.Linear(pre, chl, bias=False), nn.BatchNorm1d(chl), nn.ReLU(inplace=True)]
                if roi_dropout_rate > 1e-4:
                    fc.append(nn.Dropout(roi_dropout_rate))
                pre = chl
            self.roi_mlp = nn.Sequential(*fc)
        self.heatmap_head = nn.Sequential(ConvModule(hc, hc, 3, padding=1),
                                          nn.Conv2d(hc, num_classes, 3, padding=1, bias=bool(bias)))   # :202-221
        if input_img or iterbev_wo_img:
            if self.multistage_heatmap:
                self.heatmap_head_img = nn.ModuleList()
                for i in range(self.multistage_heatmap):
                    self.heatmap_head_img.append(None if (i == 0 and reuse_first_heatmap)
                                                 else copy.deepcopy(self.heatmap_head))
            else:
                self.heatmap_head_img = copy.deepcopy(self.heatmap_head)
        self.class_encoding = nn.Conv1d(num_classes, hc, 1)
        self.decoder = nn.ModuleList()
        self.pos_embed_learned = nn.ModuleList()
        dc = {k: v for k, v in decoder_cfg.items() if k != "type"}
        for _ in range(num_decoder_layers):
            self.decoder.append(DeformableDetrTransformerDecoder(**dc))
            self.pos_embed_learned.append(MLP(256, hc, hc, 2))
        self.prediction_heads = nn.ModuleList()
        for _ in range(num_decoder_layers):
            heads = copy.deepcopy(common_heads)
            if self.classaware_reg:                                              # :317-319
                heads = {k: [v[0] * num_classes, v[1]] for k, v in heads.items()}
            heads.update(dict(heatmap=(num_classes, num_heatmap_convs)))
            self.prediction_heads.append(PredFFN(hc, heads))
        xs = test_cfg["grid_size"][0] // test_cfg["out_size_factor"]
        ys = test_cfg["grid_size"][1] // test_cfg["out_size_factor"]
        self.bev_pos = self.create_2D_grid(xs, ys)

    @staticmethod
    def create_2D_grid(x_size, y_size):                                          # :337-344
        by, bx = torch.meshgrid(torch.linspace(0, x_size - 1, x_size), torch.linspace(0, y_size - 1, y_size),
                                indexing="ij")
        bx, by = bx + 0.5, by + 0.5
        return torch.cat([bx[None], by[None]], dim=0)[None].view(1, 2, -1).permute(0, 2, 1)

    def _nms(self, heatmap):                                                     # :669-685
        pad = self.nms_kernel_size // 2
        local_max = torch.zeros_like(heatmap)
        inner = F.max_pool2d(heatmap, kernel_size=self.nms_kernel_size, stride=1, padding=0)
        if pad:
            local_max[:, :, pad:-pad, pad:-pad] = inner
        else:
            local_max = inner
        ds = self.test_cfg["dataset"]
        ex = (8, 9) if ds == "nuScenes" else (1, 2) if ds == "Waymo" else ()
        for c in ex:
            local_max[:, c] = heatmap[:, c]
        return heatmap * (heatmap == local_max)

    def _exempt(self):
        ds = self.test_cfg["dataset"]
        return slice(8, 10) if ds == "nuScenes" else slice(1, 3) if ds == "Waymo" else slice(0, 0)

    def forward(self, pts_inputs, img_inputs=None, img_metas=None, topk_fn=canonical_topk):
        self.num_proposals = self.num_proposals_ori
        lidar_feat = pts_inputs[0]
        stage_list = list(pts_inputs[1]) if isinstance(pts_inputs[1], (list, tuple)) else pts_inputs[1]
        if self.extra_feat:
            extra_feats = stage_list.pop(-1)                                     # :526-528
        B, C = lidar_feat.shape[:2]
        lidar_feat_flatten = lidar_feat.view(B, C, -1)
        dev = lidar_feat.device
        bev_pos = self.bev_pos.repeat(B, 1, 1).to(dev)
        if self.multiscale:
            s = lidar_feat.shape[2]
            bev_pos_2 = self.create_2D_grid(s // 2, s // 2).repeat(B, 1, 1).to(dev) * 2
            bev_pos_4 = self.create_2D_grid(s // 4, s // 4).repeat(B, 1, 1).to(dev) * 4
        query_box = None
        dbg = {}
        if not self.multistage_heatmap:                                          # :539-586
            dense_heatmap = self.heatmap_head(lidar_feat)
            if self.input_img or self.iterbev_wo_img:
                new_lidar_feat = stage_list[-1] if isinstance(stage_list, list) else stage_list
                lidar_feat_flatten = new_lidar_feat.view(*lidar_feat_flatten.shape)
                dense_heatmap_img = self.heatmap_head_img(new_lidar_feat.view(lidar_feat.shape))
                heatmap = (dense_heatmap.sigmoid() + dense_heatmap_img.sigmoid()) / 2
                heatmap_train = [dense_heatmap, dense_heatmap_img]
            else:
                heatmap = dense_heatmap.sigmoid()
                new_lidar_feat = lidar_feat
                heatmap_train = dense_heatmap
            heatmap = self._nms(heatmap).view(B, heatmap.shape[1], -1)
            top = torch.argsort(heatmap.view(B, -1), dim=-1, descending=True, stable=True)[..., :self.num_proposals]
            cls, pos = top // heatmap.shape[-1], top % heatmap.shape[-1]
            query_feat = lidar_feat_flatten.gather(index=pos[:, None, :].expand(-1, C, -1), dim=-1)
            self.query_labels = cls
            query_feat = query_feat + self.class_encoding(F.one_hot(cls, self.num_classes).permute(0, 2, 1).float())
            query_pos = bev_pos.gather(index=pos[:, :, None].expand(-1, -1, 2), dim=1)
            query_heatmap_score = heatmap.gather(index=pos[:, None, :].expand(-1, self.num_classes, -1), dim=-1)
            dbg["top_proposals"], dbg["nms_heatmap"] = [top], [heatmap]
        else:
            dense_heatmap = self.heatmap_head(lidar_feat)                        # :588
            multistage_feats = stage_list
            if self.reuse_first_heatmap:
                multistage_feats.insert(0, lidar_feat)                           # :591-592
            q_labels, q_feats, q_poses, q_scores = [], [], [], []
            acc_masks = torch.ones_like(dense_heatmap).view(B, -1)
            heatmap_train, dbg["top_proposals"], dbg["nms_heatmap"], dbg["acc_masks"] = [], [], [], []
            for i in range(self.multistage_heatmap):
                if i == 0 and self.reuse_first_heatmap:
                    heatmap = dense_heatmap.sigmoid()                            # :631
                    heatmap_train.append(dense_heatmap)
                    heatmap = heatmap * acc_masks.view(*heatmap.shape)           # :634
                else:
                    dense_heatmap_img = self.heatmap_head_img[i](multistage_feats[i])   # :637
                    heatmap = dense_heatmap_img.sigmoid()                        # :662
                    if i == 0:
                        heatmap_train.append(dense_heatmap)
                    heatmap = heatmap * acc_masks.view(*heatmap.shape)           # :666
                    heatmap_train.append(dense_heatmap_img)
                lidar_feat_flatten = multistage_feats[i].reshape(B, C, -1)       # :669
                heatmap = self._nms(heatmap)
                heatmap = heatmap.view(B, heatmap.shape[1], -1)
                top = topk_fn(heatmap.view(B, -1), self.num_proposals)           # :688
                cls, pos = top // heatmap.shape[-1], top % heatmap.shape[-1]     # :690-691
                qf = lidar_feat_flatten.gather(index=pos[:, None, :].expand(-1, C, -1), dim=-1)
                q_labels.append(cls)
                qf = qf + self.class_encoding(F.one_hot(cls, self.num_classes).permute(0, 2, 1).float())  # :697-700
                q_feats.append(qf)
                q_poses.append(bev_pos.gather(index=pos[:, :, None].expand(-1, -1, 2), dim=1))            # :701
                q_scores.append(heatmap.gather(index=pos[:, None, :].expand(-1, self.num_classes, -1), dim=-1))  # :702
                assert self.mask_heatmap_mode == "poscls"
                sel = acc_masks.new_zeros(B, self.num_classes * heatmap.shape[-1])
                sel.scatter_(index=top, dim=1, src=torch.ones_like(top, dtype=acc_masks.dtype))           # :729-731
                sel = sel.reshape(*dense_heatmap.shape)
                k = self.nms_kernel_size
                selk = F.max_pool2d(sel, kernel_size=k, stride=1, padding=k // 2)                         # :776
                ex = self._exempt()
                selk[:, ex] = sel[:, ex]                                                                  # :777-780
                dbg["top_proposals"].append(top)
                dbg["nms_heatmap"].append(heatmap)
                dbg["acc_masks"].append(acc_masks.clone())
                acc_masks = acc_masks * (1.0 - selk).view(*acc_masks.shape)                               # :782
            self.query_labels = torch.cat(q_labels, dim=1)
            query_feat = torch.cat(q_feats, dim=2)
            query_pos = torch.cat(q_poses, dim=1)
            query_heatmap_score = torch.cat(q_scores, dim=2)
            self.num_proposals = self.num_proposals_ori * self.multistage_heatmap
        dbg["query_feat0"], dbg["query_pos0"] = query_feat.clone(), query_pos.clone()

        if self.multiscale:                                                      # :810-823
            if not self.multistage_heatmap:
                lidar_feat = new_lidar_feat
            else:
                lidar_feat = extra_feats if self.extra_feat else multistage_feats[-1]
            ms = [lidar_feat]
            ms.append(self.dconv(ms[-1]))
            ms.append(self.dconv2(ms[-1]))
            ms_flat = torch.cat([m.flatten(2, 3) for m in ms], dim=-1)
        elif self.multistage_heatmap:
            lidar_feat = extra_feats if self.extra_feat else multistage_feats[-1]
        else:
            lidar_feat = new_lidar_feat

        ret_dicts = []
        nq = self.num_proposals
        dbg["stage_query_feat"], dbg["roi_feat"], dbg["value"] = [], [], []
        for i in range(self.num_decoder_layers):                                 # :826-958
            if not self.multiscale:
                spatial_shapes = torch.as_tensor([list(lidar_feat.shape[-2:])], dtype=torch.long, device=dev)
                lidar_feat_flatten = lidar_feat.flatten(2, 3)
                ms = [lidar_feat]
            else:
                spatial_shapes = torch.as_tensor([list(m.shape[2:]) for m in ms], dtype=torch.long, device=dev)
                lidar_feat_flatten = ms_flat
                if self.bevpos and i == 0:
                    bev_pos = torch.cat([bev_pos, bev_pos_2, bev_pos_4], dim=1)  # :846-848
            level_start_index = torch.cat([spatial_shapes.new_zeros(1), spatial_shapes.prod(1).cumsum(0)[:-1]])
            WH = torch.flip(spatial_shapes[:1], dims=(1,))[:, None].float()
            reference_points = query_pos / WH                                    # :869
            query_pos_embed = self.pos_embed_learned[i](gen_sineembed_for_position(reference_points[:, :, :2]))
            if self.bevpos:
                bev_pos_embed = self.pos_embed_learned[i](gen_sineembed_for_position((bev_pos / WH)[:, :, :2]))
                value_in = lidar_feat_flatten + bev_pos_embed.transpose(1, 2)    # :883-886
            else:
                value_in = lidar_feat_flatten
            if self.roi_feats and query_box is not None:                         # :890-922
                rot, dim, center, height, vel = (query_box[:, 6:8], query_box[:, 3:6], query_box[:, 0:2],
                                                 query_box[:, 2:3], query_box[:, 8:])
                std = self.bbox_coder.decode_box(rot.clone(), dim.clone() * self.roi_expand_ratio[i], center.clone(),
                                                 height.clone(), vel.clone() if vel.shape[1] else None)
                std = std.reshape(B * nq, std.shape[-1])
                gp = self.get_dense_grid_points(std, B * nq, self.roi_feats)
                gp = rotation_3d_in_axis_z(gp, std[:, 6])
                gp = gp + std[:, None, :2]
                gp = gp.view(B, nq, self.roi_feats ** 2, 2)
                if self.test_cfg["dataset"] == "nuScenes":
                    pcr = torch.tensor([-54, -54, -5.0, 54, 54, 3.0], device=dev)
                else:
                    pcr = torch.tensor([-75.2, -75.2, -2, 75.2, 75.2, 4], device=dev)
                gp = (gp - pcr[:2]) / (pcr[3:5] - pcr[:2])
                gp = (gp * 2.0 - 1.0).clip(min=-2.0, max=2.0)
                rf = torch.cat([F.grid_sample(f, gp, mode="bilinear", align_corners=False) for f in ms], dim=1)
                rf = rf.permute(0, 2, 1, 3).reshape(B * nq, -1)
                dbg["roi_feat"].append(rf)
                rf = self.roi_mlp(rf).view(B, nq, C).transpose(1, 2)
                query_feat = query_feat + rf                                     # :922
            dbg["value"].append(value_in)
            query_feat, reference_points = self.decoder[i](
                query=query_feat.permute(2, 0, 1), key=None, value=value_in.permute(2, 0, 1),
                query_pos=query_pos_embed.permute(1, 0, 2), reference_points=reference_points,
                spatial_shapes=spatial_shapes, level_start_index=level_start_index,
                valid_ratios=torch.ones((B, 1, 2), device=dev), key_padding_mask=None, attn_masks=None)   # :927-933
            query_feat = query_feat.permute(1, 2, 0)
            dbg["stage_query_feat"].append(query_feat)
            query_pos = reference_points * WH
            res = self.prediction_heads[i](query_feat)                           # :939
            if self.classaware_reg:                                              # :940-943: keep the query's own class
                P = res["center"].shape[-1]
                lab = self.query_labels[:, None, None, :].clip(0, self.num_classes - 1)
                for k in ("center", "height", "dim", "rot"):
                    t = res[k].view(B, self.num_classes, -1, P)
                    res[k] = t.gather(index=lab.expand(-1, -1, t.shape[2], -1), dim=1)[:, 0]
            res["center"] = res["center"] + query_pos.permute(0, 2, 1)           # :945
            query_pos = res["center"].detach().clone().permute(0, 2, 1)
            if self.roi_based_reg and query_box is not None:                     # :949-951
                res["dim"] = torch.cat([res["dim"][:, :2] + query_box[:, 3:5], res["dim"][:, 2:]], dim=1)
                res["rot"] = res["rot"] + query_box[:, 6:8]
            qb = [res["center"], res["height"], res["dim"], res["rot"]]
            if "vel" in res:
                qb.append(res["vel"])
            query_box = torch.cat(qb, dim=1).detach()                            # :954-957
            ret_dicts.append(res)
        new_res = {}
        for key in ret_dicts[0].keys():
            new_res[key] = torch.cat([r[key] for r in ret_dicts], dim=-1)        # :960-992
        new_res["query_heatmap_score"] = query_heatmap_score
        new_res["dense_heatmap"] = heatmap_train
        self.debug = dbg
        return [[new_res]]

    @staticmethod
    def get_dense_grid_points(rois, n, grid_size):                               # :1655-1664
        ii, jj = torch.meshgrid(torch.arange(grid_size, device=rois.device), torch.arange(grid_size, device=rois.device),
                                indexing="ij")
        dense_idx = torch.stack([ii.reshape(-1), jj.reshape(-1)], dim=1)[None].repeat(n, 1, 1).float()
        size = rois.view(n, -1)[:, 3:5]
        return (dense_idx + 0.5) / grid_size * size[:, None] - size[:, None] / 2

    def get_bboxes(self, preds_dicts):
        """:1313-1413 with nms_type=None; generalised over the batch (the reference asserts bs==1)."""
        p = preds_dicts[0][0]
        nq = self.num_proposals
        score = p["heatmap"][..., -nq:].sigmoid()
        one_hot = F.one_hot(self.query_labels, num_classes=self.num_classes).permute(0, 2, 1)
        score = score * p["query_heatmap_score"] * one_hot
        vel = p["vel"][..., -nq:].clone() if "vel" in p else None
        temp = self.bbox_coder.decode(score, p["rot"][..., -nq:].clone(), p["dim"][..., -nq:].clone(),
                                      p["center"][..., -nq:].clone(), p["height"][..., -nq:].clone(), vel, filter=True)
        nms_type = self.test_cfg.get("nms_type")
        if self.test_cfg["dataset"] == "nuScenes":                                # :1333-1338
            tasks = [dict(indices=[0, 1, 2, 3, 4, 5, 6, 7], radius=-1), dict(indices=[8], radius=0.175), dict(indices=[9], radius=0.175)]
        else:                                                                     # 'Waymo' :1339-1344
            tasks = [dict(indices=[0], radius=0.7), dict(indices=[1], radius=0.7), dict(indices=[2], radius=0.7)]
        out = []
        for t in temp:
            b, s, l = t["bboxes"], t["scores"], t["labels"]
            if nms_type is not None:                                              # :1352-1385
                keep_mask = torch.zeros_like(s)
                for task in tasks:
                    task_mask = torch.zeros_like(s)
                    for cls_idx in task["indices"]:
                        task_mask += l == cls_idx
                    task_mask = task_mask.bool()
                    if task["radius"] > 0:
                        if nms_type == "circle":
                            dets = torch.cat([b[task_mask][:, :2], s[:, None][task_mask]], dim=1).numpy()
                            keep_idx = torch.tensor(circle_nms(dets, task["radius"]), dtype=torch.long)
                        else:
                            keep_idx = nms_rotated_bev(xywhr2xyxyr(b[task_mask][:, [0, 1, 3, 4, 6]]), s[task_mask], task["radius"],
                                                       self.test_cfg.get("pre_maxsize"), self.test_cfg.get("post_maxsize"))
                    else:
                        keep_idx = torch.arange(int(task_mask.sum()))
                    if keep_idx.shape[0] != 0:
                        keep_mask[torch.where(task_mask != 0)[0][keep_idx]] = 1
                keep_mask = keep_mask.bool()
                full_keep = t["keep"].clone()
                full_keep[t["keep"].nonzero().flatten()[~keep_mask]] = False
                b, s, l = b[keep_mask], s[keep_mask], l[keep_mask]
                t = dict(t, keep=full_keep)
            if len(b) > 200:
                inds = s.argsort(descending=True, stable=True)[:200]
                b, s, l = b[inds], s[inds], l[inds]
            out.append(dict(boxes_3d=b, scores_3d=s, labels_3d=l.int(), keep=t["keep"]))
        return out


# ------------------------------------------------------------------------------------------------------------------
# [upstream] post-processing the head calls when test_cfg.nms_type is set (focal_decoder.py:1352-1385) and the TTA merge
# (core/post_processing/merge_augs.py:111-184): mmdet3d v0.17.1 core.circle_nms, ops.iou3d nms_gpu / boxes_iou_bev.
def circle_nms(dets, thresh, post_max_size=83):
    """mmdet3d.core.post_processing.circle_nms (numba): dets [n, 3] = (x, y, score); a kept box suppresses every lower
    scored box whose SQUARED centre distance is <= thresh; at most post_max_size boxes are kept."""
    import numpy as np
    x1, y1, scores = dets[:, 0], dets[:, 1], dets[:, 2]
    order = scores.argsort()[::-1].astype(np.int32)
    ndets = dets.shape[0]
    suppressed = np.zeros((ndets), dtype=np.int32)
    keep = []
    for _i in range(ndets):
        i = order[_i]
        if suppressed[i] == 1:
            continue
        keep.append(int(i))
        for _j in range(_i + 1, ndets):
            j = order[_j]
            if suppressed[j] == 1:
                continue
            dist = (x1[i] - x1[j]) ** 2 + (y1[i] - y1[j]) ** 2
            if dist <= thresh:
                suppressed[j] = 1
    return keep[:post_max_size] if post_max_size < len(keep) else keep


def xywhr2xyxyr(b):
    """mmdet3d.core.bbox.xywhr2xyxyr: (x, y, w, h, r) -> (x1, y1, x2, y2, r)."""
    o = torch.zeros_like(b)
    hw, hh = b[:, 2] / 2, b[:, 3] / 2
    o[:, 0], o[:, 1], o[:, 2], o[:, 3], o[:, 4] = b[:, 0] - hw, b[:, 1] - hh, b[:, 0] + hw, b[:, 1] + hh, b[:, 4]
    return o


def _rot_overlap(a, b):
    """iou3d_kernel.cu box_overlap: intersection area of two rotated rectangles (x1, y1, x2, y2, angle) -- corners of
    each box inside the other plus the edge/edge intersection points, sorted by angle around their centre, fan area."""
    import math

    def corners(r):
        cx, cy = (r[0] + r[2]) / 2, (r[1] + r[3]) / 2
        c, s = math.cos(r[4]), math.sin(r[4])
        pts = []
        for x, y in ((r[0], r[1]), (r[2], r[1]), (r[2], r[3]), (r[0], r[3])):
            dx, dy = x - cx, y - cy
            pts.append((cx + dx * c - dy * s, cy + dx * s + dy * c))
        return pts, (cx, cy, c, s, (r[2] - r[0]) / 2, (r[3] - r[1]) / 2)

    def inside(p, f):
        cx, cy, c, s, hx, hy = f
        dx, dy = p[0] - cx, p[1] - cy
        u, v = dx * c + dy * s, -dx * s + dy * c
        return abs(u) <= hx + 1e-9 and abs(v) <= hy + 1e-9

    def seg_x(p0, p1, q0, q1):
        d = (p1[0] - p0[0]) * (q1[1] - q0[1]) - (p1[1] - p0[1]) * (q1[0] - q0[0])
        if abs(d) < 1e-12:
            return None
        t = ((q0[0] - p0[0]) * (q1[1] - q0[1]) - (q0[1] - p0[1]) * (q1[0] - q0[0])) / d
        u = ((q0[0] - p0[0]) * (p1[1] - p0[1]) - (q0[1] - p0[1]) * (p1[0] - p0[0])) / d
        if 0 <= t <= 1 and 0 <= u <= 1:
            return (p0[0] + t * (p1[0] - p0[0]), p0[1] + t * (p1[1] - p0[1]))
        return None
    ca, fa = corners(a)
    cb, fb = corners(b)
    pts = [p for p in ca if inside(p, fb)] + [p for p in cb if inside(p, fa)]
    for i in range(4):
        for j in range(4):
            x = seg_x(ca[i], ca[(i + 1) % 4], cb[j], cb[(j + 1) % 4])
            if x is not None:
                pts.append(x)
    if len(pts) < 3:
        return 0.0
    mx, my = sum(p[0] for p in pts) / len(pts), sum(p[1] for p in pts) / len(pts)
    pts.sort(key=lambda p: math.atan2(p[1] - my, p[0] - mx))
    area = 0.0
    for i in range(len(pts)):
        p, q = pts[i], pts[(i + 1) % len(pts)]
        area += (p[0] - mx) * (q[1] - my) - (q[0] - mx) * (p[1] - my)
    return abs(area) / 2


def boxes_iou_bev(a, b):
    """mmdet3d.ops.iou3d boxes_iou_bev: [n, 5] x [m, 5] (x1, y1, x2, y2, r) -> IoU [n, m]."""
    out = torch.zeros(a.shape[0], b.shape[0])
    al, bl = a.double().tolist(), b.double().tolist()
    for i, x in enumerate(al):
        sa = (x[2] - x[0]) * (x[3] - x[1])
        for j, y in enumerate(bl):
            sb = (y[2] - y[0]) * (y[3] - y[1])
            inter = _rot_overlap(x, y)
            out[i, j] = inter / max(sa + sb - inter, 1e-8)
    return out


def nms_rotated_bev(boxes, scores, thresh, pre_maxsize=None, post_max_size=None):
    """mmdet3d.ops.iou3d nms_gpu: score-descending greedy NMS on rotated BEV IoU (> thresh suppresses)."""
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    bl = boxes[order]
    iou = boxes_iou_bev(bl, bl)
    n = bl.shape[0]
    sup, keep = [False] * n, []
    for i in range(n):
        if sup[i]:
            continue
        keep.append(i)
        for j in range(i + 1, n):
            if not sup[j] and iou[i, j] > thresh:
                sup[j] = True
    keep = order[torch.tensor(keep, dtype=torch.long)] if keep else order[:0]
    if post_max_size is not None:
        keep = keep[:post_max_size]
    return keep


def merge_aug_bboxes_3d(aug_results, img_metas, nms_thr=0.1, max_num=500, vote_iou_thresh=0.65):
    """merge_augs.py:14-184 (the live branch: aug_results given, rotate NMS, box voting without score voting).
    aug_results: list of dict(boxes_3d [n, 7|9] tensor, scores_3d, labels_3d); img_metas: list of
    dict(pcd_scale_factor, pcd_horizontal_flip, pcd_vertical_flip)."""
    import math
    rec = []
    for r, m in zip(aug_results, img_metas):
        b = r["boxes_3d"].clone()
        if m.get("pcd_horizontal_flip", False):
            b[:, 1::7] = -b[:, 1::7]
            b[:, 6] = -b[:, 6] + math.pi
        if m.get("pcd_vertical_flip", False):
            b[:, 0::7] = -b[:, 0::7]
            b[:, 6] = -b[:, 6]
        sf = 1.0 / m.get("pcd_scale_factor", 1.0)
        b[:, :6] *= sf
        b[:, 7:] *= sf
        rec.append(b)
    boxes = torch.cat(rec)
    scores = torch.cat([r["scores_3d"] for r in aug_results])
    labels = torch.cat([r["labels_3d"] for r in aug_results]).long()
    if labels.numel() == 0:
        return dict(boxes_3d=boxes, scores_3d=scores, labels_3d=labels)
    nmsb = xywhr2xyxyr(boxes[:, [0, 1, 3, 4, 6]])
    mb, ms, ml = [], [], []
    for c in range(int(labels.max()) + 1):
        ci = labels == c
        if int(ci.sum()) == 0:
            continue
        bi, ni, si, li = boxes[ci], nmsb[ci], scores[ci], labels[ci]
        sel = nms_rotated_bev(ni, si, nms_thr)
        iou = boxes_iou_bev(xywhr2xyxyr(bi[sel][:, [0, 1, 3, 4, 6]]), ni)
        iou[iou < vote_iou_thresh] = 0.0
        voted = (iou[:, :, None] * bi[None]).sum(1) / (iou[:, :, None].sum(1) + 1e-6)
        voted[:, 6] = torch.atan2((iou * torch.sin(bi[None, :, 6])).sum(1) / (iou.sum(1) + 1e-6),
                                  (iou * torch.cos(bi[None, :, 6])).sum(1) / (iou.sum(1) + 1e-6))
        mb.append(voted)
        ms.append(si[sel])
        ml.append(li[sel])
    mb, ms, ml = torch.cat(mb), torch.cat(ms), torch.cat(ml)
    order = ms.sort(0, descending=True)[1][:min(max_num, boxes.shape[0])]
    return dict(boxes_3d=mb[order], scores_3d=ms[order], labels_3d=ml[order])
