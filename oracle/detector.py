"""Oracle: FocalFormer3D detector assembly (LiDAR-only eval forward).  TEST INFRASTRUCTURE.

Follows ``projects/mmdet3d_plugin/models/detectors/focalformer3d.py``: extract_pts_feat :155-175,
extract_feat :177-187, voxelize :189-209, simple_test_pts :306-319, simple_test :321-332.
Module attribute names reproduce the reference state-dict keys so one
state dict drives both this oracle and the CUDA product.
"""
import torch
from torch import nn

from .voxelize import voxelize_batch, dynamic_voxelize_mean, HardSimpleVFE, HardVFE
from .sparse import SparseEncoder
from .bev import SECOND, SECONDFPN, FocalEncoder
from .head import FocalDecoder


def _strip(d):
    return {k: v for k, v in d.items() if k != "type"}


class FocalFormer3D(nn.Module):
    def __init__(self, pts_voxel_layer=None, pts_voxel_encoder=None, pts_middle_encoder=None, pts_backbone=None,
                 pts_neck=None, imgpts_neck=None, pts_bbox_head=None, train_cfg=None, test_cfg=None,
                 input_img=True, input_pts=True, **unused):
        super().__init__()
        self.input_img, self.input_pts = input_img, input_pts
        head = _strip(pts_bbox_head)
        head["test_cfg"] = test_cfg["pts"] if test_cfg and "pts" in test_cfg else test_cfg
        if input_img:
            # image tower (focalformer3d.py:133-153); camera-only (DeformFormer3D_C_R50) stops here, the fusion config
            # (FocalFormer3D_LC) also builds the LiDAR tower below
            from .camera import ResNet50, FPN, CameraFocalEncoder
            self.img_backbone = ResNet50(**unused.get("img_backbone", {}))
            self.img_neck = FPN(**_strip(unused["img_neck"]))
            if not input_pts:
                self.imgpts_neck = CameraFocalEncoder(**_strip(imgpts_neck))
                self.pts_bbox_head = FocalDecoder(**head)
                return
        self.voxel_cfg = pts_voxel_layer
        ve = _strip(pts_voxel_encoder)
        self.dynamic = "Dynamic" in pts_voxel_encoder["type"]             # focalformer3d.py:80
        if self.dynamic:
            assert pts_voxel_encoder["type"] == "DynamicSimpleVFE"
            self.pts_voxel_encoder = None                                 # no parameters; see dynamic_voxelize_mean
        elif pts_voxel_encoder["type"] == "HardSimpleVFE":
            self.pts_voxel_encoder = HardSimpleVFE(**ve)
        else:
            self.pts_voxel_encoder = HardVFE(in_channels=ve["in_channels"], feat_channels=ve["feat_channels"])
        self.pts_middle_encoder = SparseEncoder(**_strip(pts_middle_encoder))
        self.pts_backbone = SECOND(**_strip(pts_backbone))
        self.pts_neck = SECONDFPN(**_strip(pts_neck))
        self.imgpts_neck = FocalEncoder(**_strip(imgpts_neck))      # the neck's own input_img default (True) applies
        self.pts_bbox_head = FocalDecoder(**head)

    def voxelize(self, points):
        mv = self.voxel_cfg["max_voxels"]
        mv = mv[1] if isinstance(mv, (tuple, list)) else mv              # eval picks max_voxels[1]
        return voxelize_batch(points, self.voxel_cfg["voxel_size"], self.voxel_cfg["point_cloud_range"],
                              self.voxel_cfg["max_num_points"], mv)

    @torch.no_grad()
    def extract_pts_feat(self, points, stages=None):
        dev = next(self.parameters()).device           # .cuda() turns the oracle into the stock-PyTorch GPU stand-in
        if self.dynamic:                               # focalformer3d.py:159-163
            vf, coors = dynamic_voxelize_mean(points, self.voxel_cfg["voxel_size"], self.voxel_cfg["point_cloud_range"])
            vf, coors = vf.to(dev), coors.to(dev)
            voxels = num_points = None
        else:
            voxels, num_points, coors = self.voxelize(points)
            voxels, num_points, coors = voxels.to(dev), num_points.to(dev), coors.to(dev)
            vf = self.pts_voxel_encoder(voxels, num_points, coors)
        batch_size = int(coors[-1, 0]) + 1
        x = self.pts_middle_encoder(vf, coors, batch_size)
        if stages is not None:
            stages.update(voxels=voxels, num_points=num_points, coors=coors, voxel_features=vf, middle=x)
        x = self.pts_backbone(x)
        if stages is not None:
            stages["backbone"] = x
        x = self.pts_neck(x)
        if stages is not None:
            stages["neck"] = x[0]
        return x

    @torch.no_grad()
    def forward_raw(self, points, stages=None, img=None, img_metas=None):
        """points: list[B] of [Ni, F] (or None for camera-only) -> (head output dict, list of per-scene result dicts).
        img: [B, N, 3, H, W]; img_metas: list[B] of dict(lidar2img=[N, 4, 4])."""
        feats = [None]
        if self.input_img:
            B, N, C, H, W = img.shape
            for m in img_metas:
                m.update(input_shape=(H, W))                                                  # focalformer3d.py:136-139
            c = self.img_backbone(img.view(B * N, C, H, W).float())                          # focalformer3d.py:133-153
            feats = self.img_neck(c)
            if stages is not None:
                stages["img_backbone"], stages["img_feat"] = c, feats[0]
        if self.input_img and not self.input_pts:
            _, new_pts = self.imgpts_neck(feats[0], None, img_metas)                          # only level 0 (:186)
        elif self.input_img:
            pts_feats = self.extract_pts_feat(points, stages)
            _, new_pts = self.imgpts_neck(feats[0], pts_feats[0], img_metas)                  # :177-187
        else:
            pts_feats = self.extract_pts_feat(points, stages)
            _, new_pts = self.imgpts_neck(None, pts_feats[0], None)
        second = list(new_pts[1]) if isinstance(new_pts[1], (list, tuple)) else new_pts[1]
        if stages is not None:
            stages["conv_feat"] = new_pts[0]
            stages["stage_feats"] = list(second) if isinstance(second, list) else [second]
        outs = self.pts_bbox_head([new_pts[0], second], None, None)
        res = self.pts_bbox_head.get_bboxes(outs)
        return outs[0][0], res

    def simple_test(self, points, img_metas=None, img=None, rescale=False):
        _, res = self.forward_raw(points)
        return [dict(pts_bbox=dict(boxes_3d=r["boxes_3d"], scores_3d=r["scores_3d"], labels_3d=r["labels_3d"]))
                for r in res]


def build_oracle(model_cfg):
    cfg = _strip(model_cfg)
    m = FocalFormer3D(**cfg)
    m.eval()
    return m
