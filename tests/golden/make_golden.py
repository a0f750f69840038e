"""Generate golden fixtures by running the REAL reference modules (in-tree part of the hot path) on the CPU.

Run in the build container, where /root/reference exists:   python tests/golden/make_golden.py
The fixtures (tests/golden/*.pt) travel with the repo; the reference does not.

What runs from /root/reference, unmodified:
  projects/mmdet3d_plugin/models/utils/utils.py               (MLP, gen_sineembed_for_position)
  projects/mmdet3d_plugin/core/bbox/coders/transfusion_bbox_coder.py   (TransFusionBBoxCoder)
  projects/mmdet3d_plugin/models/utils/decoder_utils.py        (FFN prediction heads)
  projects/mmdet3d_plugin/models/utils/encoder_utils.py        (ConvBNReLU)
  projects/mmdet3d_plugin/models/necks/focal_encoder.py        (FocalEncoder, FocalEncoderLayer)
  projects/mmdet3d_plugin/models/dense_heads/focal_decoder.py  (FocalDecoder.forward / get_bboxes)
  projects/mmdet3d_plugin/models/necks/lss.py                  (LiftSplatShoot, through FocalEncoder's camera path)
What is stubbed (absent upstream packages mmcv / mmdet / mmdet3d): ConvModule, build_conv_layer,
build_transformer_layer_sequence (-> oracle/transformer.py), rotation_3d_in_axis (-> oracle restatement), registries,
losses, box containers, and the reference's hard-coded device='cuda' (torch.as_tensor / torch.ones wrappers).
The weights are the repo's seeded synthetic state dict, loaded with strict=True into the real modules -- which also
pins the checkpoint-key contract of the in-tree modules.
"""
import importlib
import os
import sys
import types

import torch
from torch import nn

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.dirname(os.path.abspath(__file__))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name, path):
    m = _mod(name)
    m.__path__ = [path]
    return m


class _Registry:
    def __init__(self):
        self.map = {}

    def register_module(self, *a, **k):
        def deco(cls):
            self.map[cls.__name__] = cls
            return cls
        return deco

    def build(self, cfg, **kw):
        cfg = dict(cfg)
        return self.map[cfg.pop("type")](**cfg, **kw)


def install_stubs():
    sys.path.insert(0, ROOT)
    from oracle import head as ohead, transformer as otr

    class ConvModule(nn.Module):                       # mmcv ConvModule: conv(bias = not norm) + BN + ReLU
        def __init__(self, cin, cout, kernel_size, stride=1, padding=0, bias="auto", conv_cfg=None, norm_cfg=None, **kw):
            super().__init__()
            conv = nn.Conv1d if (conv_cfg or {}).get("type") == "Conv1d" else nn.Conv2d
            bn = nn.BatchNorm1d if (norm_cfg or {}).get("type") == "BN1d" else nn.BatchNorm2d
            assert norm_cfg is not None
            self.conv = conv(cin, cout, kernel_size, stride=stride, padding=padding, bias=False)
            self.bn = bn(cout)

        def forward(self, x):
            return torch.relu(self.bn(self.conv(x)))

    def build_conv_layer(cfg, *args, **kwargs):
        conv = nn.Conv1d if (cfg or {}).get("type") == "Conv1d" else nn.Conv2d
        return conv(*args, **kwargs)

    def build_transformer_layer_sequence(cfg):
        cfg = dict(cfg)
        assert cfg.pop("type") == "DeformableDetrTransformerDecoder"
        return otr.DeformableDetrTransformerDecoder(**cfg)

    def force_fp32(*a, **k):
        return (lambda f: f)

    class _Loss(nn.Module):
        def __init__(self, **kw):
            super().__init__()

    class LiDARInstance3DBoxes:
        def __init__(self, tensor, box_dim=7, **kw):
            self.tensor, self.box_dim = tensor, box_dim

    def rotation_3d_in_axis(points, angles, axis=0):
        assert axis == 2
        xy = ohead.rotation_3d_in_axis_z(points[..., :2], angles)
        return torch.cat([xy, points[..., 2:]], dim=-1)

    HEADS, NECKS, BBOX_CODERS, TRANSFORMER = _Registry(), _Registry(), _Registry(), _Registry()
    _mod("mmcv")
    _mod("mmcv.cnn", ConvModule=ConvModule, build_conv_layer=build_conv_layer, kaiming_init=lambda *a, **k: None,
         Linear=nn.Linear, build_activation_layer=None, build_norm_layer=None, xavier_init=None)
    _mod("mmcv.runner", force_fp32=force_fp32)
    _mod("mmcv.cnn.bricks"); _mod("mmcv.cnn.bricks.transformer", build_transformer_layer_sequence=build_transformer_layer_sequence)
    _mod("mmdet"); _mod("mmdet.core", build_bbox_coder=lambda cfg: BBOX_CODERS.build(cfg), multi_apply=None,
                        build_assigner=None, build_sampler=None, AssignResult=None)
    _mod("mmdet.core.bbox", BaseBBoxCoder=object)
    _mod("mmdet.core.bbox.builder", BBOX_CODERS=BBOX_CODERS)
    _mod("mmdet.models"); _mod("mmdet.models.utils"); _mod("mmdet.models.utils.builder", TRANSFORMER=TRANSFORMER)
    _mod("mmdet3d"); _mod("mmdet3d.models", builder=types.SimpleNamespace())
    _mod("mmdet3d.core", circle_nms=None, draw_heatmap_gaussian=None, gaussian_radius=None, xywhr2xyxyr=None,
         PseudoSampler=None, LiDARInstance3DBoxes=LiDARInstance3DBoxes)
    _mod("mmdet3d.core.bbox"); _mod("mmdet3d.core.bbox.structures")
    _mod("mmdet3d.core.bbox.structures.utils", rotation_3d_in_axis=rotation_3d_in_axis)
    _mod("mmdet3d.models.builder", HEADS=HEADS, NECKS=NECKS, build_loss=lambda cfg: _Loss())
    _mod("mmdet3d.models.utils", clip_sigmoid=None)
    # test-time: no point-cloud augmentation recorded in img_metas -> the transformation is the identity
    _mod("mmdet3d.models.fusion_layers", apply_3d_transformation=lambda pts, coord, meta, reverse=False: pts)
    _mod("matplotlib"); _mod("matplotlib.pyplot"); _mod("mpl_toolkits"); _mod("mpl_toolkits.mplot3d", Axes3D=None)
    _mod("mmdet3d.ops"); _mod("mmdet3d.ops.iou3d"); _mod("mmdet3d.ops.iou3d.iou3d_utils", nms_gpu=None)
    # parent packages of the reference modules, WITHOUT executing their __init__ (they import the whole plugin)
    base = os.path.join(REF, "projects", "mmdet3d_plugin")
    _pkg("projects", os.path.join(REF, "projects"))
    _pkg("projects.mmdet3d_plugin", base)
    for sub in ("models", "models/utils", "models/dense_heads", "models/necks", "core", "core/bbox", "core/bbox/coders"):
        _pkg("projects.mmdet3d_plugin." + sub.replace("/", "."), os.path.join(base, sub))
    # locatt_ops is a JIT-compiled CUDA extension (needs a GPU + nvcc at import): its two forward entry points are
    # served by the oracle's restatement of kernels.cuh (oracle/bev.py local_similar / local_weighting)
    from oracle import bev as obev
    locatt = types.SimpleNamespace(localattention=types.SimpleNamespace(
        similar_forward=obev.local_similar, weighting_forward=obev.local_weighting))
    _mod("projects.mmdet3d_plugin.models.utils.ops", locatt_ops=locatt)
    # the copied-from-mmdet transformer.py is star-imported by focal_decoder.py but nothing of it is used
    _mod("projects.mmdet3d_plugin.models.utils.transformer")
    return LiDARInstance3DBoxes


class no_cuda_device:
    """The reference hard-codes device='cuda' (focal_decoder.py:837-863,904): run those calls on the CPU."""

    def __enter__(self):
        self.saved = (torch.as_tensor, torch.ones)
        self.saved_cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t          # lss.py:191-193, focal_encoder.py:184 call .cuda()

        def strip(fn):
            def w(*a, **k):
                if k.get("device") == "cuda":
                    k.pop("device")
                return fn(*a, **k)
            return w
        torch.as_tensor, torch.ones = strip(torch.as_tensor), strip(torch.ones)

    def __exit__(self, *a):
        torch.as_tensor, torch.ones = self.saved
        torch.Tensor.cuda = self.saved_cuda


def main():
    assert os.path.isdir(REF), "run where /root/reference exists"
    box_cls = install_stubs()
    from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
    from focalformer3d_b200.synth import make_state_dict
    fd = importlib.import_module("projects.mmdet3d_plugin.models.dense_heads.focal_decoder")
    fe = importlib.import_module("projects.mmdet3d_plugin.models.necks.focal_encoder")
    ut = importlib.import_module("projects.mmdet3d_plugin.models.utils.utils")
    importlib.import_module("projects.mmdet3d_plugin.core.bbox.coders.transfusion_bbox_coder")

    cfg = scaled_model_cfg(load_config(default_config_path())["model"], bev=16, num_proposals=12)
    sd = make_state_dict(cfg, seed=3)
    g = torch.Generator().manual_seed(11)
    out = {"cfg_bev": 16, "cfg_num_proposals": 12, "weights_seed": 3}

    # ---- utils.py
    pos = torch.rand(2, 7, 2, generator=g) * 1.3 - 0.1
    out["sine_in"], out["sine_out"] = pos, ut.gen_sineembed_for_position(pos)

    # ---- FocalEncoder (real) on a synthetic SECONDFPN output
    ne = dict(cfg["imgpts_neck"]); ne.pop("type")
    enc = fe.FocalEncoder(**ne).eval()
    enc.load_state_dict({k[len("imgpts_neck."):]: v for k, v in sd.items() if k.startswith("imgpts_neck.")}, strict=True)
    neck = torch.randn(2, 512, 16, 16, generator=g) * 0.3
    with torch.no_grad():
        _, (conv_feat, stage_list) = enc(None, neck, [dict(), dict()])
    out["enc_in"], out["enc_conv_feat"], out["enc_stage_feats"] = neck, conv_feat, [t.clone() for t in stage_list]

    # ---- FocalDecoder (real): forward (B=2) and forward + get_bboxes (B=1, the reference asserts bs == 1)
    hd = dict(cfg["pts_bbox_head"]); hd.pop("type")
    head = fd.FocalDecoder(**hd, test_cfg=dict(cfg["test_cfg"]["pts"])).eval()
    head.load_state_dict({k[len("pts_bbox_head."):]: v for k, v in sd.items() if k.startswith("pts_bbox_head.")}, strict=True)
    for B, tag in ((2, "b2"), (1, "b1")):
        feats = [torch.randn(B, 128, 16, 16, generator=g) * 0.5 for _ in range(3)]      # conv_feat, stage feat, extra
        with torch.no_grad(), no_cuda_device():
            res = head([feats[0].clone(), [feats[1].clone(), feats[2].clone()]], None, [dict()] * B)
            r = res[0][0]
            out[f"head_{tag}_in"] = feats
            out[f"head_{tag}_out"] = {k: (v.clone() if torch.is_tensor(v) else [t.clone() for t in v])
                                      for k, v in r.items() if k != "multistage_masks"}
            out[f"head_{tag}_query_labels"] = head.query_labels.clone()
            if B == 1:
                boxes, scores, labels = head.get_bboxes(res, [dict(box_type_3d=box_cls)])[0]
                out["bboxes_b1"] = dict(boxes=boxes.tensor.clone(), scores=scores.clone(), labels=labels.clone())
    torch.save(out, os.path.join(OUT, "focalformer3d_l_intree.pt"))
    n = sum(t.numel() for t in _tensors(out))
    print(f"wrote {os.path.join(OUT, 'focalformer3d_l_intree.pt')} ({n} values)")
    camera_golden(fe)
    fusion_golden(fe)
    head_variants_golden(fd, box_cls)
    proj_golden(fe)


def head_variants_golden(fd, box_cls):
    """Real FocalDecoder.forward / get_bboxes for the head branches the flagship fixture does not reach:
    single-stage DeformFormer3D_L (focal_decoder.py:539-586), HIP without heatmap re-use (FocalFormer3D_LC, :588-664) and
    class-aware regression heads with 14x14 ROI grids (FocalFormer3D_Waymo15_L, :317-319,940-943)."""
    from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
    from focalformer3d_b200.synth import make_state_dict
    out = {}
    for name, cfg_name, n_stage_feats, seed in (("deformformer_l", "deformformer3d_l", 0, 21),
                                                ("fusion_lc", "focalformer3d_lc", 3, 22),
                                                ("waymo15_l", "focalformer3d_waymo15_l", 3, 23)):
        full = load_config(default_config_path(cfg_name))["model"]
        cfg = scaled_model_cfg(full, bev=16, num_proposals=10)
        sd = make_state_dict(cfg, seed=seed)
        g = torch.Generator().manual_seed(100 + seed)
        hd = dict(cfg["pts_bbox_head"]); hd.pop("type")
        head = fd.FocalDecoder(**hd, test_cfg=dict(cfg["test_cfg"]["pts"])).eval()
        head.load_state_dict({k[len("pts_bbox_head."):]: v for k, v in sd.items() if k.startswith("pts_bbox_head.")}, strict=True)
        rec = dict(weights_seed=seed, cfg_name=cfg_name)
        for B, tag in ((2, "b2"), (1, "b1")):
            conv = torch.randn(B, 128, 16, 16, generator=g) * 0.5
            stage = [torch.randn(B, 128, 16, 16, generator=g) * 0.5 for _ in range(n_stage_feats)]
            second = [t.clone() for t in stage] if n_stage_feats else conv.clone()      # DeformFormer: the same tensor twice
            with torch.no_grad(), no_cuda_device():
                res = head([conv.clone(), second], None, [dict()] * B)
                r = res[0][0]
                rec[f"{tag}_in"] = dict(conv=conv, stage=stage)
                rec[f"{tag}_out"] = {k: (v.clone() if torch.is_tensor(v) else [t.clone() for t in v])
                                     for k, v in r.items() if k != "multistage_masks"}
                rec[f"{tag}_query_labels"] = head.query_labels.clone()
                if B == 1:
                    boxes, scores, labels = head.get_bboxes(res, [dict(box_type_3d=box_cls)])[0]
                    rec["bboxes_b1"] = dict(boxes=boxes.tensor.clone(), scores=scores.clone(), labels=labels.clone())
        out[name] = rec
    path = os.path.join(OUT, "head_variants.pt")
    torch.save(out, path)
    print(f"wrote {path} ({sum(t.numel() for t in _tensors(out))} values)")


def fusion_golden(fe):
    """Real FocalEncoder, LiDAR + camera path of FocalFormer3D_LC (focal_encoder.py:171-219: cam_lss camera BEV,
    shared_conv_pts, two 'bevfusion' FocalEncoderLayers = LocalContextAttentionBlock + 1x1 ConvBN fusion + BasicBlock on
    the camera BEV, extra_output) on a reduced size.  The locatt CUDA extension is served by the oracle restatement."""
    from focalformer3d_b200.config import load_config, default_config_path, scaled_fusion_cfg
    from focalformer3d_b200.synth import make_state_dict, synth_cameras
    img_hw, bev = (32, 64), 16
    cfg = scaled_fusion_cfg(load_config(default_config_path("focalformer3d_lc"))["model"], bev=bev, img_hw=img_hw, num_proposals=12)
    sd = make_state_dict(cfg, seed=8)
    g = torch.Generator().manual_seed(13)
    ne = dict(cfg["imgpts_neck"]); ne.pop("type")
    with no_cuda_device():
        enc = fe.FocalEncoder(**ne).eval()
        enc.load_state_dict({k[len("imgpts_neck."):]: v for k, v in sd.items() if k.startswith("imgpts_neck.")}, strict=True)
        B, N = 2, 6
        feat = torch.randn(B * N, 256, img_hw[0] // 4, img_hw[1] // 4, generator=g) * 0.7
        neck = torch.randn(B, 512, bev, bev, generator=g) * 0.3
        metas = [dict(lidar2img=synth_cameras(N, img_hw, seed=60 + b)) for b in range(B)]
        with torch.no_grad():
            new_img, (conv_feat, stage_list) = enc(feat, neck, metas)
    out = dict(img_hw=img_hw, bev=bev, weights_seed=8, feat=feat, neck=neck,
               lidar2img=torch.stack([torch.as_tensor(m["lidar2img"]) for m in metas]), conv_feat=conv_feat.clone(),
               stage_feats=[t.clone() for t in stage_list], new_img_feat=new_img.clone())
    path = os.path.join(OUT, "focalformer3d_lc_encoder.pt")
    torch.save(out, path)
    print(f"wrote {path} ({sum(t.numel() for t in _tensors(out))} values)")


def proj_golden(fe):
    """Real FocalEncoder of FocalFormer3D_LC_Proj (no Lift-Splat-Shoot): shared_conv_img on the image-plane feature, the
    layer-0 I2P projection block (encoder_utils.py:184-262: pillar sample points -> lidar2img -> grid_sample -> camera
    average -> single-head attention), then the same 'bevfusion' layers as FocalFormer3D_LC."""
    from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
    from focalformer3d_b200.synth import make_state_dict, synth_cameras
    img_hw, bev = (32, 64), 16
    cfg = scaled_model_cfg(load_config(default_config_path("focalformer3d_lc_proj"))["model"], bev=bev, num_proposals=12)
    sd = make_state_dict(cfg, seed=9)
    g = torch.Generator().manual_seed(14)
    ne = dict(cfg["imgpts_neck"]); ne.pop("type")
    with no_cuda_device():
        enc = fe.FocalEncoder(**ne).eval()
        enc.load_state_dict({k[len("imgpts_neck."):]: v for k, v in sd.items() if k.startswith("imgpts_neck.")}, strict=True)
        B, N = 2, 6
        feat = torch.randn(B * N, 256, img_hw[0] // 4, img_hw[1] // 4, generator=g) * 0.7
        neck = torch.randn(B, 512, bev, bev, generator=g) * 0.3
        # cameras 20 m behind the rig centre so that a good part of the hard-coded +-54 m pillar grid projects into them
        metas = [dict(lidar2img=synth_cameras(N, img_hw, seed=80 + b), input_shape=img_hw) for b in range(B)]
        with torch.no_grad():
            new_img, (conv_feat, stage_list) = enc(feat, neck, metas)
            i2p = enc.fusion_blocks[0].I2P_block(conv_feat, enc.shared_conv_img(feat).view(B, N, 128, img_hw[0] // 4, img_hw[1] // 4), metas)
    out = dict(img_hw=img_hw, bev=bev, weights_seed=9, feat=feat, neck=neck,
               lidar2img=torch.stack([torch.as_tensor(m["lidar2img"]) for m in metas]), conv_feat=conv_feat.clone(),
               stage_feats=[t.clone() for t in stage_list], new_img_feat=new_img.clone(), i2p=i2p.clone())
    path = os.path.join(OUT, "focalformer3d_lc_proj_encoder.pt")
    torch.save(out, path)
    seen = (i2p.abs().sum(1) > 0).float().mean().item()
    print(f"wrote {path} ({sum(t.numel() for t in _tensors(out))} values; {100 * seen:.0f} % of the BEV cells see a camera)")


def camera_golden(fe):
    """Real FocalEncoder camera path (focal_encoder.py:171-197 -> lss.py LiftSplatShoot.forward) of the
    DeformFormer3D_C_R50 config on a reduced image / BEV size: lidar2img inversion, frustum geometry, truncating voxel
    indices, sort + cumsum pooling, s2c, bevencode.  Weights: the repo's seeded synthetic state dict, strict load."""
    from focalformer3d_b200.config import load_config, default_config_path, scaled_camera_cfg
    from focalformer3d_b200.synth import make_state_dict, synth_cameras
    img_hw, bev = (32, 64), 16
    cfg = scaled_camera_cfg(load_config(default_config_path("deformformer3d_c_r50"))["model"], bev=bev, img_hw=img_hw,
                            num_proposals=12)
    sd = make_state_dict(cfg, seed=6)
    g = torch.Generator().manual_seed(12)
    ne = dict(cfg["imgpts_neck"]); ne.pop("type")
    with no_cuda_device():
        enc = fe.FocalEncoder(**ne).eval()
        enc.load_state_dict({k[len("imgpts_neck."):]: v for k, v in sd.items() if k.startswith("imgpts_neck.")}, strict=True)
        B, N = 2, 6
        feat = torch.randn(B * N, 256, img_hw[0] // 4, img_hw[1] // 4, generator=g) * 0.7
        metas = [dict(lidar2img=synth_cameras(N, img_hw, seed=40 + b)) for b in range(B)]
        with torch.no_grad():
            none, (bev_a, bev_b) = enc(feat, None, metas)
            lss = enc.cam_lss
            # the intermediate the kernels are compared on: pooled volume before bevencode, from the same real module
            rots = torch.stack([torch.stack([torch.Tensor(m).inverse()[:3, :3] for m in md["lidar2img"]]) for md in metas])
            trans = torch.stack([torch.stack([torch.Tensor(m).inverse()[:3, 3] for m in md["lidar2img"]]) for md in metas])
            vox, depth = lss.get_voxels(feat.view(B, N, *feat.shape[1:]), rots, trans, img_metas=metas)
    assert none is None and bev_a is bev_b
    pooled = lss.s2c(vox).contiguous()                       # [B, c*Z + z, Y, X]; stored as (flat index, value) pairs
    nz = pooled.flatten().nonzero().flatten()
    out = dict(img_hw=img_hw, bev=bev, weights_seed=6, feat=feat, lidar2img=torch.stack([torch.as_tensor(m["lidar2img"]) for m in metas]),
               bev_out=bev_a.clone(), depth=depth.clone(), pooled_shape=tuple(pooled.shape), pooled_idx=nz.int(),
               pooled_val=pooled.flatten()[nz].clone())
    path = os.path.join(OUT, "deformformer3d_c_r50_lss.pt")
    torch.save(out, path)
    print(f"wrote {path} ({sum(t.numel() for t in _tensors(out))} values, {nz.numel()} non-zeros of the pooled volume)")


def _tensors(o):
    if torch.is_tensor(o):
        yield o
    elif isinstance(o, dict):
        for v in o.values():
            yield from _tensors(v)
    elif isinstance(o, (list, tuple)):
        for v in o:
            yield from _tensors(v)


if __name__ == "__main__":
    main()
