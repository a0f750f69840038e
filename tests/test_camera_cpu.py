"""Camera branch, CPU side: the oracle's Lift-Splat-Shoot restatement against the golden fixture produced by the REAL
reference modules (tests/golden/make_golden.py ran focal_encoder.py's camera path -> lss.py LiftSplatShoot on the CPU),
the shipped DeformFormer3D_C_R50 config, the checkpoint-key contract and host-side weight packing."""
import os
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "deformformer3d_c_r50_lss.pt")
REF_CFG = "/root/reference/projects/configs/focalformer3d/DeformFormer3D_C_R50.py"


@pytest.fixture(scope="module")
def gold():
    if not os.path.exists(GOLD):
        pytest.skip("golden fixture missing")
    return torch.load(GOLD, map_location="cpu")


def camera_cfg(gold):
    from focalformer3d_b200.config import load_config, default_config_path, scaled_camera_cfg
    return scaled_camera_cfg(load_config(default_config_path("deformformer3d_c_r50"))["model"], bev=gold["bev"],
                             img_hw=gold["img_hw"], num_proposals=12)


def golden_pooled(gold):
    dense = torch.zeros(int(torch.tensor(gold["pooled_shape"]).prod()))
    dense[gold["pooled_idx"].long()] = gold["pooled_val"]
    return dense.view(*gold["pooled_shape"])


def test_oracle_lss_matches_reference(gold):
    from focalformer3d_b200.synth import make_state_dict
    from oracle.camera import CameraFocalEncoder
    cfg = camera_cfg(gold)
    sd = make_state_dict(cfg, seed=gold["weights_seed"])
    ne = {k: v for k, v in cfg["imgpts_neck"].items() if k != "type"}
    enc = CameraFocalEncoder(**ne).eval()
    enc.load_state_dict({k[len("imgpts_neck."):]: v for k, v in sd.items() if k.startswith("imgpts_neck.")}, strict=True)
    metas = [dict(lidar2img=m.numpy()) for m in gold["lidar2img"]]
    with torch.no_grad():
        none, (bev, bev2) = enc(gold["feat"], None, metas)
    assert none is None and bev is bev2
    lss = enc.cam_lss
    assert (lss.debug["depth"].view_as(gold["depth"]) - gold["depth"]).abs().max().item() < 1e-6
    ref = golden_pooled(gold)
    got = lss.debug["pooled"]
    # the reference sums a voxel's points by sort + global fp32 cumsum + differences, the oracle by index_add_: same
    # support (identical truncated voxel indices), values equal up to the cumsum's rounding
    assert torch.equal(got != 0, ref != 0)
    assert (got - ref).abs().max().item() < 2e-4 * max(1.0, ref.abs().max().item())
    assert ((bev - gold["bev_out"]).abs() / (1 + gold["bev_out"].abs())).max().item() < 1e-3


def test_oracle_geometry_forms_agree():
    from focalformer3d_b200.synth import synth_cameras
    from oracle.camera import LiftSplatShoot, lidar2img_to_rots_trans
    lss = LiftSplatShoot(img_scale=(64, 96), pc_range=[-7.2, -7.2, -5, 7.2, 7.2, 3], grid=0.6)
    r, t = lidar2img_to_rots_trans(synth_cameras(6, (64, 96), seed=3))
    g1, g2 = lss.get_geometry(r[None], t[None]), lss.get_geometry_matmul(r[None], t[None])
    assert (g1 - g2).abs().max().item() < 1e-4
    assert (lss.voxel_indices(g1) != lss.voxel_indices(g2)).any(-1).float().mean().item() < 1e-4
    # synthetic rig sanity: camera k looks along yaw 60k degrees; its central ray at 10 m lands ~10 m out at 1.5 m height
    c = g1[0, :, 6, 8, 12]                                   # depth bin 6 -> 10 m, centre pixel
    assert (c[:, :2].norm(dim=1) - 10.3).abs().max().item() < 0.6 and (c[:, 2] - 1.5).abs().max().item() < 0.5


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference tree not present on this box")
def test_c_r50_config_mirrors_reference():
    from focalformer3d_b200.config import load_config, default_config_path
    from focalformer3d_b200.synth import param_spec
    ref, mine = load_config(REF_CFG)["model"], load_config(default_config_path("deformformer3d_c_r50"))["model"]
    for part in ("img_backbone", "img_neck"):
        assert dict(ref[part]) == dict(mine[part]), part
    for part in ("imgpts_neck", "pts_bbox_head"):
        for k, v in dict(mine[part]).items():
            assert dict(ref[part])[k] == v, (part, k)
    assert (ref["input_img"], ref["input_pts"]) == (mine["input_img"], mine["input_pts"]) == (True, False)
    assert set(param_spec(ref)) == set(param_spec(mine))


def test_c_r50_state_dict_contract_and_packing(gold):
    from focalformer3d_b200.synth import make_state_dict, param_spec
    from focalformer3d_b200.model import build_model
    from oracle.detector import build_oracle
    cfg = camera_cfg(gold)
    sd = make_state_dict(cfg, 1)
    osd = build_oracle(cfg).state_dict()
    assert set(osd) == set(sd) == set(param_spec(cfg))
    # [upstream] torchvision / mmdet key names the released checkpoints use
    for k in ("img_backbone.layer3.5.conv3.weight", "img_backbone.layer2.0.downsample.1.running_var",
              "img_neck.lateral_convs.3.conv.weight", "img_neck.fpn_convs.0.conv.bias", "imgpts_neck.cam_lss.frustum",
              "imgpts_neck.cam_lss.camencode.depthnet.bias", "imgpts_neck.cam_lss.bevencode.9.weight",
              "imgpts_neck.cam_lss.bevencode.10.num_batches_tracked", "pts_bbox_head.heatmap_head_img.1.bias"):
        assert k in sd, k
    assert not any(k.startswith("pts_middle_encoder") or k.startswith("imgpts_neck.shared_conv_pts") for k in sd)
    model = build_model(cfg)
    model.load_state_dict(sd, strict=True)
    model.prepare("cpu")
    lss = model.imgpts_neck.pk["lss"]
    assert lss.nxyz == [gold["bev"], gold["bev"], 13] and lss.D == 41
    w0 = lss.bevenc[0][0]
    assert tuple(w0.shape) == (9, 832, 896) and w0.bn == 128          # 832 -> 896 output channels: 7 N tiles of 128
    assert torch.count_nonzero(w0.w[:, :, 832:]).item() == 0
    dn_w, dn_b = lss.depthnet
    assert tuple(dn_w.shape) == (1, 256, 128)                         # 64 context | 41 depth | zero padding
    ref_w = sd["imgpts_neck.cam_lss.camencode.depthnet.weight"].reshape(105, 256)
    assert torch.equal(dn_w.w[0, :, :64], ref_w[41:].t()) and torch.equal(dn_w.w[0, :, 64:105], ref_w[:41].t())
    stem = model.img_backbone.pk["stem"][0]
    assert tuple(stem.shape) == (49, 8, 64) and torch.count_nonzero(stem.w[:, 3:]).item() == 0
    with pytest.raises(ValueError):
        model.forward_raw(None)                                       # camera config without images: loud failure
