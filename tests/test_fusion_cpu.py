"""LiDAR + camera fusion (FocalFormer3D_LC), CPU side: the oracle's FocalEncoder fusion path against the golden fixture
produced by the REAL reference FocalEncoder (tests/golden/make_golden.py: focal_encoder.py + lss.py + encoder_utils.py on
the CPU, the JIT locatt extension served by the oracle restatement), an independent formulation of the local attention,
the shipped config and the checkpoint-key contract."""
import math
import os
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "focalformer3d_lc_encoder.pt")
REF_CFG = "/root/reference/projects/configs/focalformer3d/FocalFormer3D_LC.py"


@pytest.fixture(scope="module")
def gold():
    if not os.path.exists(GOLD):
        pytest.skip("golden fixture missing")
    return torch.load(GOLD, map_location="cpu")


def fusion_cfg(gold):
    from focalformer3d_b200.config import load_config, default_config_path, scaled_fusion_cfg
    return scaled_fusion_cfg(load_config(default_config_path("focalformer3d_lc"))["model"], bev=gold["bev"],
                             img_hw=gold["img_hw"], num_proposals=12)


def test_local_attention_restatement_vs_explicit_loops():
    """unfold formulation == the literal per-pixel / per-offset loops of kernels.cuh (cc2k, ck2c_ori)."""
    from oracle.bev import local_similar, local_weighting
    g = torch.Generator().manual_seed(1)
    B, C, H, W, K = 1, 6, 5, 7, 3
    q, k, v = (torch.randn(B, C, H, W, generator=g) for _ in range(3))
    sim = local_similar(q, k, K, K)
    w = torch.softmax(sim / math.sqrt(C), -1)
    out = local_weighting(v, w, K, K)
    R = K // 2
    for y in range(H):
        for x in range(W):
            acc = torch.zeros(C, dtype=torch.float64)
            for kk in range(K * K):
                ny, nx = y + kk // K - R, x + kk % K - R
                inside = 0 <= ny < H and 0 <= nx < W
                s = float((q[0, :, y, x].double() * k[0, :, ny, nx].double()).sum()) if inside else 0.0
                assert abs(float(sim[0, y, x, kk]) - s) < 1e-5
                if inside:
                    acc += v[0, :, ny, nx].double() * float(w[0, y, x, kk])
            assert (out[0, :, y, x].double() - acc).abs().max().item() < 1e-5
    # corner pixel: 5 of the 9 neighbours are outside the map, score exactly 0 and still receive soft-max mass
    assert (sim[0, 0, 0, [0, 1, 2, 3, 6]] == 0).all() and (w[0, 0, 0, [0, 1, 2, 3, 6]] > 0).all()


def test_oracle_fusion_encoder_matches_reference(gold):
    from focalformer3d_b200.synth import make_state_dict
    from oracle.bev import FocalEncoder
    cfg = fusion_cfg(gold)
    sd = make_state_dict(cfg, seed=gold["weights_seed"])
    ne = {k: v for k, v in cfg["imgpts_neck"].items() if k != "type"}
    enc = FocalEncoder(**ne).eval()
    enc.load_state_dict({k[len("imgpts_neck."):]: v for k, v in sd.items() if k.startswith("imgpts_neck.")}, strict=True)
    metas = [dict(lidar2img=m.numpy()) for m in gold["lidar2img"]]
    with torch.no_grad():
        new_img, (conv_feat, stages) = enc(gold["feat"], gold["neck"], metas)
    close = lambda a, b: ((a - b).abs() / (1 + b.abs())).max().item()
    assert close(conv_feat, gold["conv_feat"]) < 1e-5
    assert len(stages) == len(gold["stage_feats"]) == 3                 # two fused stage features + extra_output
    for a, b in zip(stages, gold["stage_feats"]):
        assert close(a, b) < 1e-3                                        # the camera BEV carries the cumsum-trick noise
    assert close(new_img, gold["new_img_feat"]) < 1e-3
    w = enc.fusion_blocks[0].P_IML.debug["weight"]
    assert w.max(-1)[0].mean().item() > 0.2                              # the fixture exercises a peaked soft-max


@pytest.mark.skipif(not os.path.exists(REF_CFG), reason="reference tree not present on this box")
def test_lc_config_mirrors_reference():
    from focalformer3d_b200.config import load_config, default_config_path
    from focalformer3d_b200.synth import param_spec
    ref, mine = load_config(REF_CFG)["model"], load_config(default_config_path("focalformer3d_lc"))["model"]
    for part in ("img_backbone", "img_neck", "pts_voxel_layer", "pts_voxel_encoder", "pts_middle_encoder", "pts_backbone",
                 "pts_neck", "imgpts_neck"):
        assert dict(ref[part]) == dict(mine[part]), part
    for k, v in dict(mine["pts_bbox_head"]).items():
        assert dict(ref["pts_bbox_head"])[k] == v, k
    assert dict(ref["test_cfg"]["pts"]) == dict(mine["test_cfg"]["pts"])
    assert set(param_spec(ref)) == set(param_spec(mine))


def test_lc_state_dict_contract_and_packing(gold):
    from focalformer3d_b200.synth import make_state_dict, param_spec
    from focalformer3d_b200.model import build_model
    from oracle.detector import build_oracle
    cfg = fusion_cfg(gold)
    sd = make_state_dict(cfg, 2)
    assert set(build_oracle(cfg).state_dict()) == set(sd) == set(param_spec(cfg))
    for k in ("imgpts_neck.fusion_blocks.1.P_IML.query_project.1.conv.weight", "imgpts_neck.fusion_blocks.0.P_IML.value_project.bn.running_var",
              "imgpts_neck.fusion_blocks.0.P_out_proj.conv.weight", "imgpts_neck.fusion_blocks.1.iterimg_conv.0.bn2.weight",
              "imgpts_neck.cam_lss.bevencode.0.weight", "imgpts_neck.shared_conv_pts.bias", "imgpts_neck.extra_output.bn.bias",
              "img_backbone.layer4.2.conv3.weight", "pts_middle_encoder.conv_out.0.weight",
              "pts_bbox_head.heatmap_head_img.1.1.bias"):
        assert k in sd, k
    model = build_model(cfg)
    model.load_state_dict(sd, strict=True)
    model.prepare("cpu")
    blk = model.imgpts_neck.pk["fusion_blocks.0"]
    assert tuple(blk["qkv1"][0].shape) == (1, 128, 384)                  # query.0 | key.0 | value projections in one GEMM
    assert "img1" in blk and "img1" not in model.imgpts_neck.pk["fusion_blocks.1"]
    with pytest.raises(ValueError):
        model.forward_raw([torch.zeros(4, 5)])                           # fusion config without images: loud failure


# ------------------------------------------------------------------------------------------------------------------
# FocalFormer3D_LC_Proj ('proj' variant: I2P projection block instead of Lift-Splat-Shoot) -- oracle side only
PROJ = os.path.join(ROOT, "tests", "golden", "focalformer3d_lc_proj_encoder.pt")
REF_PROJ_CFG = "/root/reference/projects/configs/focalformer3d/FocalFormer3D_LC_Proj.py"


def test_oracle_proj_encoder_matches_reference():
    """oracle/bev.py I2P + FocalEncoder against the REAL FocalEncoder / I2P of the reference (golden fixture)."""
    if not os.path.exists(PROJ):
        pytest.skip("golden fixture missing")
    from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
    from focalformer3d_b200.synth import make_state_dict
    from oracle.bev import FocalEncoder
    gold = torch.load(PROJ, map_location="cpu")
    cfg = scaled_model_cfg(load_config(default_config_path("focalformer3d_lc_proj"))["model"], bev=gold["bev"], num_proposals=12)
    sd = make_state_dict(cfg, seed=gold["weights_seed"])
    ne = {k: v for k, v in cfg["imgpts_neck"].items() if k != "type"}
    enc = FocalEncoder(**ne).eval()
    enc.load_state_dict({k[len("imgpts_neck."):]: v for k, v in sd.items() if k.startswith("imgpts_neck.")}, strict=True)
    metas = [dict(lidar2img=m.numpy(), input_shape=tuple(gold["img_hw"])) for m in gold["lidar2img"]]
    B, N = len(metas), gold["lidar2img"].shape[1]
    with torch.no_grad():
        new_img, (conv_feat, stages) = enc(gold["feat"], gold["neck"], metas)
        img_plane = enc.shared_conv_img(gold["feat"])
        i2p = enc.fusion_blocks[0].I2P_block(conv_feat, img_plane.view(B, N, *img_plane.shape[1:]), metas)
    close = lambda a, b: ((a - b).abs() / (1 + b.abs())).max().item()
    assert close(conv_feat, gold["conv_feat"]) < 1e-5
    assert close(i2p, gold["i2p"]) < 1e-4                       # the projection block on its own
    assert (gold["i2p"].abs().sum(1) > 0).float().mean().item() > 0.5
    for a, b in zip(stages, gold["stage_feats"]):
        assert close(a, b) < 1e-4
    assert close(new_img, gold["new_img_feat"]) < 1e-4


@pytest.mark.skipif(not os.path.exists(REF_PROJ_CFG), reason="reference tree not present on this box")
def test_lc_proj_config_mirrors_reference_and_product_refuses_it():
    from focalformer3d_b200.config import load_config, default_config_path
    from focalformer3d_b200.synth import param_spec
    from focalformer3d_b200.model import build_model
    ref, mine = load_config(REF_PROJ_CFG)["model"], load_config(default_config_path("focalformer3d_lc_proj"))["model"]
    for part in ("img_backbone", "img_neck", "pts_middle_encoder", "pts_backbone", "pts_neck", "imgpts_neck"):
        assert dict(ref[part]) == dict(mine[part]), part
    assert set(param_spec(ref)) == set(param_spec(mine))
    assert "imgpts_neck.fusion_blocks.0.I2P_block.learnedAlign.in_proj_weight" in param_spec(mine)
    with pytest.raises(NotImplementedError):                     # no silent fallback: the CUDA product does not build it yet
        build_model(mine)
