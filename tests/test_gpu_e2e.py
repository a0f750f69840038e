"""End-to-end parity of the CUDA path against the CPU oracle on seeded synthetic scenes (reduced BEV so the
oracle finishes in seconds), stage by stage, plus size-independent properties at the full FocalFormer3D_L size.

Bars (BASELINE.json north_star): top-k query indices / class ids bit-exact (as sets per HIP stage: the reference's
torch.topk(sorted=False) leaves the order undefined), heatmaps and box regressions within 1e-3 abs."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-3          # final head outputs: absolute (north-star bar)


def _nchw(t):
    return t.permute(0, 3, 1, 2).cpu()


def _close(a, b, what=""):
    """Intermediate feature maps span several orders of magnitude: mixed tolerance |a-b| <= 1e-3 * (1 + |b|)."""
    err = ((a - b).abs() / (1.0 + b.abs())).max().item()
    assert err < TOL, f"{what}: max mixed abs/rel err {err}"


def _waymo_tiny():
    from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
    from focalformer3d_b200.synth import make_state_dict, synth_points
    cfg = scaled_model_cfg(load_config(default_config_path("focalformer3d_waymo_l"))["model"], bev=24, num_proposals=16)
    pts = [torch.from_numpy(synth_points(n, cfg["pts_voxel_layer"]["point_cloud_range"], seed=10 + s, n_beams=64))
           for s, n in enumerate((7000, 5000))]
    return cfg, make_state_dict(cfg, 1), pts


def _deform_tiny():
    from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
    from focalformer3d_b200.synth import make_state_dict, synth_points
    cfg = scaled_model_cfg(load_config(default_config_path("deformformer3d_l"))["model"], bev=24, num_proposals=20)
    pts = [torch.from_numpy(synth_points(n, cfg["pts_voxel_layer"]["point_cloud_range"], seed=20 + s))
           for s, n in enumerate((6000, 6500))]
    return cfg, make_state_dict(cfg, 2), pts


def _waymo15_tiny():
    """FocalFormer3D_Waymo15_L: 14x14 ROI grids (roi_mlp K = 75264) and class-aware regression heads."""
    from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
    from focalformer3d_b200.synth import make_state_dict, synth_points
    cfg = scaled_model_cfg(load_config(default_config_path("focalformer3d_waymo15_l"))["model"], bev=24, num_proposals=16)
    pts = [torch.from_numpy(synth_points(n, cfg["pts_voxel_layer"]["point_cloud_range"], seed=40 + s, n_beams=64))
           for s, n in enumerate((7000, 5000))]
    return cfg, make_state_dict(cfg, 5), pts


def _dynamic_tiny():
    """DeformFormer3D_L_dynamic: dynamic voxelisation (no point / voxel caps) + DynamicSimpleVFE (mean of all points)."""
    import copy
    from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
    from focalformer3d_b200.synth import make_state_dict, synth_points
    base = copy.deepcopy(load_config(default_config_path("deformformer3d_l"))["model"])
    base["pts_voxel_layer"].update(max_num_points=-1, max_voxels=(-1, -1))      # DeformFormer3D_L_dynamic.py:189-196
    base["pts_voxel_encoder"] = dict(type="DynamicSimpleVFE", voxel_size=base["pts_voxel_layer"]["voxel_size"],
                                     point_cloud_range=base["pts_voxel_layer"]["point_cloud_range"])
    cfg = scaled_model_cfg(base, bev=24, num_proposals=20)
    pts = [torch.from_numpy(synth_points(n, cfg["pts_voxel_layer"]["point_cloud_range"], seed=60 + s))
           for s, n in enumerate((9000, 6500))]
    return cfg, make_state_dict(cfg, 6), pts


def _fusion_tiny():
    """FocalFormer3D_LC: LiDAR tower + image tower + Lift-Splat-Shoot + two 'bevfusion' layers (9x9 local attention)."""
    from focalformer3d_b200.config import load_config, default_config_path, scaled_fusion_cfg
    from focalformer3d_b200.synth import make_state_dict, synth_points, synth_cameras
    hw = (64, 96)
    cfg = scaled_fusion_cfg(load_config(default_config_path("focalformer3d_lc"))["model"], bev=24, img_hw=hw, num_proposals=16)
    pts = [torch.from_numpy(synth_points(n, cfg["pts_voxel_layer"]["point_cloud_range"], seed=30 + s))
           for s, n in enumerate((6000, 5000))]
    img = torch.randn(2, 6, 3, *hw, generator=torch.Generator().manual_seed(7))
    metas = [dict(lidar2img=synth_cameras(6, hw, seed=50 + b)) for b in range(2)]
    return cfg, make_state_dict(cfg, 4), pts, img, metas


@pytest.fixture(scope="module", params=["nuscenes_l", "waymo_l", "deformformer_l", "fusion_lc", "waymo15_l",
                                        "deformformer_l_dynamic"])
def pair(request, tiny_cfg, tiny_sd, tiny_points):
    from focalformer3d_b200.model import build_model
    from oracle.detector import build_oracle
    if request.param == "waymo_l":           # HardVFE, 3 classes, 3 HIP stages, no velocity head, 2 encoder layers
        tiny_cfg, tiny_sd, tiny_points = _waymo_tiny()
    if request.param == "deformformer_l":    # no HIP: single averaged heatmap, one decoder stage, no ROI, no fusion layers
        tiny_cfg, tiny_sd, tiny_points = _deform_tiny()
    kw, okw = {}, {}
    if request.param == "waymo15_l":
        tiny_cfg, tiny_sd, tiny_points = _waymo15_tiny()
    if request.param == "deformformer_l_dynamic":
        tiny_cfg, tiny_sd, tiny_points = _dynamic_tiny()
    if request.param == "fusion_lc":
        tiny_cfg, tiny_sd, tiny_points, img, metas = _fusion_tiny()
        kw, okw = dict(img=img.cuda(), img_metas=metas), dict(img=img, img_metas=metas)
    model = build_model(tiny_cfg)
    model.load_state_dict(tiny_sd, strict=True)
    model.prepare("cuda")
    res, det, st = model.forward_raw([p.cuda() for p in tiny_points], keep_stages=True, **kw)
    torch.cuda.synchronize()
    oracle = build_oracle(tiny_cfg)
    oracle.load_state_dict(tiny_sd, strict=True)
    ost = {}
    ref, rdet = oracle.forward_raw(tiny_points, ost, **okw)
    return dict(model=model, res=res, det=det, st=st, oracle=oracle, ref=ref, rdet=rdet, ost=ost, points=tiny_points,
                name=request.param, kw=kw)


def test_voxel_stage(pair):
    st, ost = pair["st"], pair["ost"]
    n = int(st["vox"]["n_dev"][0].item())
    assert n == ost["coors"].shape[0]
    if pair["name"] == "deformformer_l_dynamic":
        # dynamic voxelisation: same voxel SET (order is implementation-defined), every voxel = mean of ALL its points
        def keyed(c):
            c = c.long()
            return ((c[:, 0] * 64 + c[:, 1]) * 4096 + c[:, 2]) * 4096 + c[:, 3]
        ka, kb = keyed(st["vox"]["coors"][:n].cpu()), keyed(ost["coors"])
        ia, ib = ka.argsort(), kb.argsort()
        assert torch.equal(ka[ia], kb[ib])
        assert (st["vox"]["mean"][:n, :5].cpu()[ia] - ost["voxel_features"][ib]).abs().max().item() < 1e-5
        assert int(st["vox"]["num_points"][:n].max().item()) > 10           # the 10-point cap of the hard voxeliser is gone
        return
    assert torch.equal(st["vox"]["coors"][:n].cpu(), ost["coors"].int())
    assert torch.equal(st["vox"]["num_points"][:n].cpu(), ost["num_points"].int())
    assert torch.equal(st["vox"]["voxels"][:n].cpu(), ost["voxels"])
    nf = ost["voxel_features"].shape[1]                      # 5 (mean VFE) or 64 (HardVFE)
    assert (st["vox"]["mean"][:n, :nf].cpu() - ost["voxel_features"]).abs().max().item() < 1e-4
    assert int(st["overflow"].item()) == 0


def test_sparse_encoder_stage(pair):
    bev, ref = pair["st"]["bev"], pair["ost"]["middle"]          # ours [B,H,W,d*C+c]; reference [B, c*D+d, H, W]
    B, H, W, DC = bev.shape
    D = DC // 128
    ours = bev.view(B, H, W, D, 128).permute(0, 4, 3, 1, 2).reshape(B, DC, H, W).cpu()
    assert torch.equal(ours != 0, ref != 0) or ((ours != 0) != (ref != 0)).float().mean().item() < 1e-4
    _close(ours, ref, "sparse encoder output")


def test_bev_stages(pair):
    st, ost = pair["st"], pair["ost"]
    for i, (a, b) in enumerate(zip(st["backbone"], ost["backbone"])):
        _close(_nchw(a), b, f"SECOND stage {i}")
    _close(_nchw(st["neck"]), ost["neck"], "SECONDFPN")
    _close(_nchw(st["conv_feat"]), ost["conv_feat"], "shared conv")
    feats = ost["stage_feats"]
    for a, b in zip(st["stage_feats"], feats[:-1]):
        _close(_nchw(a), b, "FocalEncoder stage feature")
    _close(_nchw(st["extra"]), feats[-1], "FocalEncoder extra feature")
    if pair["name"] == "fusion_lc":
        for i, (a, b) in enumerate(zip(st["cam"]["img_backbone"], ost["img_backbone"])):
            _close(_nchw(a), b, f"ResNet-50 stage {i}")
        _close(_nchw(st["cam"]["img_feat"]), ost["img_feat"], "FPN level 0")
        _close(_nchw(st["cam"]["img_bev"]), pair["oracle"].imgpts_neck.debug["img_bev"], "camera BEV (Lift-Splat-Shoot)")


def test_hip_stage_outputs(pair):
    res, dbg = pair["res"], pair["oracle"].pts_bbox_head.debug
    assert len(res["dense_heatmap"]) == len(pair["ref"]["dense_heatmap"])
    for a, b in zip(res["dense_heatmap"], pair["ref"]["dense_heatmap"]):
        assert (a.cpu() - b).abs().max().item() < TOL
        assert (a.cpu().sigmoid() - b.sigmoid()).abs().max().item() < TOL
    for s, (top, otop) in enumerate(zip(res["_top_proposals"], dbg["top_proposals"])):
        for b in range(top.shape[0]):
            assert set(top[b].cpu().tolist()) == set(otop[b].tolist()), f"HIP stage {s} scene {b}: top-k sets differ"
    dense = [d.cpu() for d in res["dense_heatmap"]]
    single = len(res["_top_proposals"]) == 1 and len(dense) == 2          # DeformFormer: averaged sigmoids (:547-549)
    heat = (dense[0].sigmoid() + dense[1].sigmoid()) / 2 if single else dense[-1].sigmoid()
    _nms_maps_agree(res["_nms_heatmap"][-1].cpu().flatten(2), dbg["nms_heatmap"][-1], heat.flatten(2), heat.shape[-1])


def _nms_maps_agree(a, b, heat, W):
    """3x3 local-max maps agree within TOL, except where the pre-NMS heat map has a near-tie (two cells of a 3x3 window
    within 1e-5): fp32 rounding may then pick the other cell as the local maximum."""
    d = (a - b).abs()
    bad = (d >= TOL).nonzero()
    assert bad.shape[0] <= 4, f"{bad.shape[0]} local-max decisions differ"
    H = a.shape[-1] // W
    for bi, c, pos in bad.tolist():
        y, x = pos // W, pos % W
        v = heat[bi, c, pos].item()
        nb = [heat[bi, c, yy * W + xx].item() for yy in range(max(y - 1, 0), min(y + 2, H)) for xx in range(max(x - 1, 0), min(x + 2, W))
              if (yy, xx) != (y, x)]
        assert any(abs(v - t) < 1e-5 for t in nb), f"local-max mismatch at {(bi, c, y, x)} is not a near-tie"


def _match(res, oracle, ref):
    """queries are compared after sorting each HIP stage's proposals by flat index (set semantics)."""
    dbg = oracle.pts_bbox_head.debug
    k = res["_top_proposals"][0].shape[1]
    perm_o, perm_m = [], []
    for s, (top, otop) in enumerate(zip(res["_top_proposals"], dbg["top_proposals"])):
        perm_m.append(top.cpu().long().argsort(1) + s * k)
        perm_o.append(otop.argsort(1) + s * k)
    return torch.cat(perm_m, 1), torch.cat(perm_o, 1)


def test_decoder_outputs(pair):
    res, ref, oracle = pair["res"], pair["ref"], pair["oracle"]
    pm, po = _match(res, oracle, ref)
    nq = pm.shape[1]
    assert torch.equal(res["query_labels"].cpu().gather(1, pm), oracle.pts_bbox_head.query_labels.gather(1, po))
    for key in ("center", "height", "dim", "rot", "vel", "heatmap"):
        if key not in ref:
            assert key == "vel" and key not in res           # Waymo heads have no velocity branch
            continue
        a, b = res[key].cpu(), ref[key]
        n_stage = a.shape[-1] // nq
        for s in range(n_stage):
            aa = a[..., s * nq:(s + 1) * nq].gather(2, pm[:, None].expand(-1, a.shape[1], -1))
            bb = b[..., s * nq:(s + 1) * nq].gather(2, po[:, None].expand(-1, b.shape[1], -1))
            err = (aa - bb).abs().max().item()
            assert err < TOL, f"{key} decoder stage {s}: max abs err {err}"
    nc = ref["query_heatmap_score"].shape[1]
    a = res["query_heatmap_score"].cpu().gather(2, pm[:, None].expand(-1, nc, -1))
    b = ref["query_heatmap_score"].gather(2, po[:, None].expand(-1, nc, -1))
    assert (a - b).abs().max().item() < TOL


def test_final_boxes(pair):
    boxes, scores, labels, keep = (t.cpu() for t in pair["det"])
    pm, po = _match(pair["res"], pair["oracle"], pair["ref"])
    for b, r in enumerate(pair["rdet"]):
        ok = r["keep"]
        assert torch.equal(keep[b][pm[b]].bool(), ok[po[b]])
        sel_m = pm[b][keep[b][pm[b]].bool()]
        # oracle rows after the keep filter are in original order; recover them through the keep mask
        ref_boxes = torch.zeros(ok.shape[0], r["boxes_3d"].shape[1])
        if r["boxes_3d"].shape[0] == int(ok.sum()):
            ref_boxes[ok] = r["boxes_3d"]
            ref_scores = torch.zeros(ok.shape[0]); ref_scores[ok] = r["scores_3d"]
            sel_o = po[b][ok[po[b]]]
            assert (boxes[b][sel_m] - ref_boxes[sel_o]).abs().max().item() < TOL
            assert (scores[b][sel_m] - ref_scores[sel_o]).abs().max().item() < TOL


def test_simple_test_signature(pair):
    tiny_points = pair["points"]
    out = pair["model"].simple_test([p.cuda() for p in tiny_points], **pair["kw"])
    assert len(out) == len(tiny_points)
    for o in out:
        d = o["pts_bbox"]
        assert d["boxes_3d"].shape[1] == (7 if pair["name"].startswith("waymo") else 9) and d["boxes_3d"].shape[0] == d["scores_3d"].shape[0] == d["labels_3d"].shape[0]
        assert d["boxes_3d"].device.type == "cpu" and d["boxes_3d"].shape[0] <= 200


def test_determinism_and_batch_independence(pair):
    """Size-independent properties: same inputs -> bit-identical outputs; a scene's result does not depend on its
    batch neighbours (scenes never exchange data: the basis of scene-level data parallelism)."""
    model, tiny_points = pair["model"], pair["points"]
    kw = pair["kw"]
    r1, d1, _ = model.forward_raw([p.cuda() for p in tiny_points], **kw)
    r2, d2, _ = model.forward_raw([p.cuda() for p in tiny_points], **kw)
    for k in ("center", "dim", "heatmap"):
        if kw:      # camera BEV pooling sums with atomics: the addition order (last bits) is not fixed run to run
            assert (r1[k] - r2[k]).abs().max().item() < 1e-4
        else:
            assert torch.equal(r1[k], r2[k])
    kw1 = dict(img=kw["img"][1:2], img_metas=kw["img_metas"][1:2]) if kw else {}
    r3, d3, _ = model.forward_raw([tiny_points[1].cuda()], **kw1)
    assert torch.equal(r3["_top_proposals"][0][0], r1["_top_proposals"][0][1])
    assert (r3["center"][0] - r1["center"][1]).abs().max().item() < 1e-4


def test_full_size_properties():
    """FocalFormer3D_L at BASELINE.json's full size (bs=2 here to bound test time): shapes, finiteness, unique
    proposals per stage, accumulated-mask exclusion between HIP stages, no capacity overflow."""
    from focalformer3d_b200.config import load_config, default_config_path
    from focalformer3d_b200.synth import make_state_dict, synth_points
    from focalformer3d_b200.model import build_model
    cfg = load_config(default_config_path())["model"]
    model = build_model(cfg)
    model.load_state_dict(make_state_dict(cfg, 0), strict=True)
    model.cuda().prepare("cuda")
    pts = [torch.from_numpy(synth_points(300000, cfg["pts_voxel_layer"]["point_cloud_range"], seed=s)).cuda() for s in range(2)]
    res, (boxes, scores, labels, keep), st = model.forward_raw(pts, keep_stages=True)
    torch.cuda.synchronize()
    assert int(st["overflow"].item()) == 0
    assert res["center"].shape == (2, 2, 1200) and res["heatmap"].shape == (2, 10, 1200)
    for k in ("center", "height", "dim", "rot", "vel", "heatmap"):
        assert torch.isfinite(res[k]).all()
    t0, t1 = res["_top_proposals"]
    for b in range(2):
        s0, s1 = set(t0[b].tolist()), set(t1[b].tolist())
        assert len(s0) == 300 and len(s1) == 300 and not (s0 & s1)      # stage-2 picks exclude stage-1 positives
    n = [int(x.item()) for x in st["level_sizes"]]
    assert n[0] == int(st["vox"]["n_dev"][0].item()) and all(v > 0 for v in n)
    assert boxes.shape == (2, 600, 9) and torch.isfinite(boxes).all()


def test_voxel_cap_parity(tiny_cfg, tiny_sd, tiny_points):
    """max_voxels smaller than the number of occupied voxels: voxels created after the cap are dropped in
    first-appearance order (reference semantics) -- end-to-end parity with the cap active."""
    import copy
    from focalformer3d_b200.model import build_model
    from oracle.detector import build_oracle
    cfg = copy.deepcopy(tiny_cfg)
    cfg["pts_voxel_layer"]["max_voxels"] = (1000, 1500)
    model = build_model(cfg)
    model.load_state_dict(tiny_sd, strict=True)
    model.prepare("cuda")
    res, det, st = model.forward_raw([p.cuda() for p in tiny_points], keep_stages=True)
    oracle = build_oracle(cfg)
    oracle.load_state_dict(tiny_sd, strict=True)
    ost = {}
    ref, _ = oracle.forward_raw(tiny_points, ost)
    n = int(st["vox"]["n_dev"][0].item())
    assert n == ost["coors"].shape[0] == 2 * 1500                       # both scenes hit the cap
    assert torch.equal(st["vox"]["coors"][:n].cpu(), ost["coors"].int())
    pm, po = _match(res, oracle, ref)
    for key in ("center", "dim", "heatmap"):
        a = res[key].cpu()[..., -pm.shape[1]:].gather(2, pm[:, None].expand(-1, res[key].shape[1], -1))
        b = ref[key][..., -po.shape[1]:].gather(2, po[:, None].expand(-1, ref[key].shape[1], -1))
        assert (a - b).abs().max().item() < TOL, key


def test_batch_with_an_empty_scene_runs(tiny_cfg, tiny_sd, tiny_points):
    """A scene with no in-range points contributes no voxels; its queries come from an all-background BEV map.
    (The reference derives the batch size from the last voxel's batch index, focalformer3d.py:167, and would
    mis-size the batch here; we keep the true batch size.)"""
    from focalformer3d_b200.model import build_model
    model = build_model(tiny_cfg)
    model.load_state_dict(tiny_sd, strict=True)
    model.prepare("cuda")
    far = torch.full((50, 5), 1000.0)
    res, (boxes, scores, labels, keep), _ = model.forward_raw([tiny_points[0].cuda(), far.cuda()])
    torch.cuda.synchronize()
    assert boxes.shape[0] == 2 and torch.isfinite(boxes).all() and torch.isfinite(res["heatmap"]).all()
    single, _, _ = model.forward_raw([tiny_points[0].cuda()])
    assert torch.equal(single["_top_proposals"][0][0], res["_top_proposals"][0][0])     # scene 0 is unaffected


def test_waymo_full_size_properties():
    """FocalFormer3D_Waymo_L at its full geometry (1536^2 x 40 voxels, 192^2 BEV, 3 HIP stages x 200)."""
    from focalformer3d_b200.config import load_config, default_config_path
    from focalformer3d_b200.synth import make_state_dict, synth_points
    from focalformer3d_b200.model import build_model
    cfg = load_config(default_config_path("focalformer3d_waymo_l"))["model"]
    model = build_model(cfg)
    model.load_state_dict(make_state_dict(cfg, 0), strict=True)
    model.prepare("cuda")
    pts = [torch.from_numpy(synth_points(180000, cfg["pts_voxel_layer"]["point_cloud_range"], seed=s, n_beams=64, n_sweeps=1)).cuda()
           for s in range(2)]
    res, (boxes, scores, labels, keep), st = model.forward_raw(pts, keep_stages=True)
    torch.cuda.synchronize()
    assert int(st["overflow"].item()) == 0
    assert res["center"].shape == (2, 2, 1200) and res["heatmap"].shape == (2, 3, 1200) and "vel" not in res
    assert boxes.shape == (2, 600, 7) and torch.isfinite(boxes).all()
    tops = [set(t[0].tolist()) for t in res["_top_proposals"]]
    assert all(len(t) == 200 for t in tops) and not (tops[0] & tops[1]) and not (tops[1] & tops[2]) and not (tops[0] & tops[2])


@pytest.mark.parametrize("nms_type", ["circle", "rotate"])
def test_get_bboxes_with_nms(tiny_cfg, tiny_sd, tiny_points, nms_type):
    """test_cfg.nms_type = 'circle' / 'rotate' (focal_decoder.py:1352-1385): the per-task NMS keep flags and the final
    box list of the CUDA head equal the oracle's."""
    import copy
    from focalformer3d_b200.model import build_model
    from oracle.detector import build_oracle
    from oracle import parity
    cfg = copy.deepcopy(tiny_cfg)
    cfg["test_cfg"]["pts"].update(nms_type=nms_type, pre_maxsize=1000, post_maxsize=83)
    model = build_model(cfg)
    model.load_state_dict(tiny_sd, strict=True)
    model.prepare("cuda")
    res, det, _ = model.forward_raw([p.cuda() for p in tiny_points])
    oracle = build_oracle(cfg)
    oracle.load_state_dict(tiny_sd, strict=True)
    ref, rdet = oracle.forward_raw(tiny_points)
    rep = parity.head_report(res, det, oracle.pts_bbox_head, ref, rdet)
    assert rep["topk_sets_equal"] and rep["keep_equal"] and rep["box_labels_equal"], rep
    assert rep["max_abs"]["boxes"] < TOL and rep["max_abs"]["scores"] < TOL, rep
    base = build_model(tiny_cfg)
    base.load_state_dict(tiny_sd, strict=True)
    base.prepare("cuda")
    _, det0, _ = base.forward_raw([p.cuda() for p in tiny_points])
    assert int(det[3].sum()) <= int(det0[3].sum())                       # NMS only ever removes boxes
