"""Parity at the sizes the benchmark is quoted on (BASELINE.json configs[1..3]) -- the CUDA path against the CPU oracle on
the same seeded synthetic scenes, stage by stage: voxel coordinates exact, sparse-encoder BEV, per-stage top-k SETS
exact (focal_decoder.py:688 set semantics), class ids exact, heads within 1e-3 abs, keep masks / boxes
(focal_decoder.py:1313-1413).  Plus the direct CUDA <-> real-reference golden fixture for the LiDAR flagship.

Shapes that only exist at full size are exercised here: 128-channel 256-row sparse tiles, 2^21-slot hashes, the 3x level-2
capacity growth, roi_mlp K = 18816 at 2400 rows, the 324 000-way radix select with > 16 384 contenders."""
import os
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _run_pair(cfg, sd, pts, **kw):
    from focalformer3d_b200.model import build_model
    from oracle.detector import build_oracle
    model = build_model(cfg)
    model.load_state_dict(sd, strict=True)
    model.prepare("cuda")
    gkw = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
    res, det, st = model.forward_raw([p.cuda() for p in pts], keep_stages=True, **gkw)
    torch.cuda.synchronize()
    assert int(st["overflow"].item()) == 0
    oracle = build_oracle(cfg)
    oracle.load_state_dict(sd, strict=True)
    ost = {}
    ref, rdet = oracle.forward_raw(pts, ost, **kw)
    return model, res, det, st, oracle, ref, rdet, ost


def _check(res, det, st, oracle, ref, rdet, ost, min_scenes, check_encoder=True, logit_tol=TOL):
    from oracle import parity
    srep = parity.stage_report(st, ost)
    hrep = parity.head_report(res, det, oracle.pts_bbox_head, ref, rdet)
    print("stage parity:", srep)
    print("head parity:", hrep)
    assert srep["voxel_coors_equal"] and srep.get("voxel_num_points_equal", True)
    assert srep["voxel_feature_max_abs"] < 1e-4
    # intermediate feature maps: rounding-level differences, amplified by the 21 stacked sparse convs (values up to ~1e3).
    # Against a float64 run of the oracle the CUDA path is within 2e-3 there and the fp32 oracle itself within 2e-4
    # (tests/test_gpu_accuracy.py); the north-star bars (1e-3 abs) apply to the heads and heat maps below.
    assert srep["sparse_bev_support_mismatch_frac"] < 1e-4 and srep["sparse_bev_mixed_err"] < 5 * TOL
    for k in ("second_mixed_err", "secondfpn_mixed_err", "shared_conv_mixed_err") + (("focal_encoder_mixed_err",) if check_encoder else ()):
        assert srep[k] < TOL, (k, srep[k])
    # top-k sets: exact, except swaps across a near-tie of the k-th value (reported, bounded, never "unexplained")
    assert hrep["topk_unexplained"] == 0, hrep
    assert hrep["topk_near_tie_swaps"] <= 2, hrep
    assert hrep["scenes_compared"] >= min_scenes, hrep
    assert hrep["labels_equal"] and hrep["keep_equal"] and hrep["box_labels_equal"], hrep
    assert hrep["max_abs"]["dense_heatmap"] < logit_tol and hrep["max_abs"]["dense_heatmap_sigmoid"] < TOL, hrep
    assert hrep["max_abs_heads"] < TOL and hrep["max_abs"]["query_heatmap_score"] < TOL, hrep
    assert hrep["max_abs"]["boxes"] < TOL and hrep["max_abs"]["scores"] < TOL, hrep
    return srep, hrep


def test_focalformer3d_l_full_size_parity():
    """BASELINE.json configs[1]: FocalFormer3D_L, bs = 4, 300k-point 10-sweep clouds, seeds 0-3 (the bench's scenes)."""
    from focalformer3d_b200.config import load_config, default_config_path
    from focalformer3d_b200.synth import make_state_dict, synth_points
    cfg = load_config(default_config_path())["model"]
    sd = make_state_dict(cfg, 0)
    pts = [torch.from_numpy(synth_points(300000, cfg["pts_voxel_layer"]["point_cloud_range"], seed=s)) for s in range(4)]
    model, res, det, st, oracle, ref, rdet, ost = _run_pair(cfg, sd, pts)
    _check(res, det, st, oracle, ref, rdet, ost, min_scenes=3)
    assert res["center"].shape == (4, 2, 1200)
    n = [int(x.item()) for x in st["level_sizes"]]
    assert n[0] > 400000 and n[-2] > 100000          # full-size levels really were exercised


def test_waymo_l_full_size_parity():
    """BASELINE.json configs[3]: FocalFormer3D_Waymo_L, 64-beam 180k-point clouds, 1536^2 x 40 voxels, 192^2 BEV."""
    from focalformer3d_b200.config import load_config, default_config_path
    from focalformer3d_b200.synth import make_state_dict, synth_points
    cfg = load_config(default_config_path("focalformer3d_waymo_l"))["model"]
    sd = make_state_dict(cfg, 0)
    pts = [torch.from_numpy(synth_points(180000, cfg["pts_voxel_layer"]["point_cloud_range"], seed=s, n_beams=64, n_sweeps=1))
           for s in range(2)]
    model, res, det, st, oracle, ref, rdet, ost = _run_pair(cfg, sd, pts)
    _check(res, det, st, oracle, ref, rdet, ost, min_scenes=1)
    assert res["heatmap"].shape == (2, 3, 1200) and "vel" not in res


def test_fusion_lc_full_size_parity():
    """BASELINE.json configs[2]: FocalFormer3D_LC, bs = 2, 300k-point clouds + 6 x 448 x 800 images per scene (the size
    the shipped ScaleImageMultiViewImage hands the model)."""
    from focalformer3d_b200.config import load_config, default_config_path
    from focalformer3d_b200.synth import make_state_dict, synth_points, synth_cameras
    from oracle import parity
    cfg = load_config(default_config_path("focalformer3d_lc"))["model"]
    sd = make_state_dict(cfg, 0)
    H, W = cfg["imgpts_neck"]["img_scale"]
    B = 2
    pts = [torch.from_numpy(synth_points(300000, cfg["pts_voxel_layer"]["point_cloud_range"], seed=s)) for s in range(B)]
    img = torch.randn(B, 6, 3, H, W, generator=torch.Generator().manual_seed(0))
    metas = [dict(lidar2img=synth_cameras(6, (H, W), seed=b)) for b in range(B)]
    model, res, det, st, oracle, ref, rdet, ost = _run_pair(cfg, sd, pts, img=img, img_metas=metas)
    # The fused encoder (two 9x9 local-attention layers on features of magnitude ~100) amplifies rounding: against a float64
    # run the fp32 ORACLE itself is off by 3.6e-3 (mixed) on the last stage feature and 1.2e-3 on its heat-map logits
    # (tests/test_gpu_accuracy.py, gpurun_out/accuracy_vs_fp64_focalformer3d_lc_f16.json), so those two intermediate maps
    # get bars at that scale; the heat-map PROBABILITIES, top-k sets, class ids, heads and boxes keep the 1e-3 bar.
    srep, hrep = _check(res, det, st, oracle, ref, rdet, ost, min_scenes=1, check_encoder=False, logit_tol=2e-2)
    assert srep["focal_encoder_mixed_err"] < 5e-2, srep
    cam_err = ((st["cam"]["img_bev"].permute(0, 3, 1, 2).cpu() - oracle.imgpts_neck.debug["img_bev"]).abs()
               / (1.0 + oracle.imgpts_neck.debug["img_bev"].abs())).max().item()
    assert cam_err < TOL, f"camera BEV (Lift-Splat-Shoot) mixed err {cam_err}"


def test_cuda_lidar_flagship_matches_reference_golden():
    """CUDA FocalEncoder ('bevfusionmb2') + FocalDecoder + get_bboxes against tests/golden/focalformer3d_l_intree.pt, the
    outputs of the REAL reference modules (focal_encoder.py:171-222, focal_decoder.py:522-992,1313-1413)."""
    from focalformer3d_b200 import ops
    from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
    from focalformer3d_b200.synth import make_state_dict
    from focalformer3d_b200.model import build_model
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "focalformer3d_l_intree.pt")
    gold = torch.load(path, map_location="cpu")
    cfg = scaled_model_cfg(load_config(default_config_path())["model"], bev=gold["cfg_bev"],
                           num_proposals=gold["cfg_num_proposals"])
    model = build_model(cfg)
    model.load_state_dict(make_state_dict(cfg, seed=gold["weights_seed"]), strict=True)
    model.prepare("cuda")
    head = model.pts_bbox_head

    def nhwc(t):
        return t.permute(0, 2, 3, 1).contiguous().cuda()

    def nchw(t):
        return t.permute(0, 3, 1, 2).cpu()
    # --- encoder
    neck = nhwc(gold["enc_in"])
    B, H, W, _ = neck.shape
    geom = ops.LevelGeom([(H >> l, W >> l) for l in range(head.n_levels)])
    ms_value = torch.empty((B, geom.n_tokens, head.hc), device="cuda")
    conv_feat, stages, extra = model.imgpts_neck(neck, extra_out=ms_value[:, :H * W].view(B, H, W, head.hc))
    assert (nchw(conv_feat) - gold["enc_conv_feat"]).abs().max().item() < TOL
    for a, b in zip(stages + [extra], gold["enc_stage_feats"]):
        assert ((nchw(a) - b).abs() / (1 + b.abs())).max().item() < TOL
    # --- head on the reference's own inputs
    for tag in ("b2", "b1"):
        f = gold[f"head_{tag}_in"]
        cf, sf, ex = nhwc(f[0]), nhwc(f[1]), nhwc(f[2])
        B = cf.shape[0]
        ms_value = torch.empty((B, geom.n_tokens, head.hc), device="cuda")
        ms_value[:, :H * W].view(B, H, W, head.hc).copy_(ex)
        res = head(cf, [sf], ms_value, geom)
        ref = gold[f"head_{tag}_out"]
        nq = res["query_labels"].shape[1]
        k = nq // len(res["_top_proposals"])
        # the fixture's query order is torch.topk's; ours is canonical: match through (label, flat index) sets
        ref_lab = gold[f"head_{tag}_query_labels"]
        ours_idx = torch.cat([t.cpu().long() for t in res["_top_proposals"]], 1)               # [B, nq] class*HW + pos
        # reference positions are recoverable from its query labels and ... only as sets of labels per stage: compare
        # the label multiset per stage, then the sorted regression outputs through the canonical per-stage ordering
        for s in range(nq // k):
            for b in range(B):
                a = sorted((ours_idx[b, s * k:(s + 1) * k] // (H * W)).tolist())
                r = sorted(ref_lab[b, s * k:(s + 1) * k].tolist())
                assert a == r, f"{tag}: stage {s} scene {b} class ids differ from the reference's"
        for a, b in zip(res["dense_heatmap"], ref["dense_heatmap"]):
            assert (a.cpu() - b).abs().max().item() < TOL
        # per-query outputs: match queries by their (unique) query_heatmap_score column + label
        qs_o, qs_m = ref["query_heatmap_score"], res["query_heatmap_score"].cpu()
        for b in range(B):
            for s in range(nq // k):
                sl = slice(s * k, (s + 1) * k)
                key_o = qs_o[b, :, sl].max(0).values + ref_lab[b, sl].float() * 10
                key_m = qs_m[b, :, sl].max(0).values + res["query_labels"][b, sl].cpu().float() * 10
                io, im = key_o.argsort(), key_m.argsort()
                assert (key_o[io] - key_m[im]).abs().max().item() < TOL
                for key in ("center", "height", "dim", "rot", "vel", "heatmap"):
                    n_stage = ref[key].shape[-1] // nq
                    for ds in range(n_stage):
                        ao = ref[key][b, :, ds * nq:(ds + 1) * nq][:, sl][:, io]
                        am = res[key][b, :, ds * nq:(ds + 1) * nq][:, sl].cpu()[:, im]
                        assert (ao - am).abs().max().item() < TOL, (tag, key, ds)
