"""Oracle-side parity report: product outputs (libff3d.so path) vs the CPU oracle on the same inputs.

TEST INFRASTRUCTURE (see oracle/__init__.py): imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg only.  Everything here is a *comparison*; nothing computes anything for the product.

Bars (BASELINE.json north_star): per-HIP-stage top-k query index SETS bit-exact (torch.topk(sorted=False) at
focal_decoder.py:688 leaves the order undefined), class ids bit-exact, heatmaps and box regressions within 1e-3 abs.
A top-k set may legitimately differ only at a near-tie of the k-th value (two candidates within ``tie_eps`` of each
other across the cut): such cases are reported separately (``topk_near_tie_swaps``), never silently accepted as equal.
"""
import torch

TOL = 1e-3
HEAD_KEYS = ("center", "height", "dim", "rot", "vel", "heatmap")


def _cpu(t):
    return t.detach().cpu() if isinstance(t, torch.Tensor) else t


def topk_report(res, dbg, tie_eps=2e-5):
    """Per stage and scene: product top-k set vs oracle top-k set.  Returns (stages list, all_equal, swaps, worst_gap).
    A mismatching element counts as a near-tie swap when its ORACLE post-NMS heat value is within tie_eps of the oracle's
    k-th value."""
    stages, all_equal, swaps, unexplained = [], True, 0, 0
    for s, (top, otop) in enumerate(zip(res["_top_proposals"], dbg["top_proposals"])):
        top, otop = _cpu(top).long(), otop.long()
        heat = dbg["nms_heatmap"][s] if s < len(dbg["nms_heatmap"]) else dbg["nms_heatmap"][-1]   # [B, nc, H*W]
        flat = heat.reshape(heat.shape[0], -1)
        per_scene = []
        for b in range(top.shape[0]):
            sm, so = set(top[b].tolist()), set(otop[b].tolist())
            diff = sorted(sm ^ so)
            ok = not diff
            if diff:
                all_equal = False
                kth = flat[b][otop[b]].min().item()
                near = [abs(flat[b][i].item() - kth) <= tie_eps for i in diff]
                swaps += sum(near) // 2
                unexplained += len(diff) - sum(near)
            per_scene.append(ok)
        stages.append(per_scene)
    return stages, all_equal, swaps, unexplained


def _perms(res, dbg):
    """queries matched through each HIP stage's proposals sorted by flat index (set semantics)."""
    k = res["_top_proposals"][0].shape[1]
    pm, po = [], []
    for s, (top, otop) in enumerate(zip(res["_top_proposals"], dbg["top_proposals"])):
        pm.append(_cpu(top).long().argsort(1) + s * k)
        po.append(otop.argsort(1) + s * k)
    return torch.cat(pm, 1), torch.cat(po, 1)


def head_report(res, det, oracle_head, ref, rdet):
    """Compare one forward.  res/det: product head dict + (boxes, scores, labels, keep); oracle_head: the oracle's
    FocalDecoder after its forward (debug + query_labels); ref: oracle head dict; rdet: oracle get_bboxes list.
    Scenes whose top-k sets differ are excluded from the per-query comparisons (their query sets differ by
    construction) and counted in ``scenes_compared``."""
    dbg = oracle_head.debug
    stages, all_equal, swaps, unexplained = topk_report(res, dbg)
    B = res["_top_proposals"][0].shape[0]
    scene_ok = [all(st[b] for st in stages) for b in range(B)]
    pm, po = _perms(res, dbg)
    nq = pm.shape[1]
    out = {"topk_sets_equal": bool(all_equal), "topk_near_tie_swaps": int(swaps), "topk_unexplained": int(unexplained),
           "scenes": B, "scenes_compared": int(sum(scene_ok)), "max_abs": {}}
    sel = [b for b in range(B) if scene_ok[b]]
    dense = [(a, b) for a, b in zip(res["dense_heatmap"], ref["dense_heatmap"])]
    out["max_abs"]["dense_heatmap"] = max((_cpu(a) - b).abs().max().item() for a, b in dense) if dense else 0.0
    out["max_abs"]["dense_heatmap_sigmoid"] = max((_cpu(a).sigmoid() - b.sigmoid()).abs().max().item() for a, b in dense) if dense else 0.0
    if not sel:
        out["labels_equal"] = False
        return out
    idx = torch.tensor(sel)
    pm_s, po_s = pm[idx], po[idx]
    lab_m = _cpu(res["query_labels"])[idx].gather(1, pm_s)
    lab_o = oracle_head.query_labels[idx].gather(1, po_s)
    out["labels_equal"] = bool(torch.equal(lab_m, lab_o))
    worst = 0.0
    for key in HEAD_KEYS:
        if key not in ref:
            continue
        a, b = _cpu(res[key])[idx], ref[key][idx]
        n_stage = a.shape[-1] // nq
        err = 0.0
        for s in range(n_stage):
            aa = a[..., s * nq:(s + 1) * nq].gather(2, pm_s[:, None].expand(-1, a.shape[1], -1))
            bb = b[..., s * nq:(s + 1) * nq].gather(2, po_s[:, None].expand(-1, b.shape[1], -1))
            err = max(err, (aa - bb).abs().max().item())
        out["max_abs"][key] = err
        worst = max(worst, err)
    nc = ref["query_heatmap_score"].shape[1]
    a = _cpu(res["query_heatmap_score"])[idx].gather(2, pm_s[:, None].expand(-1, nc, -1))
    b = ref["query_heatmap_score"][idx].gather(2, po_s[:, None].expand(-1, nc, -1))
    out["max_abs"]["query_heatmap_score"] = (a - b).abs().max().item()
    out["max_abs_heads"] = worst
    # final boxes (get_bboxes + coder decode): keep mask exact; then the reference's own output list (score-truncated to
    # 200 boxes, focal_decoder.py:1395-1400) against ours truncated the same way, matched box by box
    if det is not None and rdet is not None:
        boxes, scores, labels, keep = (_cpu(t) for t in det)
        keep_equal, box_err, score_err, labels_ok, n_boxes = True, 0.0, 0.0, True, 0
        for b in sel:
            r = rdet[b]
            ok = r["keep"]
            km = keep[b][pm[b]].bool()
            keep_equal &= bool(torch.equal(km, ok[po[b]]))
            m = keep[b].bool()
            bx, sc, lb = boxes[b][m], scores[b][m], labels[b][m]
            if bx.shape[0] > 200:
                inds = sc.argsort(descending=True, stable=True)[:200]
                bx, sc, lb = bx[inds], sc[inds], lb[inds]
            rb_, rs_, rl_ = r["boxes_3d"], r["scores_3d"], r["labels_3d"].int()
            if bx.shape[0] != rb_.shape[0]:
                keep_equal = False
                continue
            if rb_.shape[0] == 0:
                continue
            # nearest match in (box, score) space: the 200-cut and the order may differ only between equal scores
            d = torch.cdist(torch.cat([rb_, rs_[:, None]], 1).double(), torch.cat([bx, sc[:, None]], 1).double(), p=float("inf"))
            best, arg = d.min(1)
            box_err = max(box_err, best.max().item())
            score_err = max(score_err, (rs_ - sc[arg]).abs().max().item())
            labels_ok &= bool(torch.equal(rl_, lb[arg].int()))
            n_boxes += int(rb_.shape[0])
        out["keep_equal"], out["box_labels_equal"], out["boxes_compared"] = bool(keep_equal), bool(labels_ok), n_boxes
        out["max_abs"]["boxes"], out["max_abs"]["scores"] = box_err, score_err
    return out


def stage_report(st, ost, n_feat=None):
    """Stage-by-stage comparison of the LiDAR tower: voxel coordinates exact, sparse-encoder BEV / SECOND / FPN / encoder
    features as max mixed error |a-b| / (1+|b|)."""
    def mixed(a, b):
        return ((a - b).abs() / (1.0 + b.abs())).max().item()

    def nchw(t):
        return _cpu(t).permute(0, 3, 1, 2)
    out = {}
    if "vox" in st and st["vox"] is not None and "coors" in ost:
        n = int(st["vox"]["n_dev"][0].item())
        out["n_voxels"] = n
        out["voxel_coors_equal"] = bool(n == ost["coors"].shape[0] and torch.equal(_cpu(st["vox"]["coors"][:n]), ost["coors"].int()))
        if out["voxel_coors_equal"] and ost.get("num_points") is not None:
            out["voxel_num_points_equal"] = bool(torch.equal(_cpu(st["vox"]["num_points"][:n]), ost["num_points"].int()))
            nf = ost["voxel_features"].shape[1]
            out["voxel_feature_max_abs"] = (_cpu(st["vox"]["mean"][:n, :nf]) - ost["voxel_features"]).abs().max().item()
    if "bev" in st and "middle" in ost:
        bev, ref = st["bev"], ost["middle"]
        B, H, W, DC = bev.shape
        C = 128
        D = DC // C
        ours = _cpu(bev).view(B, H, W, D, C).permute(0, 4, 3, 1, 2).reshape(B, DC, H, W)
        out["sparse_bev_support_mismatch_frac"] = ((ours != 0) != (ref != 0)).float().mean().item()
        out["sparse_bev_mixed_err"] = mixed(ours, ref)
    if "backbone" in st and "backbone" in ost:
        out["second_mixed_err"] = max(mixed(nchw(a), b) for a, b in zip(st["backbone"], ost["backbone"]))
    if "neck" in st and "neck" in ost:
        out["secondfpn_mixed_err"] = mixed(nchw(st["neck"]), ost["neck"])
    if "conv_feat" in st and "conv_feat" in ost:
        out["shared_conv_mixed_err"] = mixed(nchw(st["conv_feat"]), ost["conv_feat"])
        feats = ost["stage_feats"]
        errs = [mixed(nchw(a), b) for a, b in zip(st["stage_feats"], feats[:-1])]
        if st.get("extra") is not None:
            errs.append(mixed(nchw(st["extra"]), feats[-1]))
        if errs:
            out["focal_encoder_mixed_err"] = max(errs)
    return out


def passes(rep, tol=TOL, min_scenes=None):
    """The pass criterion used by the tests and the bench line."""
    min_scenes = rep["scenes"] if min_scenes is None else min_scenes
    ok = rep["topk_unexplained"] == 0 and rep["scenes_compared"] >= min_scenes and rep.get("labels_equal", False)
    ok = ok and rep.get("max_abs_heads", 1.0) < tol and rep["max_abs"]["dense_heatmap_sigmoid"] < tol
    ok = ok and rep.get("keep_equal", True) and rep["max_abs"].get("boxes", 0.0) < tol
    return bool(ok)
