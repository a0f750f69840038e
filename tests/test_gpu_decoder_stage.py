"""The one-launch decoder stage (csrc/decstage.cu: 3 decoder layers + prediction heads, query block resident in shared
memory, mma.sync hi/lo split, K / V exchanged through global memory in fragment order) against the unfused path (one
launch per projection / attention / LayerNorm, the path every parity test of round 1 pinned to the oracle), on the same
model and inputs.  Query counts that do not fill the 32-row blocks and scenes with several blocks are both covered."""
import pytest
import torch

pytestmark = pytest.mark.gpu

KEYS = ("center", "height", "dim", "rot", "vel", "heatmap")


def _run(cfg_name, num_proposals, n_scenes, seed, bev=24, **synth_kw):
    from focalformer3d_b200 import ops
    from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
    from focalformer3d_b200.model import build_model
    from focalformer3d_b200.synth import make_state_dict, synth_points
    cfg = scaled_model_cfg(load_config(default_config_path(cfg_name))["model"], bev=bev, num_proposals=num_proposals)
    pts = [torch.from_numpy(synth_points(5000 + 700 * s, cfg["pts_voxel_layer"]["point_cloud_range"], seed=seed + s, **synth_kw)).cuda()
           for s in range(n_scenes)]
    model = build_model(cfg)
    model.load_state_dict(make_state_dict(cfg, seed), strict=True)
    assert ops.FUSED_DECODER, "FF3D_FUSED_DECODER=0 in the environment: nothing to compare"
    model.prepare("cuda")
    assert all("fused" in st for st in model.pts_bbox_head.pk["stage"]), "the fused stage was not prepared for this config"
    out = {}
    try:
        for fused in (True, False):
            ops.FUSED_DECODER = fused
            res, det = model.forward_raw(pts)[:2]
            torch.cuda.synchronize()
            out[fused] = ({k: v.clone() for k, v in res.items() if k in KEYS}, [t.clone() for t in det[:2]])
    finally:
        ops.FUSED_DECODER = True
    model.check_flags()
    return out


@pytest.mark.parametrize("cfg_name,num_proposals,n_scenes,kw", [
    ("focalformer3d_l", 16, 2, {}),             # nq = 32: exactly one block per scene
    ("focalformer3d_l", 50, 3, {}),             # nq = 100: four blocks, the last with 4 queries
    ("focalformer3d_waymo_l", 30, 2, dict(n_beams=64)),   # three HIP stages (nq = 90), no velocity head
    ("deformformer3d_l", 40, 2, {}),            # single-stage head: one decoder stage, nq = 40
    ("focalformer3d_l", 300, 8, {}),            # nq = 600 x 8 scenes: 19 blocks per scene -> two cooperative launches (7 + 1 scenes)
])
def test_fused_stage_matches_unfused(cfg_name, num_proposals, n_scenes, kw):
    out = _run(cfg_name, num_proposals, n_scenes, seed=11, **kw)
    (fa, da), (fb, db) = out[True], out[False]
    for k in fa:
        err = (fa[k] - fb[k]).abs().max().item()
        assert err < 2e-4, f"{k}: fused vs unfused max abs {err}"
    assert (da[0] - db[0]).abs().max().item() < 2e-4 and (da[1] - db[1]).abs().max().item() < 2e-5
