"""GPU stand-in for "the reference's own GPU build" (BASELINE.md section 4): the oracle restatement run on the B200 through
stock PyTorch kernels (cuDNN / cuBLAS, torch index ops for the spconv-v1 style rulebook: gather + mm + scatter-add per
offset).  The reference itself needs mmcv / mmdet3d / spconv, which are not installable offline.  Voxelisation stays
on the host (numpy), like a data-loader stage, and is excluded from the timed region."""
import json
import os
import sys
import time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from focalformer3d_b200.config import load_config, default_config_path
from focalformer3d_b200.synth import make_state_dict, synth_points
from oracle.detector import build_oracle

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = load_config(default_config_path())["model"]
o = build_oracle(cfg)
o.load_state_dict(make_state_dict(cfg, 0), strict=True)
o.cuda()
pts = [torch.from_numpy(synth_points(300000, cfg["pts_voxel_layer"]["point_cloud_range"], seed=s)) for s in range(bs)]
out = {}
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    vox = o.voxelize(pts)
    vox = tuple(t.cuda() for t in vox)

    def fwd():
        vf = o.pts_voxel_encoder(*vox)
        x = o.pts_middle_encoder(vf, vox[2], bs)
        x = o.pts_neck(o.pts_backbone(x))
        _, new = o.imgpts_neck(None, x[0], None)
        outs = o.pts_bbox_head([new[0], list(new[1])], None, None)
        return o.pts_bbox_head.get_bboxes(outs)

    with torch.no_grad():
        for _ in range(2):
            fwd()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fwd()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out["tf32_on" if tf32 else "fp32"] = {"ms_per_step": ms, "scenes_per_s": bs / ms * 1e3}
print(json.dumps({"gpu_standin": "oracle restatement on stock PyTorch CUDA kernels (voxelisation excluded)", "bs": bs,
                  "steps": steps, **out}))
