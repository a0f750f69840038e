"""One tiny FocalFormer3D_L forward (BEV 32, 2 scenes x 8000 points) -- the target command for compute-sanitizer
(memcheck / racecheck / synccheck) and other slow tools.  Usage: tiny_forward.py [config] [n_forward]."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
from focalformer3d_b200.synth import make_state_dict, synth_points
from focalformer3d_b200.model import build_model

name = sys.argv[1] if len(sys.argv) > 1 else "focalformer3d_l"
n_fwd = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = scaled_model_cfg(load_config(default_config_path(name))["model"], bev=32, num_proposals=24)
model = build_model(cfg)
model.load_state_dict(make_state_dict(cfg, 0), strict=True)
model.prepare("cuda")
pts = [torch.from_numpy(synth_points(n, cfg["pts_voxel_layer"]["point_cloud_range"], seed=s)).cuda()
       for s, n in enumerate((8000, 6000))]
for _ in range(n_fwd):
    res, det, _ = model.forward_raw(pts)
torch.cuda.synchronize()
print("tiny forward ok", float(res["heatmap"].abs().sum()))
