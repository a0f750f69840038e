"""Run a few full-size FocalFormer3D_L forwards (bs=4) -- the target command for ncu captures."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from focalformer3d_b200.config import load_config, default_config_path
from focalformer3d_b200.synth import make_state_dict, synth_points
from focalformer3d_b200.model import build_model

n_fwd = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = load_config(default_config_path())["model"]
model = build_model(cfg)
model.load_state_dict(make_state_dict(cfg, 0), strict=True)
model.prepare("cuda")          # weights are packed on the host and copied once: no torch kernels before the forward
pts = [torch.from_numpy(synth_points(300000, cfg["pts_voxel_layer"]["point_cloud_range"], seed=s)).cuda() for s in range(4)]
for _ in range(n_fwd):
    model.forward_raw(pts)
torch.cuda.synchronize()
print("done")
