"""Full-size camera-only (DeformFormer3D_C_R50) or LiDAR + camera (FocalFormer3D_LC) forwards on one B200: per-stage ms
and per-layer achieved TFLOP/s (fp32-equivalent) from CUDA events.  Usage: profile_camera.py [batch] [reps] [config].
Writes gpurun_out/<config>_profile.json; also an ncu target."""
import json
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from focalformer3d_b200 import ops
from focalformer3d_b200.config import load_config, default_config_path
from focalformer3d_b200.synth import make_state_dict, synth_cameras, synth_points
from focalformer3d_b200.model import build_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
name = sys.argv[3] if len(sys.argv) > 3 else "deformformer3d_c_r50"
cfg = load_config(default_config_path(name))["model"]
model = build_model(cfg)
model.load_state_dict(make_state_dict(cfg, 0), strict=True)
model.prepare("cuda")
H, W = cfg["imgpts_neck"]["img_scale"]
img = torch.randn(B, 6, 3, H, W, generator=torch.Generator().manual_seed(0)).pin_memory()
metas = [dict(lidar2img=synth_cameras(6, (H, W), seed=b)) for b in range(B)]
dimg = img.cuda()
pts = None
if cfg.get("input_pts", True):
    pts = [torch.from_numpy(synth_points(300000, cfg["pts_voxel_layer"]["point_cloud_range"], seed=s)).cuda() for s in range(B)]
for _ in range(3):
    model.forward_raw(pts, img=dimg, img_metas=metas)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    model.forward_raw(pts, img=dimg, img_metas=metas)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
e0.record()
for _ in range(reps):
    model.simple_test(pts, img_metas=metas, img=img.cuda(non_blocking=True))
e1.record()
torch.cuda.synchronize()
ms_e2e = e0.elapsed_time(e1) / reps
stage = {}
ops.prof.start(records=False)
for _ in range(reps):
    model.forward_raw(pts, img=dimg, img_metas=metas)
torch.cuda.synchronize()
ops.prof.stop()
for k, v in ops.prof.stage_ms().items():
    stage[k] = round(v / reps, 3)
ops.prof.start(records=True)
model.forward_raw(pts, img=dimg, img_metas=metas)
torch.cuda.synchronize()
ops.prof.stop()
layers = {}
for label, s, e, flops, nbytes, n_dev, meta in ops.prof.records:
    d = layers.setdefault(label, dict(n=0, ms=0.0, flops=0.0, bytes=0.0))
    d["n"] += 1
    d["ms"] += s.elapsed_time(e)
    d["flops"] += flops or 0.0
    d["bytes"] += nbytes or 0.0
top = sorted(layers.items(), key=lambda kv: -kv[1]["ms"])[:25]
out = dict(config=name, batch=B, ms_per_forward=round(ms, 3), frames_per_s=round(B * 1000.0 / ms, 2),
           ms_per_forward_e2e=round(ms_e2e, 3), stage_ms=stage,
           top_layers=[dict(label=k, n=v["n"], ms=round(v["ms"], 3), tflops=round(v["flops"] / max(v["ms"], 1e-6) / 1e9, 1),
                            gbps=round(v["bytes"] / max(v["ms"], 1e-6) / 1e6, 1)) for k, v in top])
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/{name}_profile.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out))
