"""Input side of the path on the GPU: the reference's CPU data pipeline between the files and
``simple_test``, as device kernels (csrc/preprocess.cu) so that an 8-GPU box is not bound by host pre-processing.

``assemble_sweeps``   [upstream] mmdet3d v0.17.1 LoadPointsFromMultiSweeps (test mode) + PointsRangeFilter as configured at
                      projects/configs/focalformer3d/FocalFormer3D_L.py:100-111
``preprocess_images`` LoadMultiViewImageFromFiles(to_float32) + ScaleImageMultiViewImage + NormalizeMultiviewImage +
                      PadMultiViewImage + DefaultFormatBundle3D (datasets/pipelines/transform_3d.py:125-249,
                      FocalFormer3D_LC.py:84-97), incl. the lidar2img rescale of ScaleImageMultiViewImage (:239-246)
"""
import ctypes as C
import numpy as np
import torch

from . import lib as L
from .lib import lib, check
from .runtime import PAD_VALUE


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def assemble_sweeps(key_points, sweeps, timestamp, sweeps_num=10, close_radius=1.0, point_range=None, device="cuda", out=None):
    """key_points: [N0, >=4] float32 (host or device) key-frame cloud; sweeps: list of dict(points [Mi, >=4] float32,
    sensor2lidar_rotation [3,3], sensor2lidar_translation [3], timestamp (microseconds)) as the nuScenes info files hold
    them; timestamp: key-frame time in seconds.  Returns [N0 + sum Mi, 5] float32 on the device: (x, y, z, intensity, time
    lag); removed points (close to the sensor, outside ``point_range``) are overwritten with an out-of-range pad value and
    discarded later by the voxeliser (same surviving points, same order, fixed size)."""
    sweeps = list(sweeps)[:sweeps_num]                       # test mode: the first sweeps_num sweeps
    clouds = [key_points] + [s["points"] for s in sweeps]
    n_feat = int(clouds[0].shape[1])
    offs = [0]
    for c in clouds:
        if int(c.shape[1]) != n_feat:
            raise L.Ff3dError("assemble_sweeps: every cloud needs the same number of columns")
        offs.append(offs[-1] + int(c.shape[0]))
    raw = torch.cat([torch.as_tensor(c, dtype=torch.float32).to(device, non_blocking=True) for c in clouds]).contiguous()
    ns = len(clouds)
    rot = np.tile(np.eye(3, dtype=np.float64).reshape(1, 9), (ns, 1))
    trans = np.zeros((ns, 3), dtype=np.float64)
    dt = np.zeros((ns,), dtype=np.float32)
    rm = np.zeros((ns,), dtype=np.uint8)
    tf = np.zeros((ns,), dtype=np.uint8)
    for i, s in enumerate(sweeps, start=1):
        rot[i] = np.asarray(s["sensor2lidar_rotation"], dtype=np.float64).reshape(9)
        trans[i] = np.asarray(s["sensor2lidar_translation"], dtype=np.float64).reshape(3)
        dt[i] = np.float32(timestamp - s["timestamp"] / 1e6)
        rm[i], tf[i] = 1, 1
    if out is None:
        out = torch.empty((offs[-1], 5), dtype=torch.float32, device=device)
    rng = L.float_array(point_range) if point_range is not None else None
    check(lib.ff3d_assemble_sweeps(C.c_void_p(raw.data_ptr()), n_feat, L.int_array(offs), ns,
                                   rot.ctypes.data_as(C.POINTER(C.c_double)), trans.ctypes.data_as(C.POINTER(C.c_double)),
                                   dt.ctypes.data_as(C.POINTER(C.c_float)), rm.ctypes.data_as(C.POINTER(C.c_ubyte)),
                                   tf.ctypes.data_as(C.POINTER(C.c_ubyte)), float(close_radius), rng, float(PAD_VALUE),
                                   C.c_void_p(out.data_ptr()), _stream()), "ff3d_assemble_sweeps")
    return out


def preprocess_images(frames, img_scale=(800, 448), mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375), to_rgb=True,
                      size_divisor=32, lidar2img=None, device="cuda"):
    """frames: uint8 [n, H, W, 3] BGR (host or device).  Returns (img float32 [n, 3, H', W'] on the device, rescaled
    lidar2img or None).  img_scale = (width, height) as in the configs (ScaleImageMultiViewImage swaps it)."""
    frames = torch.as_tensor(frames)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise L.Ff3dError("preprocess_images: uint8 [n, H, W, 3] frames expected")
    frames = frames.to(device, non_blocking=True).contiguous()
    n, H, W, _ = frames.shape
    # ScaleImageMultiViewImage (:215-246): scales = (h, w) after the swap; sizes are truncated products
    sh, sw = float(img_scale[1]) / H, float(img_scale[0]) / W
    oh, ow = int(H * sh), int(W * sw)
    ph, pw = -(-oh // size_divisor) * size_divisor, -(-ow // size_divisor) * size_divisor
    out = torch.empty((n, 3, ph, pw), dtype=torch.float32, device=device)
    check(lib.ff3d_image_preprocess(C.c_void_p(frames.data_ptr()), n, H, W, oh, ow, L.float_array(mean), L.float_array(std),
                                    1 if to_rgb else 0, size_divisor, C.c_void_p(out.data_ptr()), ph, pw, _stream()),
          "ff3d_image_preprocess")
    l2i = None
    if lidar2img is not None:
        sf = np.eye(4)
        sf[0, 0] *= sw
        sf[1, 1] *= sh
        l2i = [sf @ np.asarray(m, dtype=np.float64) for m in lidar2img]
    return out, l2i
