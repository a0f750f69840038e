"""Oracle: the reference's CPU data pipeline between the files and simple_test.  TEST INFRASTRUCTURE (see oracle/__init__.py).

[upstream] mmdet3d v0.17.1 ``LoadPointsFromMultiSweeps`` (datasets/pipelines/loading.py), ``PointsRangeFilter``
(transforms_3d.py) as configured at projects/configs/focalformer3d/FocalFormer3D_L.py:100-111, and the in-tree image
transforms projects/mmdet3d_plugin/datasets/pipelines/transform_3d.py:125-249 on [upstream] mmcv 1.3.18 ``imresize``
(cv2.INTER_LINEAR) / ``imnormalize`` / ``impad_to_multiple``.  cv2 is not installed here: the bilinear resize restates
OpenCV's float INTER_LINEAR arithmetic (resize.cpp: fx = (dx + 0.5) * scale - 0.5, border clamps, horizontal then vertical
pass) -- PARITY UNPINNED for that one function (no fixture of cv2's own output is available offline).
"""
import numpy as np


def remove_close(points, radius=1.0):
    """LoadPointsFromMultiSweeps._remove_close"""
    x_filt = np.abs(points[:, 0]) < radius
    y_filt = np.abs(points[:, 1]) < radius
    return points[np.logical_not(np.logical_and(x_filt, y_filt))]


def load_points_from_multi_sweeps(key_points, sweeps, timestamp, sweeps_num=10, use_dim=(0, 1, 2, 3, 4)):
    """test_mode=True branch: the first sweeps_num sweeps, in order."""
    points = np.array(key_points, dtype=np.float32, copy=True)
    points[:, 4] = 0
    out = [points]
    for sweep in list(sweeps)[:sweeps_num]:
        ps = remove_close(np.array(sweep["points"], dtype=np.float32, copy=True))
        sweep_ts = sweep["timestamp"] / 1e6
        ps[:, :3] = ps[:, :3] @ np.asarray(sweep["sensor2lidar_rotation"], dtype=np.float64).T
        ps[:, :3] += np.asarray(sweep["sensor2lidar_translation"], dtype=np.float64)
        ps[:, 4] = timestamp - sweep_ts
        out.append(ps)
    return np.concatenate(out)[:, list(use_dim)]


def points_range_filter(points, rng):
    """[upstream] PointsRangeFilter -> BasePoints.in_range_3d (strict inequalities)."""
    m = ((points[:, 0] > rng[0]) & (points[:, 1] > rng[1]) & (points[:, 2] > rng[2]) & (points[:, 0] < rng[3])
         & (points[:, 1] < rng[4]) & (points[:, 2] < rng[5]))
    return points[m]


def resize_linear_f32(img, ow, oh):
    """cv2.resize(img float32 [H, W, C], (ow, oh), interpolation=INTER_LINEAR)."""
    H, W, _ = img.shape
    sx, sy = np.float32(W / ow), np.float32(H / oh)

    def taps(n_out, n_in, scale):
        f = (np.arange(n_out, dtype=np.float32) + np.float32(0.5)) * scale - np.float32(0.5)
        i0 = np.floor(f).astype(np.int32)
        f = (f - i0.astype(np.float32)).astype(np.float32)
        lo = i0 < 0
        f[lo], i0[lo] = 0, 0
        hi = i0 >= n_in - 1
        f[hi], i0[hi] = 0, n_in - 1
        return i0, np.minimum(i0 + 1, n_in - 1), (np.float32(1) - f).astype(np.float32), f
    x0, x1, a0, a1 = taps(ow, W, sx)
    y0, y1, b0, b1 = taps(oh, H, sy)
    img = img.astype(np.float32)
    h = img[:, x0] * a0[None, :, None] + img[:, x1] * a1[None, :, None]            # horizontal pass, float32
    return (h[y0] * b0[:, None, None] + h[y1] * b1[:, None, None]).astype(np.float32)


def image_pipeline(frames_u8, img_scale=(800, 448), mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375), to_rgb=True,
                   size_divisor=32, lidar2img=None):
    """frames uint8 [n, H, W, 3] BGR -> (float32 [n, 3, H', W'], rescaled lidar2img)."""
    n, H, W, _ = frames_u8.shape
    scales = np.array([img_scale[1], img_scale[0]])                                # the swap in __init__ (:205-208)
    rand_scale = scales / np.array([H, W])
    y_size, x_size = int(H * rand_scale[0]), int(W * rand_scale[1])
    mean32, stdinv = np.asarray(mean, np.float32), (1 / np.float64(np.asarray(std, np.float32))).astype(np.float32)
    outs = []
    for f in frames_u8:
        img = resize_linear_f32(f.astype(np.float32), x_size, y_size)              # to_float32, then mmcv.imresize
        if to_rgb:
            img = img[..., ::-1]
        img = ((img - mean32) * stdinv).astype(np.float32)                         # mmcv.imnormalize
        ph, pw = -(-y_size // size_divisor) * size_divisor, -(-x_size // size_divisor) * size_divisor
        pad = np.zeros((ph, pw, 3), np.float32)                                    # mmcv.impad_to_multiple, pad_val 0
        pad[:y_size, :x_size] = img
        outs.append(pad.transpose(2, 0, 1))                                        # DefaultFormatBundle3D: HWC -> CHW
    l2i = None
    if lidar2img is not None:
        sf = np.eye(4)
        sf[0, 0] *= rand_scale[1]
        sf[1, 1] *= rand_scale[0]
        l2i = [sf @ np.asarray(m, dtype=np.float64) for m in lidar2img]
    return np.stack(outs), l2i
