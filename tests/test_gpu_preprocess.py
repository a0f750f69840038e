"""Input side on the GPU against the oracle's restatement of the reference's CPU data pipeline."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sweeps(seed, n_sweeps=10, n_pts=30000):
    rng = np.random.default_rng(seed)
    key = rng.uniform(-60, 60, (n_pts, 5)).astype(np.float32)
    key[:, 2] = rng.uniform(-6, 4, n_pts)
    key[:, 4] = 123.0                                         # whatever the file holds: the loader zeroes the key frame's lag
    sweeps = []
    t0 = 1.6e9
    for s in range(n_sweeps + 2):                             # more sweeps than sweeps_num: only the first 10 are used
        pts = rng.uniform(-60, 60, (n_pts - 50 * s, 5)).astype(np.float32)
        pts[:, 2] = rng.uniform(-6, 4, pts.shape[0])
        pts[:200, :2] = rng.uniform(-1.5, 1.5, (200, 2))      # some points close to the sensor
        a = rng.uniform(-0.05, 0.05)
        R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], dtype=np.float64)
        sweeps.append(dict(points=pts, sensor2lidar_rotation=R, sensor2lidar_translation=rng.uniform(-1, 1, 3),
                           timestamp=(t0 - 0.05 * (s + 1)) * 1e6))
    return key, sweeps, t0


def test_assemble_sweeps_matches_reference_loader():
    """Same surviving points in the same order as LoadPointsFromMultiSweeps (+ PointsRangeFilter), bit-exact copies and
    time lags, transformed coordinates within one float32 ulp (numpy's float64 matmul may fuse differently)."""
    from focalformer3d_b200.preprocess import assemble_sweeps
    from focalformer3d_b200.runtime import PAD_VALUE
    from oracle.preprocess import load_points_from_multi_sweeps, points_range_filter
    key, sweeps, t0 = _sweeps(0)
    rngf = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
    for use_range in (False, True):
        got = assemble_sweeps(key, sweeps, t0, point_range=rngf if use_range else None).cpu().numpy()
        want = load_points_from_multi_sweeps(key, sweeps, t0)
        if use_range:
            want = points_range_filter(want, rngf)
        assert got.shape[0] == key.shape[0] + sum(s["points"].shape[0] for s in sweeps[:10])       # fixed size, no compaction
        kept = got[got[:, 0] != np.float32(PAD_VALUE)]
        assert kept.shape == want.shape
        assert np.array_equal(kept[:, 3:], want[:, 3:])                                           # intensity, time lag: exact
        assert np.array_equal(kept[:key.shape[0] if not use_range else 0, :3], want[:key.shape[0] if not use_range else 0, :3])
        ulp = np.spacing(np.abs(want[:, :3]).astype(np.float32))
        assert (np.abs(kept[:, :3] - want[:, :3]) <= ulp).all()
        assert (kept[:, :3] == want[:, :3]).mean() > 0.999


def test_assembled_cloud_voxelises_like_the_compacted_one():
    """The pad-instead-of-compact contract: hard voxelisation of the padded cloud == voxelisation of the reference's
    compacted cloud (same voxels, same order, same points)."""
    from focalformer3d_b200 import ops
    from focalformer3d_b200.preprocess import assemble_sweeps
    from oracle.preprocess import load_points_from_multi_sweeps
    key, sweeps, t0 = _sweeps(1, n_pts=8000)
    padded = assemble_sweeps(key, sweeps, t0)
    compact = torch.from_numpy(load_points_from_multi_sweeps(key, sweeps, t0)).cuda()
    # use the coordinates the kernel produced (1-ulp differences would move points across voxel borders)
    kept = padded[padded[:, 0] != padded.new_tensor(1.0e9)]
    vs, rg = [0.075, 0.075, 0.2], [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
    a = ops.voxelize(padded.contiguous(), [0, padded.shape[0]], vs, rg, 10, 160000, want_voxels=True)
    b = ops.voxelize(kept.contiguous(), [0, kept.shape[0]], vs, rg, 10, 160000, want_voxels=True)
    na, nb = int(a["n_dev"][0].item()), int(b["n_dev"][0].item())
    assert na == nb and na > 1000 and kept.shape[0] == compact.shape[0]
    for k in ("coors", "num_points", "voxels"):
        assert torch.equal(a[k][:na], b[k][:nb])


@pytest.mark.parametrize("H,W,scale", [(900, 1600, (800, 448)), (450, 801, (400, 224)), (64, 96, (96, 64))])
def test_image_preprocess_matches_reference_pipeline(H, W, scale):
    from focalformer3d_b200.preprocess import preprocess_images
    from oracle.preprocess import image_pipeline
    rng = np.random.default_rng(H)
    frames = rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8)
    l2i = [rng.normal(size=(4, 4)) for _ in range(3)]
    got, gl = preprocess_images(frames, img_scale=scale, lidar2img=l2i)
    want, wl = image_pipeline(frames, img_scale=scale, lidar2img=l2i)
    assert tuple(got.shape) == want.shape
    assert np.abs(got.cpu().numpy() - want).max() < 2e-5
    assert all(np.array_equal(a, b) for a, b in zip(gl, wl))
