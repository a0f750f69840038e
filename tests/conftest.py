import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # FF3D_GEMM=tf32 runs the whole suite on the TF32 hi/lo operand format instead of the default fp16 hi/lo split
    # (focalformer3d_b200/ops.py GEMM_KIND); both are validated on the B200 every round.


@pytest.fixture(scope="session")
def tiny_cfg():
    from focalformer3d_b200.config import load_config, default_config_path, scaled_model_cfg
    cfg = load_config(default_config_path())
    return scaled_model_cfg(cfg["model"], bev=32, num_proposals=24)


@pytest.fixture(scope="session")
def tiny_sd(tiny_cfg):
    from focalformer3d_b200.synth import make_state_dict
    return make_state_dict(tiny_cfg, 0)


@pytest.fixture(scope="session")
def tiny_points(tiny_cfg):
    import torch
    from focalformer3d_b200.synth import synth_points
    r = tiny_cfg["pts_voxel_layer"]["point_cloud_range"]
    return [torch.from_numpy(synth_points(n, r, seed=s)) for s, n in enumerate((9000, 7000))]
