"""Config loading + type registry: the host-side boundary the reference exposes through mmcv.

Reference: ``tools/test.py:129-159`` (``Config.fromfile`` then ``importlib.import_module(plugin_dir)``),
``tools/test.py:202-203`` (``build_model(cfg.model, test_cfg=cfg.get('test_cfg'))``) and the registry
decorators at ``projects/mmdet3d_plugin/models/detectors/focalformer3d.py:26``,
``models/necks/focal_encoder.py:89``, ``models/dense_heads/focal_decoder.py:33``,
``core/bbox/coders/transfusion_bbox_coder.py:7``.  The shipped configs are plain python (no ``_base_``,
no imports), so ``exec`` reproduces ``Config.fromfile`` for them.
"""
import copy
import os


class Registry:
    """Minimal stand-in for mmcv.utils.Registry: same ``type``-string lookup, same decorator."""

    def __init__(self, name):
        self.name = name
        self._map = {}

    def register_module(self, name=None, force=False):
        def deco(cls):
            key = name or cls.__name__
            if key in self._map and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._map[key] = cls
            return cls
        return deco

    def get(self, key):
        return self._map.get(key)

    def build(self, cfg, **default_args):
        if cfg is None:
            return None
        args = dict(cfg)
        typ = args.pop("type")
        cls = self._map.get(typ)
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        for k, v in default_args.items():
            args.setdefault(k, v)
        return cls(**args)

    def __contains__(self, key):
        return key in self._map


DETECTORS = Registry("detector")
NECKS = Registry("neck")
HEADS = Registry("head")
BBOX_CODERS = Registry("bbox_coder")
VOXEL_ENCODERS = Registry("voxel_encoder")
MIDDLE_ENCODERS = Registry("middle_encoder")
BACKBONES = Registry("backbone")


class ConfigDict(dict):
    """dict with attribute access (mmcv ConfigDict behaviour used by the reference: cfg.model, cfg.get)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, list):
        return [_wrap(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_wrap(x) for x in v)
    return v


def load_config(path):
    """Execute an mmcv-style python config and return its public names (like ``Config.fromfile``)."""
    with open(path, "r") as f:
        src = f.read()
    ns = {"__file__": os.path.abspath(path)}
    exec(compile(src, path, "exec"), ns)
    import types
    out = {k: v for k, v in ns.items() if not k.startswith("__") and not isinstance(v, types.ModuleType)}
    return _wrap(out)


def default_config_path(name="focalformer3d_l"):
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", name + ".py")


def scaled_camera_cfg(model_cfg, bev, img_hw, num_proposals=None):
    """Smaller variant of a camera-only config: BEV cells per side and image size (H, W multiples of 32)."""
    m = copy.deepcopy(model_cfg)
    osf = m["pts_bbox_head"]["bbox_coder"]["out_size_factor"]
    vs = m["pts_bbox_head"]["bbox_coder"]["voxel_size"]
    old = m["imgpts_neck"]["pc_range"]
    half = bev * osf * vs[0] / 2.0
    rng = [-half, -half, old[2], half, half, old[5]]
    m["imgpts_neck"]["pc_range"] = rng
    m["imgpts_neck"]["img_scale"] = tuple(img_hw)
    m["pts_bbox_head"]["bbox_coder"]["pc_range"] = rng[:2]
    m["pts_bbox_head"]["bbox_coder"]["post_center_range"] = [rng[0] * 1.2, rng[1] * 1.2, -10.0, rng[3] * 1.2, rng[4] * 1.2, 10.0]
    if num_proposals is not None:
        m["pts_bbox_head"]["num_proposals"] = num_proposals
    t = m["test_cfg"]["pts"]
    t["grid_size"] = [bev * osf, bev * osf, t["grid_size"][2]]
    t["pc_range"] = rng[:2]
    return _wrap(m)


def scaled_fusion_cfg(model_cfg, bev, img_hw, num_proposals=None, max_voxels=None):
    """Smaller variant of a LiDAR + camera config: scaled_model_cfg for the LiDAR tower / head plus the image size and
    the Lift-Splat-Shoot BEV range of the neck."""
    m = scaled_model_cfg(model_cfg, bev, num_proposals=num_proposals, max_voxels=max_voxels)
    m["imgpts_neck"]["pc_range"] = list(m["pts_voxel_layer"]["point_cloud_range"])
    m["imgpts_neck"]["img_scale"] = tuple(img_hw)
    return m


def scaled_model_cfg(model_cfg, bev, z_cells=None, num_proposals=None, max_voxels=None):
    """Derive a geometrically smaller variant of a LiDAR config (same layers, same voxel size, smaller
    range) for parity tests the CPU oracle finishes in seconds.  ``bev`` = BEV cells per side."""
    m = copy.deepcopy(model_cfg)
    osf = m["pts_bbox_head"]["bbox_coder"]["out_size_factor"]
    vs = m["pts_voxel_layer"]["voxel_size"]
    old = m["pts_voxel_layer"]["point_cloud_range"]
    nx = bev * osf
    half = nx * vs[0] / 2.0
    nz = int(round((old[5] - old[2]) / vs[2])) if z_cells is None else z_cells
    rng = [-half, -half, old[2], half, half, old[2] + nz * vs[2]]
    m["pts_voxel_layer"]["point_cloud_range"] = rng
    if max_voxels is not None:
        m["pts_voxel_layer"]["max_voxels"] = max_voxels
    m["pts_middle_encoder"]["sparse_shape"] = [nz + 1, nx, nx]
    m["pts_bbox_head"]["bbox_coder"]["pc_range"] = rng[:2]
    m["pts_bbox_head"]["bbox_coder"]["post_center_range"] = [rng[0] * 1.2, rng[1] * 1.2, -10.0, rng[3] * 1.2, rng[4] * 1.2, 10.0]
    if num_proposals is not None:
        m["pts_bbox_head"]["num_proposals"] = num_proposals
    t = m["test_cfg"]["pts"]
    t["grid_size"] = [nx, nx, nz]
    t["pc_range"] = rng[:2]
    return _wrap(m)
