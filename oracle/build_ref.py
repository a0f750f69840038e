"""Build the reference's OWN CUDA extensions for the checker (TEST INFRASTRUCTURE; see oracle/__init__.py).

The two in-tree extensions of the reference's LiDAR + camera path compile from their own few source files with nothing but
torch headers, so they are compiled from the sources WHERE THEY LIE under /root/reference (nothing is copied into this
repository) into oracle/_ref/ (git-ignored, but shipped to the GPU box like the other built artefacts):

  locatt_ops   projects/mmdet3d_plugin/models/utils/ops/locatt_ops/{similar.cu, weighting.cu, localAttention.cpp}
               -> oracle/_ref/ref_localattention.so   (similar_forward / weighting_forward, kernels.cuh:4-80)
  bev_pool     projects/mmdet3d_plugin/models/utils/ops/bev_pool/src/{bev_pool.cpp, bev_pool_cuda.cu}
               -> oracle/_ref/ref_bev_pool.so          (bev_pool_forward, bev_pool_cuda.cu:20-42)

They validate oracle/bev.py (local_similar / local_weighting) and oracle/camera.py (voxel pooling) -- and through them the
CUDA product kernels -- against the reference's real kernels on the GPU box (tests/test_gpu_reference_ext.py).
Run in the build container (where /root/reference exists):   python oracle/build_ref.py
"""
import os
import shutil
import sys

REF = "/root/reference/projects/mmdet3d_plugin/models/utils/ops"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def up_to_date():
    outs = [os.path.join(OUT, n + ".so") for n in ("ref_localattention", "ref_bev_pool")]
    return all(os.path.exists(o) for o in outs)


def build(verbose=False, force=False):
    if not os.path.isdir(REF):
        return False                                   # GPU box: the prebuilt .so files travel with the snapshot
    if up_to_date() and not force:
        return True
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils import cpp_extension
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
             "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"]
    jobs = [("ref_localattention", [f"{REF}/locatt_ops/similar.cu", f"{REF}/locatt_ops/weighting.cu",
                                    f"{REF}/locatt_ops/localAttention.cpp"]),
            ("ref_bev_pool", [f"{REF}/bev_pool/src/bev_pool.cpp", f"{REF}/bev_pool/src/bev_pool_cuda.cu"])]
    for name, srcs in jobs:
        bdir = os.path.join(OUT, name + "_build")
        os.makedirs(bdir, exist_ok=True)
        cpp_extension.load(name, sources=srcs, build_directory=bdir, extra_cuda_cflags=flags, verbose=verbose,
                           is_python_module=False)
        so = os.path.join(bdir, name + ".so")
        if os.path.exists(so):
            os.replace(so, os.path.join(OUT, name + ".so"))
        shutil.rmtree(bdir, ignore_errors=True)        # objects / ninja files are not needed on the GPU box
    return True


def load(name):
    """torch.ops-free loader: returns the python extension module built above (or None when absent)."""
    import importlib.util
    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = build(verbose="-v" in sys.argv, force="--force" in sys.argv)
    print("built" if ok else "reference tree not present: nothing to build", os.listdir(OUT) if os.path.isdir(OUT) else [])
