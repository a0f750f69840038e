"""focalformer3d_b200: B200-native (sm_100a) implementation of FocalFormer3D's per-scene forward path.

``config``/``synth`` are importable anywhere; ``model``/``ops``/``lib`` need the built ``libff3d.so`` (they raise
loudly if it is missing -- there is no CPU fallback) and a CUDA device to run.
"""
__all__ = ["config", "synth"]
