"""In-container-only helper: import the UNMODIFIED reference from /root/reference on CPU.

TEST INFRASTRUCTURE.  Only `oracle/make_golden.py` uses this, and only in the build
container (the GPU box has no /root/reference).  It applies the six import-time shims of
SURVEY.md §8c and changes no semantics:
  (1) a stub top-level module `depth_rasterization`  (mesh/cuda_kernel/__init__.py:1)
  (2) np.float = float                                (kinematicsTransformation.py:118)
  (3) stub matplotlib / matplotlib.pyplot / cv2-free  (mesh/bone_length.py:7,15)
  (4) Tensor.cuda / Module.cuda -> identity on a CPU box
  (5) torch.load default map_location='cpu'           (network/pose_vae.py:20)
  (6) CWD = reference root                            (network/constants.py:4)
"""
import os
import sys
import types

REF_ROOT = os.environ.get("SPHEREHAND_REFERENCE", "/root/reference")


def install():
    import numpy as np
    import torch

    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    if "depth_rasterization" not in sys.modules:
        stub = types.ModuleType("depth_rasterization")

        def _forward(width, height, vertices):
            raise RuntimeError("reference CUDA rasteriser is not available on CPU")

        stub.forward = _forward
        sys.modules["depth_rasterization"] = stub
    if not hasattr(np, "float"):
        np.float = float
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        _load = torch.load

        def _cpu_load(f, *a, **k):
            k.setdefault("map_location", "cpu")
            k.setdefault("weights_only", False)
            return _load(f, *a, **k)

        torch.load = _cpu_load
    os.chdir(REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
