"""Recipe: compile the REFERENCE's own CUDA rasteriser into oracle/_ref/ (git-ignored).

TEST INFRASTRUCTURE — the result is only ever used as a checker (tests/) or as the
"reference" timing arm of bench.py; the product never loads it.

The reference sources are compiled where they lie under /root/reference
(mesh/cuda_kernel/depth_rasterization_cuda.cpp and ..._cuda_kernel.cu).  As shipped the .cu
does not compile against torch >= 2.x: line 124 passes `vertices.type()` to
AT_DISPATCH_FLOATING_TYPES.  The recipe pre-processes a scratch copy under a temp dir with
that single token changed to `vertices.scalar_type()` (no arithmetic is touched) and runs
nvcc/g++ on it directly; nothing from the reference is written into this repository.

Output: oracle/_ref/depth_rasterization_ref.so  (a torch extension module exporting
`forward(width, height, vertices) -> Tensor`, sm_100 SASS).  The GPU box only uses the
pre-built file; this script is a no-op there.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SPHEREHAND_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "mesh", "cuda_kernel")
OUT = os.path.join(HERE, "_ref")
NAME = "depth_rasterization_ref"


def build(force=False):
    target = os.path.join(OUT, NAME + ".so")
    if os.path.exists(target) and not force:
        return target
    if not os.path.isdir(SRC):
        return None  # GPU box: nothing to build from
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="sh_ref_")
    try:
        cu = open(os.path.join(SRC, "depth_rasterization_cuda_kernel.cu")).read()
        assert cu.count("vertices.type()") == 1
        cu = cu.replace("vertices.type()", "vertices.scalar_type()")
        open(os.path.join(tmp, "k.cu"), "w").write(cu)
        shutil.copy(os.path.join(SRC, "depth_rasterization_cuda.cpp"), os.path.join(tmp, "b.cpp"))
        inc = []
        for p in ce.include_paths("cuda"):
            inc += ["-I", p]
        inc += ["-I", sysconfig.get_paths()["include"]]
        defs = ["-DTORCH_EXTENSION_NAME=" + NAME, "-DTORCH_API_INCLUDE_EXTENSION_H",
                "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
        nvcc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
        subprocess.check_call([nvcc, "-c", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
                               "-gencode", "arch=compute_100,code=sm_100", "-Xcompiler", "-fPIC",
                               "-w"] + defs + inc + [os.path.join(tmp, "k.cu"), "-o", os.path.join(tmp, "k.o")])
        subprocess.check_call(["g++", "-c", "-O2", "-std=c++17", "-fPIC", "-w"] + defs + inc +
                              [os.path.join(tmp, "b.cpp"), "-o", os.path.join(tmp, "b.o")])
        libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
        subprocess.check_call(["g++", "-shared", os.path.join(tmp, "k.o"), os.path.join(tmp, "b.o"),
                               "-o", target, "-L", libdir, "-Wl,-rpath," + libdir,
                               "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda",
                               "-ltorch_python", "-L/usr/local/cuda/lib64", "-lcudart"])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return target


REF_TREE = os.path.join(OUT, "reference")
_ASSETS = ("mesh/model/preprocessed_hand.pkl", "mesh/model/pose_prior.pkl", "mesh/model/pose_vae.pth",
           "mesh/model/pose_denoiser.pth", "pretrained/synthetic.pth")


def stage_reference(force=False):
    """Copy the reference's Python modules of the path and the assets they open at import time (network/constants.py:4-8,
    network/pose_vae.py:20, engine.py:72) into oracle/_ref/reference/ (git-ignored, NOT gpurun-ignored: it travels to the GPU
    box like the compiled kernel).  Byte-for-byte copies, nothing edited; only bench.py's reference-GPU leg (oracle/ref_gpu.py)
    and the drop-in tests read them.  A no-op on the GPU box."""
    if not os.path.isdir(os.path.join(REF, "network")):
        return REF_TREE if os.path.isdir(REF_TREE) else None
    stamp = os.path.join(REF_TREE, ".staged")
    if os.path.exists(stamp) and not force:
        return REF_TREE
    for pkg in ("mesh", "network", "dataset"):
        dst = os.path.join(REF_TREE, pkg)
        os.makedirs(dst, exist_ok=True)
        for fn in sorted(os.listdir(os.path.join(REF, pkg))):
            if fn.endswith(".py"):
                shutil.copyfile(os.path.join(REF, pkg, fn), os.path.join(dst, fn))
    os.makedirs(os.path.join(REF_TREE, "mesh", "cuda_kernel"), exist_ok=True)
    shutil.copyfile(os.path.join(SRC, "__init__.py"), os.path.join(REF_TREE, "mesh", "cuda_kernel", "__init__.py"))
    for a in _ASSETS:
        os.makedirs(os.path.dirname(os.path.join(REF_TREE, a)), exist_ok=True)
        shutil.copyfile(os.path.join(REF, a), os.path.join(REF_TREE, a))
    open(stamp, "w").write("staged from %s\n" % REF)
    return REF_TREE


def load():
    """Import the pre-built module (needs a CUDA device to be useful)."""
    import importlib.util
    import torch  # noqa: F401  (must be imported first: the .so links libtorch)

    path = os.path.join(OUT, NAME + ".so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(NAME, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(stage_reference(force="--force" in sys.argv))
