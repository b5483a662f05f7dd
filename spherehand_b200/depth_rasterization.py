"""Drop-in for the reference's only native module, the pybind extension `depth_rasterization`
(/root/reference/mesh/cuda_kernel/depth_rasterization_cuda.cpp:15-25, built by mesh/cuda_kernel/setup.py:1-7).

    forward(width: int, height: int, vertices: Tensor[B,F,3,3] cuda float32 contiguous) -> Tensor[B,height,width]

Same contract as the reference shim: a non-CUDA or non-contiguous `vertices` raises RuntimeError (CHECK_CUDA /
CHECK_CONTIGUOUS, depth_rasterization_cuda.cpp:11-13,19); the result is a NEW tensor, 1000.0 where nothing was drawn
(depth_rasterization_cuda_kernel.cu:122).  `spherehand_b200.install()` registers this module under the reference's
top-level name so `import depth_rasterization` / `from mesh.cuda_kernel import depth_rasterization` resolve to it.
"""
import torch

from . import ops


def forward(width, height, vertices):
    if not isinstance(vertices, torch.Tensor):
        raise TypeError('vertices must be a torch.Tensor')
    if not vertices.is_cuda:
        raise RuntimeError('vertices must be a CUDA tensor')
    if not vertices.is_contiguous():
        raise RuntimeError('vertices must be contiguous')
    if vertices.dtype != torch.float32:
        raise RuntimeError('vertices must be float32 (the reference kernel reads data<float>(), .cu:130-131)')
    return ops.tri_raster_fwd(vertices, int(width), int(height))
