"""Mirror of /root/reference/mesh/cuda_kernel/__init__.py, which re-exports the top-level native module
`depth_rasterization` so that `from mesh.cuda_kernel import depth_rasterization` (mesh/render.py:6) works."""
from ... import depth_rasterization  # noqa: F401
