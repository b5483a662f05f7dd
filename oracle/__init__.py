"""CPU oracle for the sphereHand hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Every function here is a from-scratch CPU restatement (numpy / plain C / torch-fp32-on-CPU for the
floating-point heads) of the reference algorithm it cites (file:line into /root/reference).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this package, and only as the checker / the timed CPU baseline.  The product package
`spherehand_b200` never imports it and fails loudly when its CUDA library is missing.

Pinning status:
  * The reference ships NO golden vectors or known-answer tests.  The oracle is
    therefore pinned against outputs of the reference itself, generated in the build container by
    `oracle/make_golden.py` (imports the unmodified reference on CPU through `oracle/_refshim.py`)
    and committed under `tests/golden/`.
  * R1 (triangle rasteriser) has no CPU implementation in the reference; `oracle/tri_raster.c`
    restates depth_rasterization_cuda_kernel.cu:18-113, and is pinned on the GPU box against the
    reference's own kernel compiled by `oracle/build_ref.py` into `oracle/_ref/`.
"""
