"""spherehand_b200 — the per-batch training inner loop of melonwan/sphereHand on B200 (sm_100a) CUDA kernels.

Layout: `csrc/` (kernels + C ABI, include/spherehand_b200.h), `_lib`/`ops` (ctypes binding), `mesh/`, `network/`
(drop-in mirrors of the reference's modules for this path), `engine` (the fused CUDA-graph train step), `parallel`
(batch sharding + gradient all-reduce rules), `depth_rasterization` (drop-in for the reference's pybind module).
"""
import sys


def install():
    """Make the reference's import names resolve to this package, so network/engine.py and mesh/render.py of the
    reference (or user code written against them) run on the new path unchanged:
        depth_rasterization, mesh.cuda_kernel, mesh.render, mesh.multiview_utility, mesh.kinematicsTransformation,
        mesh.pointTransformation, network.hourglass, network.create_network_and_criterion, network.util_modules,
        network.pose_vae, network.pose_denoiser, dataset.joint_angle, dataset.nyu_dataset.
    Only names that are not already imported are registered (an already-imported reference module is left alone)."""
    import importlib
    from . import depth_rasterization
    table = {
        'depth_rasterization': depth_rasterization,
        'mesh': importlib.import_module('.mesh', __name__),
        'mesh.cuda_kernel': importlib.import_module('.mesh.cuda_kernel', __name__),
        'mesh.render': importlib.import_module('.mesh.render', __name__),
        'mesh.multiview_utility': importlib.import_module('.mesh.multiview_utility', __name__),
        'mesh.kinematicsTransformation': importlib.import_module('.mesh.kinematicsTransformation', __name__),
        'mesh.pointTransformation': importlib.import_module('.mesh.pointTransformation', __name__),
        'mesh.bone_length': importlib.import_module('.mesh.bone_length', __name__),
        'network': importlib.import_module('.network', __name__),
        'network.hourglass': importlib.import_module('.network.hourglass', __name__),
        'network.create_network_and_criterion': importlib.import_module('.network.create_network_and_criterion', __name__),
        'network.util_modules': importlib.import_module('.network.util_modules', __name__),
        'network.pose_vae': importlib.import_module('.network.pose_vae', __name__),
        'network.pose_denoiser': importlib.import_module('.network.pose_denoiser', __name__),
        'dataset': importlib.import_module('.dataset', __name__),
        'dataset.joint_angle': importlib.import_module('.dataset.joint_angle', __name__),
        'dataset.nyu_dataset': importlib.import_module('.dataset.nyu_dataset', __name__),
    }
    for name, mod in table.items():
        sys.modules.setdefault(name, mod)
    return table
