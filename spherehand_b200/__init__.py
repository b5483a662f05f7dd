"""spherehand_b200 — the per-batch training inner loop of melonwan/sphereHand on B200 (sm_100a) CUDA kernels.

Layout: `csrc/` (kernels + C ABI, include/spherehand_b200.h), `_lib`/`ops` (ctypes binding), `mesh/`, `network/`
(drop-in mirrors of the reference's modules for this path), `engine` (the fused CUDA-graph train step), `parallel`
(batch sharding + gradient all-reduce rules), `depth_rasterization` (drop-in for the reference's pybind module).
"""
import sys


_MIRRORED = (
    'mesh.cuda_kernel', 'mesh.render', 'mesh.multiview_utility', 'mesh.kinematicsTransformation', 'mesh.pointTransformation',
    'mesh.bone_length', 'network.hourglass', 'network.create_network_and_criterion', 'network.util_modules', 'network.pose_vae',
    'network.pose_denoiser', 'dataset.joint_angle', 'dataset.nyu_dataset',
)
_REF_PREFIX = '_spherehand_reference.'


def _reference_fallback(name, ref_root):
    """Module-level __getattr__ (PEP 562) for a mirrored module: a symbol this package does not re-implement (off-path classes
    such as DepthResample, FuseMvPose, WeightedMultiviewConsistencyLoss) is taken from the reference's own file of the same
    module, loaded once under a private name; its imports of `mesh.*` / `network.*` resolve to the installed modules."""
    import importlib.util
    import os
    path = os.path.join(ref_root, *name.split('.')) + '.py'

    def __getattr__(attr):
        if attr.startswith('__'):
            raise AttributeError(attr)
        key = _REF_PREFIX + name
        ref = sys.modules.get(key)
        if ref is None:
            if not os.path.exists(path):
                raise AttributeError('%s.%s is not part of spherehand_b200 and %s does not exist' % (name, attr, path))
            spec = importlib.util.spec_from_file_location(key, path)
            ref = importlib.util.module_from_spec(spec)
            sys.modules[key] = ref
            try:
                spec.loader.exec_module(ref)
            except BaseException:
                sys.modules.pop(key, None)
                raise
        try:
            return getattr(ref, attr)
        except AttributeError:
            raise AttributeError('neither spherehand_b200.%s nor the reference module %s defines %r' % (name, path, attr)) from None
    return __getattr__


def install(reference_root=None):
    """Make the reference's import names resolve to this package, so network/engine.py and mesh/render.py of the
    reference (or user code written against them) run on the new path unchanged:
        depth_rasterization, mesh.cuda_kernel, mesh.render, mesh.multiview_utility, mesh.kinematicsTransformation,
        mesh.pointTransformation, network.hourglass, network.create_network_and_criterion, network.util_modules,
        network.pose_vae, network.pose_denoiser, dataset.joint_angle, dataset.nyu_dataset.
    Only names that are not already imported are registered (an already-imported reference module is left alone).

    reference_root (or $SPHEREHAND_REFERENCE_ROOT): the root of a sphereHand checkout.  With it the three shadow packages
    `network`, `mesh`, `dataset` also search the reference's directories, so modules this package does not mirror — the unchanged
    callers network/engine.py, run_engine.py, constants.py, util_vis.py, utils_metric.py, mesh/joint_order.py ... — import from the
    checkout, and symbols missing from a mirrored module fall back to the reference's definition (`_reference_fallback`).
    Without it only the mirrored modules resolve (`from network.engine import Engine` then needs the checkout on sys.path
    BEFORE install(), in which case nothing of `network` is shadowed)."""
    import importlib
    import os
    from . import depth_rasterization
    reference_root = reference_root or os.environ.get('SPHEREHAND_REFERENCE_ROOT')
    table = {'depth_rasterization': depth_rasterization}
    for pkg in ('mesh', 'network', 'dataset'):
        table[pkg] = importlib.import_module('.' + pkg, __name__)
    for name in _MIRRORED:
        table[name] = importlib.import_module('.' + name, __name__)
    if reference_root:
        reference_root = os.path.abspath(reference_root)
        if not os.path.isdir(os.path.join(reference_root, 'network')):
            raise FileNotFoundError('install(reference_root=%r): no network/ directory there' % reference_root)
        for pkg in ('mesh', 'network', 'dataset'):
            d = os.path.join(reference_root, pkg)
            if os.path.isdir(d) and d not in table[pkg].__path__:
                table[pkg].__path__.append(d)
        for name in _MIRRORED:
            if '__getattr__' not in table[name].__dict__:
                table[name].__getattr__ = _reference_fallback(name, reference_root)
    for name, mod in table.items():
        sys.modules.setdefault(name, mod)
    return table
