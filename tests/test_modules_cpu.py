"""CPU-side checks of the drop-in module surfaces (no kernel runs): constructor arguments, registered buffers /
state_dict keys the reference's checkpoints and callers rely on, host-side tables, and the no-CPU-fallback rule."""
import os

import numpy as np
import pytest
import torch

from conftest import golden, mesh_dict
from oracle import hourglass as oh, losses

import spherehand_b200
from spherehand_b200.mesh import bone_length, kinematicsTransformation as kt, multiview_utility as mv, pointTransformation as pt, render
from spherehand_b200.network import create_network_and_criterion as cnc, pose_vae, util_modules
from spherehand_b200.network.hourglass import create_hourglass_network


def test_tables_match_the_reference_tables():
    a, b = losses.collision_pairs()
    col = render.CollisionLoss()
    assert col.joint_1.tolist() == a.tolist() and col.joint_2.tolist() == b.tolist() and len(a) == 690
    a, b, length = losses.bone_pairs()
    assert bone_length.joint_1 == a.tolist() and bone_length.joint_2 == b.tolist()
    assert np.allclose(bone_length.uniform_length, length) and len(length) == 35
    bl = render.BoneLengthLoss()
    assert torch.allclose(bl.min_length[0], (torch.tensor(length) * 0.8) ** 2) and torch.allclose(bl.max_length[0], (torch.tensor(length) * 1.05) ** 2)


def test_module_surfaces_and_state_dict_keys(hand_model):
    mesh = mesh_dict(hand_model)
    radii = [float(r) for r in hand_model['keypoint_radius']]
    assert set(render.BallRender(64, 48).state_dict()) == {'dist_weight', 'x_grid', 'y_grid'}
    assert render.BallRender(64, 48).x_grid.shape == (48, 64)
    d2m = render.DataToModelLoss(32, 32, mesh)
    assert d2m.num_joints == 41 and d2m.radiuses.shape == (1, 1, 1, 41) and torch.allclose(d2m.radiuses.view(-1), torch.tensor(radii))
    assert render.DataToModelLoss(32, 32, radii).num_joints == 41
    with pytest.raises(TypeError):
        render.DataToModelLoss(32, 32, np.asarray(radii))
    mpl = mv.MutualProjectionLoss(64, mesh)
    assert mpl.num_joints == 41 and 'mutual_projection.radiuses' in mpl.state_dict()
    hb = render.HandBallPrimitiveRender(mesh['bones'], 64, 64)
    assert hb.num_vertices == 41 and hb.radiuses.shape == (1, 41)
    lbs = pt.LinearBlendSkinning(mesh['vertices'], [b['weight_coeff'] for b in mesh['bones']], [b['weight_vertexid'] for b in mesh['bones']])
    assert lbs.num_vertices == 10144 and lbs.row_ptr[-1].item() == 25890 and lbs.wv.shape == (25890, 4)
    with pytest.raises(AssertionError):
        pt.LinearBlendSkinning(mesh['vertices'], [[1.0]], [[0], [1]])
    faces = mesh['faces'].copy()
    dr = render.DepthRender(mesh, 128)
    assert np.array_equal(mesh['faces'], faces)                    # no in-place mutation of the caller's faces
    assert dr.rasterizer.faces.dtype == torch.int64 and dr.rasterizer.faces.numel() == 3382 * 3
    assert dr.rasterizer.faces.view(-1, 3)[:, 0].tolist() == faces[:, 1].tolist()      # right-hand winding swap (render.py:298-300)
    assert dr.camera.k_mat.shape == (1, 4, 4) and dr.camera.k_mat[0, 0, 3] == 320
    htm = kt.HandTransformationMat([b['offset_matrix'] for b in mesh['bones']])
    assert htm.offset_mats.shape == (17, 4, 4)
    with pytest.raises(AssertionError):
        kt.HandTransformationMat([np.eye(4)] * 3)
    hs = util_modules.HandSynthesizer(mesh, 64, 16, 1.0, 0.01)
    assert hs.rand_scale.rand_scale == 0.1 and hs.depth_noiser.sigma_z == 0.05
    rec = util_modules.RecoverXYZCoordinateFromHeatmap(16, 16, 0.01)
    assert rec.depth_scale == 100.0 and set(rec.state_dict()) == {'u_grid', 'v_grid'}
    vae = pose_vae.PoseVae(123, 32)
    assert set(vae.state_dict()) == set(golden('pose_vae'))
    with pytest.raises(ValueError):
        pose_vae.PoseVae(60, 16)
    net = cnc.HeatmapEstimationNetwork(32, 0.01, 41, 2, real_aug=False)
    ref_keys = set('hg.' + k for k in oh.param_shapes(82, 2)) | {'xyz_recover.u_grid', 'xyz_recover.v_grid'}
    assert set(net.state_dict()) == ref_keys                     # the keys of pretrained/*.pth (SURVEY.md §2.2 assets)
    aug = cnc.HeatmapEstimationNetwork(32, 0.01, 41, 2)          # real_aug defaults to True like the reference (:28)
    assert aug.resize_dm is not None and set(aug.state_dict()) == ref_keys

    class Constant:
        pass
    Constant.mesh = mesh
    crit = cnc.MultiTaskLoss(True, True, True, False, False, True, True, Constant(), image_size=128, heatmap_size=32)
    assert crit.weights == {'synt_hm': 1e3, 'synt_pt': 1e-1, 'mv_consistency': 1e-3, 'mv_projection': 1, 'temporal_smooth': 1.0,
                            'prior': 1e-2, 'hm_mean': 1e-2, 'domain': 0.0, 'collision': 1.0, 'bone_length': 1.0}
    assert crit.prior_loss is None and crit.heatmap_size == 32
    with pytest.raises(NotImplementedError):
        cnc.MultiTaskLoss(True, True, True, True, False, True, True, Constant())


def test_joint_angle_host_walk_and_generator_position():
    """Host side of the batched pose sampler (spherehand_b200/dataset/joint_angle.py): the integer walk over the mode draws
    finds the same pose boundaries as the oracle restatement, and the generator is left exactly where n consecutive
    `__getitem__` calls of the reference leave it (fixture value `after`)."""
    from spherehand_b200.dataset import joint_angle as ja
    from oracle import poses as oposes
    g = golden('joint_angle')
    n = g['poses'].shape[0]
    _, offs, used = oposes.joint_angle_batch(g['u'], n)
    o, e = ja.sequential_offsets(g['u'])
    assert np.array_equal(o[:n], offs) and int(e[n - 1]) == used and ja.MAX_UNIFORMS == 44
    with pytest.raises(RuntimeError):
        ja.JointAngleDataset(device='cpu')
    # __getitem__ is the reference's DataLoader-worker protocol (a CPU tensor, no CUDA touched): 256 consecutive items == the 256
    # consecutive `__getitem__` results of the unmodified reference, bit for bit, and the generator ends where the reference leaves it
    ds = ja.JointAngleDataset()
    torch.manual_seed(int(g['seed']))
    items = [ds[i] for i in range(n)]
    assert all(not t.is_cuda and t.dtype == torch.float32 and t.shape == (26,) for t in items)
    assert np.array_equal(torch.stack(items).numpy(), g['poses'])
    assert np.array_equal(torch.rand(4).numpy(), g['after'])
    assert len(ja.JointAngleDataset.__mro__) and ja.JointAngleDataset.INDEX == 6 and ja.JointAngleDataset.THUMB == 22


def test_npy_shard_reader_and_writer(tmp_path):
    """dataset/nyu_dataset.py mirror: the committed shards (tests/golden/shard, the reference's on-disk format) read item by item
    like the UNMODIFIED reference reader did (tests/golden/shard_items.npz, oracle/make_golden_shard.py), inverse camera poses
    included bit for bit; the writer reproduces the shard files byte for byte."""
    import filecmp
    from spherehand_b200.dataset import nyu_dataset as nd
    shard_dir = os.path.join(os.path.dirname(__file__), 'golden', 'shard')
    g = golden('shard_items')
    ds = nd.create_nyu_dataset(shard_dir)
    assert len(ds) == int(g['n']) == 8 and len(ds.datasets) == 2
    for i in range(len(ds)):
        dm, jp, cp, icp = ds[i]
        assert dm.dtype == np.float32 and dm.shape == (3, 16, 16)
        for a, k in ((dm, 'dm'), (jp, 'jp'), (cp, 'cp'), (icp, 'icp')):
            assert np.array_equal(a, g['%s%d' % (k, i)])
    one = nd.NpyDataset(os.path.join(shard_dir, 'mv_data_1'), transform=lambda dm, jp, cp, icp: (dm.sum(), jp.shape))
    assert len(one) == 3 and one[2][1] == (3, 36, 3)
    src = nd.NpyDataset(os.path.join(shard_dir, 'mv_data_0'))
    nd.write_npy_shard(str(tmp_path), 'mv_data_0', np.asarray(src.dms), src.joint_poses, src.camera_poses)
    for suffix in ('_dms.bat', '_joint_poses.npy', '_camera_poses.npy', '_shape.pkl'):
        assert filecmp.cmp(os.path.join(shard_dir, 'mv_data_0' + suffix), str(tmp_path / ('mv_data_0' + suffix)), shallow=False), suffix
    with pytest.raises(RuntimeError):
        nd.ShardBatchLoader(ds, 2, device='cpu')


def test_pose_denoiser_surface():
    from spherehand_b200.network import pose_denoiser as pd
    g = golden('pose_denoiser')
    net = pd.PoseDenoiser()
    assert set(net.state_dict()) == {k[3:] for k in g if k.startswith('sd.')}          # the reference's checkpoint keys
    assert pd.input_indices == g['input_indices'].tolist() and pd.output_indices == g['output_indices'].tolist()
    assert net.input_fea == 112 and net.output_fea == 33 and net.scale_factor == 0.01
    net.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('sd.')})
    with pytest.raises(NotImplementedError):
        net.train()(torch.zeros(2, 123))
    with pytest.raises(RuntimeError):
        net.eval()(torch.zeros(2, 123))                                                # CPU tensor: no fallback
    j = torch.from_numpy(g['joints'])
    assert abs(float(net.loss(j, torch.from_numpy(g['out3']))) - float(g['loss'])) < 1e-6 * float(g['loss'])
    assert abs(pd.average_joint_error(j, torch.from_numpy(g['out3']), pd.key_points)
               - float((j[:, :11] - torch.from_numpy(g['out3'])[:, :11]).norm(dim=-1).mean())) < 1e-6


def test_no_cpu_fallback_anywhere(hand_model):
    """Every forward on CPU tensors raises (CHECK_CUDA of the reference shim, depth_rasterization_cuda.cpp:11-13): the
    product path must never silently compute on the host."""
    mesh = mesh_dict(hand_model)
    with pytest.raises(RuntimeError):
        spherehand_b200.depth_rasterization.forward(16, 16, torch.zeros(1, 2, 3, 3))
    with pytest.raises(RuntimeError):
        render.BallRender(16, 16)(torch.zeros(3, 3), torch.ones(3))
    with pytest.raises(RuntimeError):
        render.CollisionLoss()(torch.zeros(2, 3, 41, 3))
    with pytest.raises(RuntimeError):
        mv.MultiviewConsistencyLoss()(torch.eye(4).repeat(2, 3, 1, 1), torch.zeros(2, 3, 41, 3))
    with pytest.raises(RuntimeError):
        kt.HandTransformationMat([b['offset_matrix'] for b in mesh['bones']])(torch.zeros(2, 26))
    with pytest.raises(RuntimeError):
        create_hourglass_network(82, 1)(torch.zeros(1, 64, 64))
    with pytest.raises(IndexError):
        render.CollisionLoss()(torch.zeros(2, 40, 3))
    with pytest.raises(TypeError):
        spherehand_b200.depth_rasterization.forward(16, 16, np.zeros((1, 2, 3, 3), np.float32))


def test_install_table():
    import sys
    before = set(sys.modules)
    names = spherehand_b200.install()
    try:
        assert {'depth_rasterization', 'mesh.render', 'mesh.cuda_kernel', 'mesh.multiview_utility', 'mesh.kinematicsTransformation',
                'mesh.pointTransformation', 'network.hourglass', 'network.create_network_and_criterion', 'dataset.joint_angle'} <= set(names)
        import depth_rasterization
        assert callable(depth_rasterization.forward)
    finally:
        for k in set(sys.modules) - before:
            if k.split('.')[0] in ('mesh', 'network', 'depth_rasterization', 'dataset'):
                sys.modules.pop(k, None)


def test_c_abi_exports_every_declared_symbol():
    """libspherehand_b200.so loads without a GPU and exports exactly what include/spherehand_b200.h declares."""
    import subprocess
    from spherehand_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 48 and 'sh_tri_raster_fwd' in protos and 'sh_conv_fwd' in protos
    L = _lib.lib()                                         # binds every declared symbol (AttributeError if one is missing)
    assert L.sh_abi_version() == 3 and L.sh_build_arch() == b'sm_100a'
    exported = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    names = {l.split()[-1] for l in exported.splitlines() if ' T ' in l and l.split()[-1].startswith('sh_')}
    assert names == set(protos), (names ^ set(protos))
