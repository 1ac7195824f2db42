"""A/B against the reference's OWN CUDA kernels (oracle/_ref/libkf_ref.so: the three .cu files of
/root/reference compiled unmodified for sm_100a behind a header-only OpenCV shim, see
oracle/Makefile).  Same GPU, same inputs: results must be BIT-IDENTICAL for integrate, raycast,
vertex/normal maps, the model pyramid and point extraction, because the product spells the same
sequence of roundings (MUFU.RCP included); ICP sums differ only by the reference's intermediate
f32 roundings.  This is what pins parity, since the reference ships no tests."""
import numpy as np
import pytest

from conftest import make_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kref():
    from oracle import kref as m
    if not m.available():
        pytest.fail("oracle/_ref/libkf_ref.so missing: run `make -C oracle` where /root/reference exists")
    assert m.lib().ref_sizeof_voxel() == 8
    return m


def _frames(kfo, Ko, ks):
    return [(kfo.trajectory_pose(k), kfo.frontend(kfo.render_depth_mm(kfo.trajectory_pose(k), Ko), Ko, levels=1)[0][0]) for k in ks]


@pytest.mark.parametrize("dims,w,h", [(64, 320, 240), (256, 640, 480)])
def test_integrate_bit_exact(kfo, kfb, kref, dims, w, h):
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, dims, w, h)
    ctx = kfb.Context(Kb, Pb)
    rv = kref.RefVolume(dims)
    volpose = np.array(Po.volu_pose, np.float32)
    rv.upload(np.zeros((dims, dims, dims, 2), np.int16))
    for pose, dm in _frames(kfo, Ko, (0, 5, 10, 15)):
        v2c = kfo.pose_mul(kfo.pose_inv(pose), volpose)
        ctx.upload_depth_m(0, dm)
        ctx.integrate(v2c)
        rv.integrate(v2c, dm, Ko)
    ours, ref = ctx.download_volume(), rv.download()
    assert ref[..., 1].max() >= 3
    assert np.array_equal(ours, ref)


def test_integrate_bit_exact_with_holes_and_rotation(kfo, kfb, kref):
    dims = 128
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, dims, 640, 480)
    ctx = kfb.Context(Kb, Pb)
    rv = kref.RefVolume(dims)
    rv.upload(np.zeros((dims, dims, dims, 2), np.int16))
    volpose = np.array(Po.volu_pose, np.float32)
    rng = np.random.default_rng(2)
    for k in (40, 75, 110):
        pose = kfo.trajectory_pose(k)
        # extra roll so voxel rows are not aligned with image rows
        roll = kfo.pose_apply_increment(kfo.identity(), np.array([0, 0, 0.3, 0, 0, 0.0]))
        pose = kfo.pose_mul(pose, roll)
        dm = kfo.frontend(kfo.render_depth_mm(pose, Ko), Ko, levels=1)[0][0]
        dm[rng.random(dm.shape) < 0.05] = 0.0
        v2c = kfo.pose_mul(kfo.pose_inv(pose), volpose)
        ctx.upload_depth_m(0, dm)
        ctx.integrate(v2c)
        rv.integrate(v2c, dm, Ko)
    assert np.array_equal(ctx.download_volume(), rv.download())


def test_raycast_bit_exact(kfo, kfb, kref):
    dims = 128
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, dims, 640, 480)
    ctx = kfb.Context(Kb, Pb)
    rv = kref.RefVolume(dims)
    rv.upload(np.zeros((dims, dims, dims, 2), np.int16))
    volpose = np.array(Po.volu_pose, np.float32)
    for pose, dm in _frames(kfo, Ko, (0, 5, 10)):
        rv.integrate(kfo.pose_mul(kfo.pose_inv(pose), volpose), dm, Ko)
    vol = rv.download()
    ctx.upload_volume(vol)
    for k in (8, 60):
        pose = kfo.trajectory_pose(k)
        c2v = kfo.pose_mul(kfo.pose_inv(volpose), pose)
        rinv = kfo.rot_inv(c2v)
        rv_v, rv_n, _ = rv.raycast(c2v, rinv, Ko)
        ctx.raycast(c2v, rinv)
        gv, gn = ctx.download_maps(1, 0)
        assert (rv_v[..., 2] != 0).mean() > 0.8
        assert np.array_equal(gv, rv_v) and np.array_equal(gn, rv_n)
        # model pyramid
        ctx.model_pyramid()
        s_v, s_n = kref.resize_maps(rv_v, rv_n)
        g1v, g1n = ctx.download_maps(1, 1)
        assert np.array_equal(g1v, s_v) and np.array_equal(g1n, s_n)


def test_vertex_normal_bit_exact(kfo, kfb, kref):
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, 64, 640, 480)
    d = kfo.render_depth_mm(kfo.trajectory_pose(33), Ko)
    d[200:220, 300:340] = 0
    ctx = kfb.Context(Kb, Pb)
    ctx.upload_depth_mm(d)
    ctx.frontend()
    for l in range(3):
        dm = ctx.download_depth(l)           # the product's own filtered depth feeds both sides
        Kl = Ko.level(l)
        rv, rn = kref.vertex_normal(dm, Kl)
        gv, gn = ctx.download_maps(0, l)
        assert np.array_equal(gv, rv)
        assert np.array_equal(gn, rn, equal_nan=True)


def test_depth_truncation_matches(kfo, kfb, kref):
    Ko = kfo.intr()
    d = kfo.render_depth_mm(kfo.identity(), Ko)
    d[0, :50] = 5001.0
    filt = kfo.bilateral(d)
    assert np.array_equal(kref.depth_truncation(filt), kfo.truncate(filt))


def test_icp_vs_reference_kernels(kfo, kfb, kref):
    """Level 0 only: the reference's scratch indexing overruns its allocation at levels 1-2 (§9 Q11)."""
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, 64, 640, 480)
    cur = kfo.frontend(kfo.render_depth_mm(kfo.trajectory_pose(7), Ko), Ko, levels=1)[0]
    pre = kfo.frontend(kfo.render_depth_mm(kfo.trajectory_pose(6), Ko), Ko, levels=1)[0]
    ctx = kfb.Context(Kb, Pb)
    ctx.upload_maps(0, 0, cur[1], cur[2])
    ctx.upload_maps(1, 0, pre[1], pre[2])
    pose = kfo.pose_apply_increment(kfo.identity(), np.array([1e-3, -2e-3, 5e-4, 2e-3, -1e-3, 1e-3]))
    A, b, _ = kref.rigid_icp(cur[1], cur[2], pre[1], pre[2], Ko, pose)
    want = kref.sums27_from_Ab(A, b)
    got = ctx.icp_accumulate(0, pose)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6 * np.abs(want).max())
    # and the oracle's model of the reference's f32 tile rounding is exact
    orc, _ = kfo.icp_accumulate(cur[1], cur[2], pre[1], pre[2], Ko, pose)
    np.testing.assert_allclose(orc, want, rtol=1e-6, atol=1e-7 * np.abs(want).max())


def test_extract_points_same_set(kfo, kfb, kref):
    dims = 128
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, dims, 640, 480)
    ctx = kfb.Context(Kb, Pb)
    rv = kref.RefVolume(dims)
    rv.upload(np.zeros((dims, dims, dims, 2), np.int16))
    volpose = np.array(Po.volu_pose, np.float32)
    for pose, dm in _frames(kfo, Ko, (0, 20)):
        rv.integrate(kfo.pose_mul(kfo.pose_inv(pose), volpose), dm, Ko)
    ctx.upload_volume(rv.download())
    ref_pts = rv.extract_points(volpose)
    our_pts = ctx.extract_points(volpose)
    assert len(ref_pts) == len(our_pts) > 1000
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(key(ref_pts), key(our_pts))


def test_render_matches(kfo, kfb, kref):
    dims = 128
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, dims, 640, 480)
    ctx = kfb.Context(Kb, Pb)
    vd = kfo.volume_desc(dims)
    vol = kfo.new_volume(vd)
    volpose = np.array(Po.volu_pose, np.float32)
    for pose, dm in _frames(kfo, Ko, (0, 5)):
        kfo.integrate(vol, vd, kfo.pose_mul(kfo.pose_inv(pose), volpose), dm, Ko)
    ctx.upload_volume(vol)
    pose = kfo.trajectory_pose(3)
    c2v = kfo.pose_mul(kfo.pose_inv(volpose), pose)
    ctx.raycast(c2v, kfo.rot_inv(c2v))
    gv, gn = ctx.download_maps(1, 0)
    eye = pose[[3, 7, 11]]
    for phong in (True, False):
        ref = kref.render(gv, gn, eye, phong)
        ours = ctx.render_phong(eye) if phong else ctx.render_normals()
        d = np.abs(ref.astype(np.int16) - ours.astype(np.int16))
        assert d.max() <= 1 and (d > 0).mean() < 0.01   # powf ulp at a truncating cast


def test_reference_frame_loop_tracks_like_ours(kfo, kfb, kref):
    """The `--impl reference` baseline driver (the reference's kernels under its own frame loop, oracle/
    ref_harness.cu ref_kinfu_*) and the product see the same frames: poses must agree to the 1e-4 budget."""
    Ko = kfo.intr()
    Kb = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    dims = 128
    hp = kfb.default_host_params(dims)
    ours = kfb.KinectFusion(Kb, hp)
    ref = kref.RefKinfu(Kb, dims, np.array(hp.volu_pose, np.float32))
    for k in range(6):
        d = kfo.render_depth_mm(kfo.trajectory_pose(k), Ko)
        assert ours.pipeline(d) == 0
        assert ref.pipeline(d) == 0
        assert np.abs(ours.pose() - ref.pose()).max() < 1e-4, (k, ours.pose(), ref.pose())
        # (at 128^3 the reference algorithm itself drifts by centimetres from the ground truth: 23 mm voxels
        # and the raycast sign quirk, SURVEY.md 9 Q17 -- both implementations drift together)
        assert np.abs(ours.pose() - kfo.trajectory_pose(k)).max() < 0.1


def _rot(axis, deg):
    a = np.deg2rad(deg)
    c, s = np.cos(a), np.sin(a)
    R = {"x": [[1, 0, 0], [0, c, -s], [0, s, c]], "y": [[c, 0, s], [0, 1, 0], [-s, 0, c]], "z": [[c, -s, 0], [s, c, 0], [0, 0, 1]]}[axis]
    return np.array(R, np.float64)


def _pose12(R, t):
    return np.concatenate([R, np.asarray(t, np.float64)[:, None]], axis=1).astype(np.float32).reshape(12)


@pytest.mark.parametrize("sensor,case", [
    ("kinect2", "yaw40_pitch25"),        # large rotation: every frustum plane changes its z-direction
    ("realsense720", "roll90"),          # image rows along the volume's y axis, 1280x720
    ("kinect1", "inside_volume"),        # camera inside the volume: vc.z <= 0 and ~0 planes (generic path)
    ("kinect1", "looking_back"),         # volume traversed against its z axis (vc.z decreases with z)
    ("kinect2", "wide_fov"),             # fx = 80: whole volume in view (dense-update micro-config geometry)
])
def test_integrate_raycast_bit_exact_geometries(kfo, kfb, kref, sensor, case):
    """Culling (frustum interval, occlusion cut, column states, generic path) must never change a result:
    volumes and raycasts equal the reference kernels' bit for bit under awkward geometries and all sensors."""
    dims = 128
    kw = dict(kfb.SENSORS[sensor])
    if case == "wide_fov":
        kw.update(fx=80.0, fy=80.0)
    Ko, Kb = kfo.Intr(**kw), kfb.Intrinsics(**kw)
    Pb = kfb.default_params(dims)
    volpose = np.array(kfo.default_params(dims).volu_pose, np.float32)
    if case == "yaw40_pitch25":
        cams = [_pose12(_rot("y", 40) @ _rot("x", 25), [-1.0, 0.6, 0.2]), _pose12(_rot("y", -35) @ _rot("x", -20), [0.9, -0.5, 0.3])]
    elif case == "roll90":
        cams = [_pose12(_rot("z", 90), [0.05, 0.0, 0.0]), _pose12(_rot("z", 93) @ _rot("y", 4), [0.0, 0.05, 0.02])]
    elif case == "inside_volume":
        cams = [_pose12(_rot("y", 10), [0.1, 0.0, 1.0]), _pose12(_rot("x", -8), [0.0, 0.1, 1.2])]
    elif case == "looking_back":
        cams = [_pose12(_rot("y", 180), [0.0, 0.0, 4.0]), _pose12(_rot("y", 172) @ _rot("x", 5), [0.1, 0.0, 3.9])]
    else:
        cams = [kfo.identity(), kfo.trajectory_pose(20)]
    ctx = kfb.Context(Kb, Pb)
    rv = kref.RefVolume(dims)
    rv.upload(np.zeros((dims, dims, dims, 2), np.int16))
    rng = np.random.default_rng(11)
    for cam in cams:
        d = kfo.render_depth_mm(cam, Ko)
        if case == "wide_fov":
            d[:] = 4000.0
        dm = kfo.frontend(d, Ko, levels=1)[0][0]
        dm[rng.random(dm.shape) < 0.02] = 0.0
        v2c = kfo.pose_mul(kfo.pose_inv(cam), volpose)
        ctx.upload_depth_m(0, dm)
        U = ctx.integrate(v2c, count=True)
        rv.integrate(v2c, dm, Ko)
        ours, ref = ctx.download_volume(), rv.download()
        assert np.array_equal(ours, ref), (case, U)
        assert U == int((ref[..., 1] > 0).sum()) or len(cams) > 1
    assert (ref[..., 1] > 0).sum() > 1000
    for cam in cams:
        c2v = kfo.pose_mul(kfo.pose_inv(volpose), cam)
        rinv = kfo.rot_inv(c2v)
        ctx.raycast(c2v, rinv)
        gv, gn = ctx.download_maps(1, 0)
        wv, wn, _ = rv.raycast(c2v, rinv, Ko)
        assert np.array_equal(gv.view(np.int32), wv.view(np.int32)), case
        assert np.array_equal(gn.view(np.int32), wn.view(np.int32)), case


def test_ragged_sizes_bit_exact(kfo, kfb, kref):
    """Image sizes that are no multiple of any tile (331 x 250) and a volume whose size is no power of two
    (100^3, 13 bricks per axis with a partial last brick): front end against the oracle, integrate / raycast /
    model pyramid against the reference kernels, bit for bit."""
    w, h, dims = 331, 250, 100
    s = w / 640.0
    kw = dict(width=w, height=h, fx=525.0 * s, fy=525.0 * s, cx=(319.5 + 0.5) * s - 0.5, cy=(239.5 + 0.5) * h / 480.0 - 0.5)
    Ko, Kb = kfo.Intr(**kw), kfb.Intrinsics(**kw)
    Pb = kfb.default_params(dims)
    volpose = np.array(kfo.default_params(dims).volu_pose, np.float32)
    ctx = kfb.Context(Kb, Pb)
    rv = kref.RefVolume(dims)
    rv.upload(np.zeros((dims, dims, dims, 2), np.int16))
    for k in (0, 9, 18):
        cam = kfo.trajectory_pose(k)
        d = kfo.render_depth_mm(cam, Ko)
        lv = kfo.frontend(d, Ko)
        ctx.upload_depth_mm(d)
        ctx.frontend()
        for l in range(3):
            assert np.array_equal(ctx.download_raw_depth(l), lv[l][3]) if len(lv[l]) > 3 else True
            gd = ctx.download_depth(l)
            assert np.array_equal(gd == 0, lv[l][0] == 0)
            assert np.abs(gd - lv[l][0]).max() < 2e-6
        dm = ctx.download_depth(0)
        v2c = kfo.pose_mul(kfo.pose_inv(cam), volpose)
        ctx.integrate(v2c)
        rv.integrate(v2c, dm, Ko)
    assert np.array_equal(ctx.download_volume(), rv.download())
    c2v = kfo.pose_mul(kfo.pose_inv(volpose), kfo.trajectory_pose(12))
    rinv = kfo.rot_inv(c2v)
    ctx.raycast(c2v, rinv)
    ctx.model_pyramid()
    gv, gn = ctx.download_maps(1, 0)
    wv, wn, _ = rv.raycast(c2v, rinv, Ko)
    assert (wv[..., 2] != 0).mean() > 0.5
    assert np.array_equal(gv.view(np.int32), wv.view(np.int32)) and np.array_equal(gn.view(np.int32), wn.view(np.int32))
    v1, n1 = kref.resize_maps(wv, wn)
    g1v, g1n = ctx.download_maps(1, 1)
    assert np.array_equal(g1v.view(np.int32), v1.view(np.int32)) and np.array_equal(g1n.view(np.int32), n1.view(np.int32))
