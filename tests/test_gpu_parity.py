"""GPU parity tests: every stage of the hot path through the C-ABI (libkfb200.so) against the CPU
oracle on identical seeded inputs.  Each stage is fed oracle-produced inputs so errors do not
cascade.  Tolerances (BASELINE.json): pyrDown and validity masks bit-exact, int16
TSDF/weight within 1 LSB, ICP sums 1e-5 relative, raycast vertices within 1 mm, poses within
1e-4 m / 1e-4 rad.  GPU-vs-oracle differences come only from MUFU.RCP (the oracle uses the exactly
rounded reciprocal); the bit-exact GPU-vs-reference-kernel comparison is tests/test_ref_ab.py."""
import os

import numpy as np
import pytest

from conftest import ROOT, make_pair, lsb_stats

pytestmark = pytest.mark.gpu


def _ctx(kfb, Kb, Pb):
    return kfb.Context(Kb, Pb)


def _scene_frames(kfo, Ko, ks=(0, 6)):
    return [kfo.render_depth_mm(kfo.trajectory_pose(k), Ko) for k in ks]


def _nan_mask_equal(a, b):
    """Validity masks of normal maps: NaN = invalid (§9 Q6), all-zero pixel = never-written border."""
    za, zb = (a == 0).all(axis=-1), (b == 0).all(axis=-1)
    border = np.ones(za.shape, bool)
    border[1:-1, 1:-1] = False
    return (np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(za & border, zb & border)
            and za[border].all() and zb[border].all())


# ------------------------------------------------------------------------------- front end
@pytest.mark.parametrize("sensor", ["kinect1", "kinect2", "realsense720"])
def test_frontend_all_sensors(kfo, kfb, sensor):
    Ko, Kb = kfo.Intr(**kfo.SENSORS[sensor]), kfb.Intrinsics(**kfb.SENSORS[sensor])
    d = kfo.render_depth_mm(kfo.trajectory_pose(9), Ko)
    d[100:110, 200:230] = 0.0       # holes
    d[5, 7] = 7000.0                # beyond the 5 m cut
    ctx = _ctx(kfb, Kb, kfb.default_params(64))
    ctx.upload_depth_mm(d)
    ctx.frontend()
    want = kfo.frontend(d, Ko)
    for l in range(3):
        got_d = ctx.download_depth(l)
        wd, wv, wn = want[l]
        # zero / non-zero mask bit-exact; values within expf's device-vs-host ulp class
        assert np.array_equal(got_d == 0, wd == 0)
        np.testing.assert_allclose(got_d, wd, rtol=3e-6, atol=1e-7)
        gv, gn = ctx.download_maps(0, l)
        assert _nan_mask_equal(gn, wn)   # NaN = invalid normal, 0 = border (§9 Q6), bit-exact masks


def test_u16_ingest_equals_f32_ingest(kfo, kfb):
    """The 16-bit millimetre upload (sensor-native) gives exactly the front end of the f32 upload."""
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, 64, 640, 480)
    d = kfo.render_depth_mm(kfo.trajectory_pose(3), Ko)
    a, b = _ctx(kfb, Kb, Pb), _ctx(kfb, Kb, Pb)
    a.upload_depth_mm(d)
    a.frontend()
    b.upload_depth_mm_u16(d.astype(np.uint16))
    b.frontend()
    for l in range(3):
        assert np.array_equal(a.download_raw_depth(l), b.download_raw_depth(l))
        assert np.array_equal(a.download_depth(l), b.download_depth(l))
        va, na = a.download_maps(0, l)
        vb, nb = b.download_maps(0, l)
        assert np.array_equal(va.view(np.int32), vb.view(np.int32)) and np.array_equal(na.view(np.int32), nb.view(np.int32))


def test_pyrdown_bit_exact(kfo, kfb):
    """pyrDown (raw millimetre chain, REFLECT_101) is bit-exact against the oracle's FMA chain."""
    for sensor in ("kinect1", "kinect2"):
        Ko, Kb = kfo.Intr(**kfo.SENSORS[sensor]), kfb.Intrinsics(**kfb.SENSORS[sensor])
        d = kfo.render_depth_mm(kfo.trajectory_pose(3), Ko)
        rng = np.random.default_rng(7)
        d = d + rng.integers(-3, 4, d.shape).astype(np.float32)   # make the Gaussian taps non-trivial
        d[40:60, 100:140] = 0.0                                    # zeros blend in (plain Gaussian, §9 Q2)
        ctx = _ctx(kfb, Kb, kfb.default_params(64))
        ctx.upload_depth_mm(d)
        ctx.frontend()
        raw1 = kfo.pyrdown(d)
        raw2 = kfo.pyrdown(raw1)
        assert np.array_equal(ctx.download_raw_depth(0), d)
        assert np.array_equal(ctx.download_raw_depth(1), raw1)
        assert np.array_equal(ctx.download_raw_depth(2), raw2)


def test_vertex_normal_values(kfo, kfb):
    """Values of the front-end maps against the oracle.  The bilateral's 13 expf() taps differ by ulps
    between device and host libm, so depth agrees to ~1e-6 relative (not bit for bit; the bit-exact
    check of vertex/normal arithmetic is tests/test_ref_ab.py::test_vertex_normal_bit_exact)."""
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, 64, 640, 480)
    d = _scene_frames(kfo, Ko, (12,))[0]
    ctx = _ctx(kfb, Kb, Pb)
    ctx.upload_depth_mm(d)
    ctx.frontend()
    want = kfo.frontend(d, Ko)
    for l in range(3):
        got_d = ctx.download_depth(l)
        np.testing.assert_allclose(got_d, want[l][0], rtol=2e-6)
        gv, gn = ctx.download_maps(0, l)
        wv, wn = want[l][1], want[l][2]
        np.testing.assert_allclose(gv, wv, rtol=3e-6, atol=1e-7)
        assert _nan_mask_equal(gn, wn)
        ok = ~np.isnan(wn[..., 0])
        err = np.abs(gn - wn)[ok]
        # a 1-ulp depth change tilts a 2-pixel-baseline normal by ~1e-4
        assert np.percentile(err, 99.9) < 2e-3 and np.median(err) < 1e-5


# ------------------------------------------------------------------------------- integrate
def _integrate_case(kfo, kfb, dims, w, h, frames=(0, 6, 12), check=True):
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, dims, w, h)
    ctx = _ctx(kfb, Kb, Pb)
    vd = kfo.volume_desc(dims)
    vol = kfo.new_volume(vd)
    volpose = np.array(Po.volu_pose, np.float32)
    stats = []
    for k in frames:
        pose = kfo.trajectory_pose(k)
        dm = kfo.frontend(kfo.render_depth_mm(pose, Ko), Ko, levels=1)[0][0]
        v2c = kfo.pose_mul(kfo.pose_inv(pose), volpose)
        U_o = kfo.integrate(vol, vd, v2c, dm, Ko)
        ctx.upload_depth_m(0, dm)
        U_g = ctx.integrate(v2c, count=True)
        got = ctx.download_volume()
        ex, mx, out = lsb_stats(got[..., 0], vol[..., 0])
        wdiff = int((got[..., 1] != vol[..., 1]).sum())
        stats.append((k, U_o, U_g, ex, mx, out, wdiff))
        # keep both sides in lock-step so a rare pixel-rounding flip does not accumulate
        ctx.upload_volume(vol)
    return stats


def test_integrate_small(kfo, kfb):
    for k, U_o, U_g, ex, mx, out, wdiff in _integrate_case(kfo, kfb, 64, 320, 240):
        assert U_o > 0
        assert abs(U_o - U_g) <= max(4, U_o // 20000), (k, U_o, U_g)   # pixel-rounding / gate flips only
        assert ex > 0.999 and out <= max(4, U_o // 20000), (k, ex, mx, out)
        assert wdiff <= max(4, U_o // 20000)


def test_integrate_256(kfo, kfb):
    for k, U_o, U_g, ex, mx, out, wdiff in _integrate_case(kfo, kfb, 256, 640, 480, frames=(0, 30)):
        assert abs(U_o - U_g) <= U_o // 20000 + 4
        assert ex > 0.999 and out <= U_o // 20000 + 4 and wdiff <= U_o // 20000 + 4


def test_integrate_plane0_untouched_and_weight_cap(kfo, kfb):
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, 64, 320, 240, tsdf_max_weight=3)
    ctx = _ctx(kfb, Kb, Pb)
    volpose = np.array(kfo.default_params(64).volu_pose, np.float32)
    pose = kfo.identity()
    dm = kfo.frontend(kfo.render_depth_mm(pose, Ko), Ko, levels=1)[0][0]
    v2c = kfo.pose_mul(kfo.pose_inv(pose), volpose)
    ctx.upload_depth_m(0, dm)
    for _ in range(5):
        ctx.integrate(v2c)
    got = ctx.download_volume()
    assert not got[0].any()                       # z = 0 never updated (tsdf_volume.cu:53-56)
    assert got[..., 1].max() == 3                 # weight cap (made live, default 64)
    vd = kfo.volume_desc(64, max_weight=3)
    vol = kfo.new_volume(vd)
    for _ in range(5):
        kfo.integrate(vol, vd, v2c, dm, Ko)
    ex, mx, out = lsb_stats(got[..., 0], vol[..., 0])
    assert mx <= 5 and out < 50                   # 5 un-synchronised repeats: 1 LSB per step at most


def test_integrate_invalid_depth_and_empty(kfo, kfb):
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, 64, 320, 240)
    ctx = _ctx(kfb, Kb, Pb)
    volpose = np.array(Po.volu_pose, np.float32)
    ctx.upload_depth_m(0, np.zeros((240, 320), np.float32))
    assert ctx.integrate(volpose, count=True) == 0
    assert not ctx.download_volume().any()
    d = np.full((240, 320), np.nan, np.float32)
    ctx.upload_depth_m(0, d)
    assert ctx.integrate(volpose, count=True) == 0   # NaN depth fails `sdf >= -trunc`
    # camera behind the volume looking away: nothing in front
    away = kfo.identity()
    away[11] = 10.0
    ctx.upload_depth_m(0, np.full((240, 320), 2.0, np.float32))
    assert ctx.integrate(kfo.pose_mul(kfo.pose_inv(away), volpose), count=True) == 0


def test_integrate_dense_microconfig_property(kfo, kfb):
    """Dense-update micro-config: wide camera, wall behind the volume => every voxel with
    z >= 1 is updated with tsdf == 1: U = X*Y*(Z-1), stored value 32767 after one pass, weights 1."""
    dims = 128
    ki = dict(width=640, height=480, fx=80.0, fy=80.0, cx=319.5, cy=239.5)
    Kb = kfb.Intrinsics(**ki)
    Pb = kfb.default_params(dims)
    ctx = _ctx(kfb, Kb, Pb)
    ctx.upload_depth_m(0, np.full((480, 640), 4.0, np.float32))
    volpose = np.array(kfo.default_params(dims).volu_pose, np.float32)
    U = ctx.integrate(volpose, count=True)
    assert U == dims * dims * (dims - 1)
    got = ctx.download_volume()
    assert (got[1:, ..., 0] == 32767).all() and (got[1:, ..., 1] == 1).all() and not got[0].any()
    # second pass: 32767 decodes to 0.99999964, averages with 1.0 and re-encodes as 32766 (§9 Q15)
    ctx.integrate(volpose)
    got = ctx.download_volume()
    assert (got[1:, ..., 0] == 32766).all() and (got[1:, ..., 1] == 2).all()


# ------------------------------------------------------------------------------- raycast
def _build_volume(kfo, Ko, dims, ks=(0, 4, 8)):
    vd = kfo.volume_desc(dims)
    vol = kfo.new_volume(vd)
    volpose = np.array(kfo.default_params(dims).volu_pose, np.float32)
    for k in ks:
        pose = kfo.trajectory_pose(k)
        dm = kfo.frontend(kfo.render_depth_mm(pose, Ko), Ko, levels=1)[0][0]
        kfo.integrate(vol, vd, kfo.pose_mul(kfo.pose_inv(pose), volpose), dm, Ko)
    return vd, vol, volpose


@pytest.mark.parametrize("compat", [1, 0])
def test_raycast_vs_oracle(kfo, kfb, compat):
    dims = 128
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, dims, 320, 240, compat_raycast_ts_sign=compat)
    vd, vol, volpose = _build_volume(kfo, Ko, dims)
    ctx = _ctx(kfb, Kb, Pb)
    ctx.upload_volume(vol)
    pose = kfo.trajectory_pose(10)
    c2v = kfo.pose_mul(kfo.pose_inv(volpose), pose)
    wv, wn, steps = kfo.raycast(vol, vd, c2v, Ko, compat)
    ctx.raycast(c2v, kfo.rot_inv(c2v))
    ctx.model_pyramid()
    gv, gn = ctx.download_maps(1, 0)
    hit_w = wv[..., 2] != 0
    hit_g = gv[..., 2] != 0
    assert hit_w.mean() > 0.9
    assert (hit_w != hit_g).sum() <= 8                      # rays within 1 LSB of a sign change
    both = hit_w & hit_g
    dv = np.abs(gv - wv)[both]
    dn = np.abs(gn - wn)[both]
    # a ray that resolves its crossing one sample apart moves by about a voxel: count those separately
    far = (dv.max(axis=1) > 1e-3)
    assert far.sum() <= 8
    assert dv[~far].max() < 1e-3 and np.median(dv) < 2e-6   # vertices within 1 mm (typically ~ulp)
    assert np.percentile(dn.max(axis=1), 99) < 1e-3
    assert not gv[~hit_g].any() and not gn[~hit_g].any()    # explicit zeros on a miss
    # model pyramid (K4) from the product's own level-0 maps
    ov1, on1 = kfo.resize_maps(gv, gn)
    gv1, gn1 = ctx.download_maps(1, 1)
    assert np.array_equal(gv1, ov1) and np.array_equal(gn1, on1)
    ov2, on2 = kfo.resize_maps(ov1, on1)
    gv2, gn2 = ctx.download_maps(1, 2)
    assert np.array_equal(gv2, ov2) and np.array_equal(gn2, on2)


def test_raycast_miss_and_outside(kfo, kfb):
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, 64, 320, 240)
    ctx = _ctx(kfb, Kb, Pb)
    volpose = np.array(Po.volu_pose, np.float32)
    # empty volume: no crossing anywhere
    c2v = kfo.pose_mul(kfo.pose_inv(volpose), kfo.identity())
    ctx.raycast(c2v, kfo.rot_inv(c2v))
    gv, gn = ctx.download_maps(1, 0)
    assert not gv.any() and not gn.any()
    # camera looking away from the box: rays miss the volume entirely
    away = kfo.identity()
    away[0] = -1.0
    away[10] = -1.0      # 180 degrees about y
    c2v = kfo.pose_mul(kfo.pose_inv(volpose), away)
    ctx.raycast(c2v, kfo.rot_inv(c2v))
    gv, gn = ctx.download_maps(1, 0)
    assert not gv.any() and not gn.any()


# ------------------------------------------------------------------------------- ICP
@pytest.mark.parametrize("compat_rows", [1, 0])
def test_icp_sums_vs_oracle(kfo, kfb, compat_rows):
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, 64, 640, 480, compat_icp_rows=compat_rows)
    cur = kfo.frontend(kfo.render_depth_mm(kfo.trajectory_pose(7), Ko), Ko)
    pre = kfo.frontend(kfo.render_depth_mm(kfo.trajectory_pose(6), Ko), Ko)
    ctx = _ctx(kfb, Kb, Pb)
    for l in range(3):
        ctx.upload_maps(0, l, cur[l][1], cur[l][2])
        ctx.upload_maps(1, l, pre[l][1], pre[l][2])
    pose = kfo.pose_apply_increment(kfo.identity(), np.array([1e-3, -2e-3, 5e-4, 2e-3, -1e-3, 1e-3]))
    for l in range(3):
        Kl = Ko.level(l)
        want, cnt = kfo.icp_accumulate(cur[l][1], cur[l][2], pre[l][1], pre[l][2], Kl, pose, 0.015, 0.5, compat_rows)
        got = ctx.icp_accumulate(l, pose)
        assert cnt > 1000
        scale = np.abs(want).max()
        # relative 1e-5 per entry against the dominant magnitude; a handful of correspondences sit within
        # 1 ulp of a gate and may flip (MUFU.RCP vs exact reciprocal)
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-5 * scale)
        rc_g, x_g = kfo.icp_solve(got)
        rc_w, x_w = kfo.icp_solve(want)
        assert rc_g == rc_w == 0
        np.testing.assert_allclose(x_g, x_w, atol=2e-5)
    # determinism: the single-pass grid reduction sums in a fixed order
    a = ctx.icp_accumulate(0, pose)
    b = ctx.icp_accumulate(0, pose)
    assert np.array_equal(a, b)


def test_icp_schedule_equals_direct(kfo, kfb):
    """The whole-schedule kernel (kfb_icp_begin/step/end) returns bit-identical sums to the direct call for whatever
    poses the caller passes (here the numpy oracle's solve, which need not round like the kernel's replica of the
    facade's: a pose the kernel did not predict moves the rest of the schedule to ordinary launches), survives
    early termination (tracking failure) and leaves the context reusable."""
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, 64, 640, 480)
    cur = kfo.frontend(kfo.render_depth_mm(kfo.trajectory_pose(7), Ko), Ko)
    pre = kfo.frontend(kfo.render_depth_mm(kfo.trajectory_pose(6), Ko), Ko)
    ctx = _ctx(kfb, Kb, Pb)
    for l in range(3):
        ctx.upload_maps(0, l, cur[l][1], cur[l][2])
        ctx.upload_maps(1, l, pre[l][1], pre[l][2])
    iters = [4, 5, 10]
    for trial in range(3):
        pose = kfo.identity()
        ctx.icp_begin(iters)
        direct = []
        for level in (2, 1, 0):
            for i in range(iters[level]):
                got = ctx.icp_step(pose)
                direct.append((level, pose.copy(), got))
                rc, x = kfo.icp_solve(got)
                assert rc == 0
                pose = kfo.pose_apply_increment(pose, x)
        ctx.icp_end()
        for level, p, got in direct[::4]:
            assert np.array_equal(ctx.icp_accumulate(level, p), got)
    # a pose the kernel cannot have predicted, in the middle of a schedule (fresh context: after a misprediction the
    # next schedules of a context start on ordinary launches)
    ctx = _ctx(kfb, Kb, Pb)
    for l in range(3):
        ctx.upload_maps(0, l, cur[l][1], cur[l][2])
        ctx.upload_maps(1, l, pre[l][1], pre[l][2])
    ctx.icp_begin(iters)
    pose = kfo.identity()
    before = ctx.icp_mispredict_count()
    for i in range(6):
        if i == 4:
            pose = pose.copy()
            pose[3] += np.float32(1e-3)
        got = ctx.icp_step(pose)
        last = (pose.copy(), got)
        rc, x = kfo.icp_solve(got)
        pose = kfo.pose_apply_increment(pose, x)
    ctx.icp_end()
    assert np.array_equal(ctx.icp_accumulate(2, last[0]), last[1])
    assert ctx.icp_mispredict_count() > before and ctx.icp_fallback_count() >= 1
    # early exit after 3 of 19 iterations: whatever is still running must retire by itself
    ctx.icp_begin(iters)
    for i in range(3):
        ctx.icp_step(kfo.identity())
    ctx.icp_end()
    ctx.synchronize()
    a = ctx.icp_accumulate(2, kfo.identity())
    ctx.icp_begin(iters)
    b = ctx.icp_step(kfo.identity())
    ctx.icp_end()
    ctx.synchronize()
    assert np.array_equal(a, b)


def test_icp_transport_timeout_falls_back_without_losing_the_map(kfo, kfb, monkeypatch):
    """A whole-schedule ICP kernel that gives up on a poll (here: a 2 us bound instead of 1 s) is a transport
    failure, not a tracking failure: the schedule finishes on ordinary launches with bit-identical sums, the
    facade neither resets the volume nor loses the pose history, and the poses equal an undisturbed run's."""
    Ko = kfo.intr()
    Kb = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    frames = [kfo.render_depth_mm(kfo.trajectory_pose(k), Ko) for k in range(5)]

    def run():
        kf = kfb.KinectFusion(Kb, kfb.default_host_params(64))
        for d in frames:
            assert kf.pipeline(d) == 0
        return kf, np.stack([np.asarray(p) for p in kf.poses()])

    first, want = run()
    monkeypatch.setenv("KFB_ICP_TIMEOUT_NS", "2000")
    kf, got = run()
    monkeypatch.delenv("KFB_ICP_TIMEOUT_NS")
    assert kf.frame_count == len(frames) + 1 and len(got) == len(want)
    assert np.array_equal(got, want)
    assert first.context().icp_fallback_count() == 0 and first.context().icp_mispredict_count() == 0   # the undisturbed run
    assert kf.context().icp_fallback_count() >= 1
    assert kf.context().download_volume()[..., 1].max() == len(frames)


@pytest.mark.parametrize("switch", ["KFB_ICP_DIRECT", "KFB_ICP_PLAIN_LAUNCH"])
def test_icp_launch_modes_give_the_same_poses(kfo, kfb, monkeypatch, switch):
    """One ordinary launch per iteration (KFB_ICP_DIRECT: what every fallback ends in) and the whole-schedule kernel
    started by an ordinary launch behind an occupancy query (KFB_ICP_PLAIN_LAUNCH) track exactly like the default
    cooperative launch."""
    Ko = kfo.intr()
    Kb = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    frames = [kfo.render_depth_mm(kfo.trajectory_pose(k), Ko) for k in range(5)]

    def run():
        kf = kfb.KinectFusion(Kb, kfb.default_host_params(64))
        for d in frames:
            assert kf.pipeline(d) == 0
        return kf, np.stack([np.asarray(p) for p in kf.poses()])

    _, want = run()
    monkeypatch.setenv(switch, "1")
    kf, got = run()
    monkeypatch.delenv(switch)
    assert np.array_equal(got, want)
    assert kf.context().icp_mispredict_count() == 0
    assert kf.context().icp_fallback_count() == 0      # a forced mode is not a fallback


def test_icp_schedule_bounds(kfo, kfb):
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, 64, 320, 240)
    ctx = _ctx(kfb, Kb, Pb)
    with pytest.raises(kfb.KfbError):
        ctx.icp_begin([100, 100, 100])        # 300 iterations: more than the release tag can number
    ctx.icp_begin([1, 1, 1])
    ctx.icp_end()


def test_icp_degenerate_inputs(kfo, kfb):
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, 64, 320, 240)
    ctx = _ctx(kfb, Kb, Pb)
    nanmap = np.full((240, 320, 3), np.nan, np.float32)
    zero = np.zeros((240, 320, 3), np.float32)
    ctx.upload_maps(0, 0, zero, nanmap)       # every current normal invalid
    ctx.upload_maps(1, 0, zero, zero)
    got = ctx.icp_accumulate(0, kfo.identity())
    assert not got.any()
    assert kfo.icp_solve(got)[0] == 1         # => tracking failure upstream


# ------------------------------------------------------------------------------- golden fixture
def test_golden_fixture(kfo, kfb):
    g = np.load(os.path.join(ROOT, "tests", "golden", "oracle_golden.npz"))
    ki = dict(width=int(g["intr"][0]), height=int(g["intr"][1]), fx=float(g["intr"][2]), fy=float(g["intr"][3]),
              cx=float(g["intr"][4]), cy=float(g["intr"][5]))
    Kb = kfb.Intrinsics(**ki)
    dims = int(g["dims"])
    ctx = _ctx(kfb, Kb, kfb.default_params(dims))
    ctx.upload_depth_mm(g["depth_mm"])
    ctx.frontend()
    for l in range(3):
        got = ctx.download_depth(l)
        assert np.array_equal(got == 0, g[f"depth_l{l}"] == 0)
        np.testing.assert_allclose(got, g[f"depth_l{l}"], rtol=3e-6)
        _, gn = ctx.download_maps(0, l)
        assert _nan_mask_equal(gn, g[f"nmap_l{l}"])
    ctx.upload_depth_m(0, g["depth_l0"])
    ctx.upload_volume(g["volume_pre"])
    ctx.integrate(g["vol2cam"])
    got = ctx.download_volume()
    ex, mx, out = lsb_stats(got[..., 0], g["volume"][..., 0])
    assert mx <= 1 and ex > 0.995
    assert np.array_equal(got[..., 1], g["volume"][..., 1])
    ctx.upload_volume(g["volume"])
    c2v = g["cam2vol"]
    ctx.raycast(c2v, kfo.rot_inv(c2v))
    gv, gn = ctx.download_maps(1, 0)
    hit = g["ray_v"][..., 2] != 0
    assert ((gv[..., 2] != 0) != hit).sum() <= 4
    both = hit & (gv[..., 2] != 0)
    assert np.percentile(np.abs(gv - g["ray_v"])[both], 99.5) < 1e-3
    ctx.upload_maps(0, 0, g["vmap_l0"], g["nmap_l0"])
    ctx.upload_maps(1, 0, g["icp_pre_v"], g["icp_pre_n"])
    sums = ctx.icp_accumulate(0, g["icp_pose"])
    np.testing.assert_allclose(sums, g["icp_sums"], rtol=2e-4, atol=2e-5 * np.abs(g["icp_sums"]).max())
    ctx.upload_volume(g["volume"])
    pts = ctx.extract_points(g["volpose"])
    want = g["points"]
    assert len(pts) == len(want)
    key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(key(pts), key(want))     # bit-exact as sorted sets (IEEE-only arithmetic)


# ------------------------------------------------------------------------------- whole pipeline
def test_pipeline_vs_oracle_poses(kfo, kfb):
    """End to end through the C++ host facade (kf::kinectfusion::pipeline): per-frame pose within
    1e-4 m / 1e-4 rad of the oracle pipeline on the same synthetic frames."""
    from slam_kinectfusion_b200 import host
    w, h, dims = 640, 480, 128
    Ko = kfo.intr()
    Kb = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    Po = kfo.default_params(dims)
    Ph = kfb.default_host_params(dims)
    okf = kfo.Kinfu(Ko, Po)
    gkf = host.KinectFusion(Kb, Ph)
    worst_t, worst_r = 0.0, 0.0
    for k in range(12):
        d = kfo.render_depth_mm(kfo.trajectory_pose(k), Ko)
        assert okf.pipeline(d) == 0
        assert gkf.pipeline(d) == 0
        po, pg = okf.pose().reshape(3, 4), gkf.pose().reshape(3, 4)
        worst_t = max(worst_t, np.abs(po[:, 3] - pg[:, 3]).max())
        dR = po[:, :3].astype(np.float64).T @ pg[:, :3].astype(np.float64)
        # skew part (arccos of the trace amplifies float32 rounding near the identity)
        ang = np.arcsin(min(1.0, np.linalg.norm(0.5 * np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]))))
        worst_r = max(worst_r, ang)
        assert gkf.frame_count == okf.frame_count
    assert worst_t < 1e-4 and worst_r < 1e-4, (worst_t, worst_r)
    # tracking failure => reset, like the reference (kinectfusion.cpp:97-102)
    assert gkf.pipeline(np.zeros((h, w), np.float32)) == 1
    assert gkf.frame_count == 1 and len(gkf.poses()) == 1
    # render + export run and agree with the oracle's restatement on the model maps
    assert gkf.pipeline(kfo.render_depth_mm(kfo.trajectory_pose(0), Ko)) == 0
    assert gkf.pipeline(kfo.render_depth_mm(kfo.trajectory_pose(1), Ko)) == 0
    img = gkf.render()
    assert img.shape == (h, w, 3) and img.any()
    pts = gkf.extract_pointcloud()
    assert len(pts) > 1000
    # trajectory export in the reference's stream format (main.cpp:94-98)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "poses.txt")
        assert gkf.save_poses(path) == 0
        txt = open(path).read()
        mats = [m for m in txt.replace("\n", " ").split("]") if "[" in m]
        assert len(mats) == len(gkf.poses())
        first = np.array([[float(v) for v in row.split(",")] for row in mats[0].split("[")[1].split(";")])
        assert np.array_equal(first, np.eye(4))
        # volume checkpoint: a second instance loads it and raycasts the same model
        vpath = os.path.join(td, "volume.kfb")
        assert gkf.save_volume(vpath) == 0
        other = host.KinectFusion(Kb, Ph)
        assert other.load_volume(vpath) == 0
        assert np.array_equal(other.context().download_volume(), gkf.context().download_volume())
        volpose = np.array(Ph.volu_pose, np.float32)
        c2v = kfo.pose_mul(kfo.pose_inv(volpose), gkf.pose())
        for kfx in (gkf, other):
            kfx.context().raycast(c2v, kfo.rot_inv(c2v))
        va, na = gkf.context().download_maps(1, 0)
        vb, nb = other.context().download_maps(1, 0)
        assert np.array_equal(va.view(np.int32), vb.view(np.int32)) and np.array_equal(na.view(np.int32), nb.view(np.int32))
        assert other.load_volume(os.path.join(td, "poses.txt")) == 1     # not a checkpoint


@pytest.mark.parametrize("sensor", ["kinect2", "realsense720"])
def test_pipeline_other_sensors_vs_oracle_poses(kfo, kfb, sensor):
    """BASELINE configs[2]: the other sensor resolutions the reference supports (512x424, 1280x720), whole
    pipeline through the facade against the oracle pipeline; the 1e-4 m / 1e-4 rad pose budget."""
    from slam_kinectfusion_b200 import host
    kw = kfb.SENSORS[sensor]
    Ko, Kb = kfo.Intr(**kw), kfb.Intrinsics(**kw)
    dims = 128
    okf = kfo.Kinfu(Ko, kfo.default_params(dims))
    gkf = host.KinectFusion(Kb, kfb.default_host_params(dims))
    for k in range(6):
        d = kfo.render_depth_mm(kfo.trajectory_pose(k), Ko)
        assert okf.pipeline(d) == 0 and gkf.pipeline(d) == 0
        po, pg = okf.pose().reshape(3, 4).astype(np.float64), gkf.pose().reshape(3, 4).astype(np.float64)
        assert np.abs(po[:, 3] - pg[:, 3]).max() < 1e-4
        dR = po[:, :3].T @ pg[:, :3]
        # rotation angle from the antisymmetric part (arccos of the trace loses everything below 3e-4 rad in f32)
        ang = 0.5 * np.linalg.norm([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
        assert ang < 1e-4, ang


def test_large_volume_1024_properties(kfo, kfb):
    """BASELINE configs[3] size on one GPU (1024^3 = 4 GiB packed): the update count doubles per axis as
    expected (about 8x the 512^3 anchor), the predicate is idempotent, and a raycast of the integrated room
    hits nearly everywhere and reproduces the depth it was built from to within the truncation distance."""
    Ko = kfo.intr()
    Kb = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    ctx = _ctx(kfb, Kb, kfb.default_params(1024))
    d = kfo.render_depth_mm(kfo.identity(), Ko)
    ctx.upload_depth_mm(d)
    ctx.frontend()
    volpose = np.array(kfo.default_params(1024).volu_pose, np.float32)
    U = ctx.integrate(volpose, count=True)
    assert 7.6 * 35_941_951 < U < 8.2 * 35_941_951
    assert ctx.integrate(volpose, count=True) == U
    c2v = kfo.pose_mul(kfo.pose_inv(volpose), kfo.identity())
    ctx.raycast(c2v, kfo.rot_inv(c2v))
    gv, gn = ctx.download_maps(1, 0)
    hit = gv[..., 2] != 0
    assert hit.mean() > 0.98
    err = np.abs(gv[..., 2] - ctx.download_depth(0))[hit]
    assert np.median(err) < 0.0062 and np.percentile(err, 99) < 0.03


def test_full_size_properties_512(kfo, kfb):
    """BASELINE configs[1] size (640x480, 512^3) through size-independent properties: U equals the
    SURVEY anchor, plane 0 untouched, weights == number of passing integrations, raycast of the
    integrated room hits everywhere and reproduces the depth it was built from to within trunc."""
    Ko = kfo.intr()
    Kb = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    ctx = _ctx(kfb, Kb, kfb.default_params(512))
    d = kfo.render_depth_mm(kfo.identity(), Ko)
    ctx.upload_depth_mm(d)
    ctx.frontend()
    volpose = np.array(kfo.default_params(512).volu_pose, np.float32)
    U = ctx.integrate(volpose, count=True)
    assert abs(U - 35_941_951) < 20_000                     # SURVEY §8d anchor
    U2 = ctx.integrate(volpose, count=True)
    assert U2 == U                                          # idempotent predicate
    c2v = kfo.pose_mul(kfo.pose_inv(volpose), kfo.identity())
    ctx.raycast(c2v, kfo.rot_inv(c2v))
    gv, gn = ctx.download_maps(1, 0)
    hit = gv[..., 2] != 0
    assert hit.mean() > 0.98
    depth_m = ctx.download_depth(0)
    err = np.abs(gv[..., 2] - depth_m)[hit]
    assert np.median(err) < 0.0124 and np.percentile(err, 99) < 0.05   # within trunc (quirk bias <= 2 voxels); silhouettes looser
    nrm = np.linalg.norm(gn[hit], axis=1)
    assert np.abs(nrm - 1).max() < 1e-3


def test_integrate_jump_equals_replay(kfo, kfb):
    """The closed-form jump of the float running sum (kfb_common.cuh: jump_fma) against plain replay
    (KFB_INTEGRATE_NOJUMP=1): volumes must be bit-identical for poses whose vc.x / vc.y cross zero and
    binade boundaries inside the skipped prefix, for several plan chunk heights and for far z-slabs (whose states
    kernel jumps over the prefix in front of the slab)."""
    dims = 256
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, dims, 320, 240)
    volpose = np.array(Po.volu_pose, np.float32)
    rng = np.random.default_rng(7)
    poses = [kfo.trajectory_pose(k) for k in (0, 40, 75)]
    for _ in range(3):                      # extra rotations (a few degrees about random axes) + offsets
        w = rng.normal(size=3); w *= 0.12 / np.linalg.norm(w)
        th = np.linalg.norm(w); kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
        R = np.eye(3) + np.sin(th) * kx + (1 - np.cos(th)) * kx @ kx
        t = rng.normal(size=3) * 0.05
        poses.append(np.concatenate([R, t[:, None]], axis=1).astype(np.float32).reshape(12))
    Pslab = kfb.default_params(dims)
    Pslab.slab_z_begin, Pslab.slab_z_end = 192, 256
    Pmid = kfb.default_params(dims)
    Pmid.slab_z_begin, Pmid.slab_z_end = 61, 131       # not brick-aligned
    saved = {k: os.environ.get(k) for k in ("KFB_INTEGRATE_NOJUMP", "KFB_PLAN_ZCHUNK", "KFB_INTEGRATE_JUMPMIN")}
    os.environ["KFB_INTEGRATE_JUMPMIN"] = "16"
    try:
        for P, chunks in ((Pb, "8"), (Pslab, "8"), (Pslab, "16"), (Pmid, "5")):
            vols = []
            for nojump in (False, True):
                if nojump:
                    os.environ["KFB_INTEGRATE_NOJUMP"] = "1"
                else:
                    os.environ.pop("KFB_INTEGRATE_NOJUMP", None)
                os.environ["KFB_PLAN_ZCHUNK"] = chunks
                ctx = _ctx(kfb, Kb, P)
                for pose in poses:
                    ctx.upload_depth_mm(kfo.render_depth_mm(pose, Ko))
                    ctx.frontend()
                    ctx.integrate(kfo.pose_mul(kfo.pose_inv(pose), volpose))
                vols.append(ctx.download_volume())
                ctx.close()
            assert vols[0][..., 1].max() >= 3
            assert np.array_equal(vols[0], vols[1]), chunks
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_integrate_switches_bit_identical(kfo, kfb):
    """Every work-skipping device of the planned sweep (frustum interval, occlusion cut, stream items, chunk height,
    one or two streams) must leave the volume bit-identical: compare each switch against the plain sweep, in which
    every voxel of every plane goes through the per-voxel predicate (that run also has more general items than
    state slots, so the items that replay their running sums themselves are covered)."""
    dims = 256
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, dims, 640, 480)
    volpose = np.array(Po.volu_pose, np.float32)
    poses = [kfo.trajectory_pose(k) for k in (0, 25, 60)]
    frames = [kfo.render_depth_mm(p, Ko) for p in poses]
    keys = ("KFB_INTEGRATE_NOCULL", "KFB_INTEGRATE_NOOCC", "KFB_INTEGRATE_NOFAST", "KFB_PLAN_ZCHUNK", "KFB_INTEGRATE_SERIAL",
            "KFB_INTEGRATE_PERSISTENT", "KFB_GEN_MINB", "KFB_INTEGRATE_SPLITSTATES", "KFB_GEN_NOPREFETCH")
    saved = {k: os.environ.get(k) for k in keys}

    def run(env):
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        ctx = _ctx(kfb, Kb, Pb)
        n = 0
        for pose, d in zip(poses, frames):
            ctx.upload_depth_mm(d)
            ctx.frontend()
            n += ctx.integrate(kfo.pose_mul(kfo.pose_inv(pose), volpose), count=True)
            ctx.integrate(kfo.pose_mul(kfo.pose_inv(pose), volpose))
        vol = ctx.download_volume()
        ctx.close()
        return vol, n

    try:
        plain, n_plain = run({"KFB_INTEGRATE_NOCULL": "1", "KFB_INTEGRATE_NOFAST": "1", "KFB_PLAN_ZCHUNK": "16"})
        assert plain[..., 1].max() >= 6
        for env in ({}, {"KFB_INTEGRATE_NOFAST": "1"}, {"KFB_INTEGRATE_NOOCC": "1"}, {"KFB_PLAN_ZCHUNK": "3"},
                    {"KFB_PLAN_ZCHUNK": "32"}, {"KFB_INTEGRATE_SERIAL": "1"}, {"KFB_INTEGRATE_PERSISTENT": "1"},
                    {"KFB_GEN_MINB": "5"}, {"KFB_GEN_MINB": "6"}, {"KFB_INTEGRATE_SPLITSTATES": "1"}, {"KFB_GEN_NOPREFETCH": "1"},
                    {"KFB_GEN_NOPREFETCH": "1", "KFB_PLAN_ZCHUNK": "3"}):
            vol, n = run(env)
            assert n == n_plain, env
            assert np.array_equal(vol, plain), env
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


# ------------------------------------------------------------------------------- z-slab sharding (§8e)
@pytest.mark.parametrize("world", [2, 3])
def test_slab_contexts_compose_to_single_gpu_result(kfo, kfb, world):
    """Slab contexts (here all on one GPU) integrate only their planes (+halo) and raycast only their ray
    segments; volumes must equal the corresponding planes of the single-context volume bit for bit, and the
    min-key composite (kfb_composite_mask semantics) must equal the single-context raycast bit for bit."""
    import ctypes as C
    from slam_kinectfusion_b200 import sharded
    dims = 128
    Ko, Kb, Po, Pb = make_pair(kfo, kfb, dims, 320, 240)
    volpose = np.array(Po.volu_pose, np.float32)
    full = _ctx(kfb, Kb, Pb)
    slabs = []
    for r in range(world):
        P = kfb.default_params(dims)
        P.slab_z_begin, P.slab_z_end = sharded.slab_range(dims, world, r)
        slabs.append(_ctx(kfb, Kb, P))
    for k in (0, 4, 8):
        pose = kfo.trajectory_pose(k)
        d = kfo.render_depth_mm(pose, Ko)
        v2c = kfo.pose_mul(kfo.pose_inv(pose), volpose)
        for c in [full] + slabs:
            c.upload_depth_mm(d)
            c.frontend()
            c.integrate(v2c)
    fvol = full.download_volume()
    for r, c in enumerate(slabs):
        s0, s1 = sharded.stored_range(dims, world, r)
        assert np.array_equal(c.download_volume(), fvol[s0:s1])
    c2v = kfo.pose_mul(kfo.pose_inv(volpose), kfo.trajectory_pose(9))
    rinv = kfo.rot_inv(c2v)
    full.raycast(c2v, rinv)
    fv, fn = full.download_maps(1, 0)
    assert (fv[..., 2] != 0).mean() > 0.9
    h, w = fv.shape[:2]
    keys, maps = [], []
    for c in slabs:
        c.raycast(c2v, rinv)
        maps.append(c.download_maps(1, 0))
        k = np.empty(h * w, np.float32)
        c.synchronize()
        _cuda_memcpy_d2h(k, c.device_ptr(4))
        keys.append(k.reshape(h, w))
    keys = np.stack(keys)
    # (a) slab raycast == oracle slab semantics on the same volume (hits up to MUFU.RCP-level ties)
    for r, c in enumerate(slabs):
        s0, s1 = sharded.stored_range(dims, world, r)
        zb, ze = sharded.slab_range(dims, world, r)
        ov, on, ok = kfo.raycast_slab(fvol[s0:s1], kfo.volume_desc(dims), c2v, Ko, s0, s1, zb, ze)
        assert (np.isfinite(ok) != np.isfinite(keys[r])).sum() <= 8
        both = np.isfinite(ok) & np.isfinite(keys[r]) & (ov[..., 2] != 0) & (maps[r][0][..., 2] != 0)
        if both.any():
            assert np.median(np.abs(maps[r][0] - ov)[both]) < 2e-6
    # (b) composite through kfb_composite_mask + integer sum == single-context raycast, bit for bit
    min_key = keys.min(axis=0).astype(np.float32)
    dmin = _cuda_alloc_copy(min_key)
    acc_v = np.zeros((h, w, 3), np.int64)
    acc_n = np.zeros((h, w, 3), np.int64)
    for c in slabs:
        c.composite_mask(dmin)
        mv, mn = c.download_maps(1, 0)
        acc_v += mv.view(np.int32)
        acc_n += mn.view(np.int32)
    _cuda_free(dmin)
    assert np.array_equal(acc_v.astype(np.int32), fv.view(np.int32))
    assert np.array_equal(acc_n.astype(np.int32), fn.view(np.int32))
    srt = np.sort(keys, axis=0)
    assert not (np.isfinite(srt[0]) & (srt[0] == srt[1])).any()   # an event has exactly one owner


def _cudart():
    import ctypes as C
    for name in ("libcudart.so", "libcudart.so.12"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    import glob
    import torch
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    return C.CDLL(cands[0])


def _cuda_memcpy_d2h(arr, dptr):
    import ctypes as C
    rt = _cudart()
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    assert rt.cudaMemcpy(arr.ctypes.data_as(C.c_void_p), C.c_void_p(dptr), arr.nbytes, 2) == 0


def _cuda_alloc_copy(arr):
    import ctypes as C
    rt = _cudart()
    p = C.c_void_p()
    rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    assert rt.cudaMalloc(C.byref(p), arr.nbytes) == 0
    assert rt.cudaMemcpy(p, arr.ctypes.data_as(C.c_void_p), arr.nbytes, 1) == 0
    return p.value


def _cuda_free(p):
    import ctypes as C
    rt = _cudart()
    rt.cudaFree.argtypes = [C.c_void_p]
    rt.cudaFree(C.c_void_p(p))
