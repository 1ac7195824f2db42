"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares, fails loudly without a GPU, and the host facade's CPU-side logic (6x6 solve,
parameter defaults) matches the oracle's restatement."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(kfb):
    lib = kfb.load_library()
    names = kfb.exported_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/kfb200.h but not exported"
    # and nothing torch-/C++-typed leaks through the boundary: all symbols are unmangled C
    out = subprocess.check_output(["nm", "-D", "--defined-only", kfb.library_path()]).decode()
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported


def test_header_cites_reference_interfaces():
    txt = open(os.path.join(ROOT, "include", "kfb200.h")).read()
    for cite in ("device_types.hpp:113-128", "tsdf_volume.cu:103-111", "rigid_icp.cu:135-169", "kinectfusion.cpp:54-75"):
        assert cite in txt


def test_defaults_match_reference(kfb):
    p = kfb.default_params(512)
    assert p.pyramid_height == 3 and p.bfilter_kernel_size == 5
    assert list(p.icp_iter_count)[:3] == [4, 5, 10]          # kinectfusion.cpp:179
    assert abs(p.volu_trun_dist - np.float32(2.1) * np.float32(3.0) / np.float32(512)) < 1e-9
    assert p.tsdf_max_weight == 64 and p.dfilter_dist == 5.0
    hp = kfb.default_host_params(512)
    assert list(hp.volu_pose) == [1, 0, 0, -1.5, 0, 1, 0, -1.5, 0, 0, 1, 0.5]  # kinectfusion.cpp:184


def test_level_intrinsics_match_oracle(kfb, kfo):
    Kb = kfb.Intrinsics(**kfb.SENSORS["kinect2"])
    Ko = kfo.Intr(**kfo.SENSORS["kinect2"])
    for l in range(4):
        a, b = Kb.level(l), Ko.level(l)
        assert (a.width, a.height, a.fx, a.fy, a.cx, a.cy) == (b.width, b.height, b.fx, b.fy, b.cx, b.cy)


def test_invalid_arguments_are_rejected(kfb):
    lib = kfb.load_library()
    K = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    p = kfb.default_params(512)
    h = C.c_void_p()
    p.volu_dims[0] = 510  # not a multiple of 4: 128-bit voxel rows impossible
    assert lib.kfb_create(C.byref(K), C.byref(p), 0, C.byref(h)) == 3 and not h
    p = kfb.default_params(512)
    p.pyramid_height = 0
    assert lib.kfb_create(C.byref(K), C.byref(p), 0, C.byref(h)) == 1


def test_no_cpu_fallback(kfb):
    """Without a GPU the context cannot be created and says why (no silent CPU path)."""
    lib = kfb.load_library()
    if lib.kfb_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(kfb.KfbError):
        kfb.Context(kfb.Intrinsics(**kfb.SENSORS["kinect1"]), kfb.default_params(64))
    with pytest.raises(kfb.KfbError):
        kfb.KinectFusion(kfb.Intrinsics(**kfb.SENSORS["kinect1"]), kfb.default_host_params(64))


def test_product_does_not_import_oracle():
    """The product tree must not reference oracle/ (only tests, smoke and bench may)."""
    pkg = os.path.join(ROOT, "slam-kinectfusion_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "kf_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(d, f)


def test_host_solve_matches_oracle(kfb, kfo):
    from slam_kinectfusion_b200 import host
    rng = np.random.default_rng(11)
    for _ in range(20):
        J = rng.standard_normal((50, 6))
        r = rng.standard_normal(50)
        A, b = J.T @ J, J.T @ r
        s27 = np.zeros(27)
        s = 0
        for i in range(6):
            for j in range(i, 7):
                s27[s] = b[i] if j == 6 else A[i, j]
                s += 1
        rc_h, x_h = host.icp_solve(s27)
        rc_o, x_o = kfo.icp_solve(s27)
        assert rc_h == rc_o == 0
        # the facade factors A = L D L^T with reciprocals (the sequence the device mirrors for its pose
        # prediction), the oracle uses the textbook Cholesky: same solution to rounding
        np.testing.assert_allclose(x_h, x_o, rtol=1e-9, atol=1e-13)
    assert host.icp_solve(np.zeros(27))[0] == 1


def test_read_intrinsics_follows_reference_rule(tmp_path, kfb):
    """depth_sensor.cpp:23-46: nine numbers of a 3x3 K, those > 0.1 in reading order are fx, cx, fy, cy, scale."""
    from slam_kinectfusion_b200 import host
    f = tmp_path / "intr.txt"
    f.write_text("525.0 0 319.5\n0 525.0 239.5\n0 0 1\n")
    got = host.read_intrinsics(f)
    assert got is not None and np.allclose(got, [525.0, 319.5, 525.0, 239.5, 1.0])
    f.write_text("525.0 0 319.5\n")
    assert host.read_intrinsics(f) is None
    assert host.read_intrinsics(tmp_path / "missing.txt") is None
