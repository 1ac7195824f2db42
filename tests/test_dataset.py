"""Dataset frame source: depth_sensor open/getFrame and the PNG decoder behind it, checked against
OpenCV's own cv2.imread on files written by cv2 (all filter types, 8/16 bit, grey / RGB / RGBA / palette)."""
import os
import struct
import zlib

import numpy as np
import pytest

import slam_kinectfusion_b200 as kfb

cv2 = pytest.importorskip("cv2")


def _dataset(tmp_path, n=3, w=64, h=48, intr="525 0 319.5\n0 525 239.5\n0 0 1\n", seed=0):
    rng = np.random.default_rng(seed)
    (tmp_path / "color").mkdir()
    (tmp_path / "depth").mkdir()
    frames = []
    for k in range(n):
        depth = rng.integers(0, 65535, (h, w), dtype=np.uint16)
        depth[:4] = 0
        bgr = rng.integers(0, 255, (h, w, 3), dtype=np.uint8)
        assert cv2.imwrite(str(tmp_path / "depth" / f"{k:05d}.png"), depth)
        assert cv2.imwrite(str(tmp_path / "color" / f"{k:05d}.png"), bgr)
        frames.append((bgr, depth))
    if intr is not None:
        (tmp_path / "intr.txt").write_text(intr)
    return frames


def test_sensor_reads_what_cv2_wrote(tmp_path):
    frames = _dataset(tmp_path)
    s = kfb.DatasetSensor(tmp_path)
    assert (s.width, s.height) == (64, 48)
    # intr.txt: numbers > 0.1 in file order = fx cx fy cy scale (depth_sensor.cpp:27-42)
    assert (s.fx, s.cx, s.fy, s.cy, s.scale) == (525.0, 319.5, 525.0, 239.5, 1.0)
    assert s.frames_left() == 3
    for k, (bgr, depth) in enumerate(s):
        assert np.array_equal(bgr, frames[k][0])
        assert depth.dtype == np.float32 and np.array_equal(depth, frames[k][1].astype(np.float32))
    assert s.frames_left() == 0 and s.get_frame() is None


def test_sensor_without_frames_or_intrinsics(tmp_path):
    with pytest.raises(FileNotFoundError):
        kfb.DatasetSensor(tmp_path)  # "error: no camera!"
    _dataset(tmp_path, n=1, intr=None)
    s = kfb.DatasetSensor(tmp_path)
    assert (s.width, s.height) == (64, 48) and s.fx == 0.0  # no intr.txt: size known, intrinsics left unset
    (tmp_path / "intr.txt").write_text("525 0 0\n0 525 0\n0 0 1\n")  # only three numbers > 0.1: rejected as in the reference
    assert kfb.read_intrinsics(tmp_path / "intr.txt") is None


@pytest.mark.parametrize("kind", ["gray8", "gray16", "rgb8", "rgba8", "rgb16", "palette", "gray1"])
def test_png_decoder_against_cv2(tmp_path, kind):
    rng = np.random.default_rng(5)
    h, w = 37, 53  # odd sizes: partial bytes for the 1-bit rows
    (tmp_path / "color").mkdir()
    (tmp_path / "depth").mkdir()
    name = str(tmp_path / "color" / "a.png")
    if kind == "gray8":
        cv2.imwrite(name, rng.integers(0, 255, (h, w), dtype=np.uint8))
    elif kind == "gray16":
        cv2.imwrite(name, rng.integers(0, 65535, (h, w), dtype=np.uint16))
    elif kind == "rgb8":
        # smooth content makes the encoder pick Sub / Up / Average / Paeth row filters
        y, x = np.mgrid[0:h, 0:w]
        cv2.imwrite(name, np.stack([x * 4, y * 5, x + y], -1).astype(np.uint8))
    elif kind == "rgba8":
        cv2.imwrite(name, rng.integers(0, 255, (h, w, 4), dtype=np.uint8))
    elif kind == "rgb16":
        cv2.imwrite(name, rng.integers(0, 65535, (h, w, 3), dtype=np.uint16))
    else:
        # cv2 cannot write these: assemble the file by hand
        if kind == "palette":
            pal = rng.integers(0, 255, (16, 3), dtype=np.uint8)
            idx = rng.integers(0, 16, (h, w), dtype=np.uint8)
            bits, ctype, extra = 4, 3, [(b"PLTE", pal.tobytes())]
        else:
            idx = rng.integers(0, 2, (h, w), dtype=np.uint8)
            bits, ctype, extra = 1, 0, []
        per = 8 // bits
        rows = b""
        for r in idx:
            padded = np.concatenate([r, np.zeros((-len(r)) % per, np.uint8)]).reshape(-1, per)
            packed = np.zeros(len(padded), np.uint8)
            for j in range(per):
                packed |= (padded[:, j] << ((per - 1 - j) * bits)).astype(np.uint8)
            rows += b"\x00" + packed.tobytes()

        def chunk(t, d):
            return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))

        png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bits, ctype, 0, 0, 0))
        for t, d in extra:
            png += chunk(t, d)
        png += chunk(b"IDAT", zlib.compress(rows)) + chunk(b"IEND", b"")
        open(name, "wb").write(png)
    cv2.imwrite(str(tmp_path / "depth" / "a.png"), np.zeros((h, w), np.uint16))
    s = kfb.DatasetSensor(tmp_path)
    bgr, _ = s.get_frame()
    assert np.array_equal(bgr, cv2.imread(name, 1)), kind  # imread(name, 1): 8-bit BGR whatever the file holds


def test_depth_must_be_single_channel_and_files_must_be_png(tmp_path):
    _dataset(tmp_path, n=1)
    cv2.imwrite(str(tmp_path / "depth" / "00000.png"), np.zeros((48, 64, 3), np.uint8))
    s = kfb.DatasetSensor(tmp_path)
    assert s.get_frame() is None and "one channel" in s.last_error()
    open(tmp_path / "color" / "00000.png", "wb").write(b"not a png at all, just bytes" * 4)
    with pytest.raises(FileNotFoundError):
        kfb.DatasetSensor(tmp_path)  # the first colour image gives the size: undecodable => no camera


def test_png_writer_round_trip(tmp_path):
    rng = np.random.default_rng(2)
    d = rng.integers(0, 65535, (30, 40), dtype=np.uint16)
    c = rng.integers(0, 255, (30, 40, 3), dtype=np.uint8)
    assert kfb.write_png_gray16(tmp_path / "d.png", d) and kfb.write_png_rgb8(tmp_path / "c.png", c)
    assert np.array_equal(cv2.imread(str(tmp_path / "d.png"), -1), d)
    assert np.array_equal(cv2.imread(str(tmp_path / "c.png"), 1)[..., ::-1], c)


@pytest.mark.gpu
def test_pipeline_from_dataset_equals_pipeline_from_memory(tmp_path):
    """main.cpp:64-101 loop over a dataset on disk == the same frames handed over in memory (bit-identical poses)."""
    from slam_kinectfusion_b200 import synth
    K = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    frames = synth.sequence(6, K)
    (tmp_path / "color").mkdir()
    (tmp_path / "depth").mkdir()
    for k, (_, d) in enumerate(frames):
        assert kfb.write_png_gray16(tmp_path / "depth" / f"{k:04d}.png", d.astype(np.uint16))
        assert kfb.write_png_rgb8(tmp_path / "color" / f"{k:04d}.png", np.full((K.height, K.width, 3), 128, np.uint8))
    (tmp_path / "intr.txt").write_text(f"{K.fx} 0 {K.cx}\n0 {K.fy} {K.cy}\n0 0 1\n")
    s = kfb.DatasetSensor(tmp_path)
    Kd = kfb.Intrinsics(width=s.width, height=s.height, fx=s.fx, fy=s.fy, cx=s.cx, cy=s.cy)
    a = kfb.KinectFusion(Kd, kfb.default_host_params(128))
    b = kfb.KinectFusion(K, kfb.default_host_params(128))
    for (_, depth), (_, d) in zip(s, frames):
        assert a.pipeline(depth) == 0 and b.pipeline(d) == 0
    assert np.array_equal(np.asarray(a.poses()), np.asarray(b.poses()))


def _example_binary():
    """examples/kinfu_dataset.cpp -- the reference's main.cpp loop, headless -- compiled against the facade."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    kfb.build_all()
    subprocess.check_call(["make", "-C", os.path.join(root, "slam-kinectfusion_b200", "kfusion"), "example"],
                          stdout=subprocess.DEVNULL)
    exe = os.path.join(root, "slam-kinectfusion_b200", "kinfu_dataset")
    assert os.access(exe, os.X_OK)
    return exe


def test_reference_style_application_compiles_against_the_facade():
    _example_binary()


@pytest.mark.gpu
def test_reference_style_application_runs(tmp_path):
    import subprocess
    from slam_kinectfusion_b200 import synth
    exe = _example_binary()
    K = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    data, out = tmp_path / "dataset", tmp_path / "out"
    (data / "color").mkdir(parents=True)
    (data / "depth").mkdir()
    out.mkdir()
    for k, (_, d) in enumerate(synth.sequence(7, K)):
        assert kfb.write_png_gray16(data / "depth" / f"{k:04d}.png", d.astype(np.uint16))
        assert kfb.write_png_rgb8(data / "color" / f"{k:04d}.png", np.full((K.height, K.width, 3), 90, np.uint8))
    (data / "intr.txt").write_text(f"{K.fx} 0 {K.cx}\n0 {K.fy} {K.cy}\n0 0 1\n")
    r = subprocess.run([exe, str(data), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("7 frames (0 tracking failures)"), r.stdout + r.stderr
    assert len(open(out / "poses.txt").read().strip().splitlines()) == 7 * 4  # one 4x4 matrix per frame
    assert open(out / "pointcloud.ply").readline().strip() == "ply"
    view = cv2.imread(str(out / "scene.png"), 1)
    assert view.shape == (K.height, K.width, 3) and view.any()


@pytest.mark.parametrize("kind", ["rgb8", "gray16", "gray1"])
def test_png_decoder_reads_adam7_interlaced_files(tmp_path, kind):
    """Interlaced files (assembled by hand: cv2 cannot write them) against cv2.imread, sizes that leave passes empty."""
    rng = np.random.default_rng(11)
    for h, w in ((37, 53), (2, 3), (1, 1), (9, 4)):
        if kind == "rgb8":
            img, bits, ctype = rng.integers(0, 255, (h, w, 3), dtype=np.uint8), 8, 2
        elif kind == "gray16":
            img, bits, ctype = rng.integers(0, 65535, (h, w), dtype=np.uint16), 16, 0
        else:
            img, bits, ctype = rng.integers(0, 2, (h, w), dtype=np.uint8), 1, 0
        data = b""
        for x0, y0, dx, dy in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            sub = img[y0::dy, x0::dx]
            if sub.size == 0:
                continue
            for r in sub:
                if bits == 16:
                    row = r.astype(">u2").tobytes()
                elif bits == 8:
                    row = r.tobytes()
                else:
                    row = np.packbits(r).tobytes()  # most significant bit first, zero padded
                data += b"\x00" + row

        def chunk(t, d):
            return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))

        d = tmp_path / f"{kind}_{h}x{w}"
        (d / "color").mkdir(parents=True)
        (d / "depth").mkdir()
        name = str(d / "color" / "a.png")
        open(name, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bits, ctype, 0, 0, 1))
                               + chunk(b"IDAT", zlib.compress(data)) + chunk(b"IEND", b""))
        ref = cv2.imread(name, 1)
        assert ref is not None and ref.shape == (h, w, 3)
        cv2.imwrite(str(d / "depth" / "a.png"), np.zeros((h, w), np.uint16))
        bgr, _ = kfb.DatasetSensor(d).get_frame()
        assert np.array_equal(bgr, ref), (kind, h, w)


@pytest.mark.parametrize("h,w", [(1, 1), (1, 7), (7, 1), (3, 5), (48, 64)])
def test_own_writer_and_decoder_agree_on_odd_sizes(tmp_path, h, w):
    rng = np.random.default_rng(h * 100 + w)
    (tmp_path / "color").mkdir()
    (tmp_path / "depth").mkdir()
    d = rng.integers(0, 65535, (h, w), dtype=np.uint16)
    c = rng.integers(0, 255, (h, w, 3), dtype=np.uint8)
    assert kfb.write_png_gray16(tmp_path / "depth" / "0.png", d) and kfb.write_png_rgb8(tmp_path / "color" / "0.png", c)
    s = kfb.DatasetSensor(tmp_path)
    assert (s.height, s.width) == (h, w)
    bgr, depth = s.get_frame()
    assert np.array_equal(bgr[..., ::-1], c) and np.array_equal(depth, d.astype(np.float32))
