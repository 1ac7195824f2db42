"""z-slab sharding: partition logic and the N > 1 collective protocol of
slam_kinectfusion_b200.sharded over gloo / world_size 2 on CPU, with the oracle as the per-slab compute."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_partition_covers_volume_once(kfb):
    from slam_kinectfusion_b200 import sharded
    for Z in (64, 203, 512, 812, 1024, 2048):
        for world in (1, 2, 3, 4, 8):
            owned = np.zeros(Z, int)
            for r in range(world):
                zb, ze = sharded.slab_range(Z, world, r)
                s0, s1 = sharded.stored_range(Z, world, r)
                assert 0 <= s0 <= zb < ze <= s1 <= Z
                assert zb - s0 == min(sharded.HALO, zb) and s1 - ze == min(sharded.HALO, Z - ze)
                owned[zb:ze] += 1
            assert (owned == 1).all()


def test_balanced_bounds_properties(kfb):
    from slam_kinectfusion_b200 import sharded
    rng = np.random.default_rng(3)
    for Z in (64, 512, 2048):
        for world in (2, 4, 8):
            hist = np.zeros(Z)
            a, b = sorted(rng.integers(1, Z, 2))
            hist[a:b + 1] = np.linspace(10, 1000, b + 1 - a) ** 1.5       # work concentrated in a band of planes
            bd = sharded.balanced_bounds(hist, world, min_planes=4)
            assert bd[0] == 0 and bd[-1] == Z and len(bd) == world + 1
            assert all(bd[i + 1] - bd[i] >= 4 for i in range(world))
            work = hist + 0.02 * hist.max()
            per = [work[bd[i]:bd[i + 1]].sum() for i in range(world)]
            if Z >= 512:
                assert max(per) < 1.5 * (sum(per) / world) + work.max() * 8   # balanced up to plane granularity / minimum height
            s0, s1 = sharded.stored_range(Z, world, 1, bd)
            assert s0 == max(bd[1] - sharded.HALO, 0) and s1 == min(bd[2] + sharded.HALO, Z)


def test_slab_raycast_oracle_equals_full_single_process(kfo, kfb):
    """The slab semantics themselves (no collectives): min-key composite of 3 slabs == full raycast, bit for bit."""
    from slam_kinectfusion_b200 import sharded
    dims, w, h = 64, 160, 120
    s = w / 640.0
    K = kfo.Intr(width=w, height=h, fx=525.0 * s, fy=525.0 * s, cx=(319.5 + 0.5) * s - 0.5, cy=(239.5 + 0.5) * s - 0.5)
    vd = kfo.volume_desc(dims)
    volpose = np.array(kfo.default_params(dims).volu_pose, np.float32)
    vol = kfo.new_volume(vd)
    for k in range(3):
        pose = kfo.trajectory_pose(5 * k)
        depth = kfo.frontend(kfo.render_depth_mm(pose, K), K, levels=1)[0][0]
        kfo.integrate(vol, vd, kfo.pose_mul(kfo.pose_inv(pose), volpose), depth, K)
    c2v = kfo.pose_mul(kfo.pose_inv(volpose), kfo.trajectory_pose(7))
    fv, fn, _ = kfo.raycast(vol, vd, c2v, K)
    assert (fv[..., 2] != 0).mean() > 0.5
    world = 3
    parts = []
    for r in range(world):
        zb, ze = sharded.slab_range(dims, world, r)
        s0, s1 = sharded.stored_range(dims, world, r)
        parts.append(kfo.raycast_slab(vol[s0:s1], vd, c2v, K, s0, s1, zb, ze))
    keys = np.stack([p[2] for p in parts])
    finite = np.isfinite(keys)
    assert (finite.sum(0) <= np.inf).all()
    win = keys.argmin(0)
    cv = np.zeros_like(fv)
    cn = np.zeros_like(fn)
    for r in range(world):
        m = (win == r) & np.isfinite(keys[r])
        cv[m] = parts[r][0][m]
        cn[m] = parts[r][1][m]
    assert np.array_equal(cv.view(np.int32), fv.view(np.int32))
    assert np.array_equal(cn.view(np.int32), fn.view(np.int32))
    # a terminal event belongs to exactly one slab: finite keys never tie
    srt = np.sort(keys, axis=0)
    assert not (np.isfinite(srt[0]) & (srt[0] == srt[1])).any()


def test_composite_protocol_gloo_world2(tmp_path, kfo, kfb):
    out = tmp_path / "result.txt"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29500 + os.getpid() % 2000), WORLD_SIZE="2",
               OMP_NUM_THREADS="2")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_sharded_worker.py"), str(out)],
                              env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
        logs.append(o)
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    ok_v, ok_n, hits = map(int, out.read_text().split())
    assert ok_v == 1 and ok_n == 1 and hits > 1000


@pytest.mark.gpu
def test_two_gpu_sharded_run_equals_single_gpu(kfb):
    """End to end on real GPUs (skipped on a single-GPU box): two z-slab ranks under torchrun (peer-memory
    composite, balanced slabs, pose mailbox) must track EXACTLY like one GPU holding the whole volume -- the
    slab volumes and the composite are bit-identical, hence so are the ICP sums and the poses."""
    import json
    if kfb.load_library().kfb_device_count() < 2:
        pytest.skip("needs two GPUs")
    from slam_kinectfusion_b200 import synth
    dims, steps, warm = 256, 6, 3
    env = dict(os.environ)
    env.pop("RANK", None)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--dims", str(dims),
                          "--steps", str(steps), "--warmup", str(warm)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["value"] > 0
    K = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    kf = kfb.KinectFusion(K, kfb.default_host_params(dims))
    for _, d in synth.sequence(1 + warm + steps, K):
        assert kf.pipeline(d) == 0
    assert np.array_equal(np.asarray(line["final_pose"], np.float32), kf.pose())


def _mailbox_peer(name, rank, world, n, q):
    import slam_kinectfusion_b200 as kfb
    from slam_kinectfusion_b200 import host
    mb = host.NativePoseMailbox(name, rank, world)
    got = []
    for i in range(n):
        msg = np.zeros(13, np.float32)
        assert mb.exchange(msg) == 0
        got.append(msg.copy())
    mb.close()
    q.put((rank, np.array(got)))


def test_native_pose_mailbox_three_ranks(kfb):
    """kf::PoseMailbox (the facade's own broadcast_pose): rank 0 publishes 200 messages back to back, two reader
    processes receive every one of them, in order, none overwritten before both have taken it."""
    import multiprocessing as mp
    from slam_kinectfusion_b200 import host
    name, world, n = f"/kfb_test_mailbox_{os.getpid()}", 3, 200
    mb0 = host.NativePoseMailbox(name, 0, world)          # rank 0 creates the segment first
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mailbox_peer, args=(name, r, world, n, q)) for r in (1, 2)]
    for p in procs:
        p.start()
    sent = []
    for i in range(n):
        msg = np.arange(13, dtype=np.float32) + 100.0 * i
        sent.append(msg.copy())
        assert mb0.exchange(msg) == 0
    res = dict(q.get(timeout=60) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    mb0.close()
    for r in (1, 2):
        assert np.array_equal(res[r], np.array(sent))


def test_rebalance_from_measured_times(kfb):
    """sharded.rebalance: bounds stay a partition, the slow rank's slab shrinks, equal times leave the cut alone."""
    from slam_kinectfusion_b200 import sharded
    b = [0, 311, 437, 555, 639, 714, 792, 872, 1024]
    t = [0.18, 0.29, 0.33, 0.33, 0.34, 0.35, 0.36, 0.38]
    nb = sharded.rebalance(b, t)
    assert nb[0] == 0 and nb[-1] == 1024 and len(nb) == len(b)
    assert all(nb[i + 1] - nb[i] >= 8 for i in range(8))
    assert nb[1] > b[1] and (nb[8] - nb[7]) < (b[8] - b[7])          # the idle first rank takes planes, the busy last one sheds them
    # predicted times under the piecewise-constant model are closer together than the measured ones
    dens = np.concatenate([np.full(b[i + 1] - b[i], t[i] / (b[i + 1] - b[i])) for i in range(8)])
    pred = [dens[nb[i]:nb[i + 1]].sum() for i in range(8)]
    assert max(pred) - min(pred) < 0.25 * (max(t) - min(t))
    same = sharded.rebalance([0, 100, 200, 300, 400], [1.0, 1.0, 1.0, 1.0])
    assert same == [0, 100, 200, 300, 400]
