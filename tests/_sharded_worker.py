"""world_size-2 gloo worker for tests/test_sharded.py (CPU): the collective protocol of
slam_kinectfusion_b200.sharded with the ORACLE standing in for the per-slab kernels."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, out = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), sys.argv[1]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import kfo
    from slam_kinectfusion_b200 import sharded
    dims, w, h = 64, 160, 120
    s = w / 640.0
    K = kfo.Intr(width=w, height=h, fx=525.0 * s, fy=525.0 * s, cx=(319.5 + 0.5) * s - 0.5, cy=(239.5 + 0.5) * s - 0.5)
    vd = kfo.volume_desc(dims)
    P = kfo.default_params(dims)
    volpose = np.array(P.volu_pose, np.float32)
    Z = dims
    zb, ze = sharded.slab_range(Z, world, rank)
    zs0, zs1 = sharded.stored_range(Z, world, rank)
    # every rank integrates ONLY its stored planes (slab + halo), as the GPU contexts do
    vol = kfo.new_volume(vd)
    for k in range(3):
        pose = kfo.trajectory_pose(4 * k)
        depth = kfo.frontend(kfo.render_depth_mm(pose, K), K, levels=1)[0][0]
        v2c = kfo.pose_mul(kfo.pose_inv(pose), volpose)
        kfo.integrate(vol, vd, v2c, depth, K, z_begin=max(zs0, 1), z_end=zs1)
    slab = np.ascontiguousarray(vol[zs0:zs1])
    # rank 0 decides the pose; everyone receives it
    msg = np.zeros(13, np.float32)
    if rank == 0:
        msg[0] = 1.0
        msg[1:] = kfo.trajectory_pose(5)
    msg = sharded.broadcast_pose(dist, msg, torch.device("cpu"))
    assert msg[0] == 1.0 and np.array_equal(msg[1:], kfo.trajectory_pose(5))
    c2v = kfo.pose_mul(kfo.pose_inv(volpose), msg[1:])
    v, n, key = kfo.raycast_slab(slab, vd, c2v, K, zs0, zs1, zb, ze)
    v4 = np.zeros((h * w, 4), np.float32); v4[:, :3] = v.reshape(-1, 3)
    n4 = np.zeros((h * w, 4), np.float32); n4[:, :3] = n.reshape(-1, 3)
    both = np.concatenate([v4, n4])            # vertex map then normal map, contiguous like the device buffer
    v4, n4 = both[:h * w], both[h * w:]
    tm, tk = torch.from_numpy(both.view(np.int32).reshape(-1)), torch.from_numpy(key.reshape(-1))

    def mask_fn(min_keys):          # CPU model of kfb_composite_mask
        mk = min_keys.numpy()
        lose = ~(key.reshape(-1) == mk) | np.isinf(key.reshape(-1))
        v4[lose] = 0
        n4[lose] = 0

    sharded.composite(dist, tk, tm, mask_fn)
    # the shared-memory pose mailbox (same node) must deliver what rank 0 wrote
    mb = sharded.PoseMailbox(dist, rank, f"test{os.environ['MASTER_PORT']}")
    for it in range(3):
        m2 = np.zeros(13, np.float32)
        if rank == 0:
            m2[:] = np.arange(13) + it
        mb.exchange(m2)
        assert np.array_equal(m2, np.arange(13, dtype=np.float32) + it)
    dist.barrier()
    mb.close()
    if rank == 0:
        # single-GPU truth: the full volume integrated over all planes, full raycast
        full = kfo.new_volume(vd)
        for k in range(3):
            pose = kfo.trajectory_pose(4 * k)
            depth = kfo.frontend(kfo.render_depth_mm(pose, K), K, levels=1)[0][0]
            kfo.integrate(full, vd, kfo.pose_mul(kfo.pose_inv(pose), volpose), depth, K)
        fv, fn, _ = kfo.raycast(full, vd, c2v, K)
        ok_v = np.array_equal(v4[:, :3].view(np.int32), fv.reshape(-1, 3).view(np.int32))
        ok_n = np.array_equal(n4[:, :3].view(np.int32), fn.reshape(-1, 3).view(np.int32))
        hits = int((fv[..., 2] != 0).sum())
        with open(out, "w") as f:
            f.write(f"{int(ok_v)} {int(ok_n)} {hits}\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
