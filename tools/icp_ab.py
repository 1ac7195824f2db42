"""ICP timing A/B inside the pipelined frame sequence: whole-schedule kernel duration (events 54/55) and frame time for
the whole-schedule kernel (free) and one ordinary launch per iteration (direct = KFB_ICP_DIRECT=1).
    python tools/icp_ab.py [dims] [frames] [variant ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import slam_kinectfusion_b200 as kfb  # noqa: E402
from slam_kinectfusion_b200 import synth  # noqa: E402


def main():
    dims = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 44
    K = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    frames = synth.sequence(n, K)
    dev = torch.stack([torch.from_numpy(d) for _, d in frames]).cuda()
    poses = {}
    for pxt in sys.argv[3:] or ("free", "direct", "free", "direct"):
        os.environ.pop("KFB_ICP_DIRECT", None)
        if pxt == "direct":
            os.environ["KFB_ICP_DIRECT"] = "1"
        kf = kfb.KinectFusion(K, kfb.default_host_params(dims))
        ctx = kf.context()
        ctx.set_profiling(True)
        icp = []
        for i in range(n):
            if i == 8:
                ctx.synchronize()
                ctx.event_record(0)
            assert kf.pipeline_ptr(dev[i].data_ptr(), K.width, K.height) == 0
            if i >= 8 and i % 4 == 3:
                ctx.synchronize()
                icp.append(ctx.event_elapsed_ms(54, 55))
        ctx.event_record(1)
        ctx.synchronize()
        R = ctx.debug_icp_ring().astype(np.int64)[:19]
        per = np.diff(R[:, 0])
        poses[pxt] = kf.pose().copy()
        print(f"{pxt:>9s}: icp kernel {np.mean(icp) * 1e3:6.1f} us  frame {ctx.event_elapsed_ms(0, 1) / (n - 8) * 1e3:6.1f} us  "
              f"period ns L2 {np.median(per[:9]):.0f} L1 {np.median(per[10:14]):.0f} L0 {np.median(per[15:18]):.0f}  "
              f"fallbacks {ctx.icp_fallback_count()} mispredicted {ctx.icp_mispredict_count()}", flush=True)
        kf.close()
    ref = poses[next(iter(poses))]
    print("max pose difference between variants: %.3g" % max(float(np.abs(p - ref).max()) for p in poses.values()))


if __name__ == "__main__":
    main()
