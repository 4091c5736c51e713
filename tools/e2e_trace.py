import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W
import bench
pb = W.cassie_feet_pelvis_problem(); pb.finalize(0); m = pb.model()
B = 65536
qstar = W.sample_configurations(m, B)
names = W.task_frames(pb)
poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device="cuda:0"), names)
poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
tg = W.targets_from_frame_poses(pb, poses)
q0 = np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1))
h_q0 = bench.pinned_array((23, B), np.float64); h_q0[:] = q0.T
h_tg = bench.pinned_array((36, B), np.float64); h_tg[:] = tg.T
h_out = {"q": bench.pinned_array((23, B), np.float64), "success": bench.pinned_array((B,), np.uint8), "iters": bench.pinned_array((B,), np.int32), "resid": bench.pinned_array((B,), np.float64)}
for i in range(4):
    ik.dls_batch_host(pb, h_q0, h_tg, None, "f64", "soa", h_out)
os.environ["IKB_HOST_TRACE"] = "1"
for i in range(2):
    t0 = time.perf_counter(); ik.dls_batch_host(pb, h_q0, h_tg, None, "f64", "soa", h_out); print("wall %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
os.environ["IKB_HOST_PIPELINE"] = "0"; os.environ["IKB_HOST_TRACE"] = "0"
for i in range(2):
    t0 = time.perf_counter(); ik.dls_batch_host(pb, h_q0, h_tg, None, "f64", "soa", h_out); print("wall %.3f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
