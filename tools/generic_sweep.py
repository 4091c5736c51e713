"""Residency sweep of the table-driven kernel (threads per CTA x CTAs per SM): its per-thread scratch lives in local
memory, so fewer resident warps can be faster (working set in L2 instead of DRAM)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W

B = 65536
dev = torch.device("cuda:0")
m = W.cassie_model()
os.environ["IKB_FORCE_GENERIC"] = "1"
which = sys.argv[1] if len(sys.argv) > 1 else "dls"
pb = W.cassie_feet_pelvis_problem(m) if which == "dls" else W.cassie_demo_problem(m)
pb.finalize(0)
names = W.task_frames(pb)
qstar = W.sample_configurations(m, B, 12345)
poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
tg = torch.tensor(W.targets_from_frame_poses(pb, poses, qstar).T.copy(), device=dev)
q0 = torch.tensor(np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1)).T.copy(), device=dev)
prm = ik.pik_parameters(lambdas=[1e-2, 1e-1])
solve = (lambda: ik.dls_batch(pb, q0, tg)) if which == "dls" else (lambda: ik.pik_batch(pb, q0, tg, prm))
for thr, bps in ((128, 0), (128, 2), (128, 1), (64, 4), (64, 2), (64, 1), (32, 4), (32, 2), (32, 1)):
    os.environ["IKB_GENERIC_THREADS"] = str(thr)
    os.environ["IKB_GENERIC_BLOCKS_PER_SM"] = str(bps)
    out = solve()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = solve()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%s threads %3d blocks/SM %d: %8.2f ms  %.2f M solves/s" % (which, thr, bps, dt * 1e3, out["success"].sum().item() / dt / 1e6), flush=True)
