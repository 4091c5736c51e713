"""One ik::dls solve of the humanoid config (5 Full tasks, warm start) for ncu captures of the specialised kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda:0")
pb = W.humanoid_problem()
pb.finalize(0)
m = pb.model()
names = W.task_frames(pb)
qstar = W.sample_configurations(m, B)
poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
tg = torch.tensor(W.targets_from_frame_poses(pb, poses).T.copy(), device=dev)
q0 = torch.tensor(W.near_start(m, qstar).T.copy(), device=dev)
out = ik.dls_batch(pb, q0, tg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = ik.dls_batch(pb, q0, tg, None, out)
e1.record()
torch.cuda.synchronize()
print("humanoid B=%d kernel=%s %.2f ms converged %.4f mean iters %.2f max %d" % (B, pb.kernel_name(), e0.elapsed_time(e1),
      out["success"].float().mean().item(), out["iters"].float().mean().item(), out["iters"].max().item()))
