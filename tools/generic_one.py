"""One table-driven ik::dls solve of 65 536 Cassie feet+pelvis problems (for ncu captures of dls_generic_kernel)."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
os.environ["IKB_FORCE_GENERIC"] = "1"
import torch
import ik_b200 as ik
from ik_b200 import workloads as W

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda:0")
m = W.cassie_model()
pb = W.cassie_feet_pelvis_problem(m)
pb.finalize(0)
names = W.task_frames(pb)
qstar = W.sample_configurations(m, B, 12345)
poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
tg = torch.tensor(W.targets_from_frame_poses(pb, poses, qstar).T.copy(), device=dev)
q0 = torch.tensor(np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1)).T.copy(), device=dev)
out = ik.dls_batch(pb, q0, tg)
torch.cuda.synchronize()
print("converged", out["success"].float().mean().item(), "kernel", pb.kernel_name())
