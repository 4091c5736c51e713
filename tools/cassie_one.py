"""One Cassie feet+pelvis workload for profiling: python tools/cassie_one.py B reps [f64|f32]   (IKB_CASSIE_SOLVE selects the solve)"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W
B = int(sys.argv[1]); reps = int(sys.argv[2]); dt = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else torch.float64
dev = torch.device("cuda:0")
pb = W.cassie_feet_pelvis_problem(); pb.finalize(0)
m = pb.model(); names = W.task_frames(pb)
qstar = W.sample_configurations(m, B, 12345)
poses_t = torch.cat([ik.fk_batch(pb, torch.tensor(qstar[i:i + 65536].T.copy(), device=dev), names) for i in range(0, B, 65536)], dim=1)
poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
tg = W.targets_from_frame_poses(pb, poses, qstar)
q0 = np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1))
dq0, dtg = torch.tensor(q0.T.copy(), dtype=dt, device=dev), torch.tensor(tg.T.copy(), dtype=dt, device=dev)
for _ in range(reps):
    o = ik.dls_batch(pb, dq0, dtg)
torch.cuda.synchronize()
print("B=%d conv %.4f iters %.2f" % (B, o["success"].float().mean().item(), o["iters"].float().mean().item()))
