import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
os.environ["IKB_QUEUE_TRACE"] = "1"
import torch
import ik_b200 as ik
from ik_b200 import workloads as W
pb = W.cassie_feet_pelvis_problem(); pb.finalize(0); m = pb.model()
B = 65536
dev = torch.device("cuda:0")
names = W.task_frames(pb)
qstar = W.sample_configurations(m, B, 12345)
poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
tg = torch.tensor(W.targets_from_frame_poses(pb, poses).T.copy(), device=dev)
q0 = torch.tensor(np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1)).T.copy(), device=dev)
outs = [ik.dls_batch(pb, q0, tg) for _ in range(4)]
torch.cuda.synchronize()
queue = ik.SolveQueue(pb, int(sys.argv[1]) if len(sys.argv) > 1 else 3)
for k in range(6):
    queue.submit(q0, tg, None, outs[k % 4])
queue.drain()
