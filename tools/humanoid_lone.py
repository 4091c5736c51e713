"""Lone humanoid batches (BASELINE config 4, 262 144 problems, FP64), L2 flushed between launches: python tools/humanoid_lone.py [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = 262144
dev = torch.device("cuda:0")
pb = W.humanoid_problem(); pb.finalize(0)
m = pb.model(); names = W.task_frames(pb)
sets = []
for s in range(2):
    qstar = W.sample_configurations(m, B, 12345 + s)
    poses_t = torch.cat([ik.fk_batch(pb, torch.tensor(qstar[i:i + 65536].T.copy(), device=dev), names) for i in range(0, B, 65536)], dim=1)
    poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
    tg = W.targets_from_frame_poses(pb, poses, qstar)
    sets.append((torch.tensor(W.near_start(m, qstar).T.copy(), device=dev), torch.tensor(tg.T.copy(), device=dev)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for r in range(reps + 2):
    q0, tg = sets[r % 2]
    flush.fill_(r & 0xFF)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    o = ik.dls_batch(pb, q0, tg)
    e1.record()
    torch.cuda.synchronize()
    if r >= 2:
        ts.append(e0.elapsed_time(e1))
ts = np.array(ts)
print("humanoid %s B=%d lone: median %.3f ms  min %.3f  max %.3f  (%d reps)  conv %.4f iters %.2f" % (pb.kernel_name(), B, np.median(ts), ts.min(), ts.max(), len(ts), o["success"].float().mean().item(), o["iters"].float().mean().item()))
