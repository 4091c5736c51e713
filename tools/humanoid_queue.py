"""Humanoid (BASELINE config 4) through the pipelined queue: python tools/humanoid_queue.py <merge> <nst>   (depth = 2 x merge, one set per batch in flight)"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W
merge = int(sys.argv[1]); nst = int(sys.argv[2])
depth = 2 * merge
B = 262144
dev = torch.device("cuda:0")
pb = W.humanoid_problem(); pb.finalize(0)
m = pb.model(); names = W.task_frames(pb)
sets = []
for s in range(depth):
    qstar = W.sample_configurations(m, B, 12345 + s)
    poses_t = torch.cat([ik.fk_batch(pb, torch.tensor(qstar[i:i + 65536].T.copy(), device=dev), names) for i in range(0, B, 65536)], dim=1)
    poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
    tg = W.targets_from_frame_poses(pb, poses, qstar)
    o = {"q": torch.empty((m.nq, B), dtype=torch.float64, device=dev), "success": torch.empty(B, dtype=torch.uint8, device=dev),
         "iters": torch.empty(B, dtype=torch.int32, device=dev), "resid": torch.empty(B, dtype=torch.float64, device=dev)}
    sets.append((torch.tensor(W.near_start(m, qstar).T.copy(), device=dev), torch.tensor(tg.T.copy(), device=dev), o))
qx = ik.SolveQueue(pb, depth, merge, 0)
for w in range(depth):
    qx.submit(sets[w][0], sets[w][1], None, sets[w][2])
qx.drain()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(nst):
    sx = sets[k % depth]
    last, _ = qx.submit(sx[0], sx[1], None, sx[2])
qx.flush()
for t in range(max(0, last - depth + 1), last + 1):
    qx.wait_on_stream(t)
e1.record()
torch.cuda.synchronize()
qx.drain()
ms = e0.elapsed_time(e1) / nst
cv = sum(int(s[2]["success"].sum().item()) for s in sets) / depth
print("humanoid queued merge=%d depth=%d nst=%d: %.3f ms per batch  %.2f M solves/s" % (merge, depth, nst, ms, cv / ms / 1e3))
