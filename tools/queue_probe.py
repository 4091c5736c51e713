"""Steady-state throughput of the pipelined queue (device buffers) for several (depth, merge), and parity with dls_batch."""
import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W
pb = W.cassie_feet_pelvis_problem(); pb.finalize(0); m = pb.model()
B = int(os.environ.get("B", 65536))
dt = torch.float64 if os.environ.get("DT", "f64") == "f64" else torch.float32
dev = torch.device("cuda:0")
names = W.task_frames(pb)
sets = []
for s in range(5):
    qstar = W.sample_configurations(m, B, 12345 + s)
    poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
    poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
    tg = torch.tensor(W.targets_from_frame_poses(pb, poses).T.copy(), device=dev, dtype=dt)
    q0 = torch.tensor(np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1)).T.copy(), device=dev, dtype=dt)
    sets.append((q0, tg))
refs = [ik.dls_batch(pb, q0, tg) for q0, tg in sets]
torch.cuda.synchronize()
K = 40
t0 = time.perf_counter()
for k in range(K):
    ik.dls_batch(pb, *sets[k % 5], None, refs[k % 5])
torch.cuda.synchronize()
d0 = (time.perf_counter() - t0) / K
print("plain dls_batch          %.4f ms/batch  %.1f M problems/s" % (d0 * 1e3, B / d0 / 1e6))
for depth, merge in ((8, 1), (8, 2), (8, 4), (16, 8)):
    queue = ik.SolveQueue(pb, depth, merge)
    outs = [None] * 5
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(K):
            _, outs[k % 5] = queue.submit(*sets[k % 5], None, outs[k % 5])
        queue.drain()
        d = (time.perf_counter() - t0) / K
    same = all(torch.equal(refs[s][k], outs[s][k]) for s in range(5) for k in ("q", "success", "iters", "resid"))
    print("queue depth=%d merge=%d   %.4f ms/batch  %.1f M problems/s  bit-identical=%s" % (depth, merge, d * 1e3, B / d / 1e6, same))
