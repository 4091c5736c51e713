"""One plain solve of 4B problems vs one merged group of 4 batches of B (for an ncu launch list)."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W
pb = W.cassie_feet_pelvis_problem(); pb.finalize(0); m = pb.model()
B = 65536
dev = torch.device("cuda:0")
names = W.task_frames(pb)
sets = []
for s in range(4):
    qstar = W.sample_configurations(m, B, 12345 + s)
    poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
    poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
    tg = torch.tensor(W.targets_from_frame_poses(pb, poses).T.copy(), device=dev)
    q0 = torch.tensor(np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1)).T.copy(), device=dev)
    sets.append((q0, tg))
big_q0 = torch.cat([s[0] for s in sets], dim=1).contiguous()
big_tg = torch.cat([s[1] for s in sets], dim=1).contiguous()
queue = ik.SolveQueue(pb, 8, 4)
for rep in range(3):
    ik.dls_batch(pb, big_q0, big_tg)
    torch.cuda.synchronize()
    for q0, tg in sets:
        queue.submit(q0, tg)
    queue.drain()
import time
for name, fn in (("plain 4B", lambda: ik.dls_batch(pb, big_q0, big_tg)), ("merged 4xB", lambda: [queue.submit(q0, tg) for q0, tg in sets] and queue.drain())):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); print(name, "%.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
