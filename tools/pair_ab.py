"""Interleaved A/B of the TAIL launch's group pairing on lone Cassie batches: python tools/pair_ab.py [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B = 65536
dev = torch.device("cuda:0")
pb = W.cassie_feet_pelvis_problem(); pb.finalize(0)
m = pb.model(); names = W.task_frames(pb)
sets = []
for s in range(3):
    qstar = W.sample_configurations(m, B, 12345 + s)
    poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
    poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
    tg = W.targets_from_frame_poses(pb, poses, qstar)
    q0 = np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1))
    sets.append((torch.tensor(q0.T.copy(), device=dev), torch.tensor(tg.T.copy(), device=dev)))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {"0": [], "1": []}
for r in range(reps + 4):
    for pair in ("0", "1"):
        os.environ["IKB_TAIL_PAIR"] = pair
        q0, tg = sets[r % 3]
        flush.fill_(r & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ik.dls_batch(pb, q0, tg)
        e1.record()
        torch.cuda.synchronize()
        if r >= 4:
            res[pair].append(e0.elapsed_time(e1))
for pair in ("0", "1"):
    t = np.array(res[pair])
    print("IKB_TAIL_PAIR=%s: lone batch median %.4f ms  mean %.4f  p10 %.4f  p90 %.4f  (%d interleaved reps, 3 input sets)" % (pair, np.median(t), t.mean(), np.percentile(t, 10), np.percentile(t, 90), len(t)))
