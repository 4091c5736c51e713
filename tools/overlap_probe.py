"""Steady-state throughput of back-to-back batches when consecutive solves alternate between 1 / 2 / 3 user streams
(ikb_dls_solve_batch): the experiment behind DESIGN.md 4.1 "overlapping the TAIL launch with the next batch's BULK launch"
(two streams: 0.77 ms per batch instead of 1.07 -- two TAIL launches run side by side -- which is what led to merging the
batches of a stream into one kernel pair, the pipelined queue)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W
pb = W.cassie_feet_pelvis_problem(); pb.finalize(0); m = pb.model()
B = 65536
dev = torch.device("cuda:0")
names = W.task_frames(pb)
sets = []
for s in range(5):
    qstar = W.sample_configurations(m, B, 12345 + s)
    poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
    poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
    tg = torch.tensor(W.targets_from_frame_poses(pb, poses).T.copy(), device=dev)
    q0 = torch.tensor(np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1)).T.copy(), device=dev)
    sets.append((q0, tg, None))
outs = [ik.dls_batch(pb, q0, tg) for q0, tg, _ in sets]
torch.cuda.synchronize()
for nstreams in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 40
        for k in range(K):
            q0, tg, _ = sets[k % 5]
            with torch.cuda.stream(streams[k % nstreams]):
                ik.dls_batch(pb, q0, tg, None, outs[k % 5], stream=streams[k % nstreams].cuda_stream)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K
    print("streams=%d  %.4f ms/batch  %.1f M problems/s" % (nstreams, dt * 1e3, B / dt / 1e6))

