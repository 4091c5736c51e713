"""Steady-state throughput of back-to-back batches when consecutive solves alternate between two streams (the TAIL
launch of batch k can then run beside the BULK launch of batch k+1 if the SHARED-residency variants are selected)."""
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
    print("IKB_SHARED=%s streams=%d  %.4f ms/batch  %.1f M problems/s" % (os.environ.get("IKB_SHARED", "0"), nstreams, dt * 1e3, B / dt / 1e6))

# ---- the pipelined queue ----
for depth in (2, 3, 4):
    queue = ik.SolveQueue(pb, depth)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 40
        for k in range(K):
            q0, tg, _ = sets[k % 5]
            queue.submit(q0, tg, None, outs[k % 5])
        queue.drain()
        dt = (time.perf_counter() - t0) / K
    print("queue depth=%d  %.4f ms/batch  %.1f M problems/s" % (depth, dt * 1e3, B / dt / 1e6))
    ref = ik.dls_batch(pb, sets[4][0], sets[4][1]); torch.cuda.synchronize()
    print("  bit-identical to dls_batch:", all(torch.equal(ref[k], outs[4][k]) for k in ("q", "success", "iters", "resid")))
