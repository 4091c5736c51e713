#!/usr/bin/env python3
"""Small end-to-end run of every kernel configuration, meant to be run under compute-sanitizer (one tool per call):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Covers: latency configuration (B=200), BULK+TAIL pair (B=9600 > one wave), generic kernel, humanoid, arm, FK."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import ik_b200 as ik  # noqa: E402
from ik_b200 import workloads as W  # noqa: E402


def run(pb, B, start, dtype=torch.float64, iters=100):
    m = pb.model()
    pb.finalize(0)
    names = W.task_frames(pb)
    qstar = W.sample_configurations(m, B)
    dev = torch.device("cuda:0")
    poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
    poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
    tg = W.targets_from_frame_poses(pb, poses)
    q0 = np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1)) if start == "standing" else W.near_start(m, qstar)
    out = ik.dls_batch(pb, torch.tensor(q0.T.copy(), dtype=dtype, device=dev), torch.tensor(tg.T.copy(), dtype=dtype, device=dev),
                       ik.dls_parameters(max_iterations=iters))
    torch.cuda.synchronize()
    print("%-22s B=%-6d %s kernel=%s converged=%d" % (type(pb).__name__, B, str(dtype)[6:], pb.kernel_name(), int(out["success"].sum())))


run(W.cassie_feet_pelvis_problem(), 200, "standing")
run(W.cassie_feet_pelvis_problem(), 9600, "standing", iters=30)
run(W.cassie_feet_pelvis_problem(), 9600, "standing", torch.float32, iters=30)
os.environ["IKB_FORCE_GENERIC"] = "1"
run(W.cassie_feet_pelvis_problem(), 200, "standing", iters=20)
os.environ["IKB_FORCE_GENERIC"] = "0"
run(W.humanoid_problem(), 96, "near", iters=20)
run(W.manipulator_problem(), 5000, "near", iters=20)
print("sanitize_smoke: done")
