#!/usr/bin/env python3
"""The humanoid specialisation (distributed factorisation, 7 group barriers per solve) under compute-sanitizer:
    compute-sanitizer --tool racecheck python tools/sanitize_humanoid.py      (also memcheck, synccheck)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ik_b200 as ik  # noqa: E402
from ik_b200 import workloads as W  # noqa: E402

for B, dtype, iters in ((96, torch.float64, 20), (5000, torch.float64, 6), (5000, torch.float32, 6)):
    pb = W.humanoid_problem()
    pb.finalize(0)
    m = pb.model()
    names = W.task_frames(pb)
    qstar = W.sample_configurations(m, B)
    dev = torch.device("cuda:0")
    poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
    poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
    tg = W.targets_from_frame_poses(pb, poses)
    q0 = W.near_start(m, qstar)
    out = ik.dls_batch(pb, torch.tensor(q0.T.copy(), dtype=dtype, device=dev), torch.tensor(tg.T.copy(), dtype=dtype, device=dev),
                       ik.dls_parameters(max_iterations=iters))
    torch.cuda.synchronize()
    print("humanoid B=%d %s kernel=%s converged=%d" % (B, str(dtype)[6:], pb.kernel_name(), int(out["success"].sum())))
