"""Humanoid (BASELINE config 4) A/B: IKB_HUMANOID_SOLVE=uniform (dense distributed factorisation, r1) | arrow (shared / private
column split, r2) | arrowt (arrow with the Jacobian strip in tensor memory: two groups per SM in FP64).  262,144 problems,
FP64 and FP32, lone batch, CUDA events; results compared with the first mode.   python tools/humanoid_ab.py [B] [modes]"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["uniform", "arrow"]
dev = torch.device("cuda:0")
pb = W.humanoid_problem(); pb.finalize(0)
m = pb.model(); names = W.task_frames(pb)
qstar = W.sample_configurations(m, B, 12345)
poses_t = torch.cat([ik.fk_batch(pb, torch.tensor(qstar[i:i + 65536].T.copy(), device=dev), names) for i in range(0, B, 65536)], dim=1)
poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
tg = W.targets_from_frame_poses(pb, poses, qstar)
q0 = W.near_start(m, qstar)
ref = {}
for dt in (torch.float64, torch.float32):
    dq0, dtg = torch.tensor(q0.T.copy(), dtype=dt, device=dev), torch.tensor(tg.T.copy(), dtype=dt, device=dev)
    for mode in modes:
        os.environ["IKB_HUMANOID_SOLVE"] = mode
        o = ik.dls_batch(pb, dq0, dtg); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): o = ik.dls_batch(pb, dq0, dtg)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        note = ""
        if (dt, modes[0]) in ref and mode != modes[0]:
            r = ref[(dt, modes[0])]
            agree = (r["success"] == o["success"]) & (r["iters"] == o["iters"])
            err = (r["q"] - o["q"]).abs().max(dim=0).values[agree & o["success"].bool()]
            note = "vs %s: agree %.6f, |dq| p99.9 %.1e max %.1e" % (modes[0], agree.double().mean().item(), torch.quantile(err.double()[:2000000], 0.999).item(), err.max().item())
        ref[(dt, mode)] = {k: v.clone() for k, v in o.items()}
        print("humanoid %s %-7s B=%d: %8.3f ms  %6.2f M solves/s  conv %.4f iters %.2f  %s" % (str(dt)[6:], mode, B, ms, o["success"].sum().item() / ms / 1e3,
              o["success"].float().mean().item(), o["iters"].float().mean().item(), note), flush=True)
