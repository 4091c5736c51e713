"""One table-driven workload on the team-per-problem kernel, for ncu / latency probes.
usage: coop_one.py <cassie|humanoid|manipulator|pik|constraint> <B> [reps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
os.environ.setdefault("IKB_FORCE_GENERIC", "1")
import torch
import ik_b200 as ik
from ik_b200 import workloads as W

which, B = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
solve = lambda p, q, t: ik.dls_batch(p, q, t)
start = "standing"
if which == "cassie":
    make = W.cassie_feet_pelvis_problem
elif which == "humanoid":
    make, start = W.humanoid_problem, "near"
elif which == "manipulator":
    make, start = W.manipulator_problem, "near"
elif which == "pik":
    make = W.cassie_demo_posture_problem
    prm = ik.pik_parameters(lambdas=[1e-2, 1e-1])
    solve = lambda p, q, t: ik.pik_batch(p, q, t, prm)
else:
    def make():
        m = W.cassie_model()
        pb = ik.InverseKinematicsProblem(m, 0)
        pb.add_frame_task("pelvis", ik.FrameTask(m, "pelvis", ik.KinematicType.Full))
        pb.add_frame_task("fl", ik.FrameTask(m, "LeftFootFront", ik.KinematicType.Position))
        pb.add_frame_constraint("fr", ik.FrameConstraint(m, "RightFootFront", ik.KinematicType.Position))
        return pb
pb = make()
pb.finalize(0)
m = pb.model()
names = W.task_frames(pb)
qstar = W.sample_configurations(m, B, 12345)
poses_t = torch.cat([ik.fk_batch(pb, torch.tensor(qstar[i:i + 65536].T.copy(), device=dev), names) for i in range(0, B, 65536)], dim=1)
poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
tg = W.targets_from_frame_poses(pb, poses, qstar)
q0 = np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1)) if start == "standing" else W.near_start(m, qstar)
q0, tg = torch.tensor(q0.T.copy(), device=dev), torch.tensor(tg.T.copy(), device=dev)
out = solve(pb, q0, tg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = solve(pb, q0, tg)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
it = out["iters"].float()
print("%s %s B=%d: %.3f ms  %.2f M solves/s  conv %.4f  mean iters %.2f  max iters %d  -> %.2f us per iteration of the longest problem"
      % (which, pb.kernel_name(), B, ms, out["success"].sum().item() / ms / 1e3, out["success"].float().mean().item(), it.mean().item(),
         int(it.max().item()), ms * 1e3 / max(1.0, it.max().item())))
