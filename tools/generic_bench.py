"""Table-driven kernel (dls_generic.cuh) on 65 536 Cassie problems: single launch (IKB_GENERIC_CAP=0) vs the two-launch
schedule (default cap 16).  Workloads: the headline task set forced onto the generic kernel, + a FrameConstraint, + a
CentreOfMassTask, and ik::pik on the demo task set.  Results are compared between the two schedules (must be identical)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
CAPS = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "16"]
dev = torch.device("cuda:0")
m = W.cassie_model()


def targets(pb):
    names = W.task_frames(pb)
    qstar = W.sample_configurations(m, B, 12345)
    poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
    poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
    return W.targets_from_frame_poses(pb, poses, qstar), qstar


def run(name, pb, solve, tg):
    q0 = torch.tensor(np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1)).T.copy(), device=dev)
    tg = torch.tensor(tg.T.copy(), device=dev)
    res = {}
    for cap in CAPS:
        os.environ["IKB_GENERIC_CAP"] = cap
        out = solve(pb, q0, tg)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            out = solve(pb, q0, tg)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 2
        res[cap] = {k: v.clone() for k, v in out.items()}
        print("%-28s cap %2s  %8.2f ms  %6.2f M solves/s  converged %.4f  mean iters %.2f" % (
            name, cap, dt * 1e3, out["success"].sum().item() / dt / 1e6, out["success"].float().mean().item(),
            out["iters"].float().mean().item()), flush=True)
    same = all(torch.equal(res[CAPS[0]][k], res[c][k]) for k in ("q", "success", "iters") for c in CAPS[1:])
    print("%-28s two-launch results identical to single launch: %s" % (name, same), flush=True)


os.environ["IKB_FORCE_GENERIC"] = "1"
pb = W.cassie_feet_pelvis_problem(m)
pb.finalize(0)
tg, _ = targets(pb)
run("feet+pelvis (generic)", pb, lambda p, q, t: ik.dls_batch(p, q, t), tg)

pb = ik.InverseKinematicsProblem(m, 0)
pb.add_frame_task("pelvis", ik.FrameTask(m, "pelvis", ik.KinematicType.Full))
pb.add_frame_task("fl", ik.FrameTask(m, "LeftFootFront", ik.KinematicType.Position))
pb.add_frame_constraint("fr", ik.FrameConstraint(m, "RightFootFront", ik.KinematicType.Position))
pb.finalize(0)
tg, _ = targets(pb)
run("+ FrameConstraint", pb, lambda p, q, t: ik.dls_batch(p, q, t), tg)

pb = ik.InverseKinematicsProblem(m, 0)
pb.add_frame_task("fl", ik.FrameTask(m, "LeftFootFront", ik.KinematicType.Position))
com = pb.add_centre_of_mass_task(ik.CentreOfMassTask(m))
pb.add_frame_task("fr", ik.FrameTask(m, "RightFootFront", ik.KinematicType.Position))
pb.add_frame_task("pelvis", ik.FrameTask(m, "pelvis", ik.KinematicType.Orientation))
pb.finalize(0)
tg, qstar = targets(pb)
tg[:, pb.target_offset(com):pb.target_offset(com) + 3] = qstar[:, :3] + [0.0, 0.0, -0.15]  # near the pelvis: reachable enough
run("CentreOfMassTask", pb, lambda p, q, t: ik.dls_batch(p, q, t), tg)

pb = W.cassie_demo_problem(m)
pb.finalize(0)
tg, _ = targets(pb)
prm = ik.pik_parameters(lambdas=[1e-2, 1e-1])
run("ik::pik demo (2 levels)", pb, lambda p, q, t: ik.pik_batch(p, q, t, prm), tg)
