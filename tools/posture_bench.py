import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W
m = W.cassie_model()
pb = W.cassie_demo_posture_problem(m)
posture = pb.get_posture_task("posture")
pb.finalize(0)
B = 65536
dev = torch.device("cuda:0")
names = W.task_frames(pb)
qstar = W.sample_configurations(m, B, 12345)
poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
tg = W.targets_from_frame_poses(pb, poses)
off = pb.target_offset(posture)
tg[:, off:off + 16] = qstar[:, 7:]
tg = torch.tensor(tg.T.copy(), device=dev)
q0 = torch.tensor(np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1)).T.copy(), device=dev)
out = ik.dls_batch(pb, q0, tg); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): out = ik.dls_batch(pb, q0, tg, None, out)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
print("demo+posture kernel=%s B=%d  %.3f ms  %.2f M solves/s converged %.4f mean iters %.2f" % (pb.kernel_name(), B, dt * 1e3, out["success"].sum().item() / dt / 1e6, out["success"].float().mean().item(), out["iters"].float().mean().item()))
