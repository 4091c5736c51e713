"""One LONE Cassie batch (per-batch call, L2 flushed between launches) against the knobs of the two-launch schedule:
python tools/lone_sweep.py [B] [reps]
  IKB_BULK_CAP   step cap after which the BULK launch suspends a problem once the ticket queue is dry
  IKB_TAIL_PAIR  0: one 32-problem group per CTA in the TAIL launch, 1: two groups per CTA in lock step when groups > SMs
  IKB_CASSIE_TAIL 0: thread-per-problem TAIL, t: team-per-problem TAIL
The knobs are read per call, so one process sweeps them all on the same device buffers."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda:0")
pb = W.cassie_feet_pelvis_problem(); pb.finalize(0)
m = pb.model(); names = W.task_frames(pb)
qstar = W.sample_configurations(m, B, 12345)
poses_t = torch.cat([ik.fk_batch(pb, torch.tensor(qstar[i:i + 65536].T.copy(), device=dev), names) for i in range(0, B, 65536)], dim=1)
poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
tg = W.targets_from_frame_poses(pb, poses, qstar)
q0 = np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1))
dq0, dtg = torch.tensor(q0.T.copy(), device=dev), torch.tensor(tg.T.copy(), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
it = None


def run(env):
    global it
    for k in ("IKB_BULK_CAP", "IKB_TAIL_PAIR", "IKB_CASSIE_TAIL"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = []
    for r in range(reps + 3):
        flush.fill_(r & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = ik.dls_batch(pb, dq0, dtg)
        e1.record()
        torch.cuda.synchronize()
        if r >= 3:
            ts.append(e0.elapsed_time(e1))
    if it is None:
        it = o["iters"].cpu().numpy()
        ok = o["success"].cpu().numpy().astype(bool)
        print("B=%d converged %.4f mean steps %.2f; problems with >= k steps:" % (B, ok.mean(), it.mean()),
              {k: int((it >= k).sum()) for k in (8, 12, 16, 20, 24, 32, 40, 48, 64, 100)})
    else:
        assert (o["iters"].cpu().numpy() == it).all(), "step counts changed with the schedule"
    ts = np.array(ts)
    print("%-58s median %.4f ms  min %.4f  max %.4f" % (" ".join("%s=%s" % kv for kv in sorted(env.items())) or "(defaults)", np.median(ts), ts.min(), ts.max()), flush=True)


run({})
for tail in ("0", "t"):
    for pair in ("1", "0"):
        if tail == "t" and pair == "0":
            continue
        for cap in (8, 12, 16, 20, 24, 28, 32, 40, 48, 64):
            run({"IKB_BULK_CAP": str(cap), "IKB_TAIL_PAIR": pair, "IKB_CASSIE_TAIL": tail})
