"""Cassie feet+pelvis (BASELINE configs 2 / 3) A/B: IKB_CASSIE_SOLVE=dense (12 x 12 LDL^T on the solver role, r1) vs arrow
(shared / private column split on all three roles, r2).  Lone batches of 65 536 and one merged batch of 8 x 65 536 (what
the pipelined queue launches), FP64 and FP32, CUDA events; results compared between the two and, for FP64 65 536, with
the oracle.  usage: python tools/cassie_ab.py [B] [modes]"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["dense", "arrow"]
dev = torch.device("cuda:0")
pb = W.cassie_feet_pelvis_problem(); pb.finalize(0)
m = pb.model(); names = W.task_frames(pb)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def workload(nB, seed):
    qstar = W.sample_configurations(m, nB, seed)
    poses_t = torch.cat([ik.fk_batch(pb, torch.tensor(qstar[i:i + 65536].T.copy(), device=dev), names) for i in range(0, nB, 65536)], dim=1)
    poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
    tg = W.targets_from_frame_poses(pb, poses, qstar)
    q0 = np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (nB, 1))
    return q0, tg


oracle = None
if os.path.exists("oracle/libik_oracle.so") and B <= 65536:   # checker only (test infrastructure)
    try:
        from oracle.bridge import oracle_model, oracle_problem_like
        from oracle import oracle as O
        oracle = (O, oracle_problem_like(pb, oracle_model("cassie")))
    except Exception as e:  # noqa
        print("oracle unavailable:", e)

for nB, label in ((B, "lone"), (8 * B, "merged x8")):
    q0, tg = workload(nB, 12345)
    ref = {}
    for dt in (torch.float64, torch.float32):
        dq0, dtg = torch.tensor(q0.T.copy(), dtype=dt, device=dev), torch.tensor(tg.T.copy(), dtype=dt, device=dev)
        for mode in modes:
            os.environ["IKB_CASSIE_SOLVE"] = mode
            for _ in range(3):
                o = ik.dls_batch(pb, dq0, dtg)
            torch.cuda.synchronize()
            tot, n = 0.0, 10
            for _ in range(n):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); o = ik.dls_batch(pb, dq0, dtg); e1.record(); torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            ms = tot / n
            note = ""
            if (dt, modes[0]) in ref and mode != modes[0]:
                r = ref[(dt, modes[0])]
                agree = (r["success"] == o["success"]) & (r["iters"] == o["iters"])
                err = (r["q"] - o["q"]).abs().max(dim=0).values[agree & o["success"].bool()]
                note = "vs %s: flags differ %d, agree %.6f, |dq| p99.9 %.1e max %.1e" % (
                    modes[0], int((r["success"] != o["success"]).sum().item()), agree.double().mean().item(),
                    torch.quantile(err.double()[:2000000], 0.999).item(), err.max().item())
            ref[(dt, mode)] = {k: v.clone() for k, v in o.items()}
            print("cassie %s %-6s %-9s B=%d: %8.3f ms  %7.2f M solves/s  conv %.4f iters %.2f  %s" % (
                str(dt)[6:], mode, label, nB, ms, o["success"].sum().item() / ms / 1e3, o["success"].float().mean().item(),
                o["iters"].float().mean().item(), note), flush=True)
            if oracle and dt == torch.float64 and label == "lone":
                O, opb = oracle
                q_ref, ok_ref, it_ref, _ = O.dls_batch(opb, q0, tg, nthreads=os.cpu_count())
                ok = o["success"].cpu().numpy().astype(bool); it = o["iters"].cpu().numpy(); q = o["q"].cpu().numpy().T
                same = (ok == ok_ref.astype(bool)) & (it == it_ref)
                d = np.abs(q - q_ref).max(axis=1)
                print("    vs oracle: flag mismatches %d, same flags+iterations %.6f, max|dq| (converged, agreeing) %.3e, all %.3e" % (
                    int((ok != ok_ref.astype(bool)).sum()), same.mean(), d[same & ok].max(), d.max()), flush=True)
