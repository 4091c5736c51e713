import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import ik_b200 as ik
from ik_b200 import workloads as W
from oracle import oracle as O
from oracle.bridge import make_workload, oracle_model, oracle_problem_like
pb = W.cassie_feet_pelvis_problem(); om = oracle_model("cassie"); opb = oracle_problem_like(pb, om)
pb.finalize(0)
for B in (4096, 20000):
    q0, tg, _ = make_workload(pb, om, B, seed=77, standing=W.CASSIE_STANDING)
    q_ref, ok_ref, it_ref, res_ref = O.dls_batch(opb, q0, tg, nthreads=os.cpu_count())
    out = ik.dls_batch(pb, torch.tensor(q0.T.copy(), device="cuda:0"), torch.tensor(tg.T.copy(), device="cuda:0"))
    torch.cuda.synchronize()
    q = out["q"].cpu().numpy().T; ok = out["success"].cpu().numpy().astype(bool); it = out["iters"].cpu().numpy()
    d = np.abs(q - q_ref).max(axis=1)
    okr = ok_ref.astype(bool)
    print("B", B, "tail", os.environ.get("IKB_CASSIE_TAIL"), "flags equal", (ok == okr).all(), "iters equal", (it == it_ref).mean(),
          "conv max|dq| %.2e" % d[okr].max(), "nonconv max %.2e  p50 %.2e p90 %.2e p99 %.2e  n>1e-6: %d of %d" % (d[~okr].max(), *np.percentile(d[~okr], [50, 90, 99]), (d[~okr] > 1e-6).sum(), (~okr).sum()))
