import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import ik_b200 as ik
from ik_b200 import workloads as W
from oracle import oracle as O
from oracle.bridge import make_workload, oracle_model, oracle_problem_like
om = oracle_model("cassie")
B = 24000
res = {}
for gen in ("0", "1"):
    os.environ["IKB_FORCE_GENERIC"] = gen
    pb = W.cassie_demo_problem(); pb.finalize(0)
    opb = oracle_problem_like(pb, om)
    q0, tg, _ = make_workload(pb, om, B, seed=31, standing=W.CASSIE_STANDING)
    if gen == "0":
        q_ref, ok_ref, it_ref, r_ref = O.dls_batch(opb, q0, tg, nthreads=os.cpu_count())
    out = ik.dls_batch(pb, torch.tensor(q0.T.copy(), device="cuda:0"), torch.tensor(tg.T.copy(), device="cuda:0"))
    torch.cuda.synchronize()
    q = out["q"].cpu().numpy().T
    d = np.abs(q - q_ref).max(axis=1)
    okr = ok_ref.astype(bool)
    worst = np.argsort(-np.where(okr, d, 0))[:5]
    print(pb.kernel_name(), "converged-problem max diff %.3e; worst:" % d[okr].max(), [(int(b), "%.2e" % d[b], int(it_ref[b])) for b in worst],
          "n>1e-9:", int((d[okr] > 1e-9).sum()))
    res[gen] = q
print("spec vs generic max diff on converged %.3e" % np.abs(res["0"] - res["1"])[okr].max())
