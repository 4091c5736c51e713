"""e2e probe: the host-buffer queue arm of bench.py alone, lean wire format, with IKB_QUEUE_TRACE timeline summary.
usage: e2e_probe.py [depth] [merge] [steps]"""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import _capi as capi, workloads as W

depth, merge, steps = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 8), (2, 4), (3, 48)))
dev = torch.device("cuda:0")
pb = W.cassie_feet_pelvis_problem(); pb.finalize(0)
m = pb.model(); nq, tsz, csz = m.nq, pb.target_size, pb.compact_target_size
B = 65536
def pinned(shape, dt):
    n = int(np.prod(shape)) * np.dtype(dt).itemsize
    return np.frombuffer((C.c_char * n).from_address(capi.lib.ikb_host_alloc(n)), dtype=dt).reshape(shape)
names = W.task_frames(pb)
qstar = W.sample_configurations(m, B, 12345)
poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
tg = W.targets_from_frame_poses(pb, poses)
ctg = pb.compact_targets(tg)
q0 = W.standing_configuration(m, W.CASSIE_STANDING)
h_q0 = pinned((nq,), np.float64); h_q0[:] = q0
h_c = [pinned((csz, B), np.float64) for _ in range(depth)]
h_o = [{"q": pinned((nq, B), np.float64), "success": pinned((B,), np.uint8)} for _ in range(depth)]
for a in h_c: a[:] = ctg.T
prm = ik.dls_parameters()
queue = ik.SolveQueue(pb, depth, merge, 0)
lag = max(1, depth - 1)
def run(n, consume=True):
    got, tk = 0, []
    t_sub = t_wait = 0.0
    for k in range(n):
        t0 = time.perf_counter()
        t, _ = queue.submit_host(h_q0, h_c[k % depth], prm, "f64", "soa", h_o[k % depth], compact=True, outputs=("q", "success"))
        t_sub += time.perf_counter() - t0
        tk.append(t)
        if k >= lag:
            t0 = time.perf_counter()
            queue.wait(tk[k - lag])
            t_wait += time.perf_counter() - t0
            if consume: got += int(h_o[(k - lag) % depth]["success"].sum())
    for k in range(max(0, n - lag), n):
        queue.wait(tk[k])
        if consume: got += int(h_o[k % depth]["success"].sum())
    return got, t_sub, t_wait
run(depth + merge)
torch.cuda.synchronize()
t0 = time.perf_counter()
got, ts, tw = run(steps)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("depth %d merge %d: %.3f ms/step  %.1f M solves/s  (host time in submit %.3f ms/step, blocked in wait %.3f ms/step)" % (
    depth, merge, dt / steps * 1e3, got / dt / 1e6, ts / steps * 1e3, tw / steps * 1e3))
