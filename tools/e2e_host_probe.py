"""Host-side timeline of the e2e loop of bench.py (compact wire format) through ikb_queue_submit_host / ikb_queue_wait:
where does the host thread block?   python tools/e2e_host_probe.py <depth> <merge> <steps>   (IKB_QUEUE_CARRY_HOST=0|1)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ik_b200 as ik
from ik_b200 import workloads as W
import ctypes as C
from ik_b200 import _capi as capi


class bench:   # (bench.py's pinned allocator)
    @staticmethod
    def pinned_array(shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = capi.lib.ikb_host_alloc(max(n, 1))
        return np.frombuffer((C.c_char * n).from_address(ptr), dtype=dtype).reshape(shape)


depth, merge, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pb = W.cassie_feet_pelvis_problem(); pb.finalize(0); m = pb.model()
B = 65536
names = W.task_frames(pb)
csz = pb.compact_target_size
h_ctg, h_outs = [], []
for s in range(depth):
    qstar = W.sample_configurations(m, B, 12345 + s % 3)
    poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device="cuda:0"), names)
    poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
    tg = W.targets_from_frame_poses(pb, poses, qstar)
    a = bench.pinned_array((csz, B), np.float64); a[:] = pb.compact_targets(tg).T
    h_ctg.append(a)
    h_outs.append({"q": bench.pinned_array((23, B), np.float64), "success": bench.pinned_array((B,), np.uint8),
                   "iters": bench.pinned_array((B,), np.int32), "resid": bench.pinned_array((B,), np.float64)})
q0_one = bench.pinned_array((23,), np.float64); q0_one[:] = W.standing_configuration(m, W.CASSIE_STANDING)
queue = ik.SolveQueue(pb, depth, merge, 0)
lag = max(1, depth - 1)

def run(n, log):
    tickets = []
    t00 = time.perf_counter()
    for k in range(n):
        t0 = time.perf_counter()
        t, _ = queue.submit_host(q0_one, h_ctg[k % depth], None, "f64", "soa", h_outs[k % depth], compact=True, outputs=("q", "success"))
        t1 = time.perf_counter()
        tickets.append(t)
        if k >= lag:
            queue.wait(tickets[k - lag])
        t2 = time.perf_counter()
        if log:
            print("k=%2d  at %7.3f ms: submit %6.3f ms, wait(%2d) %6.3f ms" % (k, (t0 - t00) * 1e3, (t1 - t0) * 1e3, k - lag, (t2 - t1) * 1e3))
    for k in range(max(0, n - lag), n):
        queue.wait(tickets[k])
    torch.cuda.synchronize()
    return (time.perf_counter() - t00) * 1e3

run(depth + merge, False)
l0 = ik.kernel_launch_count()
tot = run(steps, True)
print("solve-kernel launches in the timed loop: %d" % (ik.kernel_launch_count() - l0))
print("carry_host=%s depth=%d merge=%d steps=%d: %.3f ms total, %.4f ms per step" % (os.environ.get("IKB_QUEUE_CARRY_HOST", "default (on when depth >= 3 x merge)"), depth, merge, steps, tot, tot / steps))
