#!/usr/bin/env python3
"""bench.py -- converged IK solves/sec on the BASELINE.json headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype f64|f32] [--batch B]

Workload (config.workload): Cassie feet+pelvis IK (pelvis Full + LeftFootFront/RightFootFront Position, world
frame), library-default solver parameters (max_iterations 100, damping 1e-2, step 1.0), batch 65,536 seeded random
reachable targets PER GPU (weak scaling: rank r solves problem indices [r*B, (r+1)*B)), FP64.

A "step" = one batched ik::dls over one batch of B problems.  The K timed steps go through the pipelined queue
(ikb_queue_*, include/ikb200.h -- the API for a stream of batches): `--merge` consecutive batches (default 8 of 16 in flight), each with
its own buffers, share ONE kernel pair -- the BULK launch suspends the few stragglers still unfinished when the ticket
queue runs dry, the TAIL launch continues them in the latency configuration -- so the stragglers'
serial chain (a problem that never converges runs all 100 steps, ~0.7 ms of mostly idle SMs) is paid once per group.
`value` = converged solves of all ranks / max-over-ranks device time (CUDA events) with inputs already resident in HBM;
`config.isolated_ms_per_batch` is the same K steps through the plain per-batch call (ikb_dls_solve_batch), one kernel
pair per batch.  `e2e` = the same metric with HOST buffers through the queue's host entry point
(ikb_queue_submit_host / ikb_queue_wait): every step's inputs are copied from pinned host memory, every step's results
(q, success, iters, resid) are copied back and read; the copies of one group run beside the kernels of its neighbours
(`--e2e-merge` 4 of `--e2e-depth` 8 in flight: this arm is bound by the PCIe link, smaller groups shorten its fill and drain).
`e2e.isolated_ms_per_batch` is the blocking per-batch host call (ikb_dls_solve_batch_host).
`roofline` is the compute roofline of the solve (all launches of the timed region -- they are one pass of the path per
step): algorithmic FLOPs (F_iter = 9,360 per evaluation, (iterations+1) evaluations per problem) / event-
timed device time per step, against the FP64 (FP32) FMA-pipe peak measured in this run by ikb_measure_fma_peak
(MEASURED_PEAKS.json carries no vector-pipe figure; the nominal 37.2 / 74.4 TFLOP/s is printed beside it).
`cpu_baseline` = the restated reference CPU path (oracle/, "port": Pinocchio/Eigen are unavailable so the reference
itself cannot be built) on this box's host cores, bounded sample.

--impl reference times that CPU path alone (rank 0 only under torchrun).
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

F_ITER_CASSIE = 9360.0  # SURVEY.md 8d, algorithmic FLOPs of one evaluate+solve+step for the Cassie problem
# dram__bytes_read.sum + dram__bytes_write.sum per FP64 step, from the committed ncu --set full capture
# profiles/r1_queue_full.txt: the BULK + TAIL launches of a group of 4 steps move 132.8 + 45.1 and 47.9 + 0.3 MB -> 56.5 MB
# per step (a lone step: 53.7 MB, profiles/r1_final_full.txt); algorithmic bytes are 43.8 MB.
NCU_TRAFFIC_BYTES_F64 = 56.5e6
METRIC = "converged IK solves/sec (Cassie, batch 65,536)"
UNIT = "solves/s"
NOMINAL_TFLOPS = {"f64": 37.2, "f32": 74.4}


def hbm_bytes_per_solve(nq, tsz, s):
    # (q0 + targets) in, (q + resid) + success(1) + iters(4) out -- SURVEY 8d "Layout"
    return (nq + tsz) * s + (nq + 1) * s + 1 + 4


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.armed = threading.Event()   # samples are recorded only while the timed region runs
        self.ready = threading.Event()   # NVML initialised (takes longer than a whole timed region)
        self.sm = []
        self.reasons = set()
        self.max_mhz = None
        self.power = []

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
            }
            self.ready.set()
            while not self.stop_flag.is_set():
                if not self.armed.is_set():
                    time.sleep(0.0005)
                    continue
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.001)
        except Exception as e:  # pragma: no cover
            self.reasons.add("sampler_error:%s" % type(e).__name__)
            self.ready.set()

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(sm),
                "power_w_max": (max(self.power) if self.power else None)}


def pinned_array(shape, dtype):
    from ik_b200 import _capi as capi

    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = capi.lib.ikb_host_alloc(max(n, 1))
    if not ptr:
        raise MemoryError("ikb_host_alloc failed")
    buf = (C.c_char * n).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr


def cpu_reference_arm(B_sample, steps, warmup, cores):
    """Times the restated reference CPU path (oracle) on a bounded sample of the workload, all host cores."""
    from ik_b200 import workloads as W
    from oracle import oracle as O
    from oracle.bridge import make_workload, oracle_model, oracle_problem_like

    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    q0, tg, _ = make_workload(pb, om, B_sample, standing=W.CASSIE_STANDING)
    for _ in range(warmup):
        O.dls_batch(opb, q0[:256], tg[:256], nthreads=cores)
    t0 = time.perf_counter()
    conv = 0
    for _ in range(steps):
        _, ok, it, _ = O.dls_batch(opb, q0, tg, nthreads=cores)
        conv += int(ok.sum())
    dt = time.perf_counter() - t0
    return conv / dt, dt / steps * 1e3, float(ok.mean()), float(it.mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--batch", type=int, default=65536, help="problems per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--merge", type=int, default=8, help="consecutive batches per kernel pair in the pipelined queue (1 = off)")
    ap.add_argument("--depth", type=int, default=16, help="batches in flight in the pipelined queue")
    ap.add_argument("--e2e-merge", type=int, default=4, help="the same for the host-buffer (e2e) arm: smaller groups keep the "
                    "PCIe pipeline's fill / drain short")
    ap.add_argument("--e2e-depth", type=int, default=8)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    workload = "Cassie feet+pelvis IK (pelvis Full + 2 foot Position tasks, world frame), batch %d per GPU, %s, " \
               "defaults max_it=100 damping=1e-2 step=1.0, seeded random reachable targets" % (args.batch, args.dtype)

    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = args.cpu_sample or args.batch
        warm = max(args.warmup, 1)
        val, ms, conv_frac, mean_it = cpu_reference_arm(sample, args.steps, warm, cores)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "sample": "%d problems per step" % sample},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": "%d problems x %d steps of the same seeded workload; restated reference CPU "
                                           "path (Pinocchio/Eigen unavailable here), pthreads over all cores"
                                           % (sample, args.steps)},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "converged_fraction": conv_frac, "mean_iterations": mean_it}
        print(json.dumps(line))
        return 0

    import torch

    import ik_b200 as ik
    from ik_b200 import _capi as capi
    from ik_b200 import workloads as W

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    if world > 1:
        import torch.distributed as dist

        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    npdt = np.float64 if args.dtype == "f64" else np.float32
    B = args.batch

    pb = W.cassie_feet_pelvis_problem()
    pb.finalize(local_rank)
    m = pb.model()
    nq, tsz = m.nq, pb.target_size
    names = W.task_frames(pb)
    q0_np = np.tile(W.standing_configuration(m, W.CASSIE_STANDING), (B, 1))

    # Distinct input/output sets, rotated every step, so the per-step inputs are not L2 hits left by the previous
    # step: NSETS * (inputs+outputs) > 126 MB of L2.
    per_set = B * hbm_bytes_per_solve(nq, tsz, np.dtype(npdt).itemsize)
    nsets = max(2, int(np.ceil(160e6 / per_set)) + 1, args.depth)  # ... and no set twice among the batches in flight
    sets = []
    for s in range(nsets):
        qstar = W.sample_configurations(m, B, seed=12345 + s, b0=rank * B)
        poses_t = ik.fk_batch(pb, torch.tensor(qstar.T.copy(), device=dev), names)
        poses = {n: poses_t[12 * i:12 * i + 12].T.cpu().numpy() for i, n in enumerate(names)}
        tg = W.targets_from_frame_poses(pb, poses)
        q0_d = torch.tensor(q0_np.T.copy(), dtype=tdt, device=dev)
        tg_d = torch.tensor(tg.T.copy(), dtype=tdt, device=dev)
        out = {"q": torch.empty((nq, B), dtype=tdt, device=dev), "success": torch.empty(B, dtype=torch.uint8, device=dev),
               "iters": torch.empty(B, dtype=torch.int32, device=dev), "resid": torch.empty(B, dtype=tdt, device=dev)}
        sets.append((q0_d, tg_d, out, tg))
    prm = ik.dls_parameters()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ----
    # A step = one batch of B problems through the solver.  The timed K steps are submitted to the pipelined queue
    # (ikb_queue_*, the API for a stream of batches): `--merge` consecutive batches share one BULK + TAIL kernel pair, so
    # the straggler chain (problems that never converge run all 100 steps, ~0.7 ms of mostly idle SMs) is paid once per
    # group.  The same K steps through the plain per-batch call (ikb_dls_solve_batch) are timed too: `isolated`.
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.ready.wait(10.0)
    for w in range(args.warmup):
        q0_d, tg_d, out, _ = sets[w % nsets]
        ik.dls_batch(pb, q0_d, tg_d, prm, out)
    barrier()
    iso = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    iso[0].record()
    for k in range(args.steps):
        q0_d, tg_d, out, _ = sets[k % nsets]
        ik.dls_batch(pb, q0_d, tg_d, prm, out)
    iso[1].record()
    barrier()
    isolated_ms = iso[0].elapsed_time(iso[1]) / args.steps

    depth = max(args.depth, args.merge)
    queue = ik.SolveQueue(pb, depth, args.merge, local_rank)
    for w in range(max(args.warmup, args.merge)):
        q0_d, tg_d, out, _ = sets[w % nsets]
        queue.submit(q0_d, tg_d, prm, out)
    queue.drain()
    barrier()
    sampler.armed.set()
    launches0 = ik.kernel_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()                       # the queue's compute stream waits for this point of the current stream
    last = None
    for k in range(args.steps):
        q0_d, tg_d, out, _ = sets[k % nsets]
        last, _ = queue.submit(q0_d, tg_d, prm, out)
    queue.flush()
    for t in range(max(0, last - depth + 1), last + 1):
        queue.wait_on_stream(t)          # ... and the current stream waits for every batch still in flight
    ev[1].record()
    barrier()
    queue.drain()
    launches = ik.kernel_launch_count() - launches0
    elapsed_ms = ev[0].elapsed_time(ev[1])
    sampler.armed.clear()
    kernel_ms = [elapsed_ms / args.steps]

    # optional final result gather: not part of the solve, timed and reported separately
    gather_ms = None
    if world > 1:
        from ik_b200.sharding import gather_results

        last_out = sets[(args.steps - 1) % nsets][2]
        loc = {k: last_out[k] for k in ("q", "success", "iters")}
        gather_results(loc, B * world, dist)      # warm-up (NCCL communicator set-up)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        gather_results(loc, B * world, dist)
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1)
    it_last = sets[(args.steps - 1) % nsets][2]["iters"].float()
    it_hist = {"p50": it_last.quantile(0.5).item(), "p90": it_last.quantile(0.9).item(), "p95": it_last.quantile(0.95).item(),
               "p99": it_last.quantile(0.99).item(), "at_max_iterations": (it_last >= 100).float().mean().item()}
    conv = 0
    evals = 0
    it_sum = 0
    for k in range(args.steps):
        out = sets[k % nsets][2]
        conv += int(out["success"].sum().item())
        it = out["iters"].to(torch.int64)
        it_sum += int(it.sum().item())
        # evaluations: a converged problem evaluated (iters+1) times, a failed one `iters` times
        evals += int((it + out["success"].to(torch.int64)).sum().item())

    # ---- e2e arm: HOST buffers.  Every step copies that step's inputs from pinned host memory and reads its results back
    # (q, success, iters, resid); the steps go through the queue's host entry point (ikb_queue_submit_host / ikb_queue_wait),
    # so the copies of one group of batches run beside the kernels of its neighbours. ----
    e2e_depth = max(args.e2e_depth, args.e2e_merge)
    queue_h = ik.SolveQueue(pb, e2e_depth, args.e2e_merge, local_rank)
    nbuf = e2e_depth
    h_in = [(pinned_array((nq, B), npdt), pinned_array((tsz, B), npdt)) for _ in range(nbuf)]
    h_outs = [{"q": pinned_array((nq, B), npdt), "success": pinned_array((B,), np.uint8),
               "iters": pinned_array((B,), np.int32), "resid": pinned_array((B,), npdt)} for _ in range(nbuf)]
    host_sets = [np.ascontiguousarray(sets[s][3].T, dtype=npdt) for s in range(min(nsets, 3))]
    for i, (hq, ht) in enumerate(h_in):
        hq[:] = q0_np.T
        ht[:] = host_sets[i % len(host_sets)]
    e2e_steps = args.steps
    lag = max(1, e2e_depth - 1)              # results of step k are consumed after step k + lag has been submitted: the
                                         # next group is fully in flight before the host blocks on the previous one

    def e2e_run(nsteps):
        got = 0
        tickets = []
        for k in range(nsteps):
            t, _ = queue_h.submit_host(h_in[k % nbuf][0], h_in[k % nbuf][1], prm, args.dtype, "soa", h_outs[k % nbuf])
            tickets.append(t)
            if k >= lag:
                queue_h.wait(tickets[k - lag])
                got += int(h_outs[(k - lag) % nbuf]["success"].sum())
        for k in range(max(0, nsteps - lag), nsteps):
            queue_h.wait(tickets[k])
            got += int(h_outs[k % nbuf]["success"].sum())
        return got

    # the workload's initial guess is ONE pose for every problem: the same steps with q0 passed once
    # (batch_stride = 0) -- reported beside the headline e2e, which copies a q0 per problem as a general caller would
    h_q0_one = pinned_array((nq,), npdt)
    h_q0_one[:] = q0_np[0]

    def e2e_run_bcast(nsteps):
        got = 0
        tickets = []
        for k in range(nsteps):
            t, _ = queue_h.submit_host(h_q0_one, h_in[k % nbuf][1], prm, args.dtype, "soa", h_outs[k % nbuf])
            tickets.append(t)
            if k >= lag:
                queue_h.wait(tickets[k - lag])
                got += int(h_outs[(k - lag) % nbuf]["success"].sum())
        for k in range(max(0, nsteps - lag), nsteps):
            queue_h.wait(tickets[k])
            got += int(h_outs[k % nbuf]["success"].sum())
        return got

    e2e_run(nbuf + args.e2e_merge)           # every slot's staging buffers exist before the timed region
    # the same steps through the blocking per-batch host call, for reference
    for k in range(2):
        ik.dls_batch_host(pb, h_in[0][0], h_in[0][1], prm, args.dtype, "soa", h_outs[0])
    t0 = time.perf_counter()
    for k in range(3):
        ik.dls_batch_host(pb, h_in[0][0], h_in[0][1], prm, args.dtype, "soa", h_outs[0])
    e2e_isolated_ms = (time.perf_counter() - t0) / 3 * 1e3
    barrier()
    sampler.armed.set()  # the clocks record covers both timed regions (device-resident steps and e2e steps)
    t0 = time.perf_counter()
    e2e_conv = e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_run_bcast(nbuf + args.e2e_merge)
    barrier()
    t0 = time.perf_counter()
    bcast_conv = e2e_run_bcast(e2e_steps)
    torch.cuda.synchronize()
    bcast_s = time.perf_counter() - t0
    sampler.armed.clear()
    sampler.stop_flag.set()
    sampler.join()

    # ---- reduce over ranks ----
    stats = torch.tensor([elapsed_ms, e2e_s, bcast_s], dtype=torch.float64, device=dev)
    sums = torch.tensor([conv, e2e_conv, evals, it_sum, bcast_conv], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    elapsed_ms_max, e2e_s_max, bcast_s_max = stats.tolist()
    conv_all, e2e_conv_all, evals_all, it_all, bcast_conv_all = sums.tolist()

    if rank == 0:
        value = conv_all / (elapsed_ms_max * 1e-3)
        e2e_value = e2e_conv_all / e2e_s_max
        itemsize = np.dtype(npdt).itemsize
        # roofline of the (single) solve kernel on rank 0
        k_ms = float(np.mean(kernel_ms))
        flops_per_launch = F_ITER_CASSIE * (evals / args.steps)
        achieved_tf = flops_per_launch / (k_ms * 1e-3) / 1e12
        peak = C.c_double(0)
        capi.check(capi.lib.ikb_measure_fma_peak(capi.F64 if args.dtype == "f64" else capi.F32, local_rank,
                                                 C.byref(peak)), "ikb_measure_fma_peak")
        hbm_peak = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                hbm_peak = json.load(f).get("hbm_gbs")
        except Exception:
            pass
        hbm_src = "MEASURED_PEAKS.json" if hbm_peak else "fallback (B200_PROFILING.md)"
        hbm_peak = hbm_peak or 6650.0
        hbm_achieved = B * hbm_bytes_per_solve(nq, tsz, itemsize) / (k_ms * 1e-3) / 1e9
        cpu = None
        if not args.no_cpu_baseline:
            sample = args.cpu_sample or args.batch
            probe, _, _, _ = cpu_reference_arm(min(sample, 4096), 1, 1, cores)
            passes = int(min(50, max(1, np.ceil(10.0 * probe / sample))))
            v, ms, cf, mi = cpu_reference_arm(sample, passes, 1, cores)
            v1, _, _, _ = cpu_reference_arm(min(sample, 4096), 1, 0, 1)   # one thread, small sample (both)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "single_thread_value": v1,
                   "sample": "%d problems x %d passes of the same seeded workload (%.1f s); restated reference CPU path "
                             "(Pinocchio/Eigen unavailable), pthreads over all %d host cores"
                             % (sample, passes, passes * ms / 1e3, cores)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload, "batch_per_gpu": B, "global_batch": B * world,
                       "kernel": pb.kernel_name(args.dtype),
                       "l2": "inputs/outputs rotate over %d distinct batches (%.0f MB > 126 MB L2)"
                             % (nsets, nsets * per_set / 1e6),
                       "pipeline": "ikb_queue, depth %d, %d consecutive batches per BULK+TAIL kernel pair" % (depth, args.merge),
                       "isolated_ms_per_batch": isolated_ms,
                       "isolated_value": conv * world / args.steps / (isolated_ms * 1e-3),
                       "converged_fraction": conv_all / (B * world * args.steps),
                       "mean_iterations": it_all / (B * world * args.steps), "iterations": it_hist,
                       "result_gather_ms": gather_ms},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * (nq + tsz) * itemsize),
                    "d2h_bytes_per_step": int(B * ((nq + 1) * itemsize + 5)), "steps": e2e_steps,
                    "api": "ikb_queue_submit_host + ikb_queue_wait (pinned host buffers; H2D, solve, D2H of every step)",
                    "pipeline": "ikb_queue, depth %d, %d consecutive batches per BULK+TAIL kernel pair" % (e2e_depth, args.e2e_merge),
                    "isolated_ms_per_batch": e2e_isolated_ms,
                    "shared_q0": {"value": bcast_conv_all / bcast_s_max, "h2d_bytes_per_step": int((B * tsz + nq) * itemsize),
                                  "note": "same steps, the workload's single initial guess passed once (q0 batch_stride = 0) "
                                          "instead of one copy per problem"},
                    "isolated_api": "ikb_dls_solve_batch_host (blocking, one batch at a time)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64_fma_pipe" if args.dtype == "f64" else "fp32_fma_pipe",
                         "achieved": achieved_tf, "peak": peak.value, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak.value if peak.value else None,
                         "traffic": NCU_TRAFFIC_BYTES_F64 if (args.dtype == "f64" and B == 65536) else None,
                         "traffic_source": "ncu --set full, profiles/r1_queue_full.txt: DRAM bytes of the BULK + TAIL launches of a group of 4 steps (177.9 + 48.2 MB) / 4; algorithmic 43.8 MB",
                         "kernels": "BULK %s + TAIL per group of %d steps; kernel_ms = device time of the timed region / steps"
                                    % (pb.kernel_name(args.dtype), args.merge),
                         "peak_source": "measured in this run (ikb_measure_fma_peak); nominal %.1f"
                                        % NOMINAL_TFLOPS[args.dtype],
                         "frac_of_nominal": achieved_tf / NOMINAL_TFLOPS[args.dtype],
                         "flops_per_launch": flops_per_launch, "kernel_ms": k_ms,
                         "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": hbm_achieved / hbm_peak, "peak_source": hbm_src}},
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
