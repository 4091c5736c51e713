"""CPU ORACLE side (test infrastructure, NOT product code): express an ik_b200 problem / workload for the oracle.

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only."""
import os

import numpy as np

import ik_b200 as ik
from ik_b200 import workloads as W
from . import oracle as O

DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ik_b200", "data")


def urdf_text(name):
    with open(os.path.join(DATA, name + ".urdf")) as f:
        return f.read()


def oracle_model(name, free_flyer=True):
    return O.Model.from_urdf(urdf_text(name), free_flyer)


def oracle_problem_like(problem, omodel):
    """Build the oracle's problem from an ik_b200.InverseKinematicsProblem (same tasks, same order)."""
    opb = O.Problem(omodel, problem.max_priority_level())
    for _, t, prio in problem._tasks:
        if isinstance(t, ik.FrameTask):
            opb.add_frame_task(t.frame, int(t.type), t.reference_frame, prio, t.weighting())
        elif isinstance(t, ik.AlignAxisTask):
            opb.add_align_axis_task(t.frame, int(t.axis), t.reference_frame, prio, t.weighting())
        elif isinstance(t, ik.CentreOfMassTask):
            opb.add_com_task(t.reference_frame, prio, t.weighting())
        else:
            opb.add_posture_task(t.nj, prio, t.weighting(), t.mask)
    for c in problem.get_all_constraints():
        opb.add_frame_constraint(c.frame, int(c.type), c.reference_frame)
    return opb


def oracle_frame_poses(omodel, qs, names):
    out = {n: np.zeros((qs.shape[0], 12)) for n in names}
    ids = {n: omodel.frame_id(n) for n in names}
    for b in range(qs.shape[0]):
        for n in names:
            out[n][b] = omodel.frame_placement(qs[b], ids[n])
    return out


def make_workload(problem, omodel, B, seed=12345, standing=None, b0=0, start="standing"):
    """q0 [B,nq], targets [B,tsz] (AoS, float64) with targets = oracle FK of seeded reachable configurations.
    start="standing": q0 = standing / neutral pose for every problem; start="near": W.near_start (warm start)."""
    m = problem.model()
    qstar = W.sample_configurations(m, B, seed, b0)
    poses = oracle_frame_poses(omodel, qstar, W.task_frames(problem))
    targets = W.targets_from_frame_poses(problem, poses, qstar)
    if start == "near":
        q0 = W.near_start(m, qstar, seed, b0)
    else:
        q0 = np.tile(W.standing_configuration(m, standing), (B, 1))
    return q0, targets, qstar
