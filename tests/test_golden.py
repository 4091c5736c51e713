"""Golden fixtures (tests/golden/*.npz, written by tools/make_golden.py from the CPU oracle -- the reference itself
cannot run and pins no vectors).

CPU (-m "not gpu"): the oracle still reproduces them bit for bit; the PRODUCT's URDF flattener (C++, through the C ABI)
yields exactly the golden topology.  GPU (-m gpu): the CUDA path reproduces the golden solves and frame placements.
"""
import os

import numpy as np
import pytest

import ik_b200 as ik
from ik_b200 import workloads as W
from oracle import oracle as O
from tests.common import oracle_model, oracle_problem_like

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TOPO = [("topology_cassie", "cassie", True), ("topology_ur5", "ur5", False), ("topology_humanoid", "humanoid", True),
        ("topology_manipulator", "manipulator", False)]
SOLVES = [("cassie_defaults", "cassie", True, W.cassie_feet_pelvis_problem),
          ("cassie_demo", "cassie", True, W.cassie_feet_pelvis_problem),
          ("manipulator_defaults", "manipulator", False, W.manipulator_problem),
          ("humanoid_defaults", "humanoid", True, W.humanoid_problem)]


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def test_cassie_golden_topology_is_the_surveyed_one():
    """SURVEY.md 8c.1 'Cassie golden topology (exact-equality target)'."""
    g = load("topology_cassie")
    assert list(g["joint_names"]) == [
        "universe", "root_joint", "LeftHipRoll", "LeftHipYaw", "LeftHipPitch", "LeftKneePitch", "LeftShinPitch",
        "LeftTarsusPitch", "LeftAchillesSpring", "LeftFootPitch", "RightHipRoll", "RightHipYaw", "RightHipPitch",
        "RightKneePitch", "RightShinPitch", "RightTarsusPitch", "RightAchillesSpring", "RightFootPitch"]
    assert list(g["parent"]) == [0, 0, 1, 2, 3, 4, 5, 6, 7, 7, 1, 10, 11, 12, 13, 14, 15, 15]
    assert list(g["idx_q"][2:]) == [7 + i for i in range(16)] and list(g["idx_v"][2:]) == [6 + i for i in range(16)]
    names = list(g["frame_names"])
    for frame, joint, xyz in (("pelvis", 1, (0, 0, 0)), ("LeftFootFront", 9, (-0.0407, 0.107, 0)),
                              ("RightFootFront", 17, (-0.0407, 0.107, 0)), ("VectorNav", 1, (0.03155, 0, -0.07996))):
        f = names.index(frame)
        assert g["frame_parent"][f] == joint
        np.testing.assert_allclose(g["frame_placement"][f][9:], xyz, atol=1e-15)


@pytest.mark.parametrize("name,robot,ff", TOPO)
def test_product_flattener_matches_golden_topology(name, robot, ff):
    """ikb_model_from_urdf (ik_b200/csrc/urdf_model.cpp): names, parents, offsets, limits EXACTLY equal; placements
    to the last bit of the decimal literals."""
    g = load(name)
    m = ik.Model.builtin(robot, free_flyer=ff)
    assert m.names == list(g["joint_names"])
    assert m.frame_names == list(g["frame_names"])
    assert np.array_equal(m.parents, g["parent"]) and np.array_equal(m.jtypes, g["jtype"])
    assert np.array_equal(m.idx_qs, g["idx_q"]) and np.array_equal(m.idx_vs, g["idx_v"])
    assert np.array_equal(m.frame_parents, g["frame_parent"])
    assert np.array_equal(m.lowerPositionLimit, g["lower"]) and np.array_equal(m.upperPositionLimit, g["upper"])
    assert np.abs(m.jointPlacements - g["placement"]).max() < 1e-15
    assert np.abs(m.framePlacements - g["frame_placement"]).max() < 1e-15


@pytest.mark.parametrize("name,robot,ff", TOPO)
def test_oracle_reproduces_golden_fk(name, robot, ff):
    g = load(name)
    om = oracle_model(robot, ff)
    for q, poses in zip(g["fk_q"], g["fk_poses"]):
        got = np.stack([om.frame_placement(q, f) for f in range(om.nframes)])
        assert np.array_equal(got, poses)


@pytest.mark.parametrize("name,robot,ff,make", SOLVES)
def test_oracle_reproduces_golden_solves(name, robot, ff, make):
    g = load(name)
    pb = make()
    opb = oracle_problem_like(pb, oracle_model(robot, ff))
    mi, step, damp, tol = g["params"]
    q, ok, it, res = O.dls_batch(opb, g["q0"], g["targets"], O.params(int(mi), step, damp, tol))
    assert np.array_equal(ok, g["success"]) and np.array_equal(it, g["iters"])
    assert np.abs(q - g["q"]).max() < 1e-12 and np.abs(res - g["resid"]).max() < 1e-14
    e0, J0 = opb.evaluate(g["q0"][0], g["targets"][0])
    assert np.abs(e0 - g["e0"][0]).max() < 1e-14 and np.abs(J0 - g["J0"][0]).max() < 1e-13


# ---- GPU ------------------------------------------------------------------------------------------------


@pytest.mark.gpu
@pytest.mark.parametrize("name,robot,ff,make", SOLVES)
@pytest.mark.parametrize("layout", ["soa", "aos"])
def test_cuda_path_reproduces_golden_solves(name, robot, ff, make, layout):
    """Through the host-buffer C-ABI call (ikb_dls_solve_batch_host): flags and iteration counts exactly, q to 1e-6."""
    g = load(name)
    pb = make()
    mi, step, damp, tol = g["params"]
    prm = ik.dls_parameters(max_iterations=int(mi), step_length=step, damping=damp, tolerance=tol)
    q0, tg = g["q0"], g["targets"]
    if layout == "soa":
        out = ik.dls_batch_host(pb, q0.T.copy(), tg.T.copy(), prm, "f64", "soa")
        q = out["q"].T
    else:
        out = ik.dls_batch_host(pb, q0, tg, prm, "f64", "aos")
        q = out["q"]
    print("%s via %s: max|q-q_gold|=%.2e" % (name, pb.kernel_name(), np.abs(q - g["q"]).max()))
    assert np.array_equal(out["success"].astype(bool), g["success"].astype(bool))
    assert np.array_equal(out["iters"], g["iters"])
    assert np.abs(q - g["q"]).max() < 1e-6
    assert np.abs(out["resid"] - g["resid"]).max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("name,robot,ff", TOPO)
def test_cuda_fk_reproduces_golden_frame_placements(name, robot, ff):
    import torch

    g = load(name)
    m = ik.Model.builtin(robot, free_flyer=ff)
    pb = ik.InverseKinematicsProblem(m, 0)
    pb.add_frame_task("t", ik.FrameTask(m, m.frame_names[-1], ik.KinematicType.Full))
    names = m.frame_names[1:]
    q = torch.tensor(g["fk_q"].T.copy(), device="cuda:0")
    for k in range(0, len(names), 16):
        chunk = names[k:k + 16]
        got = ik.fk_batch(pb, q, chunk).cpu().numpy().reshape(len(chunk), 12, -1)
        for i, n in enumerate(chunk):
            assert np.abs(got[i].T - g["fk_poses"][:, 1 + k + i]).max() < 1e-13, n
