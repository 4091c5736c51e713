"""CPU-side checks of the PRODUCT's device source (no GPU needed): the __host__ __device__ math header
(ik_b200/csrc/se3_math.cuh) and the generated topology-specialised solver bodies (ik_b200/csrc/gen/*.cuh) are compiled
with g++ by tests/cpu_harness and compared with the oracle.  The harness is test scaffolding only -- it is never
linked into libikb200.so (the product has no CPU path)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import ik_b200 as ik
from ik_b200 import workloads as W
from oracle import oracle as O
from tests.common import make_workload, oracle_model, oracle_problem_like

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "cpu_harness")
OUT = os.path.join(ROOT, "build", "cpu_harness")
_dp = C.POINTER(C.c_double)


def _pd(a):
    return a.ctypes.data_as(_dp)


def _build(name):
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(HARNESS, name + ".cpp")
    so = os.path.join(OUT, "lib" + name + ".so")
    deps = [src] + [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(ROOT, "ik_b200", "csrc")) for f in fs]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-fPIC", "-shared", "-pthread", "-x", "c++", "-ffp-contract=off",
                               "-o", so, src])
    return C.CDLL(so)


@pytest.fixture(scope="module")
def math_lib():
    return _build("harness")


@pytest.fixture(scope="module")
def spec_lib():
    return _build("spec_harness")


RNG = np.random.default_rng(7)


def test_device_exp6_log6_jlog6_match_oracle(math_lib):
    for scale in (1e-9, 1e-3, 0.7, 2.5):
        for _ in range(50):
            v = np.concatenate([RNG.uniform(-1, 1, 3), RNG.uniform(-1, 1, 3) * scale])
            M = np.zeros(12)
            math_lib.h_exp6_d(_pd(v), _pd(M))
            assert np.abs(M - O.exp6(v)).max() < 1e-12  # (1 - sin t / t) / t^2 amplifies 1-ulp sin differences
            lg = np.zeros(6)
            M = O.exp6(v)
            math_lib.h_log6_d(_pd(M), _pd(lg))
            assert np.abs(lg - O.log6(M)).max() < 1e-12
            J = np.zeros(36)
            math_lib.h_jlog6_d(_pd(M), _pd(J))
            assert np.abs(J.reshape(6, 6) - O.Jlog6(M)).max() < 1e-10


def test_device_math_float_instantiation(math_lib):
    fp = C.POINTER(C.c_float)
    for _ in range(50):
        v = RNG.uniform(-1, 1, 6)
        vf = v.astype(np.float32)
        M = np.zeros(12, dtype=np.float32)
        math_lib.h_exp6_f(vf.ctypes.data_as(fp), M.ctypes.data_as(fp))
        assert np.abs(M - O.exp6(v)).max() < 5e-6
        lg = np.zeros(6, dtype=np.float32)
        math_lib.h_log6_f(M.ctypes.data_as(fp), lg.ctypes.data_as(fp))
        assert np.abs(lg - v).max() < 2e-5


def test_device_integrate_freeflyer_matches_oracle(math_lib):
    om = oracle_model("cassie")
    for _ in range(50):
        q = W.sample_configurations(ik.Model.builtin("cassie"), 1, seed=int(RNG.integers(1 << 30)))[0]
        v = np.concatenate([RNG.uniform(-0.5, 0.5, 6), np.zeros(om.nv - 6)])
        ref = om.integrate(q, v)
        out = np.zeros(7)
        math_lib.h_integrate_ff_d(_pd(np.ascontiguousarray(q[:7])), _pd(np.ascontiguousarray(v[:6])), _pd(out))
        assert np.abs(out - ref[:7]).max() < 1e-14


def _spec_solve(lib, name, dtype, pb, q0, tg, prm):
    m = pb.model()
    fn = getattr(lib, "h_spec_%s_%s" % (name, {"f64": "d", "f32": "f", "f64p": "pd"}[dtype]))
    fn.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double, _dp, C.POINTER(C.c_int), _dp, _dp]
    lo, hi = np.ascontiguousarray(m.lowerPositionLimit), np.ascontiguousarray(m.upperPositionLimit)
    w = np.ascontiguousarray(np.concatenate([t.weighting() for _, t, _ in pb._tasks]))
    mk = np.ascontiguousarray(np.concatenate([t.mask if isinstance(t, ik.PostureTask) else np.ones(t.dimension())
                                              for _, t, _ in pb._tasks]))
    B = q0.shape[0]
    q = np.zeros((B, m.nq))
    ok = np.zeros(B, dtype=bool)
    it = np.zeros(B, dtype=np.int32)
    res = np.zeros(B)
    e0 = np.zeros((B, len(w)))
    for b in range(B):
        itc = C.c_int(0)
        r = C.c_double(0)
        ok[b] = fn(_pd(lo), _pd(hi), _pd(w), _pd(mk), _pd(np.ascontiguousarray(q0[b])), _pd(np.ascontiguousarray(tg[b])),
                   prm.max_iterations, prm.step_length, prm.damping, prm.tolerance, _pd(q[b]), C.byref(itc), C.byref(r),
                   _pd(e0[b]))
        it[b], res[b] = itc.value, r.value
    return q, ok, it, res, e0


CASES = [("cassie_feet_pelvis", "cassie", True, W.cassie_feet_pelvis_problem, W.CASSIE_STANDING),      # 3 warp roles
         ("cassie_feet_pelvis_arrow", "cassie", True, W.cassie_feet_pelvis_problem, W.CASSIE_STANDING),   # 3 roles, shared / private column split
         ("cassie_feet_pelvis_arrow_b", "cassie", True, W.cassie_feet_pelvis_problem, W.CASSIE_STANDING),   # ... factor / y in registers
         ("cassie_feet_pelvis_w1", "cassie", True, W.cassie_feet_pelvis_problem, W.CASSIE_STANDING),   # 1 role, legs interleaved
         ("cassie_feet_pelvis_w2", "cassie", True, W.cassie_feet_pelvis_problem, W.CASSIE_STANDING),   # 2 roles
         ("manipulator_tool", "manipulator", False, W.manipulator_problem, "near"),
         ("humanoid_limbs", "humanoid", True, W.humanoid_problem, "near"),                              # 5 roles, 30 rows
         ("humanoid_limbs_arrow", "humanoid", True, W.humanoid_problem, "near"),   # 5 roles, shared / private column split (no dense factor)
         ("cassie_demo", "cassie", True, W.cassie_demo_problem, W.CASSIE_STANDING),   # moving reference frame + align-axis task
         ("cassie_demo_posture", "cassie", True, W.cassie_demo_posture_problem, W.CASSIE_STANDING)]  # + masked posture, level 1


def _workload(pb, om, B, standing):
    if standing == "near":
        return make_workload(pb, om, B, start="near")
    return make_workload(pb, om, B, standing=standing)


@pytest.mark.parametrize("name,robot,ff,make,standing", CASES)
@pytest.mark.parametrize("params", ["defaults", "demo"])
def test_generated_body_f64_matches_oracle(spec_lib, name, robot, ff, make, standing, params):
    """The generated straight-line evaluate / solve / integrate reproduces the oracle's ik::dls trajectory."""
    pb = make()
    om = oracle_model(robot, ff)
    opb = oracle_problem_like(pb, om)
    B = 200 if robot != "humanoid" else 60
    q0, tg, _ = _workload(pb, om, B, standing)
    prm = O.params() if params == "defaults" else O.params(200, 1e-1, 1e-1)
    q_ref, ok_ref, it_ref, res_ref = O.dls_batch(opb, q0, tg, prm)
    q, ok, it, res, e0 = _spec_solve(spec_lib, name, "f64", pb, q0, tg, prm)
    e_ref = np.stack([opb.evaluate(q0[b], tg[b])[0] for b in range(B)])
    ok_ref = ok_ref.astype(bool)
    assert np.abs(e0 - e_ref).max() < 1e-12
    print("%s %s: converged %d/%d, mean iterations %.1f" % (name, params, ok_ref.sum(), B, it_ref.mean()))
    assert (ok == ok_ref).all()
    assert (it == it_ref).all()
    assert np.abs(q - q_ref).max() < 1e-8
    assert np.abs(res - res_ref).max() < 1e-10


@pytest.mark.parametrize("name,robot,ff,make,standing", CASES)
def test_generated_body_f32_is_close(spec_lib, name, robot, ff, make, standing):
    pb = make()
    om = oracle_model(robot, ff)
    opb = oracle_problem_like(pb, om)
    B = 200
    q0, tg, _ = _workload(pb, om, B, standing)
    q_ref, ok_ref, it_ref, _ = O.dls_batch(opb, q0, tg)
    q, ok, it, res, _ = _spec_solve(spec_lib, name, "f32", pb, q0, tg, O.params())
    same = (it == it_ref) & ok & ok_ref.astype(bool)
    assert same.mean() > 0.9
    assert (ok == ok_ref.astype(bool)).mean() > 0.98


def test_specialisation_matching_is_exact():
    """ikb_problem_specialisation (host only): the compiled fast paths are picked for exactly their (tree, task list)."""
    assert W.cassie_feet_pelvis_problem().specialisation() == "cassie_feet_pelvis"
    assert W.manipulator_problem().specialisation() == "manipulator_tool"
    assert W.humanoid_problem().specialisation() == "humanoid_limbs"
    assert W.humanoid_problem(root_task=False).specialisation() is None
    m = W.cassie_model()
    # different task type / frame / reference frame / order -> generic kernel
    pb = ik.InverseKinematicsProblem(m, 0)
    pb.add_frame_task("pelvis", ik.FrameTask(m, "pelvis", ik.KinematicType.Full))
    pb.add_frame_task("fl", ik.FrameTask(m, "LeftFootFront", ik.KinematicType.Full))
    pb.add_frame_task("fr", ik.FrameTask(m, "RightFootFront", ik.KinematicType.Position))
    assert pb.specialisation() is None
    pb = ik.InverseKinematicsProblem(m, 0)
    pb.add_frame_task("pelvis", ik.FrameTask(m, "pelvis", ik.KinematicType.Full))
    pb.add_frame_task("fl", ik.FrameTask(m, "LeftFootFront", ik.KinematicType.Position, "pelvis"))
    pb.add_frame_task("fr", ik.FrameTask(m, "RightFootFront", ik.KinematicType.Position))
    assert pb.specialisation() is None
    pb = ik.InverseKinematicsProblem(m, 0)
    pb.add_frame_task("fl", ik.FrameTask(m, "LeftFootFront", ik.KinematicType.Position))
    pb.add_frame_task("pelvis", ik.FrameTask(m, "pelvis", ik.KinematicType.Full))
    pb.add_frame_task("fr", ik.FrameTask(m, "RightFootFront", ik.KinematicType.Position))
    assert pb.specialisation() is None
    # weights are run-time constants of the specialised kernel: still the fast path
    pb = W.cassie_feet_pelvis_problem()
    pb.get_frame_task("fl").weighting()[:] = [2.0, 1.0, 0.5]
    assert pb.specialisation() == "cassie_feet_pelvis"
    # the reference demo's task set (moving reference frame + align-axis task) has its own fast path ...
    assert W.cassie_demo_problem().specialisation() == "cassie_demo"
    assert W.cassie_demo_posture_problem().specialisation() == "cassie_demo_posture"
    pb = W.cassie_demo_problem(m)
    pb.add_posture_task("posture", ik.PostureTask(m, 8), 1)      # another number of coordinates: no fast path
    assert pb.specialisation() is None
    # ... for exactly that reference frame and axis
    pb = ik.InverseKinematicsProblem(m, 1)
    pb.add_frame_task("fl", ik.FrameTask(m, "LeftFootFront", ik.KinematicType.Position, "pelvis"))
    pb.add_frame_task("pelvis", ik.FrameTask(m, "pelvis", ik.KinematicType.Full))
    pb.add_align_axis_task("align", ik.AlignAxisTask(m, "LeftFootFront", ik.AlignAxisType.AxisZ))
    assert pb.specialisation() is None
    pb = ik.InverseKinematicsProblem(m, 1)
    pb.add_frame_task("fl", ik.FrameTask(m, "LeftFootFront", ik.KinematicType.Position, "VectorNav"))
    pb.add_frame_task("pelvis", ik.FrameTask(m, "pelvis", ik.KinematicType.Full))
    pb.add_align_axis_task("align", ik.AlignAxisTask(m, "LeftFootFront", ik.AlignAxisType.AxisY))
    assert pb.specialisation() is None
    # the fixed-base Cassie is a different tree
    mf = ik.Model.builtin("cassie", free_flyer=False)
    pb = ik.InverseKinematicsProblem(mf, 0)
    pb.add_frame_task("fl", ik.FrameTask(mf, "LeftFootFront", ik.KinematicType.Position))
    assert pb.specialisation() is None


def test_branch_free_sincos_and_atan2(math_lib):
    """ik_b200/csrc/fast_math.cuh against libm over the ranges the IK path uses."""
    math_lib.h_atan2pos_d.restype = C.c_double
    math_lib.h_atan2pos_d.argtypes = [C.c_double, C.c_double]
    math_lib.h_sincos_d.argtypes = [C.c_double, _dp, _dp]
    xs = np.concatenate([RNG.uniform(-7, 7, 4000), RNG.uniform(-1e-3, 1e-3, 500), RNG.uniform(-3000, 3000, 500),
                         np.array([0.0, np.pi / 4, -np.pi / 4, np.pi / 2, np.pi, -np.pi, 1e-300, 2 * np.pi])])
    s, c = np.zeros(1), np.zeros(1)
    for x in xs:
        math_lib.h_sincos_d(float(x), _pd(s), _pd(c))
        assert abs(s[0] - np.sin(x)) < 4e-16 * max(1.0, abs(x) / 10) and abs(c[0] - np.cos(x)) < 4e-16 * max(1.0, abs(x) / 10)
    th = np.concatenate([RNG.uniform(0, np.pi, 4000), RNG.uniform(0, 1e-6, 300), np.pi - RNG.uniform(0, 1e-6, 300),
                         np.array([0.0, np.pi / 8, np.pi / 4, np.pi / 2, 3 * np.pi / 4, np.pi])])
    for t in th:
        for scale in (1.0, 1e-3, 7.0):
            got = math_lib.h_atan2pos_d(float(scale * np.sin(t)), float(scale * np.cos(t)))
            assert abs(got - np.arctan2(np.sin(t), np.cos(t))) < 5e-16
    assert math_lib.h_atan2pos_d(0.0, 0.0) == 0.0
    math_lib.h_acos_d.restype = C.c_double
    math_lib.h_acos_d.argtypes = [C.c_double]
    for x in np.concatenate([RNG.uniform(-1, 1, 4000), 1 - RNG.uniform(0, 1e-8, 500), -1 + RNG.uniform(0, 1e-8, 200),
                             np.array([1.0, -1.0, 0.0, 0.5, -0.5, 0.5000001, 1 - 1e-16])]):
        ref = np.arccos(x)
        assert abs(math_lib.h_acos_d(float(x)) - ref) <= 4e-16 * max(ref, 1e-8) + 1e-300, x


@pytest.mark.parametrize("name,robot,ff,make,standing", [c for c in CASES if c[0] in ("cassie_feet_pelvis", "cassie_feet_pelvis_w2", "cassie_feet_pelvis_arrow",
                                                                                     "cassie_feet_pelvis_arrow_b",
                                                                                     "humanoid_limbs", "humanoid_limbs_arrow")])
def test_role_distributed_solve_matches_oracle(spec_lib, name, robot, ff, make, standing):
    """psolve_w<k>: the factorisation split over the warp roles (one host thread per role, std::barrier = group barrier)
    gives the oracle's trajectory too."""
    pb = make()
    om = oracle_model(robot, ff)
    opb = oracle_problem_like(pb, om)
    B = 60
    q0, tg, _ = _workload(pb, om, B, standing)
    q_ref, ok_ref, it_ref, res_ref = O.dls_batch(opb, q0, tg)
    q, ok, it, res, _ = _spec_solve(spec_lib, name, "f64p", pb, q0, tg, O.params())
    assert (ok == ok_ref.astype(bool)).all() and (it == it_ref).all()
    assert np.abs(q - q_ref).max() < 1e-8


# ---- team-per-problem iteration (dls_team.cuh): 16 host threads = the 16 lanes of a team ----------------------
@pytest.fixture(scope="module")
def team_lib():
    return _build("team_harness")


def _team_tree(m):
    """PR[2][7][9], Pp[2][7][3], FR[2][9], Fp[2][3] of Cassie's two leg chains, from the product's flattened model."""
    chains = []
    for fname in ("LeftFootFront", "RightFootFront"):
        f = m.getFrameId(fname)
        j, ch = int(m.frame_parents[f]), []
        while j > 1:
            ch.append(j)
            j = int(m.parents[j])
        chains.append((ch[::-1], m.framePlacements[f]))
    out = []
    out += [x for ch, _ in chains for j in ch for x in m.jointPlacements[j][:9]]
    out += [x for ch, _ in chains for j in ch for x in m.jointPlacements[j][9:]]
    out += [x for _, fp in chains for x in fp[:9]]
    out += [x for _, fp in chains for x in fp[9:]]
    return np.ascontiguousarray(out, dtype=np.float64)


def _team_solve(lib, dtype, pb, q0, tg, prm):
    m = pb.model()
    fn = getattr(lib, "h_team_cassie_" + dtype)
    fn.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double, _dp, C.POINTER(C.c_int), _dp, _dp, _dp]
    tree = _team_tree(m)
    lo, hi = np.ascontiguousarray(m.lowerPositionLimit), np.ascontiguousarray(m.upperPositionLimit)
    w = np.ascontiguousarray(np.concatenate([t.weighting() for _, t, _ in pb._tasks]))
    B = q0.shape[0]
    q, ok, it, res = np.zeros((B, m.nq)), np.zeros(B, dtype=bool), np.zeros(B, dtype=np.int32), np.zeros(B)
    e0, J0 = np.zeros((B, 12)), np.zeros((B, 12, 13))
    for b in range(B):
        itc, r = C.c_int(0), C.c_double(0)
        ok[b] = fn(_pd(tree), _pd(lo), _pd(hi), _pd(w), _pd(np.ascontiguousarray(q0[b])), _pd(np.ascontiguousarray(tg[b])),
                   prm.max_iterations, prm.step_length, prm.damping, prm.tolerance, _pd(q[b]), C.byref(itc), C.byref(r),
                   _pd(e0[b]), _pd(J0[b]))
        it[b], res[b] = itc.value, r.value
    return q, ok, it, res, e0, J0


@pytest.mark.parametrize("params", ["defaults", "demo"])
def test_team_iteration_f64_matches_oracle(team_lib, params):
    """The 16-lane SPMD decomposition of ik::dls (FK by rows, LDL^T by rows, ...) follows the oracle's trajectory."""
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    B = 120
    q0, tg, _ = make_workload(pb, om, B, standing=W.CASSIE_STANDING)
    prm = O.params() if params == "defaults" else O.params(200, 1e-1, 1e-1)
    q_ref, ok_ref, it_ref, res_ref = O.dls_batch(opb, q0, tg, prm)
    q, ok, it, res, e0, J0 = _team_solve(team_lib, "d", pb, q0, tg, prm)
    # first evaluation: error and the dense task Jacobian (team slots 0-5 = free-flyer columns, 6-12 = the limb's chain)
    for b in range(0, B, 10):
        e_ref, J_ref = opb.evaluate(q0[b], tg[b])[:2]
        assert np.abs(e0[b] - e_ref).max() < 1e-12
        J = np.zeros((12, 22))
        J[:, :6] = J0[b][:, :6]
        J[6:9, [6, 7, 8, 9, 10, 11, 13]] = J0[b][6:9, 6:13]
        J[9:12, [14, 15, 16, 17, 18, 19, 21]] = J0[b][9:12, 6:13]
        assert np.abs(J - np.asarray(J_ref).reshape(12, 22)).max() < 1e-11
    assert (ok == ok_ref.astype(bool)).all()
    assert (it == it_ref).all()
    assert np.abs(q - q_ref).max() < 1e-8
    assert np.abs(res - res_ref).max() < 1e-10


def test_team_iteration_f32_is_close(team_lib):
    pb = W.cassie_feet_pelvis_problem()
    om = oracle_model("cassie")
    opb = oracle_problem_like(pb, om)
    B = 100
    q0, tg, _ = make_workload(pb, om, B, standing=W.CASSIE_STANDING)
    q_ref, ok_ref, it_ref, _ = O.dls_batch(opb, q0, tg)
    q, ok, it, res, _, _ = _team_solve(team_lib, "f", pb, q0, tg, O.params())
    same = (it == it_ref) & ok & ok_ref.astype(bool)
    assert same.mean() > 0.9
    assert (ok == ok_ref.astype(bool)).mean() > 0.98
