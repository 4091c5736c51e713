"""Oracle task errors / Jacobians from first principles, the reference's documented quirks, and the provisional
known answers of SURVEY.md 8c (two independent restatements agreeing is the check -- parity is unpinned)."""
import numpy as np
import pytest

from ik_b200 import workloads as W
from oracle import oracle as O
from tests.common import oracle_model

RNG = np.random.default_rng(7)


def cassie():
    m = oracle_model("cassie")
    q0 = m.neutral()
    q0[7:] = W.CASSIE_STANDING
    return m, q0


def feet_pelvis(m):
    pb = O.Problem(m)
    pb.add_frame_task("pelvis", O.FULL)
    pb.add_frame_task("LeftFootFront", O.POSITION)
    pb.add_frame_task("RightFootFront", O.POSITION)
    return pb


def test_fk_known_answers():
    """FK at the SRDF standing pose with identity base."""
    m, q0 = cassie()
    p = lambda n: m.frame_placement(q0, m.frame_id(n))[9:]
    assert np.allclose(p("LeftFootFront"), [0.079972781092, 0.135022763736, -1.005061948739], atol=1e-11)
    assert np.allclose(p("RightFootFront"), [0.079972781092, -0.135022763735, -1.005061948739], atol=1e-11)
    assert np.allclose(p("LeftFootBack"), [-0.078053208124, 0.135020396502, -1.004535900536], atol=1e-11)
    assert np.allclose(p("VectorNav"), [0.03155, 0.0, -0.07996], atol=1e-15)


def test_fk_hand_derived_first_joint():
    """LeftHipRoll placement is rpy=(0, pi/2, 0), xyz=(-0.049, 0.135, 0) under the pelvis (cassie.urdf:281): at q=0 the
    joint frame's z axis is the pelvis x axis and its origin is the xyz offset."""
    m = oracle_model("cassie")
    oMi = m.fk(m.neutral())
    R = oMi[2][:9].reshape(3, 3)
    assert np.allclose(R[:, 2], [1, 0, 0], atol=1e-15) and np.allclose(oMi[2][9:], [-0.049, 0.135, 0.0])
    # rotating the joint by 0.3 rad keeps its own z axis and origin
    q = m.neutral()
    q[7] = 0.3
    oMi2 = m.fk(q)
    assert np.allclose(oMi2[2][:9].reshape(3, 3)[:, 2], [1, 0, 0], atol=1e-15)
    assert np.allclose(oMi2[2][9:], oMi[2][9:])
    # and moves the child origin on a circle of the right radius around that axis
    d0, d1 = oMi[3][9:] - oMi[2][9:], oMi2[3][9:] - oMi2[2][9:]
    assert abs(np.linalg.norm(d0) - 0.09) < 1e-15 and abs(np.linalg.norm(d1) - 0.09) < 1e-15
    assert abs(np.arccos(np.clip(d0 @ d1 / 0.09 ** 2, -1, 1)) - 0.3) < 1e-12


def test_single_solve_known_answers():
    """SURVEY 8c 'Single solve (BASELINE config 1 candidate)'."""
    m, q0 = cassie()
    pb = feet_pelvis(m)
    lf = m.frame_placement(q0, m.frame_id("LeftFootFront"))[9:]
    rf = m.frame_placement(q0, m.frame_id("RightFootFront"))[9:]
    tg = np.concatenate([O.se3(), O.se3(p=lf + [0.05, 0.0, 0.10]), O.se3(p=rf)])
    e, J = pb.evaluate(q0, tg)
    assert abs(e @ e - 0.020663451724) < 1e-11
    q, ok, it, res, _ = O.dls(pb, q0, tg)
    assert ok and it == 1 and abs(res - 5.5754e-05) < 1e-8
    assert np.allclose(q[7:11], [1.034213001968e-02, -2.487453189466e-03, 4.961980803968e-01, -1.227703791318],
                       atol=1e-11)
    q, ok, it, res, _ = O.dls(pb, q0, tg, O.params(200, 0.1, 0.1))  # demo parameters, cassie.cpp:107-109
    assert ok and it == 27


def test_jacobian_support_structure():
    """supports are 6 / 13 / 13 columns, 105 non-zeros of 264 for the Cassie feet+pelvis problem."""
    m, q0 = cassie()
    pb = feet_pelvis(m)
    q = q0.copy()
    q[7:] += RNG.uniform(-0.1, 0.1, 16)
    q = m.clip(q)
    tg = np.concatenate([rand_target(), rand_target(), rand_target()])
    _, J = pb.evaluate(q, tg)
    nz = np.abs(J) > 0
    assert nz[:6].any(axis=0).sum() == 6 and nz[6:9].any(axis=0).sum() == 13 and nz[9:].any(axis=0).sum() == 13
    assert not nz[:6, 6:].any() and not nz[6:9, 14:].any() and not nz[9:, 6:14].any()
    assert not nz[:, 12].any() and not nz[:, 20].any()  # Achilles springs support no task frame


def rand_target():
    return O.exp6(np.concatenate([RNG.uniform(-0.5, 0.5, 3), RNG.uniform(-0.6, 0.6, 3)]))


@pytest.mark.parametrize("robot,ff,frame", [("cassie", True, "LeftFootFront"), ("cassie", True, "pelvis"),
                                            ("ur5", False, "tool0"), ("humanoid", True, "rarm_effector")])
@pytest.mark.parametrize("ktype", [O.POSITION, O.ORIENTATION, O.FULL])
def test_task_jacobian_is_derivative_of_error_world_reference(robot, ff, frame, ktype):
    """With a `universe` reference the reference's Jacobian is exact: e(q (+) h v) ~ e(q) + J v (SURVEY 8a notes)."""
    m = oracle_model(robot, ff)
    pb = O.Problem(m)
    pb.add_frame_task(frame, ktype, weight=RNG.uniform(0.5, 2.0, 6 if ktype == O.FULL else 3))
    q = m.neutral()
    nrev = m.nv - (6 if ff else 0)
    q[m.nq - nrev:] = RNG.uniform(-0.6, 0.6, nrev)
    if ff:
        q = m.integrate(q, np.concatenate([RNG.uniform(-0.3, 0.3, 6), np.zeros(nrev)]))
    tg = O.se3_mul(m.frame_placement(q, m.frame_id(frame)), rand_target())
    e, J = pb.evaluate(q, tg)
    h = 1e-6
    for c in range(m.nv):
        v = np.zeros(m.nv)
        v[c] = h
        ep, _ = pb.evaluate(m.integrate(q, v), tg)
        em, _ = pb.evaluate(m.integrate(q, -v), tg)
        assert np.abs((ep - em) / (2 * h) - J[:, c]).max() < 2e-8, c


def test_moving_reference_frame_quirk():
    """frame.hpp:169-181 differentiates only the task frame: with reference `pelvis` the free-flyer columns are NOT
    the derivative of the error (SURVEY 8a 'Moving reference frame is not differentiated') while leg columns are."""
    m, q0 = cassie()
    pb = O.Problem(m)
    pb.add_frame_task("LeftFootFront", O.POSITION, ref="pelvis")
    tg = O.se3(p=[0.0, 0.1, -0.7])
    e, J = pb.evaluate(q0, tg)
    h = 1e-6
    err = np.zeros(m.nv)
    for c in range(m.nv):
        v = np.zeros(m.nv)
        v[c] = h
        ep, _ = pb.evaluate(m.integrate(q0, v), tg)
        em, _ = pb.evaluate(m.integrate(q0, -v), tg)
        err[c] = np.abs((ep - em) / (2 * h) - J[:, c]).max()
    assert err[6:].max() < 1e-7 and err[:6].max() > 0.1


def test_position_task_is_linear_part_of_se3_log():
    """frame.hpp:54: a Position task is log6(fMt).linear(), NOT the position difference."""
    m, q0 = cassie()
    pb = O.Problem(m)
    pb.add_frame_task("LeftFootFront", O.POSITION)
    f = m.frame_id("LeftFootFront")
    oMf = m.frame_placement(q0, f)
    tg = O.se3(p=oMf[9:] + [0.03, -0.02, 0.05])
    e, _ = pb.evaluate(q0, tg)
    fMt = O.se3_actinv(oMf, tg)
    assert np.allclose(e, O.log6(fMt)[:3], atol=1e-15)
    assert np.abs(e - fMt[9:]).max() > 1e-3  # differs from the plain position difference


def test_weighting_scales_rows():
    m, q0 = cassie()
    w = np.array([2.0, 0.5, 3.0])
    a, b = O.Problem(m), O.Problem(m)
    a.add_frame_task("LeftFootFront", O.POSITION)
    b.add_frame_task("LeftFootFront", O.POSITION, weight=w)
    tg = O.se3(p=[0.1, 0.1, -0.9])
    ea, Ja = a.evaluate(q0, tg)
    eb, Jb = b.evaluate(q0, tg)
    assert np.allclose(eb, w * ea) and np.allclose(Jb, w[:, None] * Ja)


def test_priority_stacking_and_stop_test_use_priority_zero_only():
    """dls.cpp:20-24 stacks every level; visitor.hpp:19 tests level 0 only."""
    m, q0 = cassie()
    pb = O.Problem(m, 1)
    pb.add_frame_task("LeftFootFront", O.POSITION, priority=1)
    pb.add_frame_task("pelvis", O.FULL, priority=0)
    assert pb.e_size(0) == 6 and pb.e_size(1) == 3 and pb.rows == 9
    tg = np.concatenate([O.se3(p=[0.3, 0.3, -0.5]), O.se3()])  # insertion order: foot target first
    e, J = pb.evaluate(q0, tg)
    assert np.allclose(e[:6], 0) and np.abs(e[6:]).max() > 0.1  # stacked order: level 0 rows first
    q, ok, it, res, _ = O.dls(pb, q0, tg)
    assert ok and it == 0 and res < 1e-4 and np.array_equal(q, q0)  # level-1 error ignored by the stop test


def test_non_convergence_returns_last_iterate():
    """dls.cpp:76-77: success=false and the LAST iterate (not q0, despite dls.hpp:96-97)."""
    m, q0 = cassie()
    pb = feet_pelvis(m)
    tg = np.concatenate([O.se3(), O.se3(p=[2.0, 2.0, 2.0]), O.se3(p=[-2.0, -2.0, 2.0])])  # unreachable
    q, ok, it, res, _ = O.dls(pb, q0, tg, O.params(max_iterations=7))
    assert not ok and it == 7 and not np.allclose(q, q0)
    assert np.all(q[7:] >= m.flat["lower"][7:]) and np.all(q[7:] <= m.flat["upper"][7:])


def test_align_axis_and_posture_tasks():
    m, q0 = cassie()
    pb = O.Problem(m)
    pb.add_align_axis_task("LeftFootFront", 1)
    pb.add_posture_task(16, mask=np.r_[np.ones(8), np.zeros(8)])
    tg = np.concatenate([[2.0, 0.0, 0.0], np.zeros(16)])
    e, J = pb.evaluate(q0, tg)
    f = m.frame_id("LeftFootFront")
    R = m.frame_placement(q0, f)[:9].reshape(3, 3)
    assert abs(e[0] - (1 - R[:, 1] @ [1, 0, 0])) < 1e-15      # frame.hpp:262 (target normalised)
    assert np.allclose(e[1:9], q0[7:15]) and np.allclose(e[9:], 0)  # posture.hpp:52 with mask
    assert np.array_equal(J[1:, 6:], np.eye(16)) and not J[1:, :6].any()  # posture.hpp:64
    # d(1 - r.t)/dq against central differences on the leg columns (world reference -> exact)
    h = 1e-6
    for c in range(6, m.nv):
        v = np.zeros(m.nv)
        v[c] = h
        ep, _ = pb.evaluate(m.integrate(q0, v), tg)
        em, _ = pb.evaluate(m.integrate(q0, -v), tg)
        assert abs((ep[0] - em[0]) / (2 * h) - J[0, c]) < 1e-8


def test_frame_constraint_jacobian_and_null_space_projection():
    """FrameConstraint (frame.hpp:333-465) and its use in ik::dls (dls.cpp:26-34,44-52), from first principles: Jc v is the
    velocity of the frame relative to the reference frame, expressed in the frame (checked by differencing rMf along
    q (+) eps v, for `universe` and for a moving reference frame and every KinematicType), the projected step satisfies
    Jc dq = 0, and a solve with the right foot pinned leaves it where it was."""
    m, q0 = cassie()
    rng = np.random.default_rng(5)
    q = m.integrate(q0, 0.3 * rng.standard_normal(m.nv))
    for ref in ("universe", "pelvis"):
        for ktype, rows in ((O.POSITION, slice(0, 3)), (O.ORIENTATION, slice(3, 6)), (O.FULL, slice(0, 6))):
            pb = O.Problem(m)
            pb.add_frame_constraint("RightFootFront", ktype, ref)
            Jc = pb.constraint_jacobian(q)
            f, r = m.frame_id("RightFootFront"), m.frame_id(ref)
            rel = lambda qq: O.se3_actinv(m.frame_placement(qq, r), m.frame_placement(qq, f))
            h = 1e-6
            for _ in range(4):
                v = rng.standard_normal(m.nv)
                fd = (O.log6(O.se3_actinv(rel(q), rel(m.integrate(q, h * v)))) - O.log6(O.se3_actinv(rel(q), rel(m.integrate(q, -h * v))))) / (2 * h)
                assert np.abs(Jc @ v - fd[rows]).max() < 1e-7, (ref, ktype)
    # ik::dls with the right foot pinned (Full) while the left foot and the pelvis are asked to move
    pb = O.Problem(m)
    pb.add_frame_task("pelvis", O.FULL)
    pb.add_frame_task("LeftFootFront", O.POSITION)
    pb.add_frame_constraint("RightFootFront", O.FULL)
    lf = m.frame_placement(q0, m.frame_id("LeftFootFront"))[9:]
    tg = np.concatenate([O.se3(p=[0.0, 0.02, -0.03]), O.se3(p=lf + np.array([0.03, 0.0, 0.05]))])
    q1, ok, it, res, dq = O.dls(pb, q0, tg, O.params(max_iterations=1))
    assert np.abs(pb.constraint_jacobian(q0) @ dq).max() < 1e-12            # the step lies in the null space of Jc
    qf, ok, it, res, _ = O.dls(pb, q0, tg, O.params(step_length=0.25, max_iterations=200))
    rf0 = m.frame_placement(q0, m.frame_id("RightFootFront"))
    rff = m.frame_placement(qf, m.frame_id("RightFootFront"))
    # pinned along the whole path; the projected step (dls.cpp:52 projects AFTER the damped solve) stalls short of the
    # tolerance here, like the reference would -- the error still drops from 4.7e-3 to 1e-3
    assert np.abs(rff - rf0).max() < 1e-4 and res < 4.7e-3 / 3
    # without the constraint the right foot is dragged along by the pelvis
    pb2 = O.Problem(m)
    pb2.add_frame_task("pelvis", O.FULL)
    pb2.add_frame_task("LeftFootFront", O.POSITION)
    qu = O.dls(pb2, q0, tg, O.params(step_length=0.25, max_iterations=200))[0]
    assert np.abs(m.frame_placement(qu, m.frame_id("RightFootFront")) - rf0).max() > 1e-2


def test_centre_of_mass_task_from_first_principles():
    """CentreOfMassTask (centre_of_mass.hpp:14-52, data.cpp:31-34): the oracle's centre of mass against a brute-force sum over
    the bodies, its Jacobian against central differences, the task error / Jacobian in a moving reference frame (which
    the reference does not differentiate: J = R_r^T Jcom), and the masses read from the URDF."""
    m, q0 = cassie()
    assert abs(m.flat["mass"].sum() - 34.676752) < 1e-9 and m.flat["mass"][1] == 10.33   # cassie.urdf <inertial><mass>
    rng = np.random.default_rng(8)
    q = m.integrate(q0, 0.4 * rng.standard_normal(m.nv))
    com, Jcom = m.center_of_mass(q)
    oMi = m.fk(q).reshape(-1, 12)
    brute = sum(m.flat["mass"][j] * (oMi[j][:9].reshape(3, 3) @ m.flat["com"][j] + oMi[j][9:]) for j in range(1, m.njoints))
    assert np.abs(com - brute / m.flat["mass"].sum()).max() < 1e-14
    h = 1e-6
    for c in range(m.nv):
        v = np.zeros(m.nv)
        v[c] = h
        fd = (m.center_of_mass(m.integrate(q, v))[0] - m.center_of_mass(m.integrate(q, -v))[0]) / (2 * h)
        assert np.abs(Jcom[:, c] - fd).max() < 1e-8, c
    for ref in ("universe", "pelvis"):
        pb = O.Problem(m)
        pb.add_com_task(ref, weight=[1.0, 2.0, 0.5])
        tg = np.array([0.01, -0.02, -0.5])
        e, J = pb.evaluate(q, tg)
        M = m.frame_placement(q, m.frame_id(ref))
        R, p = M[:9].reshape(3, 3), M[9:]
        assert np.abs(e - np.array([1.0, 2.0, 0.5]) * (R.T @ (com - p) - tg)).max() < 1e-14
        assert np.abs(np.asarray(J).reshape(3, m.nv) - np.diag([1.0, 2.0, 0.5]) @ R.T @ Jcom).max() < 1e-14
    # a solve: shift the centre of mass 3 cm sideways and 2 cm down, feet pinned by position tasks
    pb = O.Problem(m)
    pb.add_com_task("universe")
    pb.add_frame_task("LeftFootFront", O.POSITION)
    pb.add_frame_task("RightFootFront", O.POSITION)
    c0 = m.center_of_mass(q0)[0]
    lf = m.frame_placement(q0, m.frame_id("LeftFootFront"))[9:]
    rf = m.frame_placement(q0, m.frame_id("RightFootFront"))[9:]
    tg = np.concatenate([c0 + [0.0, 0.03, -0.02], O.se3(p=lf), O.se3(p=rf)])
    qf, ok, it, res, _ = O.dls(pb, q0, tg)
    assert ok and np.abs(m.center_of_mass(qf)[0] - (c0 + [0.0, 0.03, -0.02])).max() < 1e-2
