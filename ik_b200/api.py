"""Host-side mirror of the reference's task / solver interface (namespace ik), over the C ABI.

Names, argument meaning and failure behaviour follow dazzmo/ik so the parity tests read like tests of the
reference: ``InverseKinematicsProblem`` (ik/ik/problem.hpp:9-206), ``FrameTask`` / ``KinematicType``
(frame.hpp:20,78-200), ``AlignAxisTask`` (frame.hpp:210-319), ``PostureTask`` (posture.hpp:17-86),
``dls_parameters`` (dls.hpp:24-28), ``dls_data`` (dls.hpp:34-65, data.hpp:8-28), ``dls`` (dls.hpp:111-114).
The batched entry points (``dls_batch`` on device tensors, ``dls_batch_host`` on host arrays) are the
extension the reference lacks.  All arithmetic happens in libikb200.so on the GPU; there is no CPU path.
"""
import ctypes as C
import enum
import os

import numpy as np

from . import _capi as capi

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class KinematicType(enum.IntEnum):  # frame.hpp:20
    Position = 0
    Orientation = 1
    Full = 2


class AlignAxisType(enum.IntEnum):  # frame.hpp:202
    AxisX = 0
    AxisY = 1
    AxisZ = 2


def _as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Model:
    """Flattened kinematic tree; stands in for ik::model_t = pinocchio::Model (common.hpp:17)."""

    def __init__(self, handle):
        self._h = handle
        lib = capi.lib
        self.njoints = lib.ikb_model_njoints(handle)
        self.nq = lib.ikb_model_nq(handle)
        self.nv = lib.ikb_model_nv(handle)
        self.nframes = lib.ikb_model_nframes(handle)
        self.names = [lib.ikb_model_joint_name(handle, j).decode() for j in range(self.njoints)]
        self.frame_names = [lib.ikb_model_frame_name(handle, f).decode() for f in range(self.nframes)]
        self.parents = np.zeros(self.njoints, dtype=np.int32)
        self.jtypes = np.zeros(self.njoints, dtype=np.int32)
        self.idx_qs = np.zeros(self.njoints, dtype=np.int32)
        self.idx_vs = np.zeros(self.njoints, dtype=np.int32)
        i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        capi.check(lib.ikb_model_get_topology(handle, i32(self.parents), i32(self.jtypes), i32(self.idx_qs),
                                              i32(self.idx_vs)), "ikb_model_get_topology")
        self.jointPlacements = np.zeros((self.njoints, 12))
        self.axes = np.zeros((self.njoints, 3))
        capi.check(lib.ikb_model_get_placements(handle, _dptr(self.jointPlacements), _dptr(self.axes)),
                   "ikb_model_get_placements")
        self.frame_parents = np.zeros(self.nframes, dtype=np.int32)
        self.frame_types = np.zeros(self.nframes, dtype=np.int32)
        self.framePlacements = np.zeros((self.nframes, 12))
        capi.check(lib.ikb_model_get_frames(handle, i32(self.frame_parents), i32(self.frame_types),
                                            _dptr(self.framePlacements)), "ikb_model_get_frames")

    def __del__(self):
        if getattr(self, "_h", None) and capi is not None and getattr(capi, "lib", None) is not None:  # (None at interpreter exit)
            capi.lib.ikb_model_free(self._h)
            self._h = None

    @classmethod
    def from_urdf(cls, xml_text, free_flyer=True):
        """pinocchio::urdf::buildModelFromXML(xml, JointModelFreeFlyer(), model) (cassie.cpp:34-35)."""
        data = xml_text.encode() if isinstance(xml_text, str) else xml_text
        h = C.c_void_p()
        capi.check(capi.lib.ikb_model_from_urdf(data, len(data), int(free_flyer), C.byref(h)), "ikb_model_from_urdf")
        return cls(h)

    @classmethod
    def builtin(cls, name, free_flyer=True):
        with open(os.path.join(_DATA, name + ".urdf")) as f:
            return cls.from_urdf(f.read(), free_flyer)

    def getFrameId(self, name):
        """model.getFrameId(name): nframes when the frame does not exist (common.hpp:50)."""
        return capi.lib.ikb_model_frame_id(self._h, name.encode())

    @property
    def lowerPositionLimit(self):
        lo = np.zeros(self.nq)
        capi.check(capi.lib.ikb_model_get_limits(self._h, _dptr(lo), None), "ikb_model_get_limits")
        return lo

    @property
    def upperPositionLimit(self):
        hi = np.zeros(self.nq)
        capi.check(capi.lib.ikb_model_get_limits(self._h, None, _dptr(hi)), "ikb_model_get_limits")
        return hi

    def set_limits(self, lower=None, upper=None):
        lo = _as_f64(lower) if lower is not None else None
        hi = _as_f64(upper) if upper is not None else None
        capi.check(capi.lib.ikb_model_set_limits(self._h, _dptr(lo) if lo is not None else None,
                                                 _dptr(hi) if hi is not None else None), "ikb_model_set_limits")

    def inertias(self):
        """(mass [njoints], centre of mass in the joint frame [njoints, 3]) of the bodies each joint supports."""
        mass, com = np.zeros(self.njoints), np.zeros((self.njoints, 3))
        capi.check(capi.lib.ikb_model_get_inertias(self._h, _dptr(mass), _dptr(com)), "ikb_model_get_inertias")
        return mass, com

    def set_inertias(self, mass, com):
        mass, com = _as_f64(mass), _as_f64(com)
        if mass.shape != (self.njoints,) or com.shape != (self.njoints, 3):
            raise ValueError("inertias: expected mass [%d] and com [%d, 3]" % (self.njoints, self.njoints))
        capi.check(capi.lib.ikb_model_set_inertias(self._h, _dptr(mass), _dptr(com)), "ikb_model_set_inertias")

    def neutral(self):
        q = np.zeros(self.nq)
        capi.check(capi.lib.ikb_model_neutral(self._h, _dptr(q)), "ikb_model_neutral")
        return q


class Task:  # task.hpp:19-57
    def __init__(self, dimension):
        self._dimension = dimension
        self._weighting = np.ones(dimension)

    def dimension(self):
        return self._dimension

    def weighting(self):
        return self._weighting


class FrameTask(Task):  # frame.hpp:78-200
    def __init__(self, model, frame, type=KinematicType.Full, reference_frame="universe"):
        self.type = KinematicType(type)
        super().__init__(6 if self.type == KinematicType.Full else 3)
        self.frame = frame
        self.reference_frame = reference_frame
        self.target = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)  # se3_t::Identity()

    create = classmethod(lambda cls, *a, **k: cls(*a, **k))
    target_size = 12


class FrameConstraint:  # frame.hpp:333-465 -- hard constraint: ik::dls keeps `frame` at rest relative to `reference_frame`
    def __init__(self, model, frame, type=KinematicType.Full, reference_frame="universe"):
        self.type = KinematicType(type)
        self._dim = 6 if self.type == KinematicType.Full else 3
        self.frame = frame
        self.reference_frame = reference_frame
        self.target = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)  # declared, not read by ik::dls

    create = classmethod(lambda cls, *a, **k: cls(*a, **k))

    def dimension(self):
        return self._dim


class AlignAxisTask(Task):  # frame.hpp:210-319
    def __init__(self, model, frame, axis, reference_frame="universe"):
        super().__init__(1)
        self.axis = AlignAxisType(axis)
        self.frame = frame
        self.reference_frame = reference_frame
        self.target = np.zeros(3)

    create = classmethod(lambda cls, *a, **k: cls(*a, **k))
    target_size = 3


class PostureTask(Task):  # posture.hpp:17-86
    def __init__(self, model, nj):
        super().__init__(nj)
        self.nj = nj
        self.target = np.zeros(nj)
        self.mask = np.ones(nj)

    create = classmethod(lambda cls, *a, **k: cls(*a, **k))

    @property
    def target_size(self):
        return self.nj


class CentreOfMassTask(Task):  # centre_of_mass.hpp:14-52
    def __init__(self, model, reference_frame="universe"):
        super().__init__(3)
        self.reference_frame = reference_frame
        self.target = np.zeros(3)  # centre of mass expressed in the reference frame

    create = classmethod(lambda cls, *a, **k: cls(*a, **k))
    target_size = 3


class dls_parameters:  # dls.hpp:24-28 + common.hpp:59-66
    def __init__(self, max_iterations=100, max_time=1.0, step_length=1.0, damping=1e-2, random_restart=False,
                 tolerance=1e-4):
        self.max_iterations = max_iterations
        self.max_time = max_time
        self.step_length = step_length
        self.damping = damping
        self.random_restart = random_restart
        self.tolerance = tolerance  # visitor.hpp:19

    def c(self):
        return capi.DlsParams(self.max_iterations, int(self.random_restart), self.max_time, self.step_length,
                              self.damping, self.tolerance)


class pik_parameters:  # pik.hpp:13-18 + pik_data::lambda (pik.hpp:31)
    def __init__(self, max_iterations=100, damping=1e-2, step_length=1.0, max_time=1.0, lambdas=None, tolerance=1e-4):
        self.max_iterations = max_iterations
        self.damping = damping          # declared by the reference, never read (pik.cpp uses pik_data::lambda)
        self.step_length = step_length
        self.max_time = max_time
        self.lambdas = list(lambdas) if lambdas is not None else []   # per priority level; missing levels: 1.0
        self.tolerance = tolerance

    def c(self):
        p = capi.PikParams()
        capi.lib.ikb_pik_params_default(C.byref(p))
        p.max_iterations = int(self.max_iterations)
        p.step_length = float(self.step_length)
        p.tolerance = float(self.tolerance)
        for i, v in enumerate(self.lambdas[:7]):
            p.lam[i] = float(v)
        return p


class inverse_kinematics_visitor:  # visitor.hpp:7-24 -- the stop test itself runs in the kernel
    tolerance = 1e-4


class InverseKinematicsProblem:  # problem.hpp:9-206
    def __init__(self, model, max_priority_level=0):
        self._model = model
        self._max_priority_level = max_priority_level
        self._tasks = []  # (name, task, priority) in insertion order
        self._constraints = []  # (name, FrameConstraint) in insertion order
        self._h = None
        self._device = None

    def __del__(self):
        if getattr(self, "_h", None) and capi is not None and getattr(capi, "lib", None) is not None:
            capi.lib.ikb_problem_free(self._h)
            self._h = None

    def model(self):
        return self._model

    def max_priority_level(self):
        return self._max_priority_level

    def _add(self, name, task, priority):
        if self._h is not None:
            raise RuntimeError("problem already finalized (device constants are immutable)")
        if priority > self._max_priority_level:
            raise IndexError("Maximum priority level exceeded!")  # problem.hpp:162-163
        self._tasks.append((name, task, priority))
        return task

    def add_frame_constraint(self, name, constraint):  # problem.hpp:107-118
        if self._h is not None:
            raise RuntimeError("problem already finalized (device constants are immutable)")
        self._constraints.append((name, constraint))
        return constraint

    def get_all_constraints(self):  # problem.hpp:167-169
        return [c for _, c in self._constraints]

    def add_frame_task(self, name, task, priority=0):  # problem.hpp:55-66
        return self._add(name, task, priority)

    def add_align_axis_task(self, name, task, priority=0):  # problem.hpp:94-105
        return self._add(name, task, priority)

    def add_posture_task(self, name, task, priority=0):  # problem.hpp:134-145
        return self._add(name, task, priority)

    def add_centre_of_mass_task(self, task, priority=0):  # problem.hpp:121-128 (one per problem, no name)
        self._com_task = self._add("centre_of_mass", task, priority)
        return task

    def get_centre_of_mass_task(self):  # problem.hpp:130-132
        return getattr(self, "_com_task", None)

    def _get(self, name, cls):
        for n, t, _ in self._tasks:
            if n == name and isinstance(t, cls):
                return t
        raise KeyError("%s %r does not exist" % (cls.__name__, name))  # the reference indexes out of range here

    def get_frame_task(self, name):  # problem.hpp:79-81
        return self._get(name, FrameTask)

    def get_align_axis_task(self, name):
        return self._get(name, AlignAxisTask)

    def get_posture_task(self, name):
        return self._get(name, PostureTask)

    def get_all_tasks(self, priority):  # problem.hpp:160-165
        if priority > self._max_priority_level:
            raise IndexError("Maximum priority level exceeded!")
        return [t for _, t, p in self._tasks if p == priority]

    def e_size(self, priority):  # problem.hpp:34-40
        return sum(t.dimension() for t in self.get_all_tasks(priority))

    def c_size(self):  # problem.hpp:47-53
        return sum(c.dimension() for _, c in self._constraints)

    @property
    def target_size(self):
        return sum(t.target_size for _, t, _ in self._tasks)

    def target_offset(self, task):
        off = 0
        for _, t, _ in self._tasks:
            if t is task:
                return off
            off += t.target_size
        raise KeyError("task is not part of this problem")

    def gather_targets(self):
        """Flatten the tasks' public ``target`` members (frame.hpp:189) into one per-problem target vector."""
        parts = [_as_f64(t.target).reshape(-1) for _, t, _ in self._tasks]
        return np.concatenate(parts) if parts else np.zeros(0)

    def _build_handle(self, device):
        """Create the C-ABI problem handle; device=None stops before ikb_problem_finalize (host-only handle)."""
        lib = capi.lib
        h = C.c_void_p()
        capi.check(lib.ikb_problem_create(self._model._h, self._max_priority_level, C.byref(h)), "ikb_problem_create")
        try:
            m = self._model
            for name, t, prio in self._tasks:
                w = _as_f64(t.weighting())
                if isinstance(t, PostureTask):
                    mask = _as_f64(t.mask)
                    capi.check_index(lib.ikb_problem_add_posture_task(h, t.nj, prio, _dptr(w), _dptr(mask)),
                                     "ikb_problem_add_posture_task")
                    continue
                if isinstance(t, CentreOfMassTask):
                    r = m.getFrameId(t.reference_frame)
                    if r >= m.nframes:
                        raise KeyError("centre of mass task: unknown frame %r" % (t.reference_frame,))
                    capi.check_index(lib.ikb_problem_add_com_task(h, r, prio, _dptr(w)), "ikb_problem_add_com_task")
                    continue
                f, r = m.getFrameId(t.frame), m.getFrameId(t.reference_frame)
                if f >= m.nframes or r >= m.nframes:
                    raise KeyError("task %r: unknown frame %r / %r" % (name, t.frame, t.reference_frame))
                if isinstance(t, FrameTask):
                    capi.check_index(lib.ikb_problem_add_frame_task(h, f, int(t.type), r, prio, _dptr(w)),
                                     "ikb_problem_add_frame_task")
                else:
                    capi.check_index(lib.ikb_problem_add_align_axis_task(h, f, int(t.axis), r, prio, _dptr(w)),
                                     "ikb_problem_add_align_axis_task")
            for name, c in self._constraints:
                f, r = m.getFrameId(c.frame), m.getFrameId(c.reference_frame)
                if f >= m.nframes or r >= m.nframes:
                    raise KeyError("constraint %r: unknown frame %r / %r" % (name, c.frame, c.reference_frame))
                capi.check_index(lib.ikb_problem_add_frame_constraint(h, f, int(c.type), r), "ikb_problem_add_frame_constraint")
            if device is not None:
                capi.check(lib.ikb_problem_finalize(h, device), "ikb_problem_finalize")
        except Exception:
            lib.ikb_problem_free(h)
            raise
        return h

    def specialisation(self):
        """Name of the compiled topology-specialised kernel matching this problem, or None (host-only query)."""
        h = self._build_handle(None)
        try:
            n = capi.lib.ikb_problem_specialisation(h)
            return n.decode() if n else None
        finally:
            capi.lib.ikb_problem_free(h)

    # ---- device side ----
    def finalize(self, device=0):
        if self._h is not None:
            return self
        self._h = self._build_handle(device)
        self._device = device
        return self

    def kernel_name(self, dtype="f64"):
        n = capi.lib.ikb_problem_kernel_name(self._h, _DT[dtype][0])
        return n.decode() if n else None


_DT = {"f64": (capi.F64, np.float64), "f32": (capi.F32, np.float32)}


class dls_data:  # dls.hpp:34-65 / data.hpp:8-28
    def __init__(self, problem):
        self.success = False
        self.q = np.zeros(problem.model().nq)
        self.iterations = 0  # dls_info::iterations (dls.hpp:71-74), never filled by the reference
        self.residual = 0.0


def dls(problem, q0, data=None, visitor=None, p=None):
    """vector_t ik::dls(problem, q0, data, visitor, p) (dls.hpp:111-114): one FP64 solve on the GPU."""
    problem.finalize(problem._device or 0)
    p = p or dls_parameters()
    data = data if data is not None else dls_data(problem)
    q0 = _as_f64(q0)
    tg = _as_f64(problem.gather_targets())
    q = np.zeros(problem.model().nq)
    ok, it, res = C.c_int(0), C.c_int(0), C.c_double(0)
    prm = p.c()
    capi.check(capi.lib.ikb_dls_solve(problem._h, C.byref(prm), _dptr(q0), _dptr(tg), _dptr(q), C.byref(ok),
                                      C.byref(it), C.byref(res)), "ikb_dls_solve")
    data.q, data.success, data.iterations, data.residual = q, bool(ok.value), it.value, res.value
    return q


def dls_batch_host(problem, q0, targets, p=None, dtype="f64", layout="soa", out=None):
    """Batched ik::dls on HOST arrays (numpy, ideally pinned): H2D + solve + D2H inside the call.

    layout "soa": q0 [nq, B], targets [tsz, B]; "aos": q0 [B, nq], targets [B, tsz].  Returns dict(q, success,
    iters, resid) in the same layout.  ``out`` may hold preallocated result arrays."""
    problem.finalize(problem._device or 0)
    p = p or dls_parameters()
    code, npdt = _DT[dtype]
    nq, tsz = problem.model().nq, problem.target_size
    q0 = np.ascontiguousarray(q0, dtype=npdt)
    targets = np.ascontiguousarray(targets, dtype=npdt)
    if layout == "soa":
        B = q0.shape[1]
        assert q0.shape == (nq, B) and targets.shape == (tsz, B)
        strides = lambda k: (B, 1)
        qshape = (nq, B)
    else:
        B = q0.shape[0]
        assert q0.shape == (B, nq) and targets.shape == (B, tsz)
        strides = lambda k: (1, k)
        qshape = (B, nq)
    out = out or {}
    q = out.get("q") if out.get("q") is not None else np.empty(qshape, dtype=npdt)
    success = out.get("success") if out.get("success") is not None else np.empty(B, dtype=np.uint8)
    iters = out.get("iters") if out.get("iters") is not None else np.empty(B, dtype=np.int32)
    resid = out.get("resid") if out.get("resid") is not None else np.empty(B, dtype=npdt)
    io = capi.BatchIO(q0.ctypes.data, *strides(nq), targets.ctypes.data, *strides(tsz), q.ctypes.data, *strides(nq),
                      success.ctypes.data, iters.ctypes.data, resid.ctypes.data)
    prm = p.c()
    capi.check(capi.lib.ikb_dls_solve_batch_host(problem._h, code, C.byref(prm), B, C.byref(io)),
               "ikb_dls_solve_batch_host")
    return dict(q=q, success=success, iters=iters, resid=resid)


def dls_batch(problem, q0, targets, p=None, out=None, stream=None):
    """Batched ik::dls on DEVICE tensors (torch, SoA): q0 [nq, B], targets [tsz, B], float64 or float32.

    Enqueues on ``stream`` (default: torch's current stream) and returns without synchronising.
    Returns dict(q [nq,B], success [B] uint8, iters [B] int32, resid [B])."""
    import torch

    p = p or dls_parameters()
    nq, tsz = problem.model().nq, problem.target_size
    assert q0.is_cuda and targets.is_cuda and q0.dtype == targets.dtype
    problem.finalize(q0.device.index or 0)
    dtype = "f64" if q0.dtype == torch.float64 else "f32"
    B = q0.shape[1]
    assert q0.shape == (nq, B) and targets.shape == (tsz, B) and q0.is_contiguous() and targets.is_contiguous()
    out = out or {}
    q = out.get("q") if out.get("q") is not None else torch.empty((nq, B), dtype=q0.dtype, device=q0.device)
    success = out.get("success") if out.get("success") is not None else torch.empty(B, dtype=torch.uint8, device=q0.device)
    iters = out.get("iters") if out.get("iters") is not None else torch.empty(B, dtype=torch.int32, device=q0.device)
    resid = out.get("resid") if out.get("resid") is not None else torch.empty(B, dtype=q0.dtype, device=q0.device)
    io = capi.BatchIO(q0.data_ptr(), B, 1, targets.data_ptr(), B, 1, q.data_ptr(), B, 1, success.data_ptr(),
                      iters.data_ptr(), resid.data_ptr())
    s = stream if stream is not None else torch.cuda.current_stream(q0.device).cuda_stream
    prm = p.c()
    capi.check(capi.lib.ikb_dls_solve_batch(problem._h, _DT[dtype][0], C.byref(prm), B, C.byref(io), C.c_void_p(s)),
               "ikb_dls_solve_batch")
    return dict(q=q, success=success, iters=iters, resid=resid)


def pik_batch(problem, q0, targets, p=None, out=None, stream=None):
    """Batched ik::pik (pik.cpp:31-96) on DEVICE tensors; arguments and results as dls_batch."""
    import torch

    p = p or pik_parameters()
    nq, tsz = problem.model().nq, problem.target_size
    assert q0.is_cuda and targets.is_cuda and q0.dtype == targets.dtype
    problem.finalize(q0.device.index or 0)
    dtype = "f64" if q0.dtype == torch.float64 else "f32"
    B = q0.shape[1]
    assert q0.shape == (nq, B) and targets.shape == (tsz, B) and q0.is_contiguous() and targets.is_contiguous()
    out = out or {}
    q = out.get("q") if out.get("q") is not None else torch.empty((nq, B), dtype=q0.dtype, device=q0.device)
    success = out.get("success") if out.get("success") is not None else torch.empty(B, dtype=torch.uint8, device=q0.device)
    iters = out.get("iters") if out.get("iters") is not None else torch.empty(B, dtype=torch.int32, device=q0.device)
    resid = out.get("resid") if out.get("resid") is not None else torch.empty(B, dtype=q0.dtype, device=q0.device)
    io = capi.BatchIO(q0.data_ptr(), B, 1, targets.data_ptr(), B, 1, q.data_ptr(), B, 1, success.data_ptr(),
                      iters.data_ptr(), resid.data_ptr())
    s = stream if stream is not None else torch.cuda.current_stream(q0.device).cuda_stream
    prm = p.c()
    capi.check(capi.lib.ikb_pik_solve_batch(problem._h, _DT[dtype][0], C.byref(prm), B, C.byref(io), C.c_void_p(s)),
               "ikb_pik_solve_batch")
    return dict(q=q, success=success, iters=iters, resid=resid)


def pik_batch_host(problem, q0, targets, p=None, dtype="f64"):
    """Batched ik::pik on HOST arrays, AoS: q0 [B, nq], targets [B, tsz]."""
    problem.finalize(problem._device or 0)
    p = p or pik_parameters()
    code, npdt = _DT[dtype]
    nq, tsz = problem.model().nq, problem.target_size
    q0 = np.ascontiguousarray(q0, dtype=npdt)
    targets = np.ascontiguousarray(targets, dtype=npdt)
    B = q0.shape[0]
    assert q0.shape == (B, nq) and targets.shape == (B, tsz)
    q = np.empty((B, nq), dtype=npdt)
    success, iters, resid = np.empty(B, dtype=np.uint8), np.empty(B, dtype=np.int32), np.empty(B, dtype=npdt)
    io = capi.BatchIO(q0.ctypes.data, 1, nq, targets.ctypes.data, 1, tsz, q.ctypes.data, 1, nq, success.ctypes.data,
                      iters.ctypes.data, resid.ctypes.data)
    prm = p.c()
    capi.check(capi.lib.ikb_pik_solve_batch_host(problem._h, code, C.byref(prm), B, C.byref(io)), "ikb_pik_solve_batch_host")
    return dict(q=q, success=success, iters=iters, resid=resid)


class pik_data(dls_data):  # pik.hpp:27-49: the user-owned per-solve record (P, da, lambda live in the kernel / parameters)
    pass


def pik(problem, q0, data=None, visitor=None, p=None):
    """vector_t ik::pik(problem, q0, data, visitor, p) (pik.hpp:51-54): one problem, targets from the tasks' `target` members."""
    p = p or pik_parameters()
    if visitor is not None:
        p.tolerance = visitor.tolerance
    out = pik_batch_host(problem, np.asarray(q0, dtype=np.float64)[None, :], problem.gather_targets()[None, :], p)
    if data is not None:
        data.success = bool(out["success"][0])
        data.iterations = int(out["iters"][0])
        data.residual = float(out["resid"][0])
        data.q = out["q"][0].copy()
    return out["q"][0].copy()


class SolveQueue:
    """Pipelined stream of batches (ikb_queue_*, include/ikb200.h): up to ``depth`` batches in flight, ``merge``
    consecutive batches per kernel pair (the ~0.7 ms straggler chain is paid once per group), host-buffer copies of one
    group beside the kernels of its neighbours.  Results are those of dls_batch / dls_batch_host (bit for bit in FP64 when
    both take the same kernels, to rounding otherwise; see include/ikb200.h).

        queue = ik.SolveQueue(problem, depth=8, merge=4)
        tickets = [queue.submit(q0_k, targets_k, out=out_k) for ...]     # device tensors (torch, SoA)
        queue.wait(tickets[0]); ... ; queue.drain()
    """

    def __init__(self, problem, depth=8, merge=4, device=0):
        problem.finalize(problem._device if problem._device is not None else device)
        self._problem = problem
        self._h = C.c_void_p()
        capi.check(capi.lib.ikb_queue_create(problem._h, depth, merge, C.byref(self._h)), "ikb_queue_create")
        self._keep = {}  # ticket -> buffers that must outlive the batch

    def __del__(self):
        if getattr(self, "_h", None) and capi is not None and getattr(capi, "lib", None) is not None:
            capi.lib.ikb_queue_free(self._h)
            self._h = None

    def submit(self, q0, targets, p=None, out=None, in_stream=None):
        """Device tensors q0 [nq, B], targets [tsz, B]; inputs must be ready in ``in_stream`` order (default: torch's
        current stream).  Returns (ticket, dict(q, success, iters, resid)); the outputs are valid after wait(ticket)."""
        import torch

        problem = self._problem
        p = p or dls_parameters()
        nq, tsz = problem.model().nq, problem.target_size
        dtype = "f64" if q0.dtype == torch.float64 else "f32"
        B = q0.shape[1]
        assert q0.is_cuda and targets.is_cuda and q0.dtype == targets.dtype
        assert q0.shape == (nq, B) and targets.shape == (tsz, B) and q0.is_contiguous() and targets.is_contiguous()
        out = out or {}
        q = out.get("q") if out.get("q") is not None else torch.empty((nq, B), dtype=q0.dtype, device=q0.device)
        success = out.get("success") if out.get("success") is not None else torch.empty(B, dtype=torch.uint8, device=q0.device)
        iters = out.get("iters") if out.get("iters") is not None else torch.empty(B, dtype=torch.int32, device=q0.device)
        resid = out.get("resid") if out.get("resid") is not None else torch.empty(B, dtype=q0.dtype, device=q0.device)
        io = capi.BatchIO(q0.data_ptr(), B, 1, targets.data_ptr(), B, 1, q.data_ptr(), B, 1, success.data_ptr(),
                          iters.data_ptr(), resid.data_ptr())
        s = in_stream if in_stream is not None else torch.cuda.current_stream(q0.device).cuda_stream
        prm = p.c()
        t = capi.check_index(capi.lib.ikb_queue_submit(self._h, _DT[dtype][0], C.byref(prm), B, C.byref(io), C.c_void_p(s)),
                             "ikb_queue_submit")
        res = dict(q=q, success=success, iters=iters, resid=resid)
        self._keep[t] = (q0, targets, res)
        return t, res

    def submit_host(self, q0, targets, p=None, dtype="f64", layout="soa", out=None):
        """Host arrays (numpy; pinned -- ikb_host_alloc -- for the copies to overlap), layouts as dls_batch_host."""
        problem = self._problem
        p = p or dls_parameters()
        code, npdt = _DT[dtype]
        nq, tsz = problem.model().nq, problem.target_size
        assert q0.dtype == npdt and targets.dtype == npdt and q0.flags.c_contiguous and targets.flags.c_contiguous
        if layout == "soa":
            B = targets.shape[1]
            assert targets.shape == (tsz, B)
            strides = lambda k: (B, 1)
            qshape = (nq, B)
        else:
            B = targets.shape[0]
            assert targets.shape == (B, tsz)
            strides = lambda k: (1, k)
            qshape = (B, nq)
        # q0 of shape (nq,): ONE initial guess for the whole batch (batch_stride = 0, include/ikb200.h) -- nq values cross
        # the host link instead of B * nq
        q0_strides = (1, 0) if q0.ndim == 1 else strides(nq)
        assert q0.shape == ((nq,) if q0.ndim == 1 else qshape)
        out = out or {}
        q = out.get("q") if out.get("q") is not None else np.empty(qshape, dtype=npdt)
        success = out.get("success") if out.get("success") is not None else np.empty(B, dtype=np.uint8)
        iters = out.get("iters") if out.get("iters") is not None else np.empty(B, dtype=np.int32)
        resid = out.get("resid") if out.get("resid") is not None else np.empty(B, dtype=npdt)
        io = capi.BatchIO(q0.ctypes.data, *q0_strides, targets.ctypes.data, *strides(tsz), q.ctypes.data, *strides(nq),
                          success.ctypes.data, iters.ctypes.data, resid.ctypes.data)
        prm = p.c()
        t = capi.check_index(capi.lib.ikb_queue_submit_host(self._h, code, C.byref(prm), B, C.byref(io)), "ikb_queue_submit_host")
        res = dict(q=q, success=success, iters=iters, resid=resid)
        self._keep[t] = (q0, targets, res)
        return t, res

    def wait(self, ticket):
        capi.check(capi.lib.ikb_queue_wait(self._h, ticket), "ikb_queue_wait")
        return self._keep.pop(ticket, (None, None, None))[2]

    def wait_on_stream(self, ticket, stream=None):
        import torch

        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        capi.check(capi.lib.ikb_queue_wait_on_stream(self._h, ticket, C.c_void_p(s)), "ikb_queue_wait_on_stream")

    def flush(self):
        capi.check(capi.lib.ikb_queue_flush(self._h), "ikb_queue_flush")

    def drain(self):
        capi.check(capi.lib.ikb_queue_drain(self._h), "ikb_queue_drain")
        self._keep.clear()


def fk_batch(problem, q, frames, out=None, stream=None):
    """Batched framesForwardKinematics (data.cpp:28-29) on device tensors: q [nq, B] -> [len(frames)*12, B]."""
    import torch

    m = problem.model()
    problem.finalize(q.device.index or 0)
    ids = np.array([m.getFrameId(f) if isinstance(f, str) else int(f) for f in frames], dtype=np.int32)
    if (ids >= m.nframes).any():
        raise KeyError("unknown frame in %r" % (frames,))
    B = q.shape[1]
    dtype = "f64" if q.dtype == torch.float64 else "f32"
    if out is None:
        out = torch.empty((len(ids) * 12, B), dtype=q.dtype, device=q.device)
    s = stream if stream is not None else torch.cuda.current_stream(q.device).cuda_stream
    capi.check(capi.lib.ikb_fk_batch(problem._h, _DT[dtype][0], B, C.c_void_p(q.data_ptr()), B, 1, len(ids),
                                     ids.ctypes.data_as(C.POINTER(C.c_int32)), C.c_void_p(out.data_ptr()),
                                     C.c_void_p(s)), "ikb_fk_batch")
    return out


def kernel_launch_count():
    return int(capi.lib.ikb_kernel_launch_count())
