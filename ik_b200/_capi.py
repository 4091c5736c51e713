"""ctypes binding of libikb200.so (include/ikb200.h).  The product has no CPU path: importing this
module fails loudly when the CUDA library has not been built."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libikb200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "ik_b200: %s is missing -- build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
        "There is no CPU fallback." % LIB_PATH
    )

lib = C.CDLL(LIB_PATH)

OK = 0
ERR_INVALID_ARG = 1
F64, F32 = 0, 1
POSITION, ORIENTATION, FULL = 0, 1, 2
TASK_FRAME, TASK_ALIGN_AXIS, TASK_POSTURE = 0, 1, 2

STATUS_NAMES = {0: "IKB_OK", 1: "IKB_ERR_INVALID_ARG", 2: "IKB_ERR_PARSE", 3: "IKB_ERR_UNKNOWN_FRAME",
                4: "IKB_ERR_UNSUPPORTED", 5: "IKB_ERR_CUDA", 6: "IKB_ERR_NO_DEVICE", 7: "IKB_ERR_NOT_FINALIZED"}


class DlsParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("random_restart", C.c_int32), ("max_time", C.c_double),
                ("step_length", C.c_double), ("damping", C.c_double), ("tolerance", C.c_double)]


class BatchIO(C.Structure):
    _fields_ = [("q0", C.c_void_p), ("q0_elem_stride", C.c_int64), ("q0_batch_stride", C.c_int64),
                ("targets", C.c_void_p), ("targets_elem_stride", C.c_int64), ("targets_batch_stride", C.c_int64),
                ("q", C.c_void_p), ("q_elem_stride", C.c_int64), ("q_batch_stride", C.c_int64),
                ("success", C.c_void_p), ("iters", C.c_void_p), ("resid", C.c_void_p),
                ("targets_format", C.c_int32), ("reserved_", C.c_int32)]


TARGETS_SE3, TARGETS_COMPACT = 0, 1


class PikParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("step_length", C.c_double), ("tolerance", C.c_double),
                ("lam", C.c_double * 7)]


class ModelDesc(C.Structure):
    _fields_ = [("njoints", C.c_int32), ("parent", C.POINTER(C.c_int32)), ("jtype", C.POINTER(C.c_int32)),
                ("placement", C.POINTER(C.c_double)), ("axis", C.POINTER(C.c_double)),
                ("lower", C.POINTER(C.c_double)), ("upper", C.POINTER(C.c_double)),
                ("joint_names", C.POINTER(C.c_char_p)), ("nframes", C.c_int32),
                ("frame_parent", C.POINTER(C.c_int32)), ("frame_placement", C.POINTER(C.c_double)),
                ("frame_names", C.POINTER(C.c_char_p))]


_vp = C.c_void_p
_i32p = C.POINTER(C.c_int32)
_dp = C.POINTER(C.c_double)

# every symbol include/ikb200.h declares: (restype, argtypes)
SIGNATURES = {
    "ikb_dls_params_default": (None, [C.POINTER(DlsParams)]),
    "ikb_model_from_urdf": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(_vp)]),
    "ikb_model_from_desc": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(_vp)]),
    "ikb_model_free": (None, [_vp]),
    "ikb_model_njoints": (C.c_int, [_vp]),
    "ikb_model_nq": (C.c_int, [_vp]),
    "ikb_model_nv": (C.c_int, [_vp]),
    "ikb_model_nframes": (C.c_int, [_vp]),
    "ikb_model_frame_id": (C.c_int, [_vp, C.c_char_p]),
    "ikb_model_joint_name": (C.c_char_p, [_vp, C.c_int]),
    "ikb_model_frame_name": (C.c_char_p, [_vp, C.c_int]),
    "ikb_model_get_topology": (C.c_int, [_vp, _i32p, _i32p, _i32p, _i32p]),
    "ikb_model_get_placements": (C.c_int, [_vp, _dp, _dp]),
    "ikb_model_get_limits": (C.c_int, [_vp, _dp, _dp]),
    "ikb_model_set_limits": (C.c_int, [_vp, _dp, _dp]),
    "ikb_model_get_frames": (C.c_int, [_vp, _i32p, _i32p, _dp]),
    "ikb_model_neutral": (C.c_int, [_vp, _dp]),
    "ikb_problem_create": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "ikb_problem_free": (None, [_vp]),
    "ikb_problem_add_frame_task": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]),
    "ikb_problem_add_align_axis_task": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]),
    "ikb_problem_add_posture_task": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp]),
    "ikb_problem_add_com_task": (C.c_int, [_vp, C.c_int, C.c_int, _dp]),
    "ikb_model_get_inertias": (C.c_int, [_vp, _dp, _dp]),
    "ikb_model_set_inertias": (C.c_int, [_vp, _dp, _dp]),
    "ikb_problem_add_frame_constraint": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "ikb_problem_c_size": (C.c_int, [_vp]),
    "ikb_problem_num_tasks": (C.c_int, [_vp]),
    "ikb_problem_task_dim": (C.c_int, [_vp, C.c_int]),
    "ikb_problem_e_size": (C.c_int, [_vp, C.c_int]),
    "ikb_problem_rows": (C.c_int, [_vp]),
    "ikb_problem_target_size": (C.c_int, [_vp]),
    "ikb_problem_task_target_offset": (C.c_int, [_vp, C.c_int]),
    "ikb_problem_finalize": (C.c_int, [_vp, C.c_int]),
    "ikb_problem_kernel_name": (C.c_char_p, [_vp, C.c_int]),
    "ikb_problem_specialisation": (C.c_char_p, [_vp]),
    "ikb_dls_solve_batch": (C.c_int, [_vp, C.c_int, C.POINTER(DlsParams), C.c_int64, C.POINTER(BatchIO), _vp]),
    "ikb_dls_solve_batch_host": (C.c_int, [_vp, C.c_int, C.POINTER(DlsParams), C.c_int64, C.POINTER(BatchIO)]),
    "ikb_dls_solve": (C.c_int, [_vp, C.POINTER(DlsParams), _dp, _dp, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int), _dp]),
    "ikb_dls_solve_ex": (C.c_int, [_vp, C.POINTER(DlsParams), _dp, _dp, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int), _dp, _dp, _dp, _dp]),
    "ikb_pik_solve_ex": (C.c_int, [_vp, C.POINTER(PikParams), _dp, _dp, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int), _dp, _dp, _dp, _dp]),
    "ikb_problem_status_string": (C.c_char_p, [_vp]),
    "ikb_load_specialisation": (C.c_int, [C.c_char_p]),
    "ikb_problem_compact_target_size": (C.c_int, [_vp]),
    "ikb_problem_task_compact_target_offset": (C.c_int, [_vp, C.c_int]),
    "ikb_multi_create": (C.c_int, [_vp, _i32p, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "ikb_multi_free": (None, [_vp]),
    "ikb_multi_device_count": (C.c_int, [_vp]),
    "ikb_multi_problem": (_vp, [_vp, C.c_int]),
    "ikb_multi_submit_host": (C.c_int64, [_vp, C.c_int, C.POINTER(DlsParams), C.c_int64, C.POINTER(BatchIO)]),
    "ikb_multi_wait": (C.c_int, [_vp, C.c_int64]),
    "ikb_multi_drain": (C.c_int, [_vp]),
    "ikb_multi_dls_solve_batch_host": (C.c_int, [_vp, C.c_int, C.POINTER(DlsParams), C.c_int64, C.POINTER(BatchIO)]),
    "ikb_multi_gather_device": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, C.c_int64, C.c_int, C.c_int, _vp]),
    "ikb_pik_params_default": (None, [C.POINTER(PikParams)]),
    "ikb_pik_solve_batch": (C.c_int, [_vp, C.c_int, C.POINTER(PikParams), C.c_int64, C.POINTER(BatchIO), _vp]),
    "ikb_pik_solve_batch_host": (C.c_int, [_vp, C.c_int, C.POINTER(PikParams), C.c_int64, C.POINTER(BatchIO)]),
    "ikb_queue_create": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(_vp)]),
    "ikb_queue_flush": (C.c_int, [_vp]),
    "ikb_queue_free": (None, [_vp]),
    "ikb_queue_submit": (C.c_int64, [_vp, C.c_int, C.POINTER(DlsParams), C.c_int64, C.POINTER(BatchIO), _vp]),
    "ikb_queue_submit_host": (C.c_int64, [_vp, C.c_int, C.POINTER(DlsParams), C.c_int64, C.POINTER(BatchIO)]),
    "ikb_queue_wait": (C.c_int, [_vp, C.c_int64]),
    "ikb_queue_wait_on_stream": (C.c_int, [_vp, C.c_int64, _vp]),
    "ikb_queue_drain": (C.c_int, [_vp]),
    "ikb_fk_batch": (C.c_int, [_vp, C.c_int, C.c_int64, _vp, C.c_int64, C.c_int64, C.c_int, _i32p, _vp, _vp]),
    "ikb_last_error": (C.c_char_p, []),
    "ikb_version": (C.c_int, []),
    "ikb_device_count": (C.c_int, []),
    "ikb_host_alloc": (_vp, [C.c_size_t]),
    "ikb_host_free": (None, [_vp]),
    "ikb_kernel_launch_count": (C.c_int64, []),
    "ikb_measure_fma_peak": (C.c_int, [C.c_int, C.c_int, _dp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = the library does not export what the header declares
    _fn.restype = _res
    _fn.argtypes = _args


class IkbError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib.ikb_last_error()
        super().__init__("%s failed: %s (%s)" % (where, STATUS_NAMES.get(code, code), msg.decode() if msg else ""))


def check(code, where):
    if code != OK:
        raise IkbError(code, where)


def check_index(value, where):
    if value < 0:
        raise IkbError(-value, where)
    return value
