"""ik_b200 -- B200-native batched inverse kinematics (drop-in for the DLS path of dazzmo/ik).

The package is a thin host-side mirror of the reference's task/solver API over libikb200.so
(hand-written sm_100a CUDA behind the C ABI of include/ikb200.h).  There is no CPU fallback.
"""
from . import _capi  # noqa: F401  (fails loudly if the CUDA library is missing)
from .api import (AlignAxisTask, AlignAxisType, CentreOfMassTask, FrameConstraint, FrameTask,
                  InverseKinematicsProblem, KinematicType, Model, MultiGPU,
                  PostureTask, SolveQueue, dls, dls_batch, dls_batch_host, dls_data, dls_parameters, fk_batch,
                  inverse_kinematics_visitor, kernel_launch_count, pik, pik_batch, pik_batch_host, pik_data,
                  pik_parameters)

__all__ = ["AlignAxisTask", "AlignAxisType", "CentreOfMassTask", "FrameConstraint", "FrameTask", "InverseKinematicsProblem", "KinematicType", "Model", "MultiGPU",
           "PostureTask", "SolveQueue", "dls", "dls_batch", "dls_batch_host", "dls_data", "dls_parameters", "fk_batch",
           "inverse_kinematics_visitor", "kernel_launch_count", "pik", "pik_batch", "pik_batch_host", "pik_data",
           "pik_parameters"]
