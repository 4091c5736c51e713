"""Test helpers (re-exported from the oracle bridge)."""
from oracle.bridge import (make_workload, oracle_frame_poses, oracle_model, oracle_problem_like,  # noqa: F401
                           urdf_text)
