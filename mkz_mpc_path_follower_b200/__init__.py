"""mkz_mpc_path_follower_b200 -- B200-native batched solver for the kinematic-bicycle
path-following MPC of govvijaycal/mkz_mpc_path_follower (hot path only).

The compute lives in libmpc_b200.so (hand-written sm_100a CUDA behind a C ABI,
include/mpc_b200.h).  This package is the Python host side: the ctypes binding (capi),
the mirror of the reference's Julia module API (mpc_path_follower), the rospy-free
reference generator (gps_ref_traj), the recorded paths (paths) and the synthetic
workloads of SURVEY.md section 8(d) (workload).
"""
from . import paths, gps_ref_traj, workload  # noqa: F401

__all__ = ["paths", "gps_ref_traj", "workload", "capi", "mpc_path_follower"]
