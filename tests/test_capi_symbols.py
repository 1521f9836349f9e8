"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/mpc_b200.h declares (no compute without a GPU), and it refuses to run without one."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpcb200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from mkz_mpc_path_follower_b200 import capi
    L = capi.lib()
    names = _declared()
    assert "mpcb200_solve_batch" in names and "mpcb200_create" in names and len(names) >= 10
    for n in names:
        assert hasattr(L, n), n
    hdr = open(os.path.join(ROOT, "include", "mpc_b200.h")).read()
    assert L.mpcb200_version() == int(re.search(r"#define MPCB200_VERSION (\d+)", hdr).group(1))


def test_default_config_matches_reference_constants():
    from mkz_mpc_path_follower_b200 import capi
    c = capi.default_config(8)
    # MKZMPCPathFollower.jl:28-48
    assert (c.N, c.dt, c.dt_control, c.L_a, c.L_b) == (8, 0.2, 0.1, 1.108, 1.742)
    assert (c.v_min, c.v_max, c.a_max, c.steer_max, c.a_dmax, c.steer_dmax) == (0.0, 20.0, 1.0, 0.5, 1.5, 0.5)
    assert c.tol == 1e-8 and c.start_mode == capi.START_ZERO


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, not fall back."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    from mkz_mpc_path_follower_b200 import capi
    with pytest.raises(capi.MpcB200Error) as e:
        capi.Solver(8)
    assert e.value.code == -2
    # nothing in the package imports the oracle
    pkg = os.path.join(ROOT, "mkz_mpc_path_follower_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("# oracle", ""), fn
