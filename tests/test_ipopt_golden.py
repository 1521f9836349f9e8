"""Real-Ipopt golden vectors (tools/ipopt_golden.py).  No Ipopt has been reachable so far (bench.py records the run-time
probe in its JSON line), so the comparison tests skip; what runs everywhere is the check that the cyipopt transcription
describes the same NLP as the oracle (its callbacks are the oracle's own evaluation functions)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ipopt_N20.npz")


def test_cyipopt_transcription_selftest():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "ipopt_golden.py"), "--selftest"], text=True)
    assert "selftest ok" in out


def test_probe_reports_what_is_reachable():
    sys.path.insert(0, ROOT)
    import bench
    p = bench.ipopt_probe()
    assert set(("julia", "ipopt", "cyipopt", "reachable")) <= set(p)
    assert p["reachable"] == bool(p["julia"] or p["ipopt"] or p["cyipopt"])


@pytest.mark.skipif(not os.path.exists(GOLD), reason="no real-Ipopt golden vectors: Ipopt has never been reachable (parity unpinned)")
def test_oracle_against_ipopt_golden():
    from oracle import oracle as O
    g = np.load(GOLD)
    N = int(g["N"])
    o = O.solve_batch(O.default_cfg(N), g["state"], g["ref"], g["v_des"], g["u_prev"], n_threads=8)
    ok = (g["ipopt_status"] <= 1) & (g["ipopt_status"] >= 0) & (o["status"] == 0)
    assert ok.mean() > 0.9
    assert np.abs(o["u0"] - g["u0"])[ok].max() <= 1e-5
    assert (np.abs(o["cost"] - g["cost"])[ok] <= 1e-6 * np.maximum(1.0, np.abs(g["cost"][ok]))).all()
