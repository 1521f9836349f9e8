#!/usr/bin/env python
"""Convert the reference's recorded drives (paths/path{1,2,3}_6_20.mat, MATLAB-5)
into .npz tables that travel with the repo (the GPU box has no /root/reference).

Data only -- no reference source is copied.  Keys kept verbatim from the .mat
schema written by the reference's scripts/analysis/parse_bag.py:51-66:
t, lat, lon, psi, x, y, v, a, df  (each (n,) float64).

Run in the build container:  python tools/convert_paths.py
"""
import os
import sys
import numpy as np
import scipy.io as sio

SRC = "/root/reference/paths"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                   "mkz_mpc_path_follower_b200", "data")
KEYS = ("t", "lat", "lon", "psi", "x", "y", "v", "a", "df")


def main():
    os.makedirs(DST, exist_ok=True)
    for i in (1, 2, 3):
        m = sio.loadmat(os.path.join(SRC, "path%d_6_20.mat" % i))
        assert str(np.ravel(m["mode"])[0]) == "Real"
        out = {k: np.ascontiguousarray(np.ravel(m[k]).astype(np.float64)) for k in KEYS}
        n = out["t"].shape[0]
        assert all(v.shape == (n,) for v in out.values())
        path = os.path.join(DST, "path%d_6_20.npz" % i)
        np.savez_compressed(path, **out)
        print(path, n, os.path.getsize(path))


if __name__ == "__main__":
    sys.exit(main())
