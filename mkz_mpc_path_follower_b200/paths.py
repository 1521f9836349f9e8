"""Recorded-drive path tables (converted from the reference's paths/*.mat by tools/convert_paths.py)."""
import os
import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

# launch/sim_path_follow.launch:18-20 -- RFS coordinate system origin
LAT0 = 37.917929
LON0 = -122.331798
YAW0 = 0.0

PATH_NAMES = ("path1_6_20", "path2_6_20", "path3_6_20")


def load_path(which):
    """which: 1, 2, 3 or a name in PATH_NAMES.  Returns dict of (n,) float64 arrays
    with the .mat schema keys t, lat, lon, psi, x, y, v, a, df."""
    name = PATH_NAMES[which - 1] if isinstance(which, int) else which
    with np.load(os.path.join(DATA_DIR, name + ".npz")) as z:
        return {k: np.ascontiguousarray(z[k]) for k in z.files}
