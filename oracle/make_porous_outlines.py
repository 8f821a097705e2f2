"""Extract the 60 obstacle outlines of the reference's porous-flow case (tests/test_bmp.mat, loaded at
tests/porous_flow.ipynb c7: x, y, id -> one closed polygon per id) into a small data file that can
travel to the GPU box (the reference tree cannot).
    python oracle/make_porous_outlines.py        (needs /root/reference)"""
import os
import numpy as np
from scipy.io import loadmat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = loadmat(os.path.join(os.environ.get("FVDBM_REFERENCE", "/root/reference"), "tests", "test_bmp.mat"))
x, y, oid = (d[k].reshape(-1) for k in ("x", "y", "id"))
out = os.path.join(ROOT, "fvdbm_jax_b200", "data", "porous_outlines.npz")
np.savez_compressed(out, x=x.astype(np.uint8), y=y.astype(np.uint8), id=oid.astype(np.uint8))
print("wrote", out, os.path.getsize(out), "bytes;", x.size, "points,", np.unique(oid).size, "outlines")
