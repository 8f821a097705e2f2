"""Extract the centre-lines of ref/ldc_Re100.mat (256^2 Navier-Stokes Re=100 cavity solution,
lid-normalised) that the reference notebooks plot against (tests/ldcFVDBM.ipynb c18-c21):
v(x, y=1/2) and u(x=1/2, y), into a small fixture that can travel to the GPU box.
    python oracle/make_ldc_fixture.py        (needs /root/reference)"""
import os
import numpy as np
from scipy.io import loadmat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = loadmat(os.path.join(os.environ.get("FVDBM_REFERENCE", "/root/reference"), "ref", "ldc_Re100.mat"))
u, v, x, y = d["u"], d["v"], d["x"].squeeze(), d["y"].squeeze()       # u[ix, iy]; lid at iy = 255
sol = np.concatenate((u.T[..., None], v.T[..., None]), axis=-1)        # notebook c18: sol[iy, ix, (u,v)]
out = os.path.join(ROOT, "tests", "golden", "ldc_re100_centerlines.npz")
np.savez_compressed(out, x=x, y=y, v_of_x=sol[128, :, 1], u_of_y=sol[:, 128, 0], nu=d["nu"].squeeze())
print("wrote", out, os.path.getsize(out), "bytes")
