"""tests/ldcFVDBM.ipynb of the reference (BASELINE configs[0]) end to end on this framework: the notebook's
hand-built structured 100x100 quad cavity (K = 4, unit cells, D2Q13, upwind; c3-c9), its loop
`env = env.step()` (c11-c12) and its check, the centre-line profiles against ref/ldc_Re100.mat (c13-c21) --
here as RMS errors instead of a plot, from the 2 x 256 reference points committed as
tests/golden/ldc_re100_centerlines.npz.

    python examples/ldc_cavity.py [--steps 500001] [--nx 100]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fvdbm_jax_b200 import D2Q13, Environment, meshgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nx", type=int, default=100)
ap.add_argument("--steps", type=int, default=500001)           # notebook: 1 + 500 000
args = ap.parse_args()

N_x = args.nx                                                  # notebook c3
Mu, Re, Tau, dt = 0.1, 100, 0.8, 0.1
U_lid = Re * Mu / N_x
dynamics = D2Q13(tau=Tau, delta_t=dt)
cells, faces, nodes = meshgen.quad_cavity(N_x, N_x, dynamics, U_lid)   # c4-c9 (Environment.create + add_items there)

env = Environment(cells, faces, nodes)                         # c10
env.init()
env.build()
t0 = time.time()
for i in range(args.steps):                                    # c11-c12, verbatim
    env = env.step()
env.sync()
wall = time.time() - t0

vel = np.asarray(env.cells.vel, dtype=np.float64).reshape(N_x, N_x, 2) / U_lid      # c13-c17
v_of_x = np.mean(vel[N_x // 2 - 1:N_x // 2 + 1, :, 1], axis=0)
u_of_y = np.mean(vel[:, N_x // 2 - 1:N_x // 2 + 1, 0], axis=1)
print(f"{N_x * N_x} cells, {args.steps} steps in {wall:.1f} s = {wall / args.steps * 1e6:.2f} us/step "
      f"({N_x * N_x * args.steps / wall / 1e6:.0f} MCUPS)")
ref_file = os.path.join(ROOT, "tests", "golden", "ldc_re100_centerlines.npz")
if os.path.exists(ref_file):                                   # c18-c21: the lid is the row y = 0 -> flip the axis
    ref = np.load(ref_file)
    xs = np.linspace(1 / (2 * N_x), 1 - 1 / (2 * N_x), N_x)
    e_v = np.sqrt(np.mean((-v_of_x - np.interp(xs, ref["x"], ref["v_of_x"])) ** 2))
    e_u = np.sqrt(np.mean((u_of_y - np.interp(np.flip(xs), ref["y"], ref["u_of_y"])) ** 2))
    print(f"RMS centre-line error vs ldc_Re100.mat: u(y) {e_u:.4f}, v(x) {e_v:.4f} (units of U_lid)")
assert np.isfinite(vel).all()
