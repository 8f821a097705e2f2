"""tests/porous_flow.ipynb of the reference (BASELINE configs[2]) end to end on this framework: the box
(-93,0)-(279,186) minus the 60 obstacle polygons of tests/test_bmp.mat (committed outline points:
fvdbm_jax_b200/data/), synthetically triangulated (no Triangle), D2Q9 Lax-Wendroff, Tau = 0.65, dt = 0.1, density
inlet 1.05 / outlet 0.95, no-slip obstacles and walls (c26), the loop `env = env.step()` (c28) and the outputs the
notebook plots (|u|, rho) plus a VTK file.

    python examples/porous_flow.py [--scale 2] [--steps 20000] [--vtk out]      # --scale 8.5 = 2 M cells
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fvdbm_jax_b200 import D2Q9, Environment, Mesher, meshgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=2.0)
ap.add_argument("--steps", type=int, default=20000)
ap.add_argument("--vtk", default="")
args = ap.parse_args()

mesh = meshgen.porous_channel(scale=args.scale)                # notebook c7-c13 (shapely + meshpy.triangle there)
mesher = Mesher()
mesher.import_meshpy(mesh)
mesher.calc_mesh_properties()
quality = mesher.verify_stencil_geometry(verbose=False)        # the reference's mesh self-check (mesher.py:386-504)

Tau, dt, rho_in, rho_out = 0.65, 0.1, 1.05, 0.95               # notebook c25
dynamics = D2Q9(tau=Tau, delta_t=dt)
cells, faces, nodes = mesher.to_env(dynamics, flux_method="lax_wendroff")           # c26
nodes = meshgen.porous_boundary_conditions(mesher, nodes, rho_in, rho_out)

env = Environment(cells, faces, nodes)                         # c27
env.init()
env.build()
t0 = time.time()
for i in range(args.steps):                                    # c28, verbatim
    env = env.step()
env.sync()
wall = time.time() - t0
vel, dens = env.cells.vel, env.cells.rho                       # c29-c31
mag = np.sqrt(np.sum(vel ** 2, axis=-1))
n = vel.shape[0]
print(f"{n} cells ({quality['interior_faces']} interior faces, mean stencil angle {quality['angle_mean_deg']:.1f} deg), "
      f"{args.steps} steps in {wall:.2f} s = {n * args.steps / wall / 1e6:.0f} MCUPS; |u| max {mag.max():.4f}, "
      f"rho in [{dens.min():.4f}, {dens.max():.4f}], non-finite values: {env.count_nonfinite()}")
assert np.isfinite(vel).all() and mag.max() < 0.5
if args.vtk:
    print("wrote", mesher.to_vtk(env, args.vtk))
