"""tests/flow_over_cyl.ipynb of the reference, end to end on this framework (no JAX, no Triangle):
synthetic triangulation of the same channel + cylinder, the same boundary conditions, the same
loop -- `env = env.step()` -- and the same outputs (cell velocity / density, VTK file).

    python examples/flow_over_cyl.py [--scale 4] [--steps 20000]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fvdbm_jax_b200 import D2Q9, Environment, Mesher, meshgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=4)
ap.add_argument("--steps", type=int, default=20000)
ap.add_argument("--vtk", default="")
args = ap.parse_args()

mesh = meshgen.cylinder_channel(scale=args.scale)              # notebook c3-c9 (meshpy.triangle there)
mesher = Mesher()
mesher.import_meshpy(mesh)
mesher.calc_mesh_properties()

Re, nu, L, dt, rho = 100, 0.05, 50, 0.1, 0.95                  # notebook c14
U_lattice = Re * nu / L
Tau = nu / (1.0 / 3.0) + 0.5
dynamics = D2Q9(tau=Tau, delta_t=dt)

cells, faces, nodes = mesher.to_env(dynamics, flux_method="lax_wendroff")          # notebook c15
nodes = mesher.set_vel_node(nodes, marker=4, velocity=np.array([U_lattice, 0.0]))
nodes = mesher.set_vel_node(nodes, marker=3, velocity=np.array([0.0, 0.0]))
nodes = mesher.set_vel_node(nodes, marker=1, velocity=np.array([0.0, 0.0]))
nodes = mesher.set_vel_node(nodes, marker=5, velocity=np.array([0.0, 0.0]))
nodes = mesher.set_rho_node(nodes, marker=2, rho=rho)

env = Environment(cells, faces, nodes)                         # notebook c16
env.init()
env.build()                                                    # engine creation (context, upload) outside the timing
t0 = time.time()
for i in range(args.steps):                                    # notebook c17, verbatim: single steps are batched by the
    env = env.step()                                           # engine (flushed every 50 steps / before any read)
env.sync()
dt_wall = time.time() - t0
vel, dens = env.cells.vel, env.cells.rho                       # notebook c18-c20
mag = np.sqrt(np.sum(vel ** 2, axis=-1))
n = vel.shape[0]
print(f"{n} cells, {args.steps} steps in {dt_wall:.2f} s = {n * args.steps / dt_wall / 1e6:.0f} MCUPS; "
      f"|u| max {mag.max():.4f}, rho in [{dens.min():.4f}, {dens.max():.4f}]")
assert np.isfinite(vel).all() and mag.max() < 0.5
if args.vtk:
    print("wrote", mesher.to_vtk(env, args.vtk))               # notebook c21
