"""Mint tests/golden/*.npz by executing the reference's own code under oracle/jaxshim.

Run in the build container (needs /root/reference):   python oracle/make_golden.py
Each fixture holds: the raw mesh, the reference Mesher's arrays (``mesher.*``), the statics handed
to Environment (``static.*``), the initial dynamic state (``init.*``) and the eight state arrays
after each checkpoint step (``s<step>.<array>``), plus meta (Q, K, tau, delta_t, scheme, float).
JAXSHIM_FLOAT=32 python oracle/make_golden.py --fp32   mints the *_f32 variants (stock-JAX-like
fp32 arithmetic: every float fp32, int*float -> fp32).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refrun                      # noqa: E402
from fvdbm_jax_b200 import meshgen             # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def perturbed_pdf(ns, dyn, centers, extent):
    """feq(rho, u) of a smooth field so no branch of the step is trivially constant."""
    x = centers[:, 0] / extent[0]
    y = centers[:, 1] / extent[1]
    rho = 1 + 0.01 * np.sin(2 * np.pi * x) * np.sin(2 * np.pi * y)
    u = 0.05 * np.stack([np.sin(2 * np.pi * y), np.sin(2 * np.pi * x)], axis=1)
    rows = [np.asarray(dyn.calc_eq(ns.jnp.asarray(rho[i]), ns.jnp.asarray(u[i]))) for i in range(rho.size)]
    return ns.jnp.asarray(np.stack(rows))


def tri_case(name, raw, scheme, bcs, steps, tau=0.8, dt=0.1, perturb=True, dim_multiplier=1, lattice="D2Q9"):
    ns = refrun.load()
    m = refrun.ref_mesher(raw)
    dyn = getattr(ns, lattice)(tau=tau, delta_t=dt)
    cells, faces, nodes = m.to_env(dyn, flux_method=scheme, dim_multiplier=dim_multiplier)
    for kind, marker, val in bcs:
        if kind == "vel":
            nodes = m.set_vel_node(nodes, marker, ns.jnp.array(val))
        else:
            nodes = m.set_rho_node(nodes, marker, val)
    if perturb:
        ext = raw.points.max(axis=0) - raw.points.min(axis=0)
        cells.pdf = perturbed_pdf(ns, dyn, np.asarray(m.cell_centers) - raw.points.min(axis=0), ext)
    env = ns.Environment(cells, faces, nodes)
    env.init()
    rec = {"raw.points": raw.points, "raw.elements": raw.elements, "raw.faces": raw.faces,
           "raw.point_markers": raw.point_markers}
    rec.update({f"mesher.{k}": v for k, v in refrun.mesher_arrays(m).items()})
    finish(name, env, rec, steps, dyn, scheme, dim_multiplier, bcs)


def quad_ldc_case(name, nx, lattice, steps, scheme="upwind", tau=0.8, dt=0.1, u_lid=0.1):
    """The hand-built structured route of tests/ldcFVDBM.ipynb c4-c10 (Environment.create +
    CustomArray.add_items, K=4 quads, ghosts in stencil slot 0 *and* slot 1), at nx x nx."""
    ns = refrun.load()
    jnp = ns.jnp
    ny = nx
    dyn = getattr(ns, lattice)(tau=tau, delta_t=dt)
    ns.Environment.dynamics = dyn
    env = ns.Environment.create(nx * ny, nx * (nx + 1) * 2, (nx + 1) * (ny + 1))
    env.faces.flux_scheme = scheme
    cell = np.arange(nx * ny).reshape(ny, nx)
    node = np.arange((nx + 1) * (ny + 1)).reshape(ny + 1, nx + 1)
    vert = np.arange(nx * (ny + 1)).reshape(ny, nx + 1)
    horz = (nx * (ny + 1) + np.arange((nx + 1) * ny)).reshape(ny + 1, nx)
    for y in range(ny):
        for x in range(nx):
            c = cell[y, x]
            env.cells.face_indices.add_items(c, jnp.asarray([vert[y, x], horz[y, x], vert[y, x + 1], horz[y + 1, x]]))
            env.cells.face_normals.add_items(c, jnp.asarray([0, 0, 1, 1]))
    env.cells.face_normals.data = jnp.where(env.cells.face_normals.data == 0, -1, env.cells.face_normals.data)
    r2 = jnp.sqrt(2)
    for y in range(ny + 1):
        for x in range(nx + 1):
            p = node[y, x]
            ring = []
            for (yy, xx) in ((y - 1, x - 1), (y - 1, x), (y, x), (y, x - 1)):
                if 0 <= yy < ny and 0 <= xx < nx:
                    ring.append(cell[yy, xx])
            boundary = y in (0, ny) or x in (0, nx)
            env.nodes.type = env.nodes.type.at[p].set(jnp.asarray(1 if boundary else 0))
            env.nodes.cells_index.add_items(p, jnp.asarray(ring))
            env.nodes.cell_dists.add_items(p, jnp.asarray([r2] * len(ring)))
            if boundary:
                lid = y == 0 and x not in (0, nx)
                env.nodes.vel = env.nodes.vel.at[p].set(jnp.asarray([u_lid if lid else 0.0, 0.0]))
    for y in range(ny):
        for x in range(nx + 1):
            j = vert[y, x]
            st = [cell[y, x - 1] if x > 0 else -2, cell[y, x] if x < nx else -2]
            env.faces.nodes_index.add_items(j, jnp.asarray([node[y, x], node[y + 1, x]]))
            env.faces.stencil_cells_index.add_items(j, jnp.asarray(st))
            env.faces.stencil_dists.add_items(j, jnp.asarray([.5, .5]))
            env.faces.n = env.faces.n.at[j].set(jnp.asarray([1, 0]))
            env.faces.L = env.faces.L.at[j].set(jnp.asarray(1))
    for y in range(ny + 1):
        for x in range(nx):
            j = horz[y, x]
            st = [cell[y - 1, x] if y > 0 else -2, cell[y, x] if y < ny else -2]
            env.faces.nodes_index.add_items(j, jnp.asarray([node[y, x], node[y, x + 1]]))
            env.faces.stencil_cells_index.add_items(j, jnp.asarray(st))
            env.faces.stencil_dists.add_items(j, jnp.asarray([.5, .5]))
            env.faces.n = env.faces.n.at[j].set(jnp.asarray([0, 1]))
            env.faces.L = env.faces.L.at[j].set(jnp.asarray(1))
    env.faces.stencil_cells_index.data = jnp.where(env.faces.stencil_cells_index.data == -2, -1,
                                                   env.faces.stencil_cells_index.data)
    env.init()
    finish(name, env, {}, steps, dyn, scheme, 1, [])


def finish(name, env, rec, steps, dyn, scheme, dim_multiplier, bcs):
    ns = refrun.load()
    rec.update({f"static.{k}": v for k, v in refrun.snapshot(env, refrun.STATIC).items()})
    rec.update({f"init.{k}": v for k, v in refrun.snapshot(env).items()})
    with contextlib.redirect_stdout(io.StringIO()):
        out = refrun.run_steps(env, steps)
    for s, arrs in out.items():
        rec.update({f"s{s}.{k}": v for k, v in arrs.items()})
    rec["meta.Q"] = np.int64(dyn.NUM_QUIVERS)
    rec["meta.K"] = np.int64(rec["static.cells.face_indices"].shape[1])
    rec["meta.tau"] = np.float64(dyn.tau)
    rec["meta.delta_t"] = np.float64(dyn.delta_t)
    rec["meta.scheme"] = np.array(str(env.faces.flux_scheme))        # "upwind" | "lax_wendroff" (cc_* use the same two)
    rec["meta.flux_method"] = np.array(scheme)
    rec["meta.steps"] = np.array(sorted(steps), dtype=np.int64)
    rec["meta.float_bits"] = np.int64(ns.float_dtype.itemsize * 8)
    rec["meta.dim_multiplier"] = np.float64(dim_multiplier)
    rec["meta.bcs"] = np.array(repr(bcs))
    suffix = "_f32" if ns.float_dtype.itemsize == 4 else ""
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{name}{suffix}.npz")
    np.savez_compressed(path, **rec)
    n = rec["static.cells.face_indices"].shape[0]
    print(f"wrote {os.path.relpath(path, ROOT)}  cells={n} steps={sorted(steps)} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


WALLS_LID = [("vel", 1, [0., 0.]), ("vel", 2, [0., 0.]), ("vel", 4, [0., 0.]), ("vel", 3, [0.1, 0.])]
CHANNEL = [("vel", 4, [0.05, 0.]), ("vel", 1, [0., 0.]), ("vel", 3, [0., 0.]), ("rho", 2, 0.95)]
CYL = [("vel", 4, [0.1, 0.]), ("vel", 3, [0., 0.]), ("vel", 1, [0., 0.]), ("vel", 5, [0., 0.]), ("rho", 2, 0.95)]
PRESSURE = [("vel", 1, [0., 0.]), ("vel", 3, [0., 0.]), ("rho", 4, 1.05), ("rho", 2, 0.95)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fp32", action="store_true", help="only the cases that have an fp32 variant")
    args = ap.parse_args()
    ns = refrun.load()
    if args.fp32 != (ns.float_dtype.itemsize == 4):
        raise SystemExit("use JAXSHIM_FLOAT=32 together with --fp32 (and neither for fp64)")
    sq = meshgen.triangulated_square
    steps = [1, 2, 5, 12]
    tri_case("ldc_tri_lw", sq(8, 6, seed=1), "lax_wendroff", WALLS_LID, steps)
    tri_case("ldc_tri_upwind", sq(8, 6, seed=1), "upwind", WALLS_LID, steps)
    tri_case("channel_lw", sq(10, 5, seed=2), "lax_wendroff", CHANNEL, steps, tau=0.65)
    tri_case("pressure_lw_dm2", sq(7, 7, seed=3), "lax_wendroff", PRESSURE, steps, tau=0.65, dim_multiplier=2.0)
    cyl = meshgen.masked_domain(24, 12, 24.0, 12.0, lambda x, y: (x - 7.0) ** 2 + (y - 6.0) ** 2 < 4.0, seed=6)
    tri_case("cylinder_lw", cyl, "lax_wendroff", CYL, [1, 2, 5], tau=0.65)
    tri_case("tri_d2q13_lw", sq(6, 5, seed=7), "lax_wendroff", WALLS_LID, [1, 2, 5], lattice="D2Q13")
    quad_ldc_case("quad_ldc_d2q13", 6, "D2Q13", [1, 2, 5, 20])
    if args.fp32:           # the fp32 set: both schemes, velocity + density nodes, dim_multiplier, obstacle,
        return              # D2Q13 and the hand-built K=4 route, each run by the reference with every float fp32
    tri_case("channel_upwind", sq(10, 5, seed=2), "upwind", CHANNEL, steps, tau=0.65)
    tri_case("nobc_lw", sq(6, 6, seed=4), "lax_wendroff", [], steps)
    tri_case("rest_upwind", sq(5, 4, seed=5), "upwind", WALLS_LID[:3] + [("vel", 3, [0., 0.])], [1, 3], perturb=False)
    tri_case("cc_ldc_upwind", sq(7, 6, seed=8), "cc_upwind", WALLS_LID, [1, 2, 5, 12])
    tri_case("cc_channel_lw", sq(9, 5, seed=9), "cc_lax_wendroff", CHANNEL, [1, 2, 5, 12], tau=0.65, dim_multiplier=1.5)
    quad_ldc_case("quad_ldc_d2q9", 6, "D2Q9", [1, 2, 5, 20])
    quad_ldc_case("quad_ldc_d2q9_lw", 5, "D2Q9", [1, 2, 5], scheme="lax_wendroff")


if __name__ == "__main__":
    main()
