"""Randomised parity: random obstacle fields, random boundary-condition assignments, schemes,
lattices and precisions, CUDA (through Environment / the C ABI) vs the NumPy oracle on all eight
state arrays.  Seeds are fixed, so failures reproduce."""
import numpy as np
import pytest

import golden

pytestmark = pytest.mark.gpu

fb = pytest.importorskip("fvdbm_jax_b200")
from fvdbm_jax_b200 import _lib, meshgen  # noqa: E402
from oracle.step_numpy import StepOracle  # noqa: E402


def random_problem(seed):
    rng = np.random.default_rng(seed)
    nx, ny = int(rng.integers(14, 40)), int(rng.integers(10, 30))
    n_obst = int(rng.integers(0, 4))
    cx, cy = rng.uniform(0.2 * nx, 0.8 * nx, n_obst), rng.uniform(0.2 * ny, 0.8 * ny, n_obst)
    r = rng.uniform(1.2, 0.18 * min(nx, ny), n_obst)

    def inside(x, y):
        out = np.zeros(x.shape, dtype=bool)
        for i in range(n_obst):
            out |= (x - cx[i]) ** 2 + (y - cy[i]) ** 2 < r[i] ** 2
        return out
    raw = meshgen.masked_domain(nx, ny, float(nx), float(ny), inside, jitter=float(rng.uniform(0, 0.25)), seed=seed)
    m = fb.Mesher()
    m.import_meshpy(raw)
    m.calc_mesh_properties()
    Q = int(rng.choice([9, 13]))
    dyn = (fb.D2Q9 if Q == 9 else fb.D2Q13)(tau=float(rng.uniform(0.6, 1.2)), delta_t=float(rng.uniform(0.03, 0.1)))
    scheme = str(rng.choice(["upwind", "lax_wendroff", "cc_upwind", "cc_lax_wendroff"]))
    cells, faces, nodes = m.to_env(dyn, flux_method=scheme, dim_multiplier=float(rng.choice([1.0, 1.0, 0.6])))
    for marker in (1, 2, 3, 4, 5):
        kind = rng.choice(["vel", "rho", "none"], p=[0.55, 0.3, 0.15])
        if kind == "vel":
            nodes = m.set_vel_node(nodes, marker, rng.uniform(-0.05, 0.05, 2))
        elif kind == "rho":
            nodes = m.set_rho_node(nodes, marker, float(rng.uniform(0.95, 1.05)))
    c = m.cell_centers
    rho = 1 + 0.02 * np.sin(2 * np.pi * c[:, 0] / nx + rng.uniform(0, 6)) * np.cos(2 * np.pi * c[:, 1] / ny)
    u = 0.04 * np.stack([np.sin(2 * np.pi * c[:, 1] / ny), np.cos(2 * np.pi * c[:, 0] / nx + rng.uniform(0, 6))], axis=1)
    cells.pdf = dyn.calc_eq(rho, u)
    steps = int(rng.integers(3, 40))
    return dyn, Q, faces.flux_scheme, cells, faces, nodes, steps


@pytest.mark.parametrize("seed", range(12))
def test_random_problem_matches_oracle(seed):
    dyn, Q, scheme, cells, faces, nodes, steps = random_problem(seed)
    static = {"cells.face_indices": cells.face_indices, "cells.face_normals": cells.face_normals,
              "faces.nodes_index": faces.nodes_index, "faces.stencil_cells_index": faces.stencil_cells_index,
              "faces.stencil_dists": faces.stencil_dists, "faces.n": faces.n, "faces.L": faces.L,
              "nodes.type": nodes.type, "nodes.cells_index": nodes.cells_index, "nodes.cell_dists": nodes.cell_dists}
    if hasattr(faces, "alpha"):
        static["faces.alpha"] = faces.alpha
    state = {"cells.pdf": cells.pdf, "nodes.pdf": nodes.pdf, "nodes.rho": nodes.rho, "nodes.vel": nodes.vel}
    for dtype, tol in ((np.float64, 1e-11), (np.float32, 1e-5)):
        o = StepOracle(static, state, Q, dyn.tau, dyn.delta_t, scheme, dtype).step(steps)
        if not (np.isfinite(o.vel).all() and np.max(np.abs(o.vel)) < 0.3):
            pytest.skip("random boundary conditions drove this case unstable (rounding differences are amplified)")
        with fb.Environment(cells, faces, nodes, dtype=dtype, reorder="hilbert" if seed % 2 else "rcm") as env:
            env.init()
            if seed % 3 == 0:                                # the record-layout kernels explicitly (the fp32 default; fp64 and
                env.set_option(_lib.OPT_VARIANT, _lib.VARIANT_REC)          # D2Q13 run the thread-per-cell kernel over records),
            elif dtype is np.float32 and seed % 3 == 1:                       # the packed pair kernel on another third
                env.set_option(_lib.OPT_VARIANT, _lib.VARIANT_PAIR)
            env = env.step(steps)
            exp = o.state()
            for name in golden.STATE:
                obj, attr = name.split(".")
                got = getattr(getattr(env, obj), attr)
                if not np.isfinite(exp[name]).all():
                    continue                       # unphysical random BCs may blow up; only compare finite runs
                err = golden.rel_err(got, exp[name])
                assert err < tol, f"seed {seed} {name} ({np.dtype(dtype).name}, Q{Q}, {scheme}, {steps} steps): {err:.3e}"
