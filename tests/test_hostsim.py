"""CPU check of the device code's per-cell functions (csrc/core.cuh) + the host planner
(csrc/plan.hpp) through tests/hostsim: same arithmetic and layout walk as k_nodes/k_fused_direct,
compared with the golden vectors minted from the reference's own code."""
import ctypes as C
import os

import numpy as np
import pytest

import golden

HERE = os.path.dirname(os.path.abspath(__file__))


def hostsim():
    lib = C.CDLL(os.path.join(HERE, "hostsim", "libhostsim.so"))
    lib.hostsim_run.restype = C.c_int
    lib.hostsim_error.restype = C.c_char_p
    return lib


def run(case, dtype, nsteps, perm=None):
    da = case.desc_arrays(dtype, perm=perm)
    real = np.dtype(dtype)
    pdf = np.zeros((da.N, da.Q), real)
    npdf = np.array(da.keep["node_pdf"], copy=True)
    nrho = np.array(da.keep["node_rho"], copy=True)
    nvel = np.array(da.keep["node_vel"], copy=True)
    prho = np.zeros((da.N,), real)
    pvel = np.zeros((da.N, 2), real)
    lib = hostsim()
    rc = lib.hostsim_run(C.byref(da.desc), nsteps, *[C.c_void_p(a.ctypes.data) for a in (pdf, npdf, nrho, nvel, prho, pvel)])
    assert rc == 0, lib.hostsim_error().decode()
    return {"cells.pdf": pdf, "nodes.pdf": npdf, "nodes.rho": nrho.reshape(-1, 1), "nodes.vel": nvel,
            "cells.rho": prho.reshape(-1, 1), "cells.vel": pvel}


@pytest.mark.parametrize("name", golden.names())
def test_hostsim_fp64_matches_reference(name):
    case = golden.Case(name)
    for s in case.steps:
        out = run(case, np.float64, s)
        for k, v in out.items():
            assert golden.rel_err(v, case.expected(s, k)) < 1e-12, (name, s, k)


@pytest.mark.parametrize("name", ["ldc_tri_lw", "channel_upwind", "cylinder_lw", "quad_ldc_d2q13"])
def test_hostsim_fp32_within_tolerance(name):
    case = golden.Case(name)
    s = case.steps[-1]
    out = run(case, np.float32, s)
    for k, v in out.items():
        assert golden.rel_err(v, case.expected(s, k)) < 1e-5, (name, s, k)


@pytest.mark.parametrize("name", ["ldc_tri_lw", "cylinder_lw"])
def test_hostsim_permutation_invariance(name):
    case = golden.Case(name)
    n = case.static["cells.face_indices"].shape[0]
    perm = np.random.default_rng(0).permutation(n).astype(np.int32)
    s = case.steps[-1]
    a = run(case, np.float64, s)
    b = run(case, np.float64, s, perm=perm)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])     # same arithmetic per cell -> bitwise equal


@pytest.mark.parametrize("seed", range(12))
def test_hostsim_random_problems_match_the_numpy_oracle(seed):
    """CPU twin of tests/test_gpu_randomized.py: random obstacle fields, boundary-condition assignments (velocity /
    density / none per marker), schemes (incl. the cc_* variants), lattices, dim_multiplier and renumberings go through
    Environment._describe -> the host planner -> the device per-cell code walked on the CPU, against the NumPy oracle."""
    import fvdbm_jax_b200 as fb
    from oracle.step_numpy import StepOracle
    from test_gpu_randomized import random_problem
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
            pytest.skip("random boundary conditions drove this case unstable")
        env = fb.Environment(cells, faces, nodes, dtype=dtype, reorder="hilbert" if seed % 2 else "rcm")
        env.init()
        da = env._describe()
        real = np.dtype(dtype)
        pdf = np.zeros((da.N, da.Q), real)
        npdf, nrho, nvel = (np.array(da.keep[k], copy=True) for k in ("node_pdf", "node_rho", "node_vel"))
        prho, pvel = np.zeros((da.N,), real), np.zeros((da.N, 2), real)
        lib = hostsim()
        rc = lib.hostsim_run(C.byref(da.desc), steps, *[C.c_void_p(a.ctypes.data) for a in (pdf, npdf, nrho, nvel, prho, pvel)])
        assert rc == 0, lib.hostsim_error().decode()
        exp = o.state()
        for name, got in (("cells.pdf", pdf), ("nodes.pdf", npdf), ("nodes.rho", nrho.reshape(-1, 1)), ("nodes.vel", nvel),
                          ("cells.rho", prho.reshape(-1, 1)), ("cells.vel", pvel)):
            if np.isfinite(exp[name]).all():
                err = golden.rel_err(got, exp[name])
                assert err < tol, f"seed {seed} {name} ({real.name}, Q{Q}, {scheme}, {steps} steps): {err:.3e}"

