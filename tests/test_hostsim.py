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


def _square_desc(nx, ny, scheme, dtype, periodic):
    import fvdbm_jax_b200 as fb
    from fvdbm_jax_b200 import meshgen
    raw = meshgen.triangulated_square(nx, ny, seed=11, periodic_x=periodic)
    m = fb.Mesher()
    m.import_meshpy(raw)
    m.calc_mesh_properties()
    dyn = fb.D2Q9(0.8, 0.1)
    cells, faces, nodes = m.to_env(dyn, scheme)
    for mk in ((1,) if periodic else (1, 2, 4)):
        nodes = m.set_vel_node(nodes, mk, np.array([0.0, 0.0]))
    nodes = m.set_vel_node(nodes, 3, np.array([0.1, 0.0]))
    c = m.cell_centers
    cells.pdf = dyn.calc_eq(1 + 0.01 * np.sin(2 * np.pi * c[:, 0] / nx) * np.sin(2 * np.pi * c[:, 1] / ny),
                            0.05 * np.stack([np.sin(2 * np.pi * c[:, 1] / ny), np.sin(2 * np.pi * c[:, 0] / nx)], axis=1))
    return fb.Environment(cells, faces, nodes, dtype=dtype, reorder="hilbert")._describe()


@pytest.mark.parametrize("nx,ny,scheme,dtype,periodic", [(40, 30, "lax_wendroff", np.float64, False),
                                                         (64, 48, "upwind", np.float32, True)])
def test_temporal_tiles_equal_two_single_steps(nx, ny, scheme, dtype, periodic):
    """Planner (levels, rings, local ids) + the two-iterations-per-pass schedule of api.cu::superstep /
    k_fused2, mirrored on the CPU: bit-identical to the single-step walk."""
    da = _square_desc(nx, ny, scheme, dtype, periodic)
    real = np.dtype(dtype)
    a, b = np.zeros((da.N, da.Q), real), np.zeros((da.N, da.Q), real)
    z = [np.zeros((da.P, da.Q), real), np.zeros(da.P, real), np.zeros((da.P, 2), real), np.zeros(da.N, real), np.zeros((da.N, 2), real)]
    lib = hostsim()
    lib.hostsim_run_temporal.restype = C.c_int
    assert lib.hostsim_run(C.byref(da.desc), 6, C.c_void_p(a.ctypes.data), *[C.c_void_p(x.ctypes.data) for x in z]) == 0
    assert lib.hostsim_run_temporal(C.byref(da.desc), 3, C.c_void_p(b.ctypes.data)) == 0, lib.hostsim_error().decode()
    np.testing.assert_array_equal(a, b)
