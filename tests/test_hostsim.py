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
