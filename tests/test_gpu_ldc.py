"""Lid-driven cavity Re=100 (BASELINE.json configs[0], tests/ldcFVDBM.ipynb): the notebook's own
hand-built 100x100 quad mesh, D2Q13, upwind, tau=0.8, dt=0.1, U_lid=0.1, driven through the drop-in
Environment, compared with the centre-lines of ref/ldc_Re100.mat exactly the way the notebook
overlays them (c13-c21) -- but with a number instead of a plot.

"To the reference's own error": the reference publishes no error norm, so the bar is established
here: (1) the CUDA result and the oracle (the reference's algorithm on the CPU) give the same
centre-line error against the .mat solution at an equal step count, (2) the converged CUDA run
reproduces the benchmark profiles to a few percent of the lid speed (first-order upwind FVDBM on
a 100^2 grid)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

fb = pytest.importorskip("fvdbm_jax_b200")
from fvdbm_jax_b200 import _lib, meshgen  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NX = 100
U_LID = 0.1                      # Re * Mu / N_x = 100 * 0.1 / 100 (notebook c3)


def centerline_errors(vel):
    """notebook c13-c21: normalise by U_lid, average the two middle rows/columns, flip the axis
    (the notebook's lid is the row y=0), compare with the 256-point reference lines."""
    ref = np.load(os.path.join(HERE, "golden", "ldc_re100_centerlines.npz"))
    v = np.asarray(vel, dtype=np.float64).reshape(NX, NX, 2) / U_LID
    outx = np.mean(v[NX // 2 - 1:NX // 2 + 1, :, 1], axis=0)          # v along x at mid height
    outy = np.mean(v[:, NX // 2 - 1:NX // 2 + 1, 0], axis=1)          # u along y at mid width
    xs = np.linspace(1 / (2 * NX), 1 - 1 / (2 * NX), NX)
    v_ref = np.interp(xs, ref["x"], ref["v_of_x"])
    u_ref = np.interp(np.flip(xs), ref["y"], ref["u_of_y"])
    e_v = float(np.sqrt(np.mean((-outx - v_ref) ** 2)))
    e_u = float(np.sqrt(np.mean((outy - u_ref) ** 2)))
    return e_u, e_v


def build(dtype):
    dyn = fb.D2Q13(tau=0.8, delta_t=0.1)
    cells, faces, nodes = meshgen.quad_cavity(NX, NX, dyn, U_LID)
    env = fb.Environment(cells, faces, nodes, dtype=dtype)
    env.init()
    env.set_option(_lib.OPT_GRAPH_STEPS, 50)
    return dyn, cells, faces, nodes, env


@pytest.mark.timeout(900)
def test_ldc_re100_against_benchmark_solution():
    dyn, cells, faces, nodes, env = build(np.float32)
    # (1) equal step count: CUDA vs the oracle's C port (the reference's algorithm on the host)
    from oracle.step_c import COracle
    static = {"cells.face_indices": cells.face_indices, "cells.face_normals": cells.face_normals,
              "faces.nodes_index": faces.nodes_index, "faces.stencil_cells_index": faces.stencil_cells_index,
              "faces.stencil_dists": faces.stencil_dists, "faces.n": faces.n, "faces.L": faces.L,
              "nodes.type": nodes.type, "nodes.cells_index": nodes.cells_index, "nodes.cell_dists": nodes.cell_dists}
    state = {"cells.pdf": cells.pdf, "nodes.pdf": nodes.pdf, "nodes.rho": nodes.rho, "nodes.vel": nodes.vel}
    n_cmp = 20000
    oracle = COracle(static, state, 13, dyn.tau, dyn.delta_t, "upwind", np.float32).step(n_cmp)
    env = env.step(n_cmp)
    eu_g, ev_g = centerline_errors(env.cells.vel)
    eu_o, ev_o = centerline_errors(oracle.vel)
    assert abs(eu_g - eu_o) < 1e-4 and abs(ev_g - ev_o) < 1e-4, (eu_g, eu_o, ev_g, ev_o)
    assert np.max(np.abs(env.cells.vel - oracle.vel)) < 5e-5
    # (2) the notebook's full run: 1 + 500 000 steps
    env = env.step(500001 - n_cmp)
    eu, ev = centerline_errors(env.cells.vel)
    print(f"LDC Re=100, 100x100 quads D2Q13 upwind, 500001 steps: rms error u(y) {eu:.4f}, v(x) {ev:.4f} (units of U_lid)")
    with open(os.path.join(HERE, "..", "gpurun_out", "ldc_re100.txt"), "w") if os.path.isdir(os.path.join(HERE, "..", "gpurun_out")) else open(os.devnull, "w") as f:
        f.write(f"steps 500001 rms_u_of_y {eu:.5f} rms_v_of_x {ev:.5f} | at {n_cmp} steps gpu ({eu_g:.5f},{ev_g:.5f}) oracle ({eu_o:.5f},{ev_o:.5f})\n")
    assert eu < 0.05 and ev < 0.05
    assert np.isfinite(env.cells.rho).all()
    env.close()
