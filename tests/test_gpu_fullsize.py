"""Value parity AT THE BASELINE.json CONFIG SIZES (run on the B200 box: pytest -m gpu).

The golden fixtures pin the oracle on small meshes; here the engine (through Environment, i.e. the
C ABI) is compared with the C/OpenMP oracle (oracle/step_c.c, fp32 = the precision of the reference's
JAX path) on the full-size problems themselves, all 8 state arrays, 1e-5 relative (north_star):
  configs[3]  synthetic triangulated square, 9 999 392 cells, x-periodic + walls      10 and 100 steps
  configs[1]  flow over a cylinder, ~194 k cells, velocity inlet / density outlet     100 and 300 steps
  configs[2]  porous obstacle field from the reference's tests/test_bmp.mat, ~2 M     100 steps
"""
import os

import numpy as np
import pytest

import golden

pytestmark = pytest.mark.gpu

fb = pytest.importorskip("fvdbm_jax_b200")
from fvdbm_jax_b200 import _lib, meshgen  # noqa: E402

TOL = 1e-5
FULL = int(os.environ.get("FVDBM_FULLSIZE_NX", "2236"))      # 2236^2 quads x 2 = 9 999 392 cells


def _static_state(cells, faces, nodes):
    static = {"cells.face_indices": cells.face_indices, "cells.face_normals": cells.face_normals,
              "faces.nodes_index": faces.nodes_index, "faces.stencil_cells_index": faces.stencil_cells_index,
              "faces.stencil_dists": faces.stencil_dists, "faces.n": faces.n, "faces.L": faces.L,
              "nodes.type": nodes.type, "nodes.cells_index": nodes.cells_index, "nodes.cell_dists": nodes.cell_dists}
    state = {"cells.pdf": cells.pdf, "nodes.pdf": nodes.pdf, "nodes.rho": nodes.rho, "nodes.vel": nodes.vel}
    return static, state


def _compare(env, oracle, label):
    exp = oracle.state()
    worst = {}
    for name in golden.STATE:
        obj, attr = name.split(".")
        got = getattr(getattr(env, obj), attr)
        assert got.shape == exp[name].shape, (label, name, got.shape, exp[name].shape)
        worst[name] = golden.rel_err(got, exp[name])
    bad = {k: v for k, v in worst.items() if not v < TOL}
    assert not bad, f"{label}: rel err >= {TOL}: {bad} (all: {worst})"
    return worst


def _run_against_oracle(mesher, dyn, cells, faces, nodes, scheme, checkpoints, label, variants=(None,)):
    from oracle.step_c import COracle, use_all_cores
    use_all_cores()
    static, state = _static_state(cells, faces, nodes)
    oracle = COracle(static, state, 9, dyn.tau, dyn.delta_t, scheme, np.float32)
    envs = []
    for v in variants:
        env = fb.Environment(cells, faces, nodes, dtype=np.float32, reorder="hilbert")
        env.init()
        if v is not None:
            env.set_option(_lib.OPT_VARIANT, v)
        envs.append(env)
    done = 0
    report = {}
    for s in checkpoints:
        oracle.step(s - done)
        for v, env in zip(variants, envs):
            env.step(s - done)
            report[(s, v)] = _compare(env, oracle, f"{label} step {s} variant {v}")
        done = s
    if len(envs) > 1:                                    # every fused variant runs the same arithmetic
        a = envs[0].cells.pdf
        for env in envs[1:]:
            np.testing.assert_array_equal(env.cells.pdf, a)
    for env in envs:
        env.close()
    print(label, {f"s{k[0]}v{k[1]}": f"{max(v.values()):.1e}" for k, v in report.items()})


@pytest.mark.timeout(1800)
def test_config3_square_10M_vs_oracle():
    """BASELINE.json configs[3] -- the bench.py workload itself (same builder)."""
    import bench
    m, dyn, cells, faces, nodes, _ = bench.build_problem(FULL, FULL, "lax_wendroff")
    assert cells.face_indices.shape[0] == 2 * FULL * FULL
    _run_against_oracle(m, dyn, cells, faces, nodes, "lax_wendroff", (10, 100), f"square nx={FULL}",
                        variants=(_lib.VARIANT_REC, _lib.VARIANT_PAIR, _lib.VARIANT_DIRECT, _lib.VARIANT_TMA))


def _cylinder(scale):
    raw = meshgen.cylinder_channel(scale=scale)
    m = fb.Mesher()
    m.import_meshpy(raw)
    m.calc_mesh_properties()
    dyn = fb.D2Q9(tau=0.65, delta_t=0.1)
    cells, faces, nodes = m.to_env(dyn, flux_method="lax_wendroff")
    # tests/flow_over_cyl.ipynb c12: inlet velocity on the left, walls + cylinder no-slip, density outlet
    nodes = m.set_vel_node(nodes, meshgen.LEFT, np.array([0.1, 0.0]))
    for mk in (meshgen.TOP, meshgen.BOTTOM, meshgen.OBSTACLE):
        nodes = m.set_vel_node(nodes, mk, np.array([0.0, 0.0]))
    nodes = m.set_rho_node(nodes, meshgen.RIGHT, 0.95)
    return m, dyn, cells, faces, nodes


@pytest.mark.timeout(900)
def test_config1_cylinder_194k_vs_oracle():
    m, dyn, cells, faces, nodes = _cylinder(int(os.environ.get("FVDBM_CYL_SCALE", "9")))
    n = cells.face_indices.shape[0]
    assert os.environ.get("FVDBM_CYL_SCALE") or 190_000 < n < 200_000
    _run_against_oracle(m, dyn, cells, faces, nodes, "lax_wendroff", (100, 300), f"cylinder {n} cells")


@pytest.mark.timeout(1800)
def test_config2_porous_2M_vs_oracle():
    """Obstacle field built from the reference's own outlines (fvdbm_jax_b200/data/porous_outlines.npz,
    extracted from tests/test_bmp.mat), BCs of tests/porous_flow.ipynb c25-c26."""
    raw = meshgen.porous_channel(scale=float(os.environ.get("FVDBM_POROUS_SCALE", "8.5")))
    m = fb.Mesher()
    m.import_meshpy(raw)
    m.calc_mesh_properties()
    dyn = fb.D2Q9(tau=0.65, delta_t=0.1)
    cells, faces, nodes = m.to_env(dyn, flux_method="lax_wendroff")
    nodes = meshgen.porous_boundary_conditions(m, nodes)
    n = cells.face_indices.shape[0]
    assert os.environ.get("FVDBM_POROUS_SCALE") or 1_900_000 < n < 2_100_000
    assert int((np.asarray(nodes.type) == 2).sum()) > 0 and int((raw.point_markers == meshgen.OBSTACLE).sum()) > 1000
    _run_against_oracle(m, dyn, cells, faces, nodes, "lax_wendroff", (100,), f"porous {n} cells")
