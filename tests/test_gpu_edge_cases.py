"""Edge cases of the hot path on the GPU, each against the NumPy oracle."""
import types

import numpy as np
import pytest

import golden

pytestmark = pytest.mark.gpu

fb = pytest.importorskip("fvdbm_jax_b200")
from fvdbm_jax_b200 import _lib, meshgen  # noqa: E402
from oracle.step_numpy import StepOracle  # noqa: E402


def static_state(cells, faces, nodes):
    static = {"cells.face_indices": cells.face_indices, "cells.face_normals": cells.face_normals,
              "faces.nodes_index": faces.nodes_index, "faces.stencil_cells_index": faces.stencil_cells_index,
              "faces.stencil_dists": faces.stencil_dists, "faces.n": faces.n, "faces.L": faces.L,
              "nodes.type": nodes.type, "nodes.cells_index": nodes.cells_index, "nodes.cell_dists": nodes.cell_dists}
    state = {"cells.pdf": cells.pdf, "nodes.pdf": nodes.pdf, "nodes.rho": nodes.rho, "nodes.vel": nodes.vel}
    return static, state


def compare(env, oracle, tol, names=golden.STATE):
    exp = oracle.state()
    for name in names:
        obj, attr = name.split(".")
        err = golden.rel_err(getattr(getattr(env, obj), attr), exp[name])
        assert err < tol, f"{name}: {err:.3e}"


def mesh_problem(raw, scheme="lax_wendroff", bcs=(("vel", 1, (0.0, 0.0)), ("vel", 3, (0.1, 0.0)))):
    m = fb.Mesher()
    m.import_meshpy(raw)
    m.calc_mesh_properties()
    dyn = fb.D2Q9(0.8, 0.1)
    cells, faces, nodes = m.to_env(dyn, scheme)
    for kind, mk, val in bcs:
        nodes = m.set_vel_node(nodes, mk, np.array(val)) if kind == "vel" else m.set_rho_node(nodes, mk, val)
    return m, dyn, cells, faces, nodes


def test_fortran_ordered_and_broadcast_inputs():
    """Initial state handed over with non-C strides (np.array / astype keep them by default): the engine and the
    lagged / untracked-node getters must still see the reference's row-major (elements, components) arrays."""
    m, dyn, cells, faces, nodes = mesh_problem(meshgen.triangulated_square(7, 5, seed=4))
    static, state = static_state(cells, faces, nodes)
    o = StepOracle(static, state, 9, dyn.tau, dyn.delta_t, "lax_wendroff", np.float64).step(5)
    cells.pdf = np.asfortranarray(cells.pdf)
    nodes.pdf = np.broadcast_to(np.asarray(nodes.pdf)[0], nodes.pdf.shape)
    nodes.vel = np.asfortranarray(nodes.vel)
    nodes.rho = np.asfortranarray(nodes.rho)
    env = fb.Environment(cells, faces, nodes, dtype=np.float64)
    env.init()
    compare(env.step(5), o, 1e-11)


@pytest.mark.parametrize("nx,ny", [(1, 1), (2, 1), (3, 2)])
@pytest.mark.parametrize("variant", [_lib.VARIANT_DIRECT, _lib.VARIANT_TMA])
def test_tiny_meshes(nx, ny, variant):
    """2-12 cells: far fewer cells than one 32-lane tile, every cell on the boundary."""
    m, dyn, cells, faces, nodes = mesh_problem(meshgen.triangulated_square(nx, ny, jitter=0.0),
                                               bcs=(("vel", 1, (0, 0)), ("vel", 2, (0, 0)), ("vel", 4, (0, 0)), ("vel", 3, (0.1, 0))))
    static, state = static_state(cells, faces, nodes)
    o = StepOracle(static, state, 9, dyn.tau, dyn.delta_t, "lax_wendroff", np.float64).step(7)
    env = fb.Environment(cells, faces, nodes, dtype=np.float64)
    env.init()
    env.set_option(_lib.OPT_VARIANT, variant)
    compare(env.step(7), o, 1e-11)
    env.close()


def fan_mesh(n_tri=40):
    """Half-disc fan: boundary node 0 is shared by n_tri (> 32) triangles -> ring wider than a warp."""
    ang = np.linspace(0.0, np.pi, n_tri + 1)
    pts = np.concatenate([[[0.0, 0.0]], np.stack([np.cos(ang), np.sin(ang)], axis=1)])
    el = np.stack([np.zeros(n_tri, int), 1 + np.arange(n_tri), 2 + np.arange(n_tri)], axis=1).astype(np.int32)
    markers = np.full(pts.shape[0], 2, dtype=np.int32)       # arc nodes
    markers[[0, 1, n_tri + 1]] = 1                            # the flat side incl. the hub
    return types.SimpleNamespace(points=pts, elements=el, faces=meshgen.unique_edges(el, pts.shape[0]), point_markers=markers)


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-11), (np.float32, 1e-5)])
def test_ring_wider_than_a_warp(dtype, tol):
    m, dyn, cells, faces, nodes = mesh_problem(fan_mesh(40), bcs=(("vel", 1, (0.05, 0.0)), ("rho", 2, 0.98)))
    assert nodes.cells_index.shape[1] == 40
    static, state = static_state(cells, faces, nodes)
    o = StepOracle(static, state, 9, dyn.tau, dyn.delta_t, "lax_wendroff", dtype).step(15)
    for mode in ("fused", "staged"):
        env = fb.Environment(cells, faces, nodes, dtype=dtype, mode=mode)
        env.init()
        compare(env.step(15), o, tol)
        env.close()


def test_general_signs_use_the_staged_kernels():
    """cells.face_normals need not be +-1 in the reference (any int multiplies the flux,
    src/containers.py:119): such meshes are routed to the reference-shaped staged kernels."""
    case = golden.Case("channel_lw")
    cells, faces, nodes = case.containers()
    sg = np.array(cells.face_normals, copy=True)
    sg[::7, 1] *= 2
    cells.face_normals = sg
    static = dict(case.static)
    static["cells.face_normals"] = sg
    o = StepOracle(static, case.init, case.Q, case.tau, case.delta_t, case.scheme, np.float64).step(6)
    env = fb.Environment(cells, faces, nodes, dtype=np.float64)
    env.init()
    env = env.step(6)
    assert env.info(_lib.INFO_MODE) == _lib.MODE_STAGED and env.info(_lib.INFO_FUSED_OK) == 0
    compare(env, o, 1e-11)
    with pytest.raises(ValueError, match="fused mode unavailable"):
        fb.Environment(cells, faces, nodes, mode="fused").step()
    env.close()


def test_boundary_values_can_change_mid_run():
    """Re-binding nodes.vel / nodes.rho between steps (what set_vel_node does before a run)."""
    case = golden.Case("channel_lw")
    cells, faces, nodes = case.containers()
    env = fb.Environment(cells, faces, nodes, dtype=np.float64)
    env.init()
    env = env.step(3)
    o = case.oracle(np.float64).step(3)
    vel = np.array(env.nodes.vel)
    inlet = np.nonzero(np.asarray(nodes.type).reshape(-1) == 1)[0]
    vel[inlet, 0] *= 0.5
    env.nodes.vel = vel
    o.nvel[inlet, 0] *= 0.5
    compare(env.step(4), o.step(4), 1e-11)
    env.close()


def test_upwind_axis_aligned_faces_where_ksi_dot_n_is_zero():
    """Unjittered mesh: many faces have KSI.n == 0 exactly -> the `>= 0` upwind tie-break matters."""
    m, dyn, cells, faces, nodes = mesh_problem(meshgen.triangulated_square(12, 9, jitter=0.0), scheme="upwind",
                                               bcs=(("vel", 1, (0, 0)), ("vel", 2, (0, 0)), ("vel", 4, (0, 0)), ("vel", 3, (0.1, 0))))
    c = m.cell_centers
    cells.pdf = dyn.calc_eq(1 + 0.02 * np.sin(c[:, 0]), 0.03 * np.stack([np.cos(c[:, 1]), np.sin(c[:, 0])], axis=1))
    static, state = static_state(cells, faces, nodes)
    o = StepOracle(static, state, 9, dyn.tau, dyn.delta_t, "upwind", np.float64).step(10)
    env = fb.Environment(cells, faces, nodes, dtype=np.float64)
    env.init()
    compare(env.step(10), o, 1e-11)
    env.close()


def test_divergence_is_detected():
    """tau < 0.5 is unstable: the run blows up and count_nonfinite() reports it."""
    case = golden.Case("ldc_tri_lw")
    env = fb.Environment(*case.containers(), dtype=np.float32)
    env.init()
    assert env.step(5).count_nonfinite() == 0
    env.set_params(0.05, 1.0)
    env = env.step(400)
    bad = env.count_nonfinite()
    assert bad > 0 and bad == int(np.count_nonzero(~np.isfinite(env.cells.pdf).all(axis=1)))
    env.close()


def test_literal_notebook_loop_is_deferred_and_equal_to_step_n():
    """`for i in range(n): env = env.step()` (tests/flow_over_cyl.ipynb c17) is batched behind the scenes and
    must give the bits of step(n); anything that observes the state flushes first (lag semantics intact)."""
    case = golden.Case("channel_lw")
    a = fb.Environment(*case.containers(), dtype=np.float32)
    a.init()
    a = a.step(137)
    b = fb.Environment(*case.containers(), dtype=np.float32)
    b.init()
    for i in range(137):
        b = b.step()
        if i == 4:                                   # observed mid-loop: golden state after 5 steps, lagged moments
            for name in golden.STATE:
                obj, attr = name.split(".")
                assert golden.rel_err(getattr(getattr(b, obj), attr), case.expected(5, name)) < 1e-5, name
    assert b._pending > 0                            # single steps are really being deferred ...
    launches_before = b._pending
    np.testing.assert_array_equal(b.cells.pdf, a.cells.pdf)      # ... and flushed by the read
    assert b._pending == 0 and launches_before < b._batch
    np.testing.assert_array_equal(b.cells.rho, a.cells.rho)
    np.testing.assert_array_equal(b.nodes.pdf, a.nodes.pdf)
    assert b.info(_lib.INFO_STEPS) == 137
    a.close(); b.close()
    with pytest.raises(RuntimeError, match="closed"):
        b.step()


def test_async_transfers_match_blocking_ones():
    import torch
    case = golden.Case("cylinder_lw")
    env = fb.Environment(*case.containers(), dtype=np.float32)
    env.init()
    env = env.step(3)
    n = case.static["cells.face_indices"].shape[0]
    pdf = torch.empty((n, 9), dtype=torch.float32).pin_memory()
    env.get_into("cells.pdf", pdf.numpy())
    ref = fb.Environment(*case.containers(), dtype=np.float32)
    ref.init()
    ref.cells.pdf = pdf.numpy().copy()
    ref = ref.step(7)
    rho = [torch.empty((n, 1), dtype=torch.float32).pin_memory() for _ in range(2)]
    vel = [torch.empty((n, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
    tickets = []
    for k in range(4):                               # pipelined: upload k+1 / download k-1 overlap step k
        env.set_cells_pdf(pdf.numpy(), wait=False)
        env.step(7)
        env.get_into("cells.rho", rho[k & 1].numpy(), wait=False)
        tickets.append(env.get_into("cells.vel", vel[k & 1].numpy(), wait=False))
        if k:
            env.wait(tickets[k - 1])
            np.testing.assert_array_equal(rho[(k - 1) & 1].numpy(), ref.cells.rho)
            np.testing.assert_array_equal(vel[(k - 1) & 1].numpy(), ref.cells.vel)
    env.wait()
    np.testing.assert_array_equal(vel[1].numpy(), ref.cells.vel)
    np.testing.assert_array_equal(env.cells.pdf, ref.cells.pdf)
    with pytest.raises(ValueError, match="ticket"):
        env.wait(10 ** 6)
    env.close(); ref.close()


@pytest.mark.parametrize("mode", ["fused", "staged"])
@pytest.mark.parametrize("scheme", ["lax_wendroff", "upwind"])
def test_optional_inverse_area_mode_matches_the_numpy_oracle(scheme, mode):
    """fvdbm_desc.cell_inv_area: the flux divergence of every cell is scaled by 1/area.  NULL is
    the reference's behaviour (every other test); here the engine is compared with the NumPy oracle's switch."""
    raw = meshgen.triangulated_square(18, 12, seed=5, lx=9.0, ly=7.0)          # non-unit cells
    m = fb.Mesher()
    m.import_meshpy(raw)
    m.calc_mesh_properties()
    dyn = fb.D2Q9(tau=0.8, delta_t=0.02)
    cells, faces, nodes = m.to_env(dyn, flux_method=scheme)
    for mk in (1, 2, 4):
        nodes = m.set_vel_node(nodes, mk, np.array([0.0, 0.0]))
    nodes = m.set_vel_node(nodes, 3, np.array([0.05, 0.0]))
    cells.inv_area = m.cell_inv_areas()
    assert cells.inv_area.min() > 2.0                                          # areas ~0.15: the factor matters
    static = {"cells.face_indices": cells.face_indices, "cells.face_normals": cells.face_normals,
              "faces.nodes_index": faces.nodes_index, "faces.stencil_cells_index": faces.stencil_cells_index,
              "faces.stencil_dists": faces.stencil_dists, "faces.n": faces.n, "faces.L": faces.L,
              "nodes.type": nodes.type, "nodes.cells_index": nodes.cells_index, "nodes.cell_dists": nodes.cell_dists,
              "cells.inv_area": cells.inv_area}
    state = {"cells.pdf": cells.pdf, "nodes.pdf": nodes.pdf, "nodes.rho": nodes.rho, "nodes.vel": nodes.vel}
    o = StepOracle(static, state, 9, dyn.tau, dyn.delta_t, scheme, np.float64).step(40)
    plain = StepOracle({k: v for k, v in static.items() if k != "cells.inv_area"}, state, 9, dyn.tau, dyn.delta_t, scheme,
                       np.float64).step(40)
    assert np.max(np.abs(o.pdf - plain.pdf)) > 1e-4                            # the switch changes the physics
    env = fb.Environment(cells, faces, nodes, dtype=np.float64, mode=mode)
    env.init()
    env = env.step(40)
    for name, ref in (("pdf", o.pdf), ("rho", o.rho), ("vel", o.vel)):
        assert golden.rel_err(getattr(env.cells, name), ref) < 1e-10, name
    env.close()
