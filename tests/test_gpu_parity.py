"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the public
Environment API, i.e. through the C ABI of libfvdbm_b200.so, and is compared with
  * the golden vectors minted from the reference's own code (tests/golden/*.npz), all 8 arrays,
  * the NumPy oracle on larger seeded meshes,
  * size-independent properties at the full 10M-cell benchmark size.
Tolerances (north_star): fp32 1e-5 relative, fp64 1e-11 relative (max-norm, per array)."""
import ctypes as C
import os
import pickle

import numpy as np
import pytest

import golden

pytestmark = pytest.mark.gpu

fb = pytest.importorskip("fvdbm_jax_b200")
from fvdbm_jax_b200 import _lib, meshgen  # noqa: E402

TOL = {np.float32: 1e-5, np.float64: 1e-11}
MODES = {"tma": ("fused", _lib.VARIANT_TMA), "direct": ("fused", _lib.VARIANT_DIRECT), "pair": ("fused", _lib.VARIANT_PAIR),
         "rec": ("fused", _lib.VARIANT_REC), "staged": ("staged", None)}


def make_env(case, dtype, mode, reorder="none"):
    cells, faces, nodes = case.containers()
    m, variant = MODES[mode]
    if variant == _lib.VARIANT_PAIR and np.dtype(dtype) != np.float32:
        pytest.skip("the packed pair kernel is fp32 only")
    env = fb.Environment(cells, faces, nodes, dtype=dtype, mode=m, reorder=reorder)
    env.init()
    if variant is not None:
        env.set_option(_lib.OPT_VARIANT, variant)
    return env


def check_state(env, case, step, tol):
    for name in golden.STATE:
        obj, attr = name.split(".")
        got = getattr(getattr(env, obj), attr)
        exp = case.expected(step, name)
        assert got.shape == exp.shape, (name, got.shape, exp.shape)
        err = golden.rel_err(got, exp)
        assert err < tol, f"{case.name} step {step} {name}: rel err {err:.3e} >= {tol}"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode", ["rec", "pair", "tma", "direct", "staged"])
@pytest.mark.parametrize("name", golden.names())
def test_golden_all_fields(name, mode, dtype):
    case = golden.Case(name)
    env = make_env(case, dtype, mode)
    done = 0
    for s in case.steps:
        env = env.step(s - done)
        done = s
        check_state(env, case, s, TOL[dtype])
    env.close()


@pytest.mark.parametrize("mode", ["rec", "pair", "direct", "tma", "staged"])
@pytest.mark.parametrize("name", golden.names(fp32=True))
def test_fp32_engine_vs_reference_run_in_fp32(name, mode):
    """*_f32 fixtures = the reference's own code executed with every float in fp32 (what stock JAX
    does, x64 off): the closest available stand-in for "the reference's JAX CPU path"; 1e-5 relative."""
    case = golden.Case(name)
    assert case.bits == 32
    env = make_env(case, np.float32, mode)
    done = 0
    for s in case.steps:
        env = env.step(s - done)
        done = s
        check_state(env, case, s, 1e-5)
    env.close()


def _square_problem(nx, ny, scheme="lax_wendroff", periodic=False, seed=11, perturb=True, lid=0.1):
    raw = meshgen.triangulated_square(nx, ny, seed=seed, periodic_x=periodic)
    m = fb.Mesher()
    m.import_meshpy(raw)
    m.calc_mesh_properties()
    dyn = fb.D2Q9(tau=0.8, delta_t=0.1)
    cells, faces, nodes = m.to_env(dyn, flux_method=scheme)
    for mk in ((1,) if periodic else (1, 2, 4)):
        nodes = m.set_vel_node(nodes, mk, np.array([0.0, 0.0]))
    nodes = m.set_vel_node(nodes, 3, np.array([lid, 0.0]))
    if perturb:
        c = m.cell_centers
        rho = 1 + 0.01 * np.sin(2 * np.pi * c[:, 0] / nx) * np.sin(2 * np.pi * c[:, 1] / ny)
        u = 0.05 * np.stack([np.sin(2 * np.pi * c[:, 1] / ny), np.sin(2 * np.pi * c[:, 0] / nx)], axis=1)
        cells.pdf = dyn.calc_eq(rho, u)
    return m, dyn, cells, faces, nodes


def _static_state(cells, faces, nodes):
    static = {"cells.face_indices": cells.face_indices, "cells.face_normals": cells.face_normals,
              "faces.nodes_index": faces.nodes_index, "faces.stencil_cells_index": faces.stencil_cells_index,
              "faces.stencil_dists": faces.stencil_dists, "faces.n": faces.n, "faces.L": faces.L,
              "nodes.type": nodes.type, "nodes.cells_index": nodes.cells_index, "nodes.cell_dists": nodes.cell_dists}
    state = {"cells.pdf": cells.pdf, "nodes.pdf": nodes.pdf, "nodes.rho": nodes.rho, "nodes.vel": nodes.vel}
    return static, state


@pytest.mark.parametrize("dtype,steps", [(np.float32, 100), (np.float32, 300), (np.float64, 100)])
@pytest.mark.parametrize("scheme", ["lax_wendroff", "upwind"])
def test_seeded_mesh_vs_oracle(dtype, steps, scheme):
    """20k-cell lid-driven mesh, N steps, CUDA vs the NumPy oracle in the SAME precision."""
    from oracle.step_numpy import StepOracle
    m, dyn, cells, faces, nodes = _square_problem(100, 100, scheme)
    static, state = _static_state(cells, faces, nodes)
    oracle = StepOracle(static, state, 9, dyn.tau, dyn.delta_t, scheme, dtype).step(steps)
    env = fb.Environment(cells, faces, nodes, dtype=dtype, reorder="hilbert")
    env.init()
    env = env.step(steps)
    exp = oracle.state()
    for name in golden.STATE:
        obj, attr = name.split(".")
        err = golden.rel_err(getattr(getattr(env, obj), attr), exp[name])
        assert err < TOL[dtype], f"{name}: {err:.3e}"
    env.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("name", ["ldc_tri_lw", "channel_upwind", "channel_lw", "cylinder_lw", "quad_ldc_d2q13", "tri_d2q13_lw",
                                  "pressure_lw_dm2"])
def test_gpu_bits_equal_the_cpu_walk_of_the_same_operation_sequence(name, dtype):
    """core.cuh spells out one canonical operation sequence (explicit add/mul/fma, fixed reduction order):
    the CUDA kernels (packed pair kernel in fp32, thread-per-cell in fp64) must reproduce, bit for bit, the
    g++ build of the same functions walking the same layout on the CPU (tests/hostsim)."""
    import test_hostsim
    case = golden.Case(name)
    s = case.steps[-1]
    cpu = test_hostsim.run(case, dtype, s)
    variants = (_lib.VARIANT_DIRECT, _lib.VARIANT_REC)
    if dtype is np.float32:
        variants = (_lib.VARIANT_PAIR, _lib.VARIANT_DIRECT, _lib.VARIANT_TMA, _lib.VARIANT_REC)
    for variant in variants:
        cells, faces, nodes = case.containers()
        env = fb.Environment(cells, faces, nodes, dtype=dtype, mode="fused", reorder="none")
        env.init()
        env.set_option(_lib.OPT_VARIANT, variant)
        env = env.step(s)
        for key in ("cells.pdf", "cells.rho", "cells.vel", "nodes.pdf", "nodes.rho", "nodes.vel"):
            obj, attr = key.split(".")
            np.testing.assert_array_equal(getattr(getattr(env, obj), attr), cpu[key], err_msg=f"{name} {key} variant {variant}")
        env.close()


def test_fused_variants_bitwise_identical():
    """direct / TMA (every tile size, pipeline depth, sweep direction, graph batching) and any
    renumbering run the same per-cell arithmetic -> bit-identical populations."""
    m, dyn, cells, faces, nodes = _square_problem(64, 48, periodic=True)
    ref = None
    configs = [dict(variant=_lib.VARIANT_DIRECT), dict(variant=_lib.VARIANT_PAIR), dict(variant=_lib.VARIANT_REC),
               dict(variant=_lib.VARIANT_REC, reverse=1, graph=6, reorder="hilbert", pdl=1), dict(variant=_lib.VARIANT_TMA, tile=128, stages=2),
               dict(variant=_lib.VARIANT_TMA, tile=256, stages=3), dict(variant=_lib.VARIANT_TMA, tile=512, stages=2),
               dict(variant=_lib.VARIANT_TMA, tile=256, stages=4, reverse=1, graph=4),
               dict(variant=_lib.VARIANT_DIRECT, reverse=1, graph=2, reorder="hilbert"),
               dict(variant=_lib.VARIANT_TMA, tile=128, stages=4, reorder="rcm", ctas=1)]
    configs.append(dict(variant=_lib.VARIANT_PAIR, reverse=1, graph=6, reorder="rcm"))
    # schedules: two-stream overlap (pdl=0) vs single-stream programmatic-dependent-launch chain (pdl=1), with / without graphs
    configs += [dict(variant=_lib.VARIANT_DIRECT, pdl=0), dict(variant=_lib.VARIANT_DIRECT, pdl=0, graph=8),
                dict(variant=_lib.VARIANT_DIRECT, pdl=1, graph=8), dict(variant=_lib.VARIANT_PAIR, pdl=1, graph=0),
                dict(variant=_lib.VARIANT_TMA, tile=128, stages=2, pdl=1, graph=4)]
    for cfg in configs:
        env = fb.Environment(cells, faces, nodes, dtype=np.float32, reorder=cfg.get("reorder", "none"))
        env.init()
        env.set_option(_lib.OPT_VARIANT, cfg["variant"])
        if "tile" in cfg:
            env.set_option(_lib.OPT_TILE_CELLS, cfg["tile"]).set_option(_lib.OPT_STAGES, cfg["stages"])
        env.set_option(_lib.OPT_REVERSE_SWEEP, cfg.get("reverse", 0)).set_option(_lib.OPT_GRAPH_STEPS, cfg.get("graph", 0))
        env.set_option(_lib.OPT_CTAS_PER_SM, cfg.get("ctas", 0))
        if "pdl" in cfg:
            env.set_option(_lib.OPT_PDL, cfg["pdl"])
        env = env.step(37)
        got = (env.cells.pdf.copy(), env.nodes.pdf.copy(), env.cells.rho.copy())
        if ref is None:
            ref = got
        else:
            for a, b in zip(ref, got):
                np.testing.assert_array_equal(a, b, err_msg=str(cfg))
        env.close()


def test_removed_temporal_option_is_refused():
    case = golden.Case("ldc_tri_lw")
    env = make_env(case, np.float32, "direct")
    env.build()
    env.set_option(_lib.OPT_TEMPORAL, 0)
    with pytest.raises(RuntimeError, match="temporal blocking was removed"):
        env.set_option(_lib.OPT_TEMPORAL, 1)
    env.close()


def test_rest_state_fixed_point_and_exact_conservation():
    m, dyn, cells, faces, nodes = _square_problem(40, 30, perturb=False, lid=0.0)
    env = fb.Environment(cells, faces, nodes, dtype=np.float64)
    env.init()
    w = np.array(cells.pdf)
    env = env.step(20)
    assert np.max(np.abs(env.cells.pdf - w)) < 1e-15
    env.close()
    # no boundary conditions + perturbation: interior fluxes cancel bit-for-bit, so the total mass
    # can only change through boundary faces; with a periodic-x mesh and untyped wall nodes the
    # boundary flux is what the ghost extrapolation gives -- compare against the oracle instead
    from oracle.step_numpy import StepOracle
    m, dyn, cells, faces, nodes = _square_problem(30, 20, periodic=True)
    static, state = _static_state(cells, faces, nodes)
    o = StepOracle(static, state, 9, dyn.tau, dyn.delta_t, "lax_wendroff", np.float64).step(50)
    env = fb.Environment(cells, faces, nodes, dtype=np.float64)
    env.init()
    env = env.step(50)
    assert abs(env.cells.pdf.sum() - o.pdf.sum()) < 1e-9 * o.pdf.sum()
    env.close()


def test_set_get_pickle_roundtrip(tmp_path):
    case = golden.Case("channel_lw")
    env = make_env(case, np.float64, "tma")
    env = env.step(2)
    # pickle mid-run, restore, continue: must land on the same golden state as an uninterrupted run
    blob = pickle.dumps(env)
    env2 = pickle.loads(blob)
    np.testing.assert_array_equal(env2.cells.rho, env.cells.rho)
    env2 = env2.step(3)
    check_state(env2, case, 5, 1e-11)
    # assigning populations behaves like rebinding cells.pdf in the reference
    pdf = env.cells.pdf.copy()
    env.cells.pdf = pdf * 1.0
    np.testing.assert_array_equal(env.cells.pdf, pdf)
    env = env.step(3)
    check_state(env, case, 5, 1e-11)
    with pytest.raises(AttributeError):
        env.cells.rho = env.cells.rho
    env.close(); env2.close()


def test_c_abi_error_conventions():
    lib = _lib.load()
    case = golden.Case("ldc_tri_lw")
    da = case.desc_arrays(np.float32)
    h = C.c_void_p()
    assert lib.fvdbm_create(C.byref(da.desc), C.byref(h)) == 0
    buf = np.zeros((da.N, da.Q), np.float32)
    assert lib.fvdbm_get(h, _lib.CELL_PDF, buf.ctypes.data, buf.nbytes - 4) == _lib.ERR_ARG
    assert b"size mismatch" in lib.fvdbm_last_error(h)
    assert lib.fvdbm_get(h, _lib.CELL_RHO, buf.ctypes.data, da.N * 4) == _lib.ERR_STATE      # no step yet
    assert lib.fvdbm_step(h, -1) == _lib.ERR_ARG
    assert lib.fvdbm_step(h, 3) == 0 and lib.fvdbm_sync(h) == 0
    v = C.c_int64()
    assert lib.fvdbm_info(h, _lib.INFO_STEPS, C.byref(v)) == 0 and v.value == 3
    assert lib.fvdbm_info(h, _lib.INFO_LAUNCHES, C.byref(v)) == 0 and v.value >= 6
    lib.fvdbm_destroy(h)
    da.desc.scheme = 7
    assert lib.fvdbm_create(C.byref(da.desc), C.byref(h)) == _lib.ERR_ARG
    assert b"Unknown flux scheme" in lib.fvdbm_last_error(None)
    with pytest.raises(ValueError, match="Unknown flux scheme"):                # reference containers.py:203
        cells, faces, nodes = case.containers()
        faces.flux_scheme = "weno"
        fb.Environment(cells, faces, nodes).step()


def test_set_params_changes_relaxation():
    case = golden.Case("ldc_tri_lw")
    a = make_env(case, np.float64, "tma").step(3)
    b = make_env(case, np.float64, "tma")
    b.set_params(case.tau, case.delta_t)
    b = b.step(3)
    np.testing.assert_array_equal(a.cells.pdf, b.cells.pdf)
    b.set_params(0.6, case.delta_t)
    b = b.step(1)
    a = a.step(1)
    assert np.max(np.abs(a.cells.pdf - b.cells.pdf)) > 1e-9
    a.close(); b.close()


FULL = int(os.environ.get("FVDBM_FULLSIZE_NX", "2236"))      # 2236^2 quads x 2 = 9 999 392 cells (config 4)


@pytest.mark.timeout(1200)
def test_full_size_properties():
    """BASELINE.json config 4 size: properties that need no CPU oracle run."""
    m, dyn, cells, faces, nodes = _square_problem(FULL, FULL, periodic=True, perturb=False, lid=0.0)
    n = cells.face_indices.shape[0]
    # (a) rest state with zero-velocity walls is a fixed point of the step
    env = fb.Environment(cells, faces, nodes, dtype=np.float32, reorder="hilbert")
    env.init()
    w = np.asarray(cells.pdf, dtype=np.float32)
    env = env.step(5)
    assert np.max(np.abs(env.cells.pdf - w)) < 1e-6
    # (b) perturbed state: TMA and direct kernels, forward and reverse sweeps agree bit for bit,
    #     moments lag one step, results stay finite and near equilibrium
    c = m.cell_centers
    rho0 = 1 + 0.01 * np.sin(2 * np.pi * c[:, 0] / FULL) * np.sin(2 * np.pi * c[:, 1] / FULL)
    u0 = 0.05 * np.stack([np.sin(2 * np.pi * c[:, 1] / FULL), np.sin(2 * np.pi * c[:, 0] / FULL)], axis=1)
    f0 = dyn.calc_eq(rho0, u0).astype(np.float32)
    env.cells.pdf = f0
    env = env.step(4)
    a = env.cells.pdf.copy()
    rho_lag = env.cells.rho.copy()
    prev = np.empty((n, 9), np.float32)
    env.get_into("cells.pdf", prev)                     # current
    assert env.info(_lib.INFO_VARIANT) == _lib.VARIANT_REC      # default kernel (fp32 D2Q9: records at every size) produced `a`
    for variant, reverse in ((_lib.VARIANT_TMA, 0), (_lib.VARIANT_DIRECT, 1), (_lib.VARIANT_PAIR, 1), (_lib.VARIANT_REC, 0)):
        env.set_option(_lib.OPT_VARIANT, variant).set_option(_lib.OPT_REVERSE_SWEEP, reverse)
        assert env.info(_lib.INFO_VARIANT) == variant
        env.cells.pdf = f0
        env = env.step(4)
        np.testing.assert_array_equal(env.cells.pdf, a)
        np.testing.assert_array_equal(env.cells.rho, rho_lag)
    assert np.isfinite(a).all()
    assert abs(float(a.sum(dtype=np.float64)) / float(f0.sum(dtype=np.float64)) - 1) < 1e-4
    env.close()
