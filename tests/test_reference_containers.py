"""Drop-in check with the REFERENCE's own objects (build container only, needs /root/reference):
containers built by the reference Mesher / Environment.create under the jax shim are accepted by
fvdbm_jax_b200.Environment and describe the same problem as our own containers."""
import numpy as np
import pytest

import golden
from oracle import refrun

pytestmark = pytest.mark.skipif(not refrun.available(), reason="reference tree not present on this box")


def test_reference_mesher_route_is_accepted():
    import fvdbm_jax_b200 as fb
    ns = refrun.load()
    case = golden.Case("channel_lw")
    m = refrun.ref_mesher(case.raw())
    dyn = ns.D2Q9(tau=case.tau, delta_t=case.delta_t)
    cells, faces, nodes = m.to_env(dyn, flux_method="lax_wendroff")
    nodes = m.set_vel_node(nodes, 4, ns.jnp.array([0.05, 0.0]))
    nodes = m.set_vel_node(nodes, 1, ns.jnp.array([0.0, 0.0]))
    nodes = m.set_vel_node(nodes, 3, ns.jnp.array([0.0, 0.0]))
    nodes = m.set_rho_node(nodes, 2, 0.95)
    env = fb.Environment(cells, faces, nodes, dtype=np.float64)       # reference objects, jax(-shim) arrays
    env.init()
    da = env._describe()
    ours = fb.Environment(*case.containers(), dtype=np.float64)._describe()
    for key in ("cell_face_idx", "cell_face_sign", "face_cell_idx", "face_dists", "face_node_idx", "face_n", "face_L",
                "node_type", "node_cell_idx", "node_cell_dist", "node_rho", "node_vel"):
        assert np.array_equal(da.keep[key], ours.keep[key]), key
    assert da.desc.Q == 9 and da.desc.K == 3 and da.desc.scheme == 1 and abs(da.desc.tau - case.tau) < 1e-15


def test_reference_create_route_with_custom_arrays_is_accepted():
    import fvdbm_jax_b200 as fb
    ns = refrun.load()
    ns.Environment.dynamics = ns.D2Q13(tau=0.8, delta_t=0.1)
    ref_env = ns.Environment.create(2, 7, 6)                         # CustomArray-backed containers
    ref_env.cells.face_indices.add_items(0, ns.jnp.asarray([0, 1, 2, 3]))
    ref_env.cells.face_indices.add_items(1, ns.jnp.asarray([2, 4, 5, 6]))
    ref_env.cells.face_normals.add_items(0, ns.jnp.asarray([0, 0, 1, 1]))   # notebook c7: 0 first, then -> -1
    ref_env.cells.face_normals.add_items(1, ns.jnp.asarray([0, 0, 1, 1]))   # (-1 is CustomArray's "empty" marker)
    ref_env.cells.face_normals.data = ns.jnp.where(ref_env.cells.face_normals.data == 0, -1, ref_env.cells.face_normals.data)
    for j, st in enumerate([(-1, 0), (-1, 0), (0, 1), (0, -1), (-1, 1), (1, -1), (1, -1)]):
        ref_env.faces.stencil_cells_index.add_items(j, ns.jnp.asarray([s if s >= 0 else -2 for s in st]))
        ref_env.faces.stencil_dists.add_items(j, ns.jnp.asarray([.5, .5]))
        ref_env.faces.nodes_index.add_items(j, ns.jnp.asarray([j % 6, (j + 1) % 6]))
    ref_env.faces.stencil_cells_index.data = ns.jnp.where(ref_env.faces.stencil_cells_index.data == -2, -1,
                                                          ref_env.faces.stencil_cells_index.data)
    for p in range(6):
        ref_env.nodes.cells_index.add_items(p, ns.jnp.asarray([p % 2]))
        ref_env.nodes.cell_dists.add_items(p, ns.jnp.asarray([1.0]))
    env = fb.Environment.define(ref_env.cells, ref_env.faces, ref_env.nodes)
    env.init()                                                       # calls the reference containers' init()
    da = env._describe()
    assert da.desc.Q == 13 and da.desc.K == 4 and da.desc.N == 2 and da.desc.F == 7 and da.desc.P == 6
    assert da.keep["face_cell_idx"].tolist()[3] == [0, -1] and da.keep["node_cell_idx"].shape == (6, 1)
    hp = fb._lib.HostPlan(da)
    assert hp.scalar("fused_ok") == 1 and hp.scalar("NB") == 6


def test_stencil_geometry_report_equals_the_reference_printout(capsys):
    """Mesher.verify_stencil_geometry (reference mesher.py:386-504): same numbers in the report, line by line (the
    lines carry 2-4 significant digits; lines whose value is rounding noise around zero are compared as numbers)."""
    import re
    import fvdbm_jax_b200 as fb
    case = golden.Case("cylinder_lw")
    ref = refrun.ref_mesher(case.raw())
    capsys.readouterr()
    ref.verify_stencil_geometry()
    want = capsys.readouterr().out.strip().splitlines()
    m = fb.Mesher()
    m.import_meshpy(case.raw())
    m.calc_mesh_properties()
    stats = m.verify_stencil_geometry()
    got = capsys.readouterr().out.strip().splitlines()
    assert len(want) == len(got) >= 14 and stats["interior_faces"] > 0
    num = re.compile(r"[-+]?\d+\.?\d*(?:e[-+]?\d+)?")
    for a, b in zip(want, got):
        assert num.sub("#", a) == num.sub("#", b)                   # same text
        for x, y in zip(num.findall(a), num.findall(b)):
            assert abs(float(x) - float(y)) <= 1e-9 + 2e-3 * abs(float(x)), (a, b)


def test_utils_helpers_equal_the_reference_functions():
    """fvdbm_jax_b200.utils against /root/reference/utils/utils.py executed under the shim."""
    from fvdbm_jax_b200 import utils as ours
    ref = refrun.load().utils
    rng = np.random.default_rng(7)
    x, w = rng.random((6, 9)), rng.random(6) + 0.1
    d = np.array([0.5, 1.5, -1.0, 2.0, -1.0, 0.7])
    assert np.allclose(ours.weighted_avg(x, w), np.asarray(ref.weighted_avg(x, w)), rtol=1e-15, atol=0)
    assert np.allclose(ours.extrapolate(x, d), np.asarray(ref.extrapolate(x, d)), rtol=1e-15, atol=0)
    assert np.allclose(ours.interp_pdf(x, np.abs(d)), np.asarray(ref.interp_pdf(x, np.abs(d))), rtol=1e-15, atol=0)
    assert np.array_equal(ours.extrap_pdf(x[0], x[1], 0.3, 0.9), np.asarray(ref.extrap_pdf(x[0], x[1], 0.3, 0.9)))
    p1, p2 = np.array([0.2, -1.0]), np.array([1.7, 0.4])
    assert np.allclose(ours.calc_normal(p1, p2), ref.calc_normal(p1, p2), rtol=1e-15) and abs(ours.calc_dist(p1, p2) - ref.calc_dist(p1, p2)) < 1e-15
    ragged = [np.array([1, 2, 3]), np.array([4]), np.array([5, 6])]
    assert np.array_equal(ours.pad_stack(ragged), np.asarray(ref.pad_stack([refrun.load().jnp.asarray(a) for a in ragged])))
    a, b = ours.CustomArray(3, dtype=np.int32), ref.CustomArray(3, dtype=np.int32)
    for arr in (a, b):
        arr.add_items(1, np.asarray([7, 8, 9]))
        arr.add_item(0, 4)
    assert np.array_equal(np.asarray(a), np.asarray(b))


VERBATIM_NOTEBOOK = r"""
ROOT, REFERENCE = %r, %r
import sys, os, json, types
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "jaxshim"))      # test-only NumPy stand-in for `jax` (not installed here)
for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules[name] = types.ModuleType(name)
import fvdbm_jax_b200.compat as compat
compat.install()
nb = json.load(open(os.path.join(REFERENCE, "tests", "ldcFVDBM.ipynb")))
cells = ["".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code"]
src = "\n".join(cells[:9]).replace("N_x = 100", "N_x = 6")
g = {"__name__": "__main__"}
exec(compile(src, "ldcFVDBM.ipynb", "exec"), g)
env = g["env"]
import numpy as np, fvdbm_jax_b200 as fb
from fvdbm_jax_b200 import meshgen
assert type(env) is fb.Environment
c, f, n = meshgen.quad_cavity(6, 6, fb.D2Q13(tau=g["Tau"], delta_t=g["dt"]), g["U_lid"])
for mine, ref, key in ((env.cells.face_indices, c.face_indices, "face_indices"), (env.cells.face_normals, c.face_normals, "face_normals"),
                       (env.faces.nodes_index, f.nodes_index, "nodes_index"), (env.faces.stencil_cells_index, f.stencil_cells_index, "stencil"),
                       (env.faces.stencil_dists, f.stencil_dists, "dists"), (env.faces.n, f.n, "n"), (env.faces.L, f.L, "L"),
                       (env.nodes.type, n.type, "type"), (env.nodes.cells_index, n.cells_index, "ring"), (env.nodes.cell_dists, n.cell_dists, "ring dists"),
                       (env.nodes.vel, n.vel, "vel")):
    a, b = np.asarray(mine, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert a.shape == b.shape and np.allclose(a, b, rtol=1e-7, atol=0), key
da = env._describe()
assert da.desc.Q == 13 and da.desc.K == 4 and da.desc.N == 36
print("notebook ok")
"""


def test_ldc_notebook_cells_run_verbatim_on_this_framework():
    """The code cells c0-c8 of the reference's tests/ldcFVDBM.ipynb (imports, parameters, Environment.create, the
    CustomArray.add_items / ``.at[...].set`` mesh construction, env.init()) are read from the reference tree and executed
    UNCHANGED (only N_x = 100 -> 6) after ``fvdbm_jax_b200.compat.install()``; the environment they build is this
    framework's and describes exactly the problem meshgen.quad_cavity builds (the mesh of the GPU LDC Re=100 test).
    ``jax`` is the oracle's NumPy shim (JAX is not installed here), matplotlib a stub."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", VERBATIM_NOTEBOOK % (root, refrun.REFERENCE)], capture_output=True, text=True,
                       cwd=os.path.join(root, "tests"), timeout=600)
    assert r.returncode == 0 and "notebook ok" in r.stdout, r.stderr[-3000:]


POST_MESH_NOTEBOOK = r"""
import sys, os, json, types
ROOT, REFERENCE, NOTEBOOK, FIRST, LAST, MESH = %r, %r, %r, %d, %d, %r
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "jaxshim"))      # test-only NumPy stand-in for `jax` (not installed here)
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.path", "matplotlib.tri", "meshpy", "meshpy.triangle",
             "meshpy.geometry", "trimesh", "shapely", "shapely.geometry", "shapely.ops"):      # plotting / mesh generation: absent
    sys.modules[name] = types.ModuleType(name)
sys.modules["meshpy"].triangle, sys.modules["meshpy"].geometry = sys.modules["meshpy.triangle"], sys.modules["meshpy.geometry"]
sys.modules["matplotlib.path"].Path = object
import fvdbm_jax_b200.compat as compat
compat.install()
nb = json.load(open(os.path.join(REFERENCE, "tests", NOTEBOOK)))
cells = ["".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code"]
g = {"__name__": "__main__"}
exec(compile(cells[0], NOTEBOOK + " c0", "exec"), g)                               # the import cell, unchanged
from fvdbm_jax_b200 import meshgen
g["mesh"] = getattr(meshgen, MESH)(scale=1)          # stands in for the meshpy.triangle / MeshRefiner cells
os.chdir(os.environ["NB_TMP"])
exec(compile("\n".join(cells[FIRST:LAST + 1]), NOTEBOOK, "exec"), g)              # Mesher ... Environment(...).init(), unchanged
import numpy as np, fvdbm_jax_b200 as fb
env, mesher = g["env"], g["mesher"]
assert type(env) is fb.Environment and type(mesher) is fb.Mesher
da = env._describe()
assert da.desc.Q == 9 and da.desc.K == 3 and da.desc.scheme == 1 and abs(da.desc.tau - 0.65) < 1e-12
t, mk = np.asarray(env.nodes.type).ravel(), np.asarray(g["mesh"].point_markers)
if MESH == "cylinder_channel":       # c11: inlet 4 velocity, walls 1/3 and cylinder 5 no-slip, outlet 2 density
    assert (t[mk == 2] == 2).all() and (t[np.isin(mk, (1, 3, 4, 5))] == 1).all() and (t[mk == 0] == 0).all()
    assert np.allclose(np.asarray(env.nodes.vel)[mk == 4], [0.1, 0.0]) and np.allclose(np.asarray(env.nodes.rho)[mk == 2], 0.95)
else:                                # c18: obstacles 5 and sides 2/4 no-slip, density 1.05 on marker 1, 0.95 on marker 3
    assert (t[np.isin(mk, (2, 4, 5))] == 1).all() and (t[np.isin(mk, (1, 3))] == 2).all()
    assert np.allclose(np.asarray(env.nodes.rho)[mk == 1], 1.05) and np.allclose(np.asarray(env.nodes.rho)[mk == 3], 0.95)
    env2, mesher2 = g["Mesher"].from_pickle("porous_temp")                    # c20 wrote it, c26 reads it back
    assert type(env2) is fb.Environment and np.array_equal(np.asarray(env2.nodes.type), np.asarray(env.nodes.type))
    assert np.array_equal(mesher2.cell_face_indices, mesher.cell_face_indices)
try:
    g["MeshRefiner"](g["mesh"], [(10, 10)])
    raise SystemExit("MeshRefiner should refuse")
except NotImplementedError:
    pass
print("notebook ok")
"""


@pytest.mark.parametrize("notebook,first,last,mesh", [("flow_over_cyl.ipynb", 9, 12, "cylinder_channel"),
                                                      ("porous_flow.ipynb", 16, 20, "porous_channel")])
def test_post_mesh_notebook_cells_run_verbatim_on_this_framework(notebook, first, last, mesh, tmp_path):
    """The import cell and the cells between mesh generation and the time loop of the two mesh-based notebooks (Mesher,
    verify_stencil_geometry, parameters, to_env, boundary conditions, Environment(...).init(), to_pickle) are read from the
    reference tree and executed UNCHANGED after ``fvdbm_jax_b200.compat.install()``; only ``mesh`` comes from this
    framework's synthetic generator instead of meshpy.triangle (absent).  They must build this framework's Environment
    with the notebook's boundary conditions."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = POST_MESH_NOTEBOOK % (root, refrun.REFERENCE, notebook, first, last, mesh)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, NB_TMP=str(tmp_path)))
    assert r.returncode == 0 and "notebook ok" in r.stdout, r.stderr[-3000:]

