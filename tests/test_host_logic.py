"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the host planner's layout (fvdbm_plan_*), Mesher parity against the reference Mesher's
arrays (recorded in the golden fixtures), the drop-in surface, and export."""
import ctypes as C
import os
import re
import types

import numpy as np
import pytest

import golden
import fvdbm_jax_b200 as fb
from fvdbm_jax_b200 import _lib, meshgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "fvdbm_b200.h")).read()
    declared = set(re.findall(r"\b(fvdbm_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/fvdbm_b200.h but not exported"
    assert declared == set(_lib.EXPORTS)
    assert _lib.load().fvdbm_abi_version() == _lib.ABI_VERSION == 2


def test_desc_struct_matches_header_layout():
    """ctypes mirror of fvdbm_desc: field order/size must follow the header."""
    hdr = open(os.path.join(ROOT, "include", "fvdbm_b200.h")).read()
    body = hdr[hdr.index("typedef struct fvdbm_desc {"):hdr.index("} fvdbm_desc;")]
    names = re.findall(r"(?:int32_t|int64_t|double|const int32_t\*|const void\*)\s+([^;]+);", body)
    flat = []
    for n in names:
        flat += [x.strip().split("[")[0].lstrip("*") for x in n.split(",")]
    assert flat == [f[0] for f in _lib.Desc._fields_]
    assert C.sizeof(_lib.Desc) == 8 * 4 + 4 * 8 + 2 * 8 + 16 * 8 + 4 * 8 + 16 * 8


@pytest.mark.parametrize("name", ["cylinder_lw", "quad_ldc_d2q13", "channel_upwind"])
def test_host_plan_layout(name):
    case = golden.Case(name)
    n = case.static["cells.face_indices"].shape[0]
    perm = np.random.default_rng(1).permutation(n).astype(np.int32)
    da = case.desc_arrays(np.float64, perm=perm)
    hp = _lib.HostPlan(da)
    N, K, Npad = hp.scalar("N"), da.K, hp.scalar("Npad")
    assert hp.scalar("fused_ok") == 1 and Npad % 512 == 0
    pos, ipos, code = hp.array("pos"), hp.array("ipos"), hp.array("ccode")
    assert sorted(pos.tolist()) == sorted(set(pos.tolist())) and pos.min() >= 0 and pos.max() < Npad
    assert np.array_equal(ipos[pos], np.arange(N))
    code = code.reshape(Npad // 32, K, 32)
    st, fi, sg = case.static["faces.stencil_cells_index"], case.static["cells.face_indices"], case.static["cells.face_normals"]
    nb_sides = 0
    Bstart = hp.scalar("Bstart")
    for c in range(N):
        p = pos[c]
        boundary = False
        for k in range(K):
            cd = int(code[p >> 5, k, p & 31])
            a, b = st[fi[c, k]]
            other = b if a == c else a
            if cd >= 0:
                assert ipos[cd >> 2] == other and (cd & 1) == (0 if a == c else 1) and ((cd >> 1) & 1) == (sg[c, k] < 0)
            else:
                assert other == -1
                boundary = True
                nb_sides += 1
        assert (p >= Bstart) == boundary            # interior cells first, border cells after Bstart
    assert nb_sides == hp.scalar("NB")
    holes = np.nonzero(ipos < 0)[0]
    assert np.all(code[holes >> 5, 0, holes & 31] == np.iinfo(np.int32).min)
    # tracked nodes: every typed node and every node of a ghost face, active ones first
    typ = case.static["nodes.type"].reshape(-1)
    ghost_faces = (st < 0).any(axis=1)
    want = set(np.nonzero(typ != 0)[0].tolist()) | set(case.static["faces.nodes_index"][ghost_faces].reshape(-1).tolist())
    tn = hp.array("tn_orig")
    assert set(tn.tolist()) == want and hp.scalar("NT") == len(want)
    na = hp.scalar("NA")
    assert np.all(typ[tn[:na]] != 0) and np.all(typ[tn[na:]] == 0)
    # boundary sides are numbered in position order and carry the tracked ids of their face's two nodes
    bf_na, bf_nb = hp.array("bf_na"), hp.array("bf_nb")
    track = hp.array("node_track")
    fnodes = case.static["faces.nodes_index"]
    b = 0
    for p in range(Bstart, Npad):
        if ipos[p] < 0:
            continue
        for k in range(K):
            cd = int(code[p >> 5, k, p & 31])
            if cd < 0 and cd != np.iinfo(np.int32).min:
                assert (-(cd + 1)) >> 2 == b
                j = fi[ipos[p], k]
                assert bf_na[b] == track[fnodes[j, 0]] and bf_nb[b] == track[fnodes[j, 1]]
                b += 1
    assert b == hp.scalar("NB")
    hp.close()


def test_plan_rejects_bad_input():
    case = golden.Case("ldc_tri_lw")
    da = case.desc_arrays(np.float32)
    bad = da.keep["cell_face_idx"].copy()
    bad[0, 0] = -1
    da.keep["cell_face_idx"] = bad
    da.desc.cell_face_idx = bad.ctypes.data
    with pytest.raises(ValueError, match="ragged"):
        _lib.HostPlan(da)
    da = case.desc_arrays(np.float32, perm=np.zeros(96, np.int32))
    with pytest.raises(ValueError, match="bijection"):
        _lib.HostPlan(da)
    with pytest.raises(ValueError, match="Unknown flux scheme"):
        case.scheme = "weno"
        case.desc_arrays(np.float32)


def test_inconsistent_mesh_falls_back_to_staged_plan():
    case = golden.Case("ldc_tri_lw")
    da = case.desc_arrays(np.float32)
    bad = da.keep["cell_face_sign"].copy()
    bad[3, 1] = 2                                     # not +-1 -> only the general staged kernels apply
    da.keep["cell_face_sign"] = bad
    da.desc.cell_face_sign = bad.ctypes.data
    hp = _lib.HostPlan(da)
    assert hp.scalar("fused_ok") == 0
    hp.close()


@pytest.mark.parametrize("backend", ["native", "numpy"])
@pytest.mark.parametrize("name", [n for n in golden.names() if not n.startswith("quad")])
def test_mesher_matches_reference_mesher(name, backend):
    """Integer connectivity bit-identical, float geometry to a few ulp (mesher.py docstring); both the native
    (fvdbm_mesh_properties, csrc/mesh.hpp) and the NumPy Mesher-equivalent against the reference Mesher's arrays."""
    case = golden.Case(name)
    g = case.g
    m = fb.Mesher()
    m.import_meshpy(case.raw())
    m.calc_mesh_properties(backend=backend)
    for key in [k[7:] for k in g.files if k.startswith("mesher.")]:
        ref, mine = g["mesher." + key], np.asarray(getattr(m, key))
        if ref.dtype.kind in "iu":
            assert np.array_equal(ref, mine), key
        elif key == "face_stencil_angles":            # arccos amplifies the last-bit differences near alpha = 0
            assert np.max(np.abs(ref - mine)) <= 1e-7, key
        else:
            assert np.max(np.abs(ref - mine)) <= 4 * np.finfo(np.float64).eps * max(1.0, np.max(np.abs(ref))), key
    # to_env + BC setters reproduce the statics the reference handed to Environment
    dyn = fb.D2Q9(case.tau, case.delta_t) if case.Q == 9 else fb.D2Q13(case.tau, case.delta_t)
    method = str(g["meta.flux_method"]) if "meta.flux_method" in g.files else case.scheme
    cells, faces, nodes = m.to_env(dyn, flux_method=method, dim_multiplier=float(g["meta.dim_multiplier"]))
    assert faces.flux_scheme == case.scheme
    if method.startswith("cc_"):
        assert np.allclose(case.static["faces.alpha"], faces.alpha, rtol=0, atol=1e-7)   # arccos near 0 amplifies ulps
    for kind, marker, val in eval(str(g["meta.bcs"])):
        nodes = m.set_vel_node(nodes, marker, np.array(val)) if kind == "vel" else m.set_rho_node(nodes, marker, val)
    got = {"cells.face_indices": cells.face_indices, "cells.face_normals": cells.face_normals,
           "faces.nodes_index": faces.nodes_index, "faces.stencil_cells_index": faces.stencil_cells_index,
           "faces.stencil_dists": faces.stencil_dists, "faces.n": faces.n, "faces.L": faces.L, "nodes.type": nodes.type,
           "nodes.cells_index": nodes.cells_index, "nodes.cell_dists": nodes.cell_dists}
    for key, val in got.items():
        ref = case.static[key]
        if ref.dtype.kind in "iu":
            assert np.array_equal(ref, np.asarray(val).reshape(ref.shape)), key
        else:
            assert np.allclose(ref, np.asarray(val).reshape(ref.shape), rtol=1e-15, atol=1e-16), key
    assert np.allclose(case.init["nodes.vel"], nodes.vel) and np.allclose(case.init["nodes.rho"], nodes.rho)


MESHER_ARRAYS = ["cells", "faces", "cell_centers", "cell_face_indices", "cell_face_normals", "cell_face_normal_signs",
                 "face_centers", "face_normals", "face_lengths", "face_cell_indices", "face_cell_center_distances",
                 "stencil_norms", "cc_stencil_dist", "face_stencil_angles", "point_cell_indices", "point_cell_center_distances"]


def _both_backends(raw):
    out = []
    for backend in ("numpy", "native"):
        m = fb.Mesher()
        m.import_meshpy(raw)
        with np.errstate(all="ignore"):
            m.calc_mesh_properties(backend=backend)
        out.append(m)
    return out


def _assert_same_mesher_arrays(a, b):
    for key in MESHER_ARRAYS:
        x, y = getattr(a, key), getattr(b, key)
        assert x.shape == y.shape and x.dtype == y.dtype, key
        if key == "face_stencil_angles":      # libm acos vs NumPy's SIMD arccos: last bit
            assert np.allclose(x, y, rtol=0, atol=4e-16, equal_nan=True), key
        else:
            assert np.array_equal(x, y, equal_nan=True), key


def _shuffled(raw, seed):
    """Same mesh with cells, faces and the vertices inside every cell / face in random order."""
    rng = np.random.default_rng(seed)
    el = np.array(raw.elements)[rng.permutation(len(raw.elements))]
    rot = rng.integers(0, 3, el.shape[0])
    el = np.stack([el[np.arange(el.shape[0]), (rot + k) % 3] for k in range(3)], axis=1)
    fa = np.array(raw.faces)[rng.permutation(len(raw.faces))]
    flip = rng.random(fa.shape[0]) < 0.5
    fa[flip] = fa[flip][:, ::-1]
    return types.SimpleNamespace(points=raw.points, elements=el, faces=fa, point_markers=raw.point_markers,
                                 point_alias=getattr(raw, "point_alias", None))


@pytest.mark.parametrize("mesh", ["square", "periodic", "cylinder", "porous", "shuffled", "shuffled_periodic"])
def test_native_mesher_is_bitwise_the_numpy_mesher(mesh):
    """fvdbm_mesh_properties (C++/OpenMP, CSR tables filled with atomics) == the sort-based NumPy sweep on every array,
    for any element / face order, with and without periodic identification, independent of the thread count."""
    raw = {"square": lambda: meshgen.triangulated_square(24, 16, seed=3),
           "periodic": lambda: meshgen.triangulated_square(30, 20, seed=1, periodic_x=True),
           "cylinder": lambda: meshgen.cylinder_channel(scale=1),
           "porous": lambda: meshgen.porous_channel(scale=1.0),
           "shuffled": lambda: _shuffled(meshgen.triangulated_square(37, 23, seed=5), 11),
           "shuffled_periodic": lambda: _shuffled(meshgen.triangulated_square(21, 34, seed=6, periodic_x=True), 12)}[mesh]()
    a, b = _both_backends(raw)
    _assert_same_mesher_arrays(a, b)
    old = os.environ.get("FVDBM_PLAN_THREADS")
    try:
        for nt in ("1", "3"):
            os.environ["FVDBM_PLAN_THREADS"] = nt
            c = fb.Mesher()
            c.import_meshpy(raw)
            c.calc_mesh_properties()
            for key in MESHER_ARRAYS:
                assert np.array_equal(getattr(b, key), getattr(c, key), equal_nan=True), (key, nt)
    finally:
        if old is None:
            os.environ.pop("FVDBM_PLAN_THREADS", None)
        else:
            os.environ["FVDBM_PLAN_THREADS"] = old


@pytest.mark.parametrize("periodic", [False, True])
def test_native_unique_edges_is_the_numpy_one(periodic):
    """fvdbm_mesh_unique_edges == np.unique-based meshgen.unique_edges: same rows, same order, point ids of the first
    cell edge that produced a face; any cell order, triangles and quads."""
    rng = np.random.default_rng(4)
    raw = meshgen.triangulated_square(23, 17, seed=8, periodic_x=periodic, with_faces=False)
    el = raw.elements[rng.permutation(raw.elements.shape[0])]
    a = meshgen.unique_edges(el, raw.points.shape[0], raw.point_alias, backend="numpy")
    b = meshgen.unique_edges(el, raw.points.shape[0], raw.point_alias, backend="native")
    assert a.dtype == b.dtype and np.array_equal(a, b)
    quads = np.array([[0, 1, 4, 3], [1, 2, 5, 4], [3, 4, 7, 6], [4, 5, 8, 7]], dtype=np.int32)
    assert np.array_equal(meshgen.unique_edges(quads, 9, backend="numpy"), meshgen.unique_edges(quads, 9, backend="native"))
    assert meshgen.unique_edges(np.zeros((0, 3), np.int32), 5).shape == (0, 2)
    with pytest.raises(ValueError):
        meshgen.unique_edges(np.array([[0, 1, 7]], dtype=np.int32), 5)


def test_native_mesher_edge_cases():
    """Duplicate faces (the last one carrying a key wins, mesher.py:129), a face no cell uses, a missing face
    (KeyError like the reference's dict lookup) and ids out of range (ValueError)."""
    raw = meshgen.triangulated_square(6, 5, seed=2)
    faces = np.array(raw.faces)
    extra = np.concatenate([faces, faces[3:7][:, ::-1], np.array([[0, len(raw.points) - 1]], dtype=faces.dtype)])
    dup = types.SimpleNamespace(points=raw.points, elements=raw.elements, faces=extra, point_markers=raw.point_markers)
    a, b = _both_backends(dup)
    _assert_same_mesher_arrays(a, b)
    assert set(np.arange(len(faces), len(faces) + 4)) <= set(b.cell_face_indices.ravel())       # the duplicates won
    assert np.array_equal(b.face_cell_indices[-1], [-1, -1]) and np.isnan(b.face_stencil_angles[-1])
    missing = types.SimpleNamespace(points=raw.points, elements=raw.elements, faces=faces[1:], point_markers=raw.point_markers)
    for backend in ("numpy", "native"):
        m = fb.Mesher()
        m.import_meshpy(missing)
        with pytest.raises(KeyError):
            m.calc_mesh_properties(backend=backend)
    bad = np.array(raw.elements).copy()
    bad[0, 0] = len(raw.points)
    m = fb.Mesher()
    m.points, m.cells, m.faces = np.array(raw.points, dtype=np.float64), bad.astype(np.int32), faces.astype(np.int32)
    m.point_markers = np.array(raw.point_markers)
    with pytest.raises(ValueError):
        m.calc_mesh_properties()
    with pytest.raises(ValueError):
        fb.Mesher().calc_mesh_properties(backend="triangle")
    empty = fb.Mesher()
    empty.points, empty.cells, empty.faces = np.zeros((0, 2)), np.zeros((0, 3), np.int32), np.zeros((0, 2), np.int32)
    empty.calc_mesh_properties()
    assert empty.point_cell_indices.shape == (0, 0) and empty.cell_face_indices.shape == (0, 3)


def test_quad_cavity_builder_equals_notebook_route():
    """meshgen.quad_cavity == tests/ldcFVDBM.ipynb c4-c9 executed through the reference API."""
    for name in ("quad_ldc_d2q9", "quad_ldc_d2q13"):
        case = golden.Case(name)
        dyn = (fb.D2Q9 if case.Q == 9 else fb.D2Q13)(case.tau, case.delta_t)
        c, f, n = meshgen.quad_cavity(6, 6, dyn, 0.1)
        for arr, key in ((c.face_indices, "cells.face_indices"), (c.face_normals, "cells.face_normals"),
                         (f.nodes_index, "faces.nodes_index"), (f.stencil_cells_index, "faces.stencil_cells_index"),
                         (f.stencil_dists, "faces.stencil_dists"), (f.n, "faces.n"), (f.L, "faces.L"),
                         (n.type, "nodes.type"), (n.cells_index, "nodes.cells_index"), (n.cell_dists, "nodes.cell_dists")):
            ref = case.static[key]
            assert np.array_equal(np.asarray(arr, dtype=np.float64).reshape(ref.shape), ref.astype(np.float64)), key
        assert np.array_equal(n.vel, case.init["nodes.vel"])


def test_environment_surface_without_gpu(tmp_path):
    """Constructor / factories / init / attribute pass-through work without touching the GPU."""
    case = golden.Case("channel_lw")
    cells, faces, nodes = case.containers()
    env = fb.Environment(cells, faces, nodes)
    env.init()
    assert env.cells.pdf.shape == (100, 9) and env.nodes.type.shape[1] == 1 and env.faces.flux_scheme == "lax_wendroff"
    env.cells.pdf = env.cells.pdf * 1.0                      # assignable before the engine exists
    assert "Environment(cells=" in repr(env)
    fb.Environment.dynamics = fb.D2Q9(0.8, 0.1)
    e2 = fb.Environment.create(4, 6, 5)
    e2.cells.face_indices.add_items(0, [1, 2, 3])
    e2.init()
    assert np.asarray(e2.cells.face_indices).shape == (4, 3) and e2.cells.face_indices[0, 2] == 3
    e3 = fb.Environment.define(cells, faces, nodes)
    assert e3.cells.rho.shape == (100, 1)
    # describe() builds the C descriptor (no device needed) with dtype-cast statics
    da = env._describe()
    assert da.desc.N == 100 and da.desc.Q == 9 and da.desc.K == 3 and da.desc.dtype == 32
    assert da.keep["face_n"].dtype == np.float32
    # VTK export of host-side state
    m = fb.Mesher()
    m.import_meshpy(case.raw())
    m.calc_mesh_properties()
    path = m.to_vtk(env, str(tmp_path / "out"), save_f=True)
    txt = open(path).read()
    assert "UNSTRUCTURED_GRID" in txt and "VECTORS Velocity double" in txt and f"CELLS 100 400" in txt and "pdf 9 100 double" in txt


def _read_legacy_vtk_binary(path):
    """Minimal reader of the BINARY legacy format (big-endian raw arrays after each header line)."""
    raw = open(path, "rb").read()
    pos = 0
    out = {}

    def line():
        nonlocal pos
        while raw[pos:pos + 1] == b"\n":
            pos += 1
        end = raw.index(b"\n", pos)
        text = raw[pos:end].decode()
        pos = end + 1
        return text

    def arr(dt, count):
        nonlocal pos
        a = np.frombuffer(raw, dtype=np.dtype(dt).newbyteorder(">"), count=count, offset=pos)
        pos += a.nbytes
        return a

    assert line().startswith("# vtk DataFile Version")
    line()
    assert line() == "BINARY" and line() == "DATASET UNSTRUCTURED_GRID"
    npts = int(line().split()[1])
    out["points"] = arr("f8", 3 * npts).reshape(npts, 3)
    _, n, tot = line().split()
    out["cells"] = arr("i4", int(tot)).reshape(int(n), -1)
    assert line() == f"CELL_TYPES {n}"
    out["types"] = arr("i4", int(n))
    assert line() == f"CELL_DATA {n}" and line() == "VECTORS Velocity double"
    out["Velocity"] = arr("f8", 3 * int(n)).reshape(int(n), 3)
    assert line() == "SCALARS Density double 1" and line() == "LOOKUP_TABLE default"
    out["Density"] = arr("f8", int(n))
    nf = int(line().split()[2])
    for _ in range(nf):
        name, comps, tuples, _ = line().split()
        out[name] = arr("f8", int(comps) * int(tuples)).reshape(int(tuples), int(comps))
    return out


def test_binary_vtk_round_trip(tmp_path):
    """to_vtk(binary=True): the BINARY legacy file pyvista's grid.save writes by default (reference mesher.py:562-598),
    read back array by array."""
    case = golden.Case("channel_lw")
    cells, faces, nodes = case.containers()
    env = fb.Environment(cells, faces, nodes)
    env.init()
    m = fb.Mesher()
    m.import_meshpy(case.raw())
    m.calc_mesh_properties()
    got = _read_legacy_vtk_binary(m.to_vtk(env, str(tmp_path / "bin"), save_f=True, save_feq=True, binary=True))
    assert np.array_equal(got["points"][:, :2], m.points) and not got["points"][:, 2].any()
    assert np.array_equal(got["cells"][:, 1:], m.cells) and (got["cells"][:, 0] == 3).all() and (got["types"] == 5).all()
    assert np.array_equal(got["Velocity"][:, :2], np.asarray(env.cells.vel)) and np.array_equal(got["Density"], np.asarray(env.cells.rho).ravel())
    assert np.array_equal(got["pdf"], np.asarray(env.cells.pdf)) and np.array_equal(got["feq"], np.asarray(env.cells.pdf_eq))
    txt = open(m.to_vtk(env, str(tmp_path / "auto"))).read()         # 100 cells: ASCII by default
    assert "ASCII" in txt


def test_stencil_geometry_report(capsys):
    """verify_stencil_geometry (reference mesher.py:386-504; pinned to the reference's printout in
    tests/test_reference_containers.py): on an unjittered grid the face-normal projections are exact, the diagonal
    faces' centres sit on the centre-to-centre midpoint and the report has the reference's 14 lines."""
    m = fb.Mesher()
    m.import_meshpy(meshgen.triangulated_square(6, 4, jitter=0.0))
    m.calc_mesh_properties()
    r = m.verify_stencil_geometry()
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip()]
    assert len(out) == 14 and out[0].startswith("Mean angle (deg):") and out[-1].startswith("Faces with sign mismatch")
    assert r["interior_faces"] == int(((m.face_cell_indices != -1).all(axis=1)).sum())
    assert r["distance_error_max"] < 1e-12 and r["bad_distance_faces"] == 0 and r["bad_angle_faces"] == 0
    assert m.verify_stencil_geometry(verbose=False) == r and capsys.readouterr().out == ""


def test_host_state_is_c_ordered_whatever_the_input_strides():
    """Fortran-ordered or broadcast inputs (np.array / astype keep such strides by default) must not reach the raw
    pointers of the C ABI: the host copies Environment hands to fvdbm_create / fvdbm_get are C-contiguous."""
    m = fb.Mesher()
    m.import_meshpy(meshgen.triangulated_square(3, 2, jitter=0.0))
    m.calc_mesh_properties()
    dyn = fb.D2Q9(0.8, 0.1)
    cells, faces, nodes = m.to_env(dyn, "lax_wendroff")
    for cont in (cells, faces, nodes):         # fresh containers: real, writable, C-ordered arrays
        assert cont.pdf.flags.c_contiguous and cont.pdf.flags.writeable
    ref = {"cells.pdf": np.array(cells.pdf), "nodes.pdf": np.array(nodes.pdf)}
    cells.pdf = np.asfortranarray(cells.pdf)
    nodes.pdf = np.broadcast_to(np.asarray(nodes.pdf)[0], nodes.pdf.shape)
    nodes.vel = np.asfortranarray(nodes.vel)
    env = fb.Environment(cells, faces, nodes, dtype=np.float32)
    env.init()
    da = env._describe()
    for name, arr in env._host.items():
        assert arr.flags.c_contiguous, name
    for name in ref:
        assert np.array_equal(env._host[name], ref[name].astype(np.float32)), name
    assert all(a.flags.c_contiguous for a in da.keep.values() if isinstance(a, np.ndarray))


def test_custom_array_semantics():
    a = fb.CustomArray(3, dtype=np.int32, default_value=-1)
    a.add_items(1, [5, 6])
    a.add_item(1, 7)
    a.add_item(0, 9)
    assert np.array_equal(np.asarray(a), [[9, -1, -1], [5, 6, 7], [-1, -1, -1]])
    assert a.shape() == (3, 3) and a[1][2] == 7


def test_cc_alt_variant_is_rejected_like_it_fails_in_the_reference():
    """CCStencilKsiFaces divides by KSI.n_PQ (0 for axis-aligned stencils): NaN in the reference
    itself ("Currently Inoperable").  We refuse it instead of producing NaN."""
    case = golden.Case("cc_ldc_upwind")
    m = fb.Mesher()
    m.import_meshpy(case.raw())
    m.calc_mesh_properties()
    with pytest.raises(ValueError, match="cc_alt_upwind"):
        m.to_env(fb.D2Q9(0.8, 0.1), flux_method="cc_alt_upwind")
    with pytest.raises(ValueError, match="Unsupported flux method"):
        m.to_env(fb.D2Q9(0.8, 0.1), flux_method="weno")
    cells, faces, nodes = case.containers()
    faces.npq = np.zeros_like(faces.n)
    with pytest.raises(ValueError, match="CCStencilKsiFaces"):
        fb.Environment(cells, faces, nodes)._describe()


COMPAT_NOTEBOOK = r"""
import sys
sys.path.append('..')                                   # the notebooks' first line: harmless here
import fvdbm_jax_b200.compat; fvdbm_jax_b200.compat.install()   # <- the one added line
from src.dynamics import *
from src.environment import *
from src.mesher import *
from src.containers import *
from src.cells import *
from src.faces import *
from src.nodes import *
from utils.test_utils import *
import numpy as np
import fvdbm_jax_b200 as fb
from fvdbm_jax_b200 import meshgen
assert Environment is fb.Environment and Mesher is fb.Mesher and Cells is fb.Cells and Faces is fb.Faces and Nodes is fb.Nodes
assert D2Q9 is fb.D2Q9 and D2Q13 is fb.D2Q13 and CustomArray is fb.CustomArray and CCStencilFaces is fb.CCStencilFaces
# tests/flow_over_cyl.ipynb c10-c16 with the synthetic triangulation standing in for meshpy.triangle
mesher = Mesher()
mesher.import_meshpy(meshgen.cylinder_channel(scale=1))
mesher.calc_mesh_properties()
dynamics = D2Q9(tau=0.65, delta_t=0.1)
cells, faces, nodes = mesher.to_env(dynamics, flux_method="lax_wendroff")
nodes = mesher.set_vel_node(nodes, marker=4, velocity=np.array([0.1, 0.0]))
nodes = mesher.set_rho_node(nodes, marker=2, rho=0.95)
env = Environment(cells, faces, nodes)
env.init()
assert env.cells.pdf.shape == (2392, 9) and "Environment(cells=" in repr(env)
# tests/ldcFVDBM.ipynb c4-c9: Environment.create + CustomArray.add_items
Environment.dynamics = D2Q13(tau=0.8, delta_t=0.1)
env2 = Environment.create(4, 12, 9)
env2.cells.face_indices.add_items(0, np.asarray([0, 1, 2, 3]))
env2.init()
assert np.asarray(env2.cells.face_indices)[0].tolist() == [0, 1, 2, 3]
cc = mesher.to_env(dynamics, flux_method="cc_upwind")[1]
assert isinstance(cc, CCStencilFaces) and cc.alpha.shape == (mesher.faces.shape[0], 1)
assert abs(extrap_pdf(2.0, 1.0, 0.5, 1.0) - 2.5) < 1e-15 and Key(0)() is not None
print("compat ok")
"""


def test_reference_module_names_resolve_to_this_framework():
    """fvdbm_jax_b200.compat: the notebooks' own ``from src.environment import *`` ... lines, unedited, bind this
    framework's classes after one added line (run in a subprocess: the oracle tests import the REAL reference under
    the same module names)."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, "-c", COMPAT_NOTEBOOK], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0 and "compat ok" in r.stdout, r.stderr[-3000:]
