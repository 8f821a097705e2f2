"""Vectorised Mesher-equivalent: the producer of the hot path's static arrays.

Mirrors the public surface of the reference ``Mesher`` (/root/reference/src/mesher.py:20-742) that
the notebooks use -- ``import_meshpy``, ``calc_mesh_properties``, ``to_env``, ``set_vel_node``,
``set_rho_node``, ``to_vtk``, ``to_pickle``/``from_pickle`` -- but replaces its per-element Python
loops (~130 us/cell) by sort-based NumPy so 10^7-cell meshes are reachable.

Contract (tests/test_host_logic.py::test_mesher_matches_reference_mesher): every integer array (``cells`` after the CCW fix, ``faces``
after the boundary flip, ``cell_face_indices``, ``cell_face_normal_signs``, ``face_cell_indices``,
``point_cell_indices``) is bit-identical to the reference Mesher's; float geometry agrees to a few
ulp (the reference's 2-vector ``np.dot``/``np.linalg.norm`` go through BLAS ddot whose FMA use is
machine dependent, so "last bit" is not defined by the reference itself).

Extension: ``point_alias`` (periodic identification).  Connectivity uses canonical point ids,
geometry is always evaluated from a cell's own vertices, so a face shared across the periodic seam
gets the right normal, length and stencil distances on both sides.
"""
from __future__ import annotations

import pickle
import numpy as np

from .containers import CCStencilFaces, Cells, Faces, Nodes
from .environment import Environment

__all__ = ["Mesher"]


def _normal(p0, p1):
    """Unit left normal of p0->p1 (reference utils/utils.py:162-173), vectorised."""
    t = p1 - p0
    nrm = np.stack([-t[:, 1], t[:, 0]], axis=1)
    return nrm / np.sqrt(nrm[:, 0] * nrm[:, 0] + nrm[:, 1] * nrm[:, 1])[:, None]


def _dot(a, b):
    return a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]


class Mesher:
    """Drop-in for the reference Mesher on the path that feeds ``Environment`` (triangles)."""

    def __init__(self):
        self.points = None
        self.cells = None
        self.neighbors = None
        self.faces = None
        self.point_markers = None
        self.point_alias = None

    # ------------------------------------------------------------------ import
    def import_meshpy(self, mesh):
        """reference mesher.py:48-61 (any object with points/elements/faces/point_markers)."""
        self.points = np.array(mesh.points, dtype=np.float64)
        self.cells = np.array(mesh.elements, dtype=np.int32)
        self.faces = np.array(mesh.faces, dtype=np.int32)
        self.point_markers = np.array(mesh.point_markers, dtype=np.int32)
        alias = getattr(mesh, "point_alias", None)
        self.point_alias = None if alias is None else np.asarray(alias, dtype=np.int32)
        self.enforce_ccw()

    import_raw = import_meshpy

    def enforce_ccw(self):
        """reference mesher.py:63-78: swap vertices 1,2 of clockwise triangles."""
        p = self.points[self.cells]
        area = 0.5 * ((p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1])
                      - (p[:, 2, 0] - p[:, 0, 0]) * (p[:, 1, 1] - p[:, 0, 1]))
        neg = area < 0
        c = self.cells.copy()
        c[neg, 1], c[neg, 2] = self.cells[neg, 2], self.cells[neg, 1]
        self.cells = c

    # ------------------------------------------------------------- connectivity
    def _canon(self, ids):
        return ids if self.point_alias is None else self.point_alias[ids]

    def _keys(self, a, b):
        a = self._canon(a).astype(np.int64)
        b = self._canon(b).astype(np.int64)
        return np.minimum(a, b) * np.int64(self.points.shape[0]) + np.maximum(a, b)

    def calc_mesh_properties(self, verbose: bool = False, backend: str = "native"):
        """Same results as reference mesher.py:319-383.  ``backend="native"`` (default): the C++/OpenMP
        Mesher-equivalent of the C-ABI library (``fvdbm_mesh_properties``, csrc/mesh.hpp; 10 M cells in about
        a second); ``backend="numpy"``: the sort-based NumPy sweep below.  Both give bit-identical arrays
        (tests/test_host_logic.py)."""
        if backend == "native":
            return self._calc_mesh_properties_native(verbose)
        if backend != "numpy":
            raise ValueError(f"unknown Mesher backend: {backend}")
        pts, cells, faces = self.points, self.cells, self.faces
        N, F, P = cells.shape[0], faces.shape[0], pts.shape[0]
        K = 3

        # cell centres (mesher.py:113-120)
        p = pts[cells]                                            # (N,3,2)
        self.cell_centers = (p[:, 0] + p[:, 1] + p[:, 2]) / 3.0

        # half-edges (c,k): vertices k -> k+1, in cell-major order
        ha = cells.reshape(-1)
        hb = np.roll(cells, -1, axis=1).reshape(-1)
        hcell = np.repeat(np.arange(N, dtype=np.int64), K)
        hkey = self._keys(ha, hb)
        fkey = self._keys(faces[:, 0], faces[:, 1])

        # face lookup: last face index carrying a key wins (dict comprehension, mesher.py:129)
        forder = np.argsort(fkey, kind="stable")
        fsorted = fkey[forder]
        pos = np.searchsorted(fsorted, hkey, side="right") - 1
        if np.any(pos < 0) or np.any(fsorted[pos] != hkey):
            raise KeyError("a cell edge is missing from `faces`")
        hface = forder[pos]
        self.cell_face_indices = hface.reshape(N, K).astype(np.int64)

        # face -> first two cells that list it, ascending cell id (mesher.py:206-220)
        horder = np.argsort(hkey, kind="stable")
        hk_s = hkey[horder]
        start = np.searchsorted(hk_s, fkey, side="left")
        stop = np.searchsorted(hk_s, fkey, side="right")
        cnt = stop - start
        he0 = horder[np.minimum(start, hk_s.size - 1)]           # first half-edge of the face
        he1 = horder[np.minimum(start + 1, hk_s.size - 1)]
        c0 = np.where(cnt > 0, hcell[he0], -1)
        c1 = np.where(cnt > 1, hcell[he1], -1)

        # boundary faces: make the stored node order give an outward normal (mesher.py:80-110).
        # Geometry from the owning cell's own vertices (identical to faces[] without aliasing).
        bnd = cnt == 1
        fp0 = pts[faces[:, 0]]
        fp1 = pts[faces[:, 1]]
        if self.point_alias is not None:
            # express the face with the point ids of its first half-edge (unwrapped coordinates)
            own0, own1 = ha[he0], hb[he0]
            same = self._canon(own0) == self._canon(faces[:, 0])
            f0 = np.where(same, own0, own1)
            f1 = np.where(same, own1, own0)
            has = cnt > 0
            faces = faces.copy()
            faces[has, 0], faces[has, 1] = f0[has], f1[has]
            fp0, fp1 = pts[faces[:, 0]], pts[faces[:, 1]]
        nb = _normal(fp0[bnd], fp1[bnd])
        mid_b = (fp0[bnd] + fp1[bnd]) / 2.0
        inward = _dot(nb, self.cell_centers[c0[bnd]] - mid_b) >= 0
        flip = np.zeros(F, dtype=bool)
        flip[np.nonzero(bnd)[0][inward]] = True
        faces = faces.copy()
        faces[flip] = faces[flip][:, ::-1]
        self.faces = faces
        fp0, fp1 = pts[faces[:, 0]], pts[faces[:, 1]]

        # face centres / normals / lengths (mesher.py:172-195)
        self.face_centers = (fp0 + fp1) / 2.0
        self.face_normals = _normal(fp0, fp1)
        t = fp1 - fp0
        self.face_lengths = np.sqrt(t[:, 0] * t[:, 0] + t[:, 1] * t[:, 1])

        # outward cell-face normals and their sign against the face normal (mesher.py:140-169)
        hp0, hp1 = pts[ha], pts[hb]
        hn = _normal(hp0, hp1)
        hmid = (hp0 + hp1) / 2.0
        hcen = np.repeat(self.cell_centers, K, axis=0)
        out = _dot(hn, hcen - hmid) < 0
        hn = np.where(out[:, None], hn, -hn)
        self.cell_face_normals = hn.reshape(N, K, 2)
        self.cell_face_normal_signs = np.sign(_dot(hn, self.face_normals[hface])).astype(np.int32).reshape(N, K)

        # face stencil: slot0 -> slot1 along the face normal, projected distances (mesher.py:222-266)
        # each cell measures from the midpoint of its *own* copy of the edge (periodic-safe)
        fn = self.face_normals
        d0 = np.where(cnt > 0, _dot(fn, self.cell_centers[np.maximum(c0, 0)] - hmid[he0]), -1.0)
        d1 = np.where(cnt > 1, _dot(fn, self.cell_centers[np.maximum(c1, 0)] - hmid[he1]), -1.0)
        interior = cnt > 1
        swap = interior & ~(d0 < d1)
        s0 = np.where(swap, c1, c0)
        s1 = np.where(swap, c0, c1)
        e0 = np.where(swap, d1, d0)
        e1 = np.where(swap, d0, d1)
        self.face_cell_indices = np.stack([s0, s1], axis=1).astype(np.int64)
        dists = np.abs(np.stack([e0, e1], axis=1))
        # ghost distance = distance of the real cell (mesher.py:268-283)
        g0 = self.face_cell_indices[:, 0] == -1
        g1 = self.face_cell_indices[:, 1] == -1
        dists[g0, 0] = dists[g0, 1]
        dists[g1, 1] = dists[g1, 0]
        self.face_cell_center_distances = dists

        # cell-centre stencil (cc_* flux methods): direction c0 -> c1 (discovery order) or cell -> face
        # centre, aligned with the face normal (mesher.py:506-544); distances projected on it with the
        # same slot order as above (mesher.py:231-256); angle to the face normal (mesher.py:546-558)
        cc0 = self.cell_centers[np.maximum(c0, 0)]
        v = np.where(interior[:, None], self.cell_centers[np.maximum(c1, 0)] - cc0, self.face_centers - cc0)
        v = np.where((cnt > 0)[:, None], v, 0.0)
        vn = np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1])
        sn = np.where((vn > 0)[:, None], v / np.where(vn > 0, vn, 1.0)[:, None], v)
        fnu = fn / np.sqrt(fn[:, 0] * fn[:, 0] + fn[:, 1] * fn[:, 1])[:, None]
        snu = sn / (np.sqrt(sn[:, 0] * sn[:, 0] + sn[:, 1] * sn[:, 1])[:, None] + 1e-14)
        sn = np.where((_dot(fnu, snu) < 0)[:, None], -sn, sn)
        self.stencil_norms = sn
        q0 = np.where(cnt > 0, _dot(sn, self.cell_centers[np.maximum(c0, 0)] - hmid[he0]), -1.0)
        q1 = np.where(cnt > 1, _dot(sn, self.cell_centers[np.maximum(c1, 0)] - hmid[he1]), -1.0)
        cc = np.abs(np.stack([np.where(swap, q1, q0), np.where(swap, q0, q1)], axis=1))
        cc[g0, 0] = cc[g0, 1]
        cc[g1, 1] = cc[g1, 0]
        self.cc_stencil_dist = cc
        snn = sn / np.sqrt(sn[:, 0] * sn[:, 0] + sn[:, 1] * sn[:, 1])[:, None]
        self.face_stencil_angles = np.arccos(np.clip(_dot(fnu, snn), -1.0, 1.0))

        # node -> ring cells (ascending), padded with -1; distances node-centroid (mesher.py:286-316)
        vpt = self._canon(cells.reshape(-1)).astype(np.int64)
        vorder = np.argsort(vpt, kind="stable")
        vp_s = vpt[vorder]
        vcell = hcell[vorder]                                     # hcell == repeat(arange(N),3)
        pstart = np.searchsorted(vp_s, np.arange(P), side="left")
        pcount = np.searchsorted(vp_s, np.arange(P), side="right") - pstart
        M = int(pcount.max()) if P else 0
        col = np.arange(vp_s.size) - pstart[vp_s]
        pci = -np.ones((P, M), dtype=np.int64)
        pcd = -np.ones((P, M), dtype=np.float64)
        pci[vp_s, col] = vcell
        own = pts[cells.reshape(-1)[vorder]] - self.cell_centers[vcell]
        pcd[vp_s, col] = np.sqrt(own[:, 0] * own[:, 0] + own[:, 1] * own[:, 1])
        self.point_cell_indices = pci
        self.point_cell_center_distances = pcd
        if verbose:
            print(f"mesh properties: {N} cells, {F} faces, {P} points, ring width {M}")

    def _calc_mesh_properties_native(self, verbose: bool = False):
        """One call into ``fvdbm_mesh_properties`` (include/fvdbm_b200.h): outputs are allocated here with the
        dtypes/shapes of the NumPy path and filled by the library."""
        from . import _lib
        lib = _lib.load()
        pts = np.ascontiguousarray(self.points, dtype=np.float64)
        cells = np.ascontiguousarray(self.cells, dtype=np.int32)
        faces = np.ascontiguousarray(self.faces, dtype=np.int32)
        if cells.ndim != 2 or cells.shape[1] != 3:
            raise ValueError("Mesher.calc_mesh_properties handles triangles (K = 3)")
        alias = None if self.point_alias is None else np.ascontiguousarray(self.point_alias, dtype=np.int32)
        N, F, P = cells.shape[0], faces.shape[0], pts.shape[0]
        M = int(lib.fvdbm_mesh_ring_width(cells.ctypes.data, None if alias is None else alias.ctypes.data, N, P))
        if M < 0:
            raise ValueError(_lib.last_error())
        out = {
            "cell_centers": np.empty((N, 2), np.float64), "cell_face_indices": np.empty((N, 3), np.int64),
            "cell_face_normals": np.empty((N, 3, 2), np.float64), "cell_face_normal_signs": np.empty((N, 3), np.int32),
            "faces_out": np.empty((F, 2), np.int32), "face_centers": np.empty((F, 2), np.float64),
            "face_normals": np.empty((F, 2), np.float64), "face_lengths": np.empty(F, np.float64),
            "face_cell_indices": np.empty((F, 2), np.int64), "face_cell_center_distances": np.empty((F, 2), np.float64),
            "stencil_norms": np.empty((F, 2), np.float64), "cc_stencil_dist": np.empty((F, 2), np.float64),
            "face_stencil_angles": np.empty(F, np.float64), "point_cell_indices": np.empty((P, M), np.int64),
            "point_cell_center_distances": np.empty((P, M), np.float64),
        }
        d = _lib.MeshDesc()
        d.N, d.F, d.P, d.M = N, F, P, M
        d.points, d.cells, d.faces = pts.ctypes.data, cells.ctypes.data, faces.ctypes.data
        d.point_alias = None if alias is None else alias.ctypes.data
        for name, arr in out.items():
            setattr(d, name, arr.ctypes.data)
        rc = lib.fvdbm_mesh_properties(d)
        if rc == _lib.ERR_STATE:
            raise KeyError(_lib.last_error())          # the reference's dict lookup raises KeyError (mesher.py:131)
        _lib.check(rc)
        self.faces = out.pop("faces_out")
        for name, arr in out.items():
            setattr(self, name, arr)
        if verbose:
            print(f"mesh properties: {N} cells, {F} faces, {P} points, ring width {M}")

    # ------------------------------------------------------- mesh quality report
    def verify_stencil_geometry(self, verbose: bool = True):
        """Mesh-quality report of reference mesher.py:386-504 over the interior faces, vectorised: angle between
        the centre-to-centre vector and the face normal, projected-distance error, face-centre offset, and the
        face-normal vs stencil-normal projected distances.  Prints the reference's lines (``verbose``) and returns
        the statistics as a dict (an extension; the reference returns None)."""
        fci = np.asarray(self.face_cell_indices)
        inner = (fci[:, 0] != -1) & (fci[:, 1] != -1)
        c0, c1 = self.cell_centers[fci[inner, 0]], self.cell_centers[fci[inner, 1]]
        n = self.face_normals[inner]
        L = self.face_lengths[inner]
        dists = self.face_cell_center_distances[inner]
        nrm = lambda a: np.sqrt(a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1])          # noqa: E731
        pct = lambda a: np.where(L > 0, 100.0 * a / np.where(L > 0, L, 1.0), 0.0)  # noqa: E731
        v = c1 - c0
        n_norm = n / nrm(n)[:, None]
        angles = np.degrees(np.arccos(np.clip(_dot(v / nrm(v)[:, None], n_norm), -1, 1)))
        dist_err = np.abs(_dot(v, n_norm) - (dists[:, 0] + dists[:, 1]))
        midpoint = (c0 + c1) / 2.0
        offsets = nrm(self.face_centers[inner] - midpoint)
        centre_vs_face = pct(nrm(v) - (np.abs(dists[:, 0]) + np.abs(dists[:, 1])))
        sn = self.stencil_norms[inner]
        s_norm = sn / nrm(sn)[:, None]
        proj = [(_dot(n_norm, c - midpoint), _dot(s_norm, c - midpoint)) for c in (c0, c1)]
        diffs = np.stack([np.abs(f - t) for f, t in proj], axis=1)              # (faces, [d0, d1]) like the reference's extend
        pdiffs = np.stack([pct(d) for d in diffs.T], axis=1)
        mismatches = int(sum(np.count_nonzero(np.sign(f) != np.sign(t)) for f, t in proj))
        mean = lambda a: float(np.mean(a)) if a.size else float("nan")          # noqa: E731
        amax = lambda a: float(np.max(a)) if a.size else float("nan")           # noqa: E731
        amin = lambda a: float(np.min(a)) if a.size else float("nan")           # noqa: E731
        r = {"interior_faces": int(inner.sum()),
             "angle_mean_deg": mean(angles), "angle_max_deg": amax(angles), "bad_angle_faces": int(np.count_nonzero(np.abs(angles) > 30)),
             "distance_error_mean": mean(dist_err), "distance_error_max": amax(dist_err),
             "distance_error_pct_mean": mean(pct(dist_err)), "distance_error_pct_max": amax(pct(dist_err)),
             "bad_distance_faces": int(np.count_nonzero(pct(dist_err) > 1.0)),
             "offset_mean": mean(offsets), "offset_max": amax(offsets), "offset_pct_mean": mean(pct(offsets)),
             "offset_pct_max": amax(pct(offsets)), "bad_offset_faces": int(np.count_nonzero(pct(offsets) > 10.0)),
             "centre_vs_face_pct_mean": mean(centre_vs_face), "centre_vs_face_pct_max": amax(centre_vs_face),
             "centre_vs_face_pct_min": amin(centre_vs_face),
             "stencil_diff_mean": mean(diffs), "stencil_diff_max": amax(diffs), "stencil_diff_pct_mean": mean(pdiffs),
             "stencil_diff_pct_max": amax(pdiffs), "stencil_large_diffs": int(np.count_nonzero(pdiffs > 1.0)),
             "stencil_sign_mismatches": mismatches}
        if verbose:
            print(f"Mean angle (deg): {r['angle_mean_deg']:.2f}, max: {r['angle_max_deg']:.2f}")
            print(f"Faces with angle > 30 deg: {r['bad_angle_faces']}")
            print(f"Mean distance error: {r['distance_error_mean']:.4f}, max: {r['distance_error_max']:.4f}")
            print(f"Mean distance error (% of face length): {r['distance_error_pct_mean']:.2f}%, max: {r['distance_error_pct_max']:.2f}%")
            print(f"Faces with distance error > 1% of face length: {r['bad_distance_faces']}")
            print(f"Mean face center offset: {r['offset_mean']:.4e}, max: {r['offset_max']:.4e}")
            print(f"Mean face center offset (% of face length): {r['offset_pct_mean']:.2f}%, max: {r['offset_pct_max']:.2f}%")
            print(f"Faces with face center offset > 10% of face length: {r['bad_offset_faces']}")
            print(f"Mean (cell center dist - face dist sum) as % of face length: {r['centre_vs_face_pct_mean']:.4e}%, "
                  f"max: {r['centre_vs_face_pct_max']:.4e}%, min: {r['centre_vs_face_pct_min']:.4e}%")
            print("\nStencil distance checks:")
            print(f"Mean |face-normal dist - stencil-normal dist|: {r['stencil_diff_mean']:.4e}, max: {r['stencil_diff_max']:.4e}")
            print(f"Mean |face-normal dist - stencil-normal dist| (% of face length): {r['stencil_diff_pct_mean']:.2f}%, "
                  f"max: {r['stencil_diff_pct_max']:.2f}%")
            print(f"Faces with |face-normal dist - stencil-normal dist| > 1.0%: {r['stencil_large_diffs']}")
            print(f"Faces with sign mismatch between face-normal and stencil-normal projected distances: {r['stencil_sign_mismatches']}")
        return r

    # ------------------------------------------------------------------ to_env
    def to_env(self, dynamics, flux_method="upwind", dim_multiplier=1):
        """reference mesher.py:610-693: "upwind", "lax_wendroff" and the cell-centre-stencil variants
        "cc_upwind" / "cc_lax_wendroff" (CCStencilFaces, src/faces.py:10-76: stencil direction as n,
        stencil-projected distances, flux * cos(alpha)).  "cc_alt_upwind" (CCStencilKsiFaces) divides by
        KSI.n_PQ, which is 0 for axis-aligned stencils, and yields NaN in the reference itself."""
        if flux_method == "cc_alt_upwind":
            raise ValueError("Unsupported flux method: cc_alt_upwind (inoperable in the reference: NaN)")
        if flux_method not in ("upwind", "lax_wendroff", "cc_upwind", "cc_lax_wendroff"):
            raise ValueError(f"Unsupported flux method: {flux_method}")
        cc = flux_method.startswith("cc_")
        if cc and self.point_alias is not None:
            raise ValueError("cc_* flux methods are not available on periodic meshes")
        cells = Cells(self.cells.shape[0], dynamics)
        cells.face_indices = np.asarray(self.cell_face_indices, dtype=np.int32)
        cells.face_normals = np.asarray(self.cell_face_normal_signs, dtype=np.int32)
        cells.centers = np.asarray(self.cell_centers, dtype=np.float64)   # extension: locality renumbering

        if cc:
            faces = CCStencilFaces(self.faces.shape[0], dynamics, flux_scheme=flux_method[3:])
            faces.alpha = np.asarray(self.face_stencil_angles, dtype=np.float64)[..., np.newaxis]
        else:
            faces = Faces(self.faces.shape[0], dynamics, flux_scheme=flux_method)
        faces.n = np.asarray(self.stencil_norms if cc else self.face_normals, dtype=np.float64)
        unit = dim_multiplier == 1               # x * 1 is the identity bit for bit: skip the passes over 10^7-row arrays
        faces.L = np.asarray(self.face_lengths, dtype=np.float64)[..., np.newaxis]
        if not unit:
            faces.L = faces.L * dim_multiplier
        faces.nodes_index = self._canon(np.asarray(self.faces, dtype=np.int32)).astype(np.int32)
        faces.stencil_cells_index = np.asarray(self.face_cell_indices, dtype=np.int32)
        faces.stencil_dists = np.asarray(self.cc_stencil_dist if cc else self.face_cell_center_distances, dtype=np.float64)
        if not unit:
            faces.stencil_dists = faces.stencil_dists * dim_multiplier

        nodes = Nodes(self.points.shape[0], dynamics)
        nodes.cells_index = np.asarray(self.point_cell_indices, dtype=np.int32)
        cd = np.asarray(self.point_cell_center_distances, dtype=np.float64)
        if not unit:
            cd = cd.copy()
            cd[cd > 0] = cd[cd > 0] * dim_multiplier
        nodes.cell_dists = cd
        nodes.type = np.zeros_like(self.point_markers[..., np.newaxis], dtype=np.int32)
        return cells, faces, nodes

    # ------------------------------------------------------- boundary conditions
    def _marker_mask(self, marker):
        mk = self.point_markers
        if self.point_alias is not None:
            # periodic duplicates are never referenced by a face and have no ring: leave them untyped
            canonical = self.point_alias == np.arange(mk.shape[0])
            return (mk[self.point_alias] == marker) & canonical
        return mk == marker

    def set_vel_node(self, nodes: Nodes, marker: int, velocity):
        """reference mesher.py:697-719: type 1 (velocity Dirichlet) on nodes with ``marker``."""
        velocity = np.asarray(velocity)
        assert velocity.shape == (nodes.dynamics.DIM,)
        m = self._marker_mask(marker)
        nodes.type = np.array(nodes.type, copy=True)
        nodes.vel = np.array(nodes.vel, copy=True)
        nodes.type[m] = 1
        nodes.vel[m] = velocity
        return nodes

    def set_rho_node(self, nodes: Nodes, marker: int, rho: float):
        """reference mesher.py:721-742: type 2 (density Dirichlet) on nodes with ``marker``."""
        assert rho > 0
        m = self._marker_mask(marker)
        nodes.type = np.array(nodes.type, copy=True)
        nodes.rho = np.array(nodes.rho, copy=True)
        nodes.type[m] = 2
        nodes.rho[m] = rho
        return nodes

    def cell_inv_areas(self, dim_multiplier=1):
        """1 / area of every cell (extension, not in the reference): assign to ``cells.inv_area`` for the
        physically consistent update f += dt((feq - f)/tau - (1/A) sum_k s_k flux_k)."""
        p = self.points[self.cells]
        area = 0.5 * np.abs((p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1])
                            - (p[:, 2, 0] - p[:, 0, 0]) * (p[:, 1, 1] - p[:, 0, 1]))
        return 1.0 / (area * dim_multiplier * dim_multiplier)

    # ------------------------------------------------------------------ decomposition
    def partition(self, nparts: int, method: str = "auto", refine: bool = True):
        """Owner rank of every cell for a multi-GPU run (north_star: "the Mesher partitions cells ...
        across the 8 B200s"): locality chunks of the cell graph + METIS-style boundary refinement.
        See fvdbm_jax_b200.partition / .distributed for local meshes, halos and the exchange."""
        from .partition import partition_cells
        return partition_cells(np.asarray(self.face_cell_indices), self.cells.shape[0], nparts,
                               centers=self.cell_centers, method=method, refine=refine)

    # ------------------------------------------------------------------ export
    def to_vtk(self, env: Environment, filename: str, save_f: bool = False, save_feq: bool = False, binary=None):
        """Legacy-VTK writer with the cell data of reference mesher.py:562-598 (no pyvista)."""
        from .export import write_vtk
        return write_vtk(self, env, filename, save_f, save_feq, binary)

    def to_pickle(self, env: Environment, filename: str):
        with open(f"{filename}.pkl", "wb") as f:
            pickle.dump((env, self), f)

    @classmethod
    def from_pickle(cls, filename: str):
        with open(f"{filename}.pkl", "rb") as f:
            env, mesh = pickle.load(f)
        return env, mesh
