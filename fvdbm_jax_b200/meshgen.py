"""Synthetic triangulations (no Triangle / meshpy needed).

The reference builds its meshes with meshpy.triangle (tests/flow_over_cyl.ipynb c3-c9,
tests/porous_flow.ipynb), which is absent here; BASELINE.json asks for "synthetically
triangulated meshes, generated without Triangle".  Every generator returns a ``RawMesh``
with the four arrays ``Mesher.import_meshpy`` consumes (/root/reference/src/mesher.py:48-61):
``points (P,2) f64``, ``elements (N,3) i32``, ``faces (F,2) i32`` (unique edges) and
``point_markers (P,) i32``.

Periodic meshes additionally carry ``point_alias`` (canonical id of every point); geometry is
always taken from a cell's own (unwrapped) vertices, connectivity from canonical ids.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

# box markers, counter-clockwise starting at the bottom like meshpy's make_box facets
BOTTOM, RIGHT, TOP, LEFT, OBSTACLE = 1, 2, 3, 4, 5


@dataclass
class RawMesh:
    points: np.ndarray
    elements: np.ndarray
    faces: np.ndarray
    point_markers: np.ndarray
    point_alias: Optional[np.ndarray] = None   # periodic identification (canonical point id)
    shape: Optional[tuple] = None              # (nx, ny) of the generating quad grid, if any

    @property
    def num_cells(self):
        return self.elements.shape[0]


def unique_edges(elements: np.ndarray, num_points: int, alias: Optional[np.ndarray] = None,
                 backend: str = "native") -> np.ndarray:
    """Unique undirected edges of a triangulation, one row per face, sorted by (min,max) key.

    With ``alias`` the key is built from canonical point ids (periodic identification) while the
    returned rows keep the point ids of the first half-edge that produced the key.
    """
    if backend == "native":                 # C++/OpenMP helper of the C-ABI library: no sort, same rows in the same order
        from . import _lib
        el = np.ascontiguousarray(elements, dtype=np.int32)
        al = None if alias is None else np.ascontiguousarray(alias, dtype=np.int32)
        out = np.empty((el.size, 2), dtype=np.int32)
        nf = int(_lib.load().fvdbm_mesh_unique_edges(el.ctypes.data, el.shape[0], el.shape[1], int(num_points),
                                                    None if al is None else al.ctypes.data, out.ctypes.data))
        if nf < 0:
            raise ValueError(_lib.last_error())
        return out[:nf].copy()
    k = elements.shape[1]
    a = elements.reshape(-1)
    b = np.roll(elements, -1, axis=1).reshape(-1)
    ca, cb = (a, b) if alias is None else (alias[a], alias[b])
    key = np.minimum(ca, cb).astype(np.int64) * np.int64(num_points) + np.maximum(ca, cb).astype(np.int64)
    _, first = np.unique(key, return_index=True)
    del k
    return np.stack([a[first], b[first]], axis=1).astype(np.int32)


def triangulated_square(nx: int, ny: int, jitter: float = 0.2, seed: int = 0,
                        periodic_x: bool = False, lx: float | None = None, ly: float | None = None,
                        with_faces: bool = True) -> RawMesh:
    """``nx x ny`` unit quads, interior vertices jittered by U(-jitter, jitter)^2, each quad split
    along alternating diagonals ((i+j)%2), triangles counter-clockwise.

    Cell numbering is row-major over quads, two triangles per quad (cell 2*(j*nx+i)+{0,1}).
    """
    lx = float(nx) if lx is None else lx
    ly = float(ny) if ly is None else ly
    xs = np.arange(nx + 1, dtype=np.float64)
    ys = np.arange(ny + 1, dtype=np.float64)
    X, Y = np.meshgrid(xs, ys)                      # (ny+1, nx+1), row = y
    if jitter > 0 and nx > 1 and ny > 1:
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter, jitter, size=(ny - 1, nx - 1, 2))
        X[1:-1, 1:-1] += d[..., 0]
        Y[1:-1, 1:-1] += d[..., 1]
    points = np.stack([X.reshape(-1) * (lx / nx), Y.reshape(-1) * (ly / ny)], axis=1)

    def pid(i, j):
        return (j * (nx + 1) + i).astype(np.int32)

    I, J = np.meshgrid(np.arange(nx), np.arange(ny))
    I = I.reshape(-1)
    J = J.reshape(-1)
    p00, p10, p11, p01 = pid(I, J), pid(I + 1, J), pid(I + 1, J + 1), pid(I, J + 1)
    even = ((I + J) % 2) == 0
    # even quads: diagonal p00-p11 ; odd quads: diagonal p10-p01 ; both CCW
    t0 = np.where(even[:, None], np.stack([p00, p10, p11], 1), np.stack([p00, p10, p01], 1))
    t1 = np.where(even[:, None], np.stack([p00, p11, p01], 1), np.stack([p10, p11, p01], 1))
    elements = np.empty((2 * nx * ny, 3), dtype=np.int32)
    elements[0::2] = t0
    elements[1::2] = t1

    markers = np.zeros((ny + 1, nx + 1), dtype=np.int32)
    if not periodic_x:
        markers[:, 0] = LEFT
        markers[:, -1] = RIGHT
    markers[0, :] = BOTTOM
    markers[-1, :] = TOP
    alias = None
    if periodic_x:
        ids = np.arange((nx + 1) * (ny + 1), dtype=np.int32).reshape(ny + 1, nx + 1)
        ids[:, -1] = ids[:, 0]
        alias = ids.reshape(-1)
    # with_faces=False: points + elements only (what the window-based decomposition needs from a global mesh)
    faces = unique_edges(elements, points.shape[0], alias) if with_faces else np.zeros((0, 2), dtype=np.int32)
    return RawMesh(points, elements, faces, markers.reshape(-1), alias, (nx, ny))


def masked_domain(nx: int, ny: int, lx: float, ly: float, inside_obstacle, jitter: float = 0.15,
                  seed: int = 0) -> RawMesh:
    """Triangulated ``lx x ly`` box with the cells whose centroid satisfies
    ``inside_obstacle(x, y)`` removed; nodes on the obstacle outline get marker 5 (the cylinder /
    porous-obstacle marker of tests/flow_over_cyl.ipynb c5 and tests/porous_flow.ipynb).
    Unreferenced points are dropped and ids compacted.
    """
    m = triangulated_square(nx, ny, jitter=jitter, seed=seed, lx=lx, ly=ly)
    cen = m.points[m.elements].mean(axis=1)
    keep = ~np.asarray(inside_obstacle(cen[:, 0], cen[:, 1]), dtype=bool)
    el = m.elements[keep]
    used = np.zeros(m.points.shape[0], dtype=bool)
    used[el.reshape(-1)] = True
    remap = -np.ones(m.points.shape[0], dtype=np.int32)
    remap[used] = np.arange(int(used.sum()), dtype=np.int32)
    el = remap[el]
    pts = m.points[used]
    mk = m.point_markers[used].copy()
    faces = unique_edges(el, pts.shape[0])
    # boundary edges = edges with exactly one adjacent cell; their unmarked nodes are obstacle nodes
    a = el.reshape(-1)
    b = np.roll(el, -1, axis=1).reshape(-1)
    key = np.minimum(a, b).astype(np.int64) * pts.shape[0] + np.maximum(a, b)
    uk, cnt = np.unique(key, return_counts=True)
    bkeys = uk[cnt == 1]
    bn = np.unique(np.concatenate([bkeys // pts.shape[0], bkeys % pts.shape[0]]))
    obst = bn[mk[bn] == 0]
    mk[obst] = OBSTACLE
    return RawMesh(pts, el.astype(np.int32), faces, mk, None, None)


def cylinder_channel(scale: int = 1, seed: int = 0) -> RawMesh:
    """Flow-over-cylinder domain of tests/flow_over_cyl.ipynb c4-c7: box (0,0)-(60,20), circle r=1
    at (10,10).  ``scale=1`` -> 60x20 quads (2.4k cells); ``scale=9`` ~ 194k cells (config 2)."""
    nx, ny = 60 * scale, 20 * scale
    return masked_domain(nx, ny, 60.0, 20.0, lambda x, y: (x - 10.0) ** 2 + (y - 10.0) ** 2 < 1.0,
                         seed=seed)


def points_in_polygon(px: np.ndarray, py: np.ndarray, poly: np.ndarray) -> np.ndarray:
    """Even-odd (ray casting) test of many points against one closed polygon ``poly (V,2)``.
    Only points inside the polygon's bounding box are tested; vectorised over points, looped over
    the (few dozen) polygon edges."""
    out = np.zeros(px.shape, dtype=bool)
    x0, y0 = poly.min(axis=0)
    x1, y1 = poly.max(axis=0)
    cand = np.nonzero((px >= x0) & (px <= x1) & (py >= y0) & (py <= y1))[0]
    if cand.size == 0:
        return out
    cx, cy = px[cand], py[cand]
    inside = np.zeros(cand.size, dtype=bool)
    ax, ay = poly[:, 0], poly[:, 1]
    bx, by = np.roll(ax, -1), np.roll(ay, -1)
    for i in range(poly.shape[0]):
        if ay[i] == by[i]:
            continue
        crosses = (ay[i] > cy) != (by[i] > cy)
        xi = ax[i] + (cy - ay[i]) * (bx[i] - ax[i]) / (by[i] - ay[i])
        inside ^= crosses & (cx < xi)
    out[cand] = inside
    return out


def porous_outlines():
    """The 60 obstacle outlines of the reference's porous case (tests/test_bmp.mat as loaded by
    tests/porous_flow.ipynb c7), one (V,2) float array per obstacle id, in the file's point order.
    Data file written by oracle/make_porous_outlines.py."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "porous_outlines.npz"))
    x, y, oid = d["x"].astype(np.float64), d["y"].astype(np.float64), d["id"]
    return [np.stack([x[oid == i], y[oid == i]], axis=1) for i in np.unique(oid)]


def porous_channel(scale: float = 1, seed: int = 0) -> RawMesh:
    """Porous-flow domain of tests/porous_flow.ipynb (c7-c13): box (-93,0)-(279,186) minus the union
    of the 60 obstacle polygons of tests/test_bmp.mat.  The notebook hands the outline to Triangle;
    here (no Triangle) cells of a jittered structured triangulation whose centroid lies inside any
    polygon are removed and the exposed nodes get the obstacle marker 5, as in c13
    (``facet_markers=5``).  ``scale=1`` -> 186x93 quads, 27.5 k cells (porosity 0.795); ``scale=8.5`` -> 1.99 M cells =
    BASELINE.json configs[2]."""
    lx, ly = 372.0, 186.0
    nx, ny = int(round(186 * scale)), int(round(93 * scale))
    polys = porous_outlines()

    def inside(x, y):
        xs = x - 93.0                       # generator frame [0,372] -> notebook frame [-93,279]
        out = np.zeros(x.shape, dtype=bool)
        for poly in polys:
            out |= points_in_polygon(xs, y, poly)
        return out
    m = masked_domain(nx, ny, lx, ly, inside, seed=seed)
    m.points[:, 0] -= 93.0
    return m


def porous_boundary_conditions(mesher, nodes, rho_in: float = 1.05, rho_out: float = 0.95):
    """Boundary conditions of tests/porous_flow.ipynb c26: obstacles (marker 5) and the two side walls
    no-slip, density inlet / outlet on the remaining two box sides.  The notebook numbers the box
    sides 1..4 in the order shapely returns the exterior ring (not reproducible here); the physical
    reading is used: walls = bottom/top, inlet = left, outlet = right."""
    for mk in (OBSTACLE, BOTTOM, TOP):
        nodes = mesher.set_vel_node(nodes, mk, np.array([0.0, 0.0]))
    nodes = mesher.set_rho_node(nodes, LEFT, rho_in)
    nodes = mesher.set_rho_node(nodes, RIGHT, rho_out)
    return nodes


# ---------------------------------------------------------------------------------------------------
# scalable strip decomposition of the synthetic square (multi-GPU weak / strong scaling)
# ---------------------------------------------------------------------------------------------------
def _hash_uniform(idx: np.ndarray, seed: int, salt: int) -> np.ndarray:
    """Counter-based U[0,1): splitmix64 of the global vertex index, so every rank reproduces the
    same jitter for the vertices it sees without generating the whole grid."""
    with np.errstate(over="ignore"):
        z = idx.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15) * np.uint64(2 * seed + salt + 1)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def strip_window(nx: int, ny_total: int, row0: int, row1: int, jitter: float = 0.2, seed: int = 0,
                 periodic_x: bool = True) -> RawMesh:
    """Quad rows [row0,row1) of the global ``nx x ny_total`` triangulated square (hash jitter, so the
    window is bit-identical to the same rows of the whole mesh).  ``cell_gid`` numbers cells as in
    the whole mesh (2*(J*nx+I)+t); ``point_gid`` likewise for vertices."""
    rows = row1 - row0
    xs = np.arange(nx + 1, dtype=np.int64)
    ys = np.arange(row0, row1 + 1, dtype=np.int64)
    GX, GY = np.meshgrid(xs, ys)
    vid = GY * (nx + 1) + GX
    X = GX.astype(np.float64)
    Y = GY.astype(np.float64)
    interior = (GX > 0) & (GX < nx) & (GY > 0) & (GY < ny_total)
    if jitter > 0:
        X = X + np.where(interior, (2 * _hash_uniform(vid, seed, 0) - 1) * jitter, 0.0)
        Y = Y + np.where(interior, (2 * _hash_uniform(vid, seed, 1) - 1) * jitter, 0.0)
    points = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)

    def pid(i, j):
        return (j * (nx + 1) + i).astype(np.int32)

    I, J = np.meshgrid(np.arange(nx), np.arange(rows))
    I = I.reshape(-1)
    J = J.reshape(-1)
    p00, p10, p11, p01 = pid(I, J), pid(I + 1, J), pid(I + 1, J + 1), pid(I, J + 1)
    even = ((I + J + row0) % 2) == 0
    t0 = np.where(even[:, None], np.stack([p00, p10, p11], 1), np.stack([p00, p10, p01], 1))
    t1 = np.where(even[:, None], np.stack([p00, p11, p01], 1), np.stack([p10, p11, p01], 1))
    elements = np.empty((2 * nx * rows, 3), dtype=np.int32)
    elements[0::2] = t0
    elements[1::2] = t1
    markers = np.zeros((rows + 1, nx + 1), dtype=np.int32)
    if not periodic_x:
        markers[:, 0] = LEFT
        markers[:, -1] = RIGHT
    if row0 == 0:
        markers[0, :] = BOTTOM
    if row1 == ny_total:
        markers[-1, :] = TOP
    alias = None
    if periodic_x:
        ids = np.arange((nx + 1) * (rows + 1), dtype=np.int32).reshape(rows + 1, nx + 1)
        ids[:, -1] = ids[:, 0]
        alias = ids.reshape(-1)
    faces = unique_edges(elements, points.shape[0], alias)
    m = RawMesh(points, elements, faces, markers.reshape(-1), alias, (nx, rows))
    gq = (J + row0).astype(np.int64) * nx + I
    gid = np.empty(2 * nx * rows, dtype=np.int64)
    gid[0::2] = 2 * gq
    gid[1::2] = 2 * gq + 1
    m.cell_gid = gid
    m.cell_row = np.repeat(J + row0, 2)
    m.point_gid = vid.reshape(-1)
    return m


# ---------------------------------------------------------------------------------------------------
# the hand-built structured cavity of tests/ldcFVDBM.ipynb (K = 4 quads), vectorised
# ---------------------------------------------------------------------------------------------------
def quad_cavity(nx: int, ny: int, dynamics, u_lid: float, flux_scheme: str = "upwind"):
    """Containers (cells, faces, nodes) of the lid-driven cavity exactly as tests/ldcFVDBM.ipynb
    c4-c9 builds them with Environment.create + CustomArray.add_items: unit quads, faces numbered
    vertical-first, cell faces [W, S, E, N] with signs [-1,-1,+1,+1], ghosts in stencil slot 0
    (west/south walls) and slot 1 (east/north walls), all boundary nodes type 1, the lid is the
    row y=0 with vel (u_lid, 0) except its two corners."""
    from .containers import Cells, Faces, Nodes
    N, F, P = nx * ny, nx * (ny + 1) + (nx + 1) * ny, (nx + 1) * (ny + 1)
    cell = np.arange(N).reshape(ny, nx)
    node = np.arange(P).reshape(ny + 1, nx + 1)
    vert = np.arange(ny * (nx + 1)).reshape(ny, nx + 1)
    horz = (ny * (nx + 1) + np.arange((ny + 1) * nx)).reshape(ny + 1, nx)
    cells, faces, nodes = Cells(N, dynamics), Faces(F, dynamics, flux_scheme=flux_scheme), Nodes(P, dynamics)
    Y, X = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    cells.face_indices = np.stack([vert[Y, X], horz[Y, X], vert[Y, X + 1], horz[Y + 1, X]], axis=-1).reshape(N, 4).astype(np.int32)
    cells.face_normals = np.tile(np.array([-1, -1, 1, 1], dtype=np.int32), (N, 1))
    cells.centers = np.stack([X.reshape(-1) + 0.5, Y.reshape(-1) + 0.5], axis=1)
    # faces
    st = -np.ones((F, 2), dtype=np.int32)
    ni = np.zeros((F, 2), dtype=np.int32)
    n = np.zeros((F, 2))
    yv, xv = np.meshgrid(np.arange(ny), np.arange(nx + 1), indexing="ij")
    fv = vert[yv, xv].reshape(-1)
    st[fv, 0] = np.where(xv > 0, cell[yv, np.maximum(xv - 1, 0)], -1).reshape(-1)
    st[fv, 1] = np.where(xv < nx, cell[yv, np.minimum(xv, nx - 1)], -1).reshape(-1)
    ni[fv, 0], ni[fv, 1] = node[yv, xv].reshape(-1), node[yv + 1, xv].reshape(-1)
    n[fv] = (1.0, 0.0)
    yh, xh = np.meshgrid(np.arange(ny + 1), np.arange(nx), indexing="ij")
    fh = horz[yh, xh].reshape(-1)
    st[fh, 0] = np.where(yh > 0, cell[np.maximum(yh - 1, 0), xh], -1).reshape(-1)
    st[fh, 1] = np.where(yh < ny, cell[np.minimum(yh, ny - 1), xh], -1).reshape(-1)
    ni[fh, 0], ni[fh, 1] = node[yh, xh].reshape(-1), node[yh, xh + 1].reshape(-1)
    n[fh] = (0.0, 1.0)
    faces.stencil_cells_index, faces.nodes_index, faces.n = st, ni, n
    faces.stencil_dists = np.full((F, 2), 0.5)
    faces.L = np.ones((F, 1))
    # nodes: ring order (y-1,x-1), (y-1,x), (y,x), (y,x-1), missing ones dropped, padded with -1
    yn, xn = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    ring = -np.ones((P, 4), dtype=np.int32)
    dist = -np.ones((P, 4))
    fill = np.zeros(P, dtype=np.int64)
    for dy, dx in ((-1, -1), (-1, 0), (0, 0), (0, -1)):
        yy, xx = (yn + dy).reshape(-1), (xn + dx).reshape(-1)
        ok = (yy >= 0) & (yy < ny) & (xx >= 0) & (xx < nx)
        p = np.nonzero(ok)[0]
        ring[p, fill[p]] = cell[yy[p], xx[p]]
        dist[p, fill[p]] = np.sqrt(2.0)
        fill[p] += 1
    boundary = ((yn == 0) | (yn == ny) | (xn == 0) | (xn == nx)).reshape(-1)
    lid = ((yn == 0) & (xn > 0) & (xn < nx)).reshape(-1)
    nodes.cells_index, nodes.cell_dists = ring, dist
    nodes.type = boundary.astype(np.int32).reshape(P, 1)
    vel = np.zeros((P, 2))
    vel[lid, 0] = u_lid
    nodes.vel = vel
    return cells, faces, nodes
