"""Domain decomposition for multi-GPU runs: cell partitioning, local meshes with a halo, exchange maps.

The reference is single-device; this is the spatial decomposition north_star adds.
Everything here is host-side NumPy on the reference's array layout and is unit-tested without a
GPU (tests/test_partition.py: halo maps against a pure-Python set construction).

Per rank r the local mesh consists of
  * owned cells   (part == r), local ids [0, No)            -- updated by this rank
  * halo cells    local ids [No, N_loc), sorted by (owner rank, global id) -- read-only copies:
       (A) face neighbours of owned cells, and
       (B) ring cells of the *active* (type 1/2) nodes on boundary faces of owned cells
           (the node kernel needs every ring cell; vertex neighbours, not only face neighbours)
  * faces         every face of an owned cell (cut faces are computed redundantly on both ranks,
                  no flux exchange: deterministic, one message per neighbour per step)
  * nodes         every endpoint of those faces; active nodes whose ring is incomplete on this rank
                  (not an endpoint of an owned cell's boundary face) are demoted to type 0 locally.
Only populations travel (Q reals per halo cell per step); rho/vel of halo cells are recomputed.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .reorder import hilbert_index, hilbert_perm, order_to_perm

__all__ = ["GlobalMesh", "LocalMesh", "partition_sfc", "partition_strips", "partition_rcm", "partition_cells",
           "refine_partition", "sfc_owner_from_raw", "window_from_raw", "local_from_raw",
           "extract_local", "exchange_lists", "edge_cut"]


@dataclass
class GlobalMesh:
    """Reference-layout arrays of one whole mesh (what Mesher.to_env + BC setters produce)."""
    face_indices: np.ndarray        # (N,K) i32   Cells.face_indices
    face_signs: np.ndarray          # (N,K) i32   Cells.face_normals
    stencil: np.ndarray             # (F,2) i32   Faces.stencil_cells_index
    stencil_dists: np.ndarray       # (F,2)
    nodes_index: np.ndarray         # (F,2) i32
    n: np.ndarray                   # (F,2)
    L: np.ndarray                   # (F,1)
    node_type: np.ndarray           # (P,1) i32
    ring: np.ndarray                # (P,M) i32
    ring_dists: np.ndarray          # (P,M)
    cell_pdf: np.ndarray            # (N,Q)
    node_pdf: np.ndarray            # (P,Q)
    node_rho: np.ndarray            # (P,1)
    node_vel: np.ndarray            # (P,2)
    centers: Optional[np.ndarray] = None   # (N,2) for locality ordering / geometric partitioning
    cell_gid: Optional[np.ndarray] = None  # (N,) global ids when this "global" mesh is itself a window

    @classmethod
    def from_containers(cls, cells, faces, nodes):
        a = np.asarray
        return cls(a(cells.face_indices, np.int32), a(cells.face_normals, np.int32),
                   a(faces.stencil_cells_index, np.int32), a(faces.stencil_dists), a(faces.nodes_index, np.int32),
                   a(faces.n), a(faces.L).reshape(-1, 1), a(nodes.type, np.int32).reshape(-1, 1),
                   a(nodes.cells_index, np.int32), a(nodes.cell_dists), a(cells.pdf), a(nodes.pdf),
                   a(nodes.rho).reshape(-1, 1), a(nodes.vel), getattr(cells, "centers", None))

    @property
    def num_cells(self):
        return self.face_indices.shape[0]


@dataclass
class LocalMesh:
    rank: int
    n_owned: int
    cell_gid: np.ndarray            # (N_loc,) global id of every local cell (owned first)
    halo_owner: np.ndarray          # (N_loc - n_owned,) owning rank of each halo cell
    face_gid: np.ndarray            # (F_loc,)
    node_gid: np.ndarray            # (P_loc,)
    node_complete: np.ndarray       # (P_loc,) bool: this rank computes the node's BC values
    mesh: GlobalMesh = None         # local arrays in reference layout (local numbering)
    perm: Optional[np.ndarray] = None   # storage permutation keeping owned cells in [0, n_owned)

    @property
    def n_local(self):
        return self.cell_gid.shape[0]


# ---------------------------------------------------------------------------------------------------
# partitioners
# ---------------------------------------------------------------------------------------------------
def partition_sfc(centers: np.ndarray, nparts: int) -> np.ndarray:
    """Balanced chunks of the Hilbert-curve ordering of the cell centroids."""
    n = centers.shape[0]
    perm = hilbert_perm(centers)                         # rank along the curve
    return (perm.astype(np.int64) * nparts // n).astype(np.int32)


def partition_strips(centers: np.ndarray, nparts: int, axis: int = 1) -> np.ndarray:
    """Balanced strips along one coordinate (1-D decomposition)."""
    n = centers.shape[0]
    order = np.argsort(centers[:, axis], kind="stable")
    part = np.empty(n, dtype=np.int32)
    part[order] = (np.arange(n, dtype=np.int64) * nparts // n).astype(np.int32)
    return part


def partition_rcm(stencil: np.ndarray, n_cells: int, nparts: int) -> np.ndarray:
    """Graph-only partitioning (no coordinates needed): balanced chunks of the reverse Cuthill-McKee
    ordering, i.e. contiguous bands of BFS levels of the cell adjacency graph."""
    from .reorder import rcm_perm
    perm = rcm_perm(stencil, n_cells)
    return (perm.astype(np.int64) * nparts // n_cells).astype(np.int32)


def partition_cells(stencil: np.ndarray, n_cells: int, nparts: int, centers=None, method: str = "auto",
                    refine: bool = True) -> np.ndarray:
    """Front door: 'sfc' (Hilbert chunks, needs centroids), 'rcm' (graph only), 'strips', or 'auto'
    (sfc when centroids are known, else rcm); optionally followed by the METIS-style greedy
    boundary refinement (edge-cut reduction under a 3 % balance constraint)."""
    if method == "auto":
        method = "sfc" if centers is not None else "rcm"
    if method == "sfc":
        part = partition_sfc(np.asarray(centers), nparts)
    elif method == "strips":
        part = partition_strips(np.asarray(centers), nparts)
    elif method == "rcm":
        part = partition_rcm(stencil, n_cells, nparts)
    else:
        raise ValueError(f"unknown partition method {method!r}")
    if refine and nparts > 1:
        part = refine_partition(stencil, part, nparts)
    return part


def edge_cut(stencil: np.ndarray, part: np.ndarray) -> int:
    m = (stencil[:, 0] >= 0) & (stencil[:, 1] >= 0)
    return int(np.count_nonzero(part[stencil[m, 0]] != part[stencil[m, 1]]))


def refine_partition(stencil: np.ndarray, part: np.ndarray, nparts: int, sweeps: int = 8,
                     imbalance: float = 1.03) -> np.ndarray:
    """METIS-style boundary refinement, fully vectorised (no per-cell Python loop): per sweep every
    boundary cell computes the gain of moving to the neighbouring part that holds most of its face
    neighbours; the candidates that are strict local maxima of (gain, -id) among their face neighbours
    form an independent set, so all of them move at once with their gains still valid; per target part
    the moves are admitted in order of decreasing gain up to ``imbalance`` x the mean size."""
    part = np.array(part, dtype=np.int32, copy=True)
    n = part.shape[0]
    m = (stencil[:, 0] >= 0) & (stencil[:, 1] >= 0)
    a = np.concatenate([stencil[m, 0], stencil[m, 1]]).astype(np.int64)
    b = np.concatenate([stencil[m, 1], stencil[m, 0]]).astype(np.int64)
    cap = int(imbalance * n / nparts) + 1
    for _ in range(sweeps):
        pa, pb = part[a], part[b]
        cut = pa != pb
        if not cut.any():
            break
        # neighbours per (boundary cell, foreign part) and per (boundary cell, own part)
        bc = np.unique(a[cut])
        on_b = np.zeros(n, dtype=bool)
        on_b[bc] = True
        sel = on_b[a]
        key = a[sel] * nparts + pb[sel]
        uk, cnt = np.unique(key, return_counts=True)
        cell, prt = uk // nparts, (uk % nparts).astype(np.int32)
        own = prt == part[cell]
        own_cnt = np.zeros(n, dtype=np.int64)
        own_cnt[cell[own]] = cnt[own]
        gain = cnt - own_cnt[cell]
        cand = (~own) & (gain > 0)
        if not cand.any():
            break
        order = np.lexsort((-gain[cand], cell[cand]))                 # best target per cell first
        c_s, p_s, g_s = cell[cand][order], prt[cand][order], gain[cand][order]
        first = np.ones(c_s.shape[0], dtype=bool)
        first[1:] = c_s[1:] != c_s[:-1]
        c_s, p_s, g_s = c_s[first], p_s[first], g_s[first]
        # independent set: drop a candidate if a face neighbour is a stronger candidate
        g_of = np.zeros(n, dtype=np.int64)
        g_of[c_s] = g_s
        ga, gb = g_of[a], g_of[b]
        weaker = (gb > ga) | ((gb == ga) & (gb > 0) & (b < a))
        blocked = np.bincount(a[weaker], minlength=n) > 0
        keep = ~blocked[c_s]
        c_m, p_m, g_m = c_s[keep], p_s[keep], g_s[keep]
        if c_m.size == 0:
            break
        # capacity: per target part admit the highest gains while it stays below the cap (departures ignored)
        sizes = np.bincount(part, minlength=nparts).astype(np.int64)
        o = np.lexsort((c_m, -g_m, p_m))
        c_m, p_m = c_m[o], p_m[o]
        start = np.searchsorted(p_m, np.arange(nparts), side="left")
        rank_in = np.arange(c_m.size) - start[p_m]
        ok = rank_in < (cap - sizes)[p_m]
        if not ok.any():
            break
        part[c_m[ok]] = p_m[ok]
    return part


# ---------------------------------------------------------------------------------------------------
# local mesh extraction
# ---------------------------------------------------------------------------------------------------
def extract_local(g: GlobalMesh, part: np.ndarray, rank: int, reorder: bool = True) -> LocalMesh:
    """Build rank ``rank``'s local mesh (owned | halo) from a global description."""
    N, K = g.face_indices.shape
    gid_of = g.cell_gid if g.cell_gid is not None else np.arange(N, dtype=np.int64)
    owned = np.nonzero(part == rank)[0]
    No = owned.shape[0]
    is_owned = np.zeros(N, dtype=bool)
    is_owned[owned] = True

    def unique_ids(ids, size):
        """Sorted unique non-negative ids < size: a mark-and-scan (O(n + size)) instead of np.unique's sort."""
        mark = np.zeros(size, dtype=bool)
        mark[ids[ids >= 0]] = True
        return np.nonzero(mark)[0]

    F_all, P_all = g.stencil.shape[0], g.node_type.shape[0]
    faces_l = unique_ids(g.face_indices[owned].reshape(-1), F_all)
    st = g.stencil[faces_l]
    nbr = st.reshape(-1)
    nbr = nbr[nbr >= 0]
    halo_a = nbr[~is_owned[nbr]]
    # (B) rings of active nodes on boundary faces of owned cells
    bfaces = faces_l[(st[:, 0] < 0) | (st[:, 1] < 0)]
    bnodes = unique_ids(g.nodes_index[bfaces].reshape(-1), P_all)
    active = bnodes[g.node_type[bnodes, 0] != 0]
    ring = g.ring[active]
    ring_ok = (ring >= 0) & (g.ring_dists[active] > 0)
    rc = ring[ring_ok]
    halo_b = rc[~is_owned[rc]]
    halo = unique_ids(np.concatenate([halo_a, halo_b]), N)
    hkey = np.lexsort((gid_of[halo], part[halo]))
    halo = halo[hkey]
    cells_l = np.concatenate([owned, halo])
    lid = -np.ones(N, dtype=np.int64)
    lid[cells_l] = np.arange(cells_l.shape[0])

    nodes_l = unique_ids(g.nodes_index[faces_l].reshape(-1), P_all)
    P = P_all
    nlid = -np.ones(P, dtype=np.int64)
    nlid[nodes_l] = np.arange(nodes_l.shape[0])
    F = g.stencil.shape[0]
    flid = -np.ones(F, dtype=np.int64)
    flid[faces_l] = np.arange(faces_l.shape[0])

    Nl = cells_l.shape[0]
    fi = np.zeros((Nl, K), dtype=np.int32)
    fs = np.ones((Nl, K), dtype=np.int32)
    fi[:No] = flid[g.face_indices[owned]]
    fs[:No] = g.face_signs[owned]
    st_l = np.where(st >= 0, lid[np.maximum(st, 0)], -1).astype(np.int32)
    # a stencil cell that is neither owned nor halo can only belong to a face no owned cell lists
    ni = g.nodes_index[faces_l]
    ni_l = np.where(ni >= 0, nlid[np.maximum(ni, 0)], -1).astype(np.int32)

    complete = np.zeros(nodes_l.shape[0], dtype=bool)
    complete[nlid[active]] = True
    ntype = g.node_type[nodes_l].copy()
    ntype[~complete] = 0
    ring_l = g.ring[nodes_l]
    ring_d = np.array(g.ring_dists[nodes_l], copy=True)
    ring_loc = np.where(ring_l >= 0, lid[np.maximum(ring_l, 0)], -1)
    lost = (ring_l >= 0) & (ring_loc < 0)
    ring_d[lost] = -1
    ring_loc = ring_loc.astype(np.int32)

    local = GlobalMesh(fi, fs, st_l, g.stencil_dists[faces_l], ni_l, g.n[faces_l], g.L[faces_l], ntype, ring_loc, ring_d,
                       g.cell_pdf[cells_l], g.node_pdf[nodes_l], g.node_rho[nodes_l], g.node_vel[nodes_l],
                       None if g.centers is None else g.centers[cells_l], gid_of[cells_l])
    perm = None
    if reorder and g.centers is not None and No > 0:
        p_owned = hilbert_perm(g.centers[owned])
        perm = np.concatenate([p_owned, np.arange(No, Nl, dtype=np.int32)]).astype(np.int32)
    return LocalMesh(rank, No, gid_of[cells_l].astype(np.int64), part[halo].astype(np.int32), faces_l.astype(np.int64),
                     nodes_l.astype(np.int64), complete, local, perm)


# ---------------------------------------------------------------------------------------------------
# scalable path: a rank's local mesh from a WINDOW of the raw mesh (nobody builds the global mesh)
# ---------------------------------------------------------------------------------------------------
def sfc_owner_from_raw(points: np.ndarray, elements: np.ndarray, nparts: int, bits: int = 16) -> np.ndarray:
    """Owner rank of every cell = balanced chunks of the Hilbert curve through the cell centroids, computed
    from the raw mesh (points, elements): O(N) keys (native OpenMP helper fvdbm_sfc_keys) + an O(N) selection
    of the nparts-1 splitters (np.partition) instead of a global sort.  Deterministic, identical on every rank."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    n = elements.shape[0]
    lo = pts.min(axis=0)
    scale = ((1 << bits) - 1) / max(float((pts.max(axis=0) - lo).max()), 1e-300)
    keys = np.empty(n, dtype=np.int64)
    from . import _lib
    el = np.ascontiguousarray(elements, dtype=np.int32)
    _lib.check(_lib.load().fvdbm_sfc_keys(pts.ctypes.data, el.ctypes.data, n, int(el.shape[1]), bits, float(lo[0]), float(lo[1]),
                                          float(scale), keys.ctypes.data))       # host-only C/OpenMP helper (no GPU)
    # ties (equal keys) are broken by cell id so that the chunks are exactly balanced
    keys = keys * np.int64(n) + np.arange(n, dtype=np.int64) if n < (1 << 31) else keys
    cuts = [(n * r) // nparts for r in range(1, nparts)]
    if not cuts:
        return np.zeros(n, dtype=np.int32)
    split = np.partition(keys, cuts)[cuts]
    return np.searchsorted(split, keys, side="right").astype(np.int32)


def window_from_raw(raw, owner: np.ndarray, rank: int, chunk: int = 1 << 22):
    """Raw sub-mesh a rank needs: its owned cells plus every cell sharing a (canonical) vertex with one of
    them -- a superset of the face neighbours and of the rings of the boundary nodes of owned cells.  Window
    cells keep ascending global ids, so shared faces get the same orientation / stencil order in every rank's
    window and cut faces are evaluated from bit-identical coefficients on both sides."""
    from .meshgen import RawMesh, unique_edges
    el = np.asarray(raw.elements)
    alias = getattr(raw, "point_alias", None)
    canon = (lambda ids: ids) if alias is None else (lambda ids: alias[ids])
    P = raw.points.shape[0]
    vmark = np.zeros(P, dtype=bool)
    owned = np.nonzero(owner == rank)[0]
    vmark[canon(el[owned]).reshape(-1)] = True
    touch = []
    for c0 in range(0, el.shape[0], chunk):
        t = vmark[canon(el[c0:c0 + chunk])].any(axis=1)
        touch.append(np.nonzero(t)[0] + c0)
    window = np.concatenate(touch) if touch else np.zeros(0, dtype=np.int64)
    el_w = el[window]
    used_ids = el_w.reshape(-1) if alias is None else np.concatenate([el_w.reshape(-1), alias[el_w.reshape(-1)]])
    used = np.zeros(P, dtype=bool)
    used[used_ids] = True
    remap = -np.ones(P, dtype=np.int64)
    pid = np.nonzero(used)[0]
    remap[pid] = np.arange(pid.shape[0])
    el_l = remap[el_w].astype(np.int32)
    alias_l = None if alias is None else remap[alias[pid]].astype(np.int32)
    faces = unique_edges(el_l, pid.shape[0], alias_l)
    w = RawMesh(np.asarray(raw.points)[pid], el_l, faces, np.asarray(raw.point_markers)[pid], alias_l, None)
    w.cell_gid = window.astype(np.int64)
    w.point_gid = pid.astype(np.int64)
    return w


def local_from_raw(raw, rank: int, nparts: int, dynamics, scheme: str, boundary_conditions=None, initial_pdf=None,
                   owner: Optional[np.ndarray] = None, dim_multiplier=1):
    """General, scalable decomposition (north_star: "the Mesher partitions cells ... across the 8 B200s"):
    owner ranks from the raw mesh (Hilbert chunks unless given), then THIS rank's window -> Mesher -> containers
    -> ``extract_local``.  Cost per rank is O(N) cheap passes over the raw arrays plus the Mesher on
    ~N/nparts cells; no rank ever holds the global connectivity.

    ``boundary_conditions(mesher, nodes) -> nodes`` applies the BC setters, ``initial_pdf(mesher) -> (n,Q)``
    the start populations (both see only the window).  Returns (LocalMesh, faces_per_owned_cell)."""
    from .mesher import Mesher
    if owner is None:
        owner = sfc_owner_from_raw(raw.points, raw.elements, nparts)
    w = window_from_raw(raw, owner, rank)
    m = Mesher()
    m.import_meshpy(w)
    m.calc_mesh_properties()
    cells, faces, nodes = m.to_env(dynamics, flux_method=scheme, dim_multiplier=dim_multiplier)
    if boundary_conditions is not None:
        nodes = boundary_conditions(m, nodes)
    if initial_pdf is not None:
        cells.pdf = initial_pdf(m)
    g = GlobalMesh.from_containers(cells, faces, nodes)
    g.cell_gid = w.cell_gid
    part_w = owner[w.cell_gid]
    local = extract_local(g, part_w, rank, reorder=True)
    owned_faces = np.unique(g.face_indices[part_w == rank].reshape(-1)).size
    return local, owned_faces / max(1, local.n_owned)


def exchange_lists(local: LocalMesh, requests_from_peers: Dict[int, np.ndarray]):
    """Turn the peers' requests (sorted global ids they need from this rank) into send lists of
    local cell ids, and this rank's halo into recv lists.  Both are ordered by (peer, global id), so
    the message from r to s is consumed by s without any reordering.

    Returns (peers_send, send_cells, send_counts, peers_recv, recv_cells, recv_counts)."""
    No = local.n_owned
    gid_owned = local.cell_gid[:No]
    order = np.argsort(gid_owned, kind="stable")
    sorted_gid = gid_owned[order]
    peers_send, send_cells, send_counts = [], [], []
    for peer in sorted(requests_from_peers):
        req = np.asarray(requests_from_peers[peer], dtype=np.int64)
        if req.size == 0:
            continue
        idx = np.searchsorted(sorted_gid, req)
        if np.any(idx >= sorted_gid.size) or np.any(sorted_gid[np.minimum(idx, sorted_gid.size - 1)] != req):
            raise ValueError(f"rank {local.rank}: peer {peer} requested cells this rank does not own")
        peers_send.append(peer)
        send_cells.append(order[idx].astype(np.int32))
        send_counts.append(req.size)
    peers_recv, recv_cells, recv_counts = [], [], []
    for peer in np.unique(local.halo_owner):
        sel = np.nonzero(local.halo_owner == peer)[0]           # already sorted by global id
        peers_recv.append(int(peer))
        recv_cells.append((No + sel).astype(np.int32))
        recv_counts.append(sel.size)
    cat = lambda xs: np.concatenate(xs).astype(np.int32) if xs else np.zeros(0, dtype=np.int32)
    return peers_send, cat(send_cells), send_counts, peers_recv, cat(recv_cells), recv_counts


def halo_requests(local: LocalMesh) -> Dict[int, np.ndarray]:
    """{owner rank: sorted global ids of the halo cells this rank needs from it}."""
    out = {}
    gids = local.cell_gid[local.n_owned:]
    for peer in np.unique(local.halo_owner):
        out[int(peer)] = gids[local.halo_owner == peer]
    return out
