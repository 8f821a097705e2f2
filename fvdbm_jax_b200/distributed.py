"""Multi-GPU stepping: one engine (C-ABI handle) per rank, one-layer halo exchange per iteration.

Per iteration and rank:

    pack owned border cells -> send buffer            (engine stream)
    NCCL send/recv with every neighbour rank          (NCCL stream, ordered after the pack)
    phase 0: update interior cells                    (engine stream, overlaps the transfer)
    wait for the transfer, unpack into the halo cells
    phase 1: node kernel + border cells, swap buffers

No data-path collective besides the neighbour send/recv; cut faces are evaluated redundantly on
both ranks from identical inputs, so a k-rank run is bit-identical to the 1-rank run
(tests/test_gpu_multi.py).  ``InProcessCluster`` runs all ranks inside one process on one device
(the exchange becomes device copies) so the halo logic is testable on a single GPU; the
``DistributedEnvironment`` is the torch.distributed (NCCL) version used by bench.py under torchrun.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from . import _lib
from .containers import Cells, Faces, Nodes
from .environment import Environment
from .partition import GlobalMesh, LocalMesh, exchange_lists, extract_local, halo_requests

__all__ = ["containers_from_mesh", "gather_requests", "HaloComm", "RankEngine", "InProcessCluster", "DistributedEnvironment", "strip_local_mesh"]


def containers_from_mesh(g: GlobalMesh, dynamics, scheme: str):
    N, F, P = g.face_indices.shape[0], g.stencil.shape[0], g.node_type.shape[0]
    cells, faces, nodes = Cells(N, dynamics), Faces(F, dynamics, flux_scheme=scheme), Nodes(P, dynamics)
    cells.face_indices, cells.face_normals, cells.pdf = g.face_indices, g.face_signs, g.cell_pdf
    if g.centers is not None:
        cells.centers = g.centers
    faces.nodes_index, faces.stencil_cells_index, faces.stencil_dists = g.nodes_index, g.stencil, g.stencil_dists
    faces.n, faces.L = g.n, g.L
    nodes.type, nodes.cells_index, nodes.cell_dists = g.node_type, g.ring, g.ring_dists
    nodes.pdf, nodes.rho, nodes.vel = g.node_pdf, g.node_rho, g.node_vel
    return cells, faces, nodes


def gather_requests(local: LocalMesh) -> Dict[int, np.ndarray]:
    """Collective: every rank publishes the global ids it needs; returns {peer: ids peer needs
    from this rank}.  Works on any backend (gloo on CPU in tests, nccl under torchrun)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    reqs: List[Optional[dict]] = [None] * world
    dist.all_gather_object(reqs, halo_requests(local))
    return {s: reqs[s][rank] for s in range(world) if rank in reqs[s]}


class HaloComm:
    """Neighbour send/recv of packed halo buffers ([cells][Q] rows, contiguous per peer)."""

    def __init__(self, peers_send, send_counts, peers_recv, recv_counts):
        self.peers_send, self.send_counts = list(peers_send), list(send_counts)
        self.peers_recv, self.recv_counts = list(peers_recv), list(recv_counts)

    def start(self, send_buf, recv_buf):
        import torch.distributed as dist
        key = (send_buf.data_ptr(), recv_buf.data_ptr())
        if getattr(self, "_key", None) != key:            # the op list is reused every iteration
            ops, off = [], 0
            for peer, cnt in zip(self.peers_send, self.send_counts):
                ops.append(dist.P2POp(dist.isend, send_buf[off:off + cnt], peer))
                off += cnt
            off = 0
            for peer, cnt in zip(self.peers_recv, self.recv_counts):
                ops.append(dist.P2POp(dist.irecv, recv_buf[off:off + cnt], peer))
                off += cnt
            self._ops, self._key = ops, key
        return dist.batch_isend_irecv(self._ops) if self._ops else []

    @staticmethod
    def finish(works):
        for w in works:
            w.wait()


class RankEngine:
    """One rank: Environment over the local mesh + halo buffers (torch CUDA tensors)."""

    def __init__(self, local: LocalMesh, dynamics, scheme: str, dtype, device: int,
                 requests_from_peers: Dict[int, np.ndarray]):
        import torch
        self.local = local
        self.Q = int(dynamics.NUM_QUIVERS)
        cells, faces, nodes = containers_from_mesh(local.mesh, dynamics, scheme)
        self.env = Environment(cells, faces, nodes, dtype=dtype, device=device, mode="fused",
                               reorder=local.perm if local.perm is not None else "none", n_owned=local.n_owned)
        self.env.init()
        self.env.build()
        (self.peers_send, send_cells, self.send_counts, self.peers_recv, recv_cells,
         self.recv_counts) = exchange_lists(local, requests_from_peers)
        self.env.halo_set_lists(send_cells, recv_cells)
        tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
        dev = torch.device("cuda", device)
        self.send_buf = torch.zeros((max(1, send_cells.size), self.Q), dtype=tdt, device=dev)
        self.recv_buf = torch.zeros((max(1, recv_cells.size), self.Q), dtype=tdt, device=dev)
        self.n_send, self.n_recv = int(send_cells.size), int(recv_cells.size)

    def send_slices(self):
        off = 0
        for peer, cnt in zip(self.peers_send, self.send_counts):
            yield peer, self.send_buf[off:off + cnt]
            off += cnt

    def recv_slices(self):
        off = 0
        for peer, cnt in zip(self.peers_recv, self.recv_counts):
            yield peer, self.recv_buf[off:off + cnt]
            off += cnt


class InProcessCluster:
    """All ranks in one process on one device; exchange = device-to-device copies."""

    def __init__(self, g: GlobalMesh, part: np.ndarray, nparts: int, dynamics, scheme: str, dtype=np.float32,
                 device: int = 0, reorder: bool = True):
        self.locals = [extract_local(g, part, r, reorder=reorder) for r in range(nparts)]
        reqs = [halo_requests(l) for l in self.locals]
        self.engines = []
        for r, l in enumerate(self.locals):
            from_peers = {s: reqs[s][r] for s in range(nparts) if r in reqs[s]}
            self.engines.append(RankEngine(l, dynamics, scheme, dtype, device, from_peers))
        self.n_global = g.num_cells
        self.Q = self.engines[0].Q
        self.dtype = np.dtype(dtype)

    def step(self, n: int = 1):
        import torch
        for _ in range(n):
            for e in self.engines:
                if e.n_send:
                    e.env.halo_pack(e.send_buf.data_ptr())
                e.env.step_phase(0)
            for e in self.engines:
                e.env.sync()
            for r, e in enumerate(self.engines):
                for peer, dst in e.recv_slices():
                    src = dict(self.engines[peer].send_slices())[r]
                    dst.copy_(src)
            torch.cuda.synchronize()
            for e in self.engines:
                if e.n_recv:
                    e.env.halo_unpack(e.recv_buf.data_ptr())
                e.env.step_phase(1)
        return self

    def gather_cells(self, name: str) -> np.ndarray:
        """Owned-cell field of every rank scattered back to global numbering."""
        width = {"pdf": self.Q, "rho": 1, "vel": 2, "pdf_eq": self.Q}[name]
        out = np.zeros((self.n_global, width), dtype=self.dtype)
        for e in self.engines:
            l = e.local
            out[l.cell_gid[:l.n_owned]] = getattr(e.env.cells, name)[:l.n_owned]
        return out

    def close(self):
        for e in self.engines:
            e.env.close()


# ---------------------------------------------------------------------------------------------------
# scalable construction of one rank's local mesh for the synthetic strip decomposition
# ---------------------------------------------------------------------------------------------------
def strip_local_mesh(nx: int, ny_per_rank: int, rank: int, world: int, dynamics, scheme: str,
                     jitter: float = 0.2, seed: int = 0, lid: float = 0.1, perturb: bool = True):
    """Rank ``rank``'s LocalMesh of the global ``nx x (ny_per_rank*world)`` x-periodic square cut
    into ``world`` strips along y, built from a window one quad row wider than the strip (nobody
    ever materialises the global mesh).  Returns (LocalMesh, faces_per_cell)."""
    from . import meshgen
    from .mesher import Mesher
    ny_total = ny_per_rank * world
    r0, r1 = rank * ny_per_rank, (rank + 1) * ny_per_rank
    w0, w1 = max(0, r0 - 1), min(ny_total, r1 + 1)
    raw = meshgen.strip_window(nx, ny_total, w0, w1, jitter=jitter, seed=seed, periodic_x=True)
    m = Mesher()
    m.import_meshpy(raw)
    m.calc_mesh_properties()
    cells, faces, nodes = m.to_env(dynamics, flux_method=scheme)
    nodes = m.set_vel_node(nodes, meshgen.BOTTOM, np.array([0.0, 0.0]))
    nodes = m.set_vel_node(nodes, meshgen.TOP, np.array([lid, 0.0]))
    if perturb:
        c = m.cell_centers
        rho = 1 + 0.01 * np.sin(2 * np.pi * c[:, 0] / nx) * np.sin(2 * np.pi * c[:, 1] / ny_total)
        u = 0.05 * np.stack([np.sin(2 * np.pi * c[:, 1] / ny_total), np.sin(2 * np.pi * c[:, 0] / nx)], axis=1)
        cells.pdf = dynamics.calc_eq(rho, u)
    g = GlobalMesh.from_containers(cells, faces, nodes)
    g.cell_gid = raw.cell_gid
    part = (raw.cell_row // ny_per_rank).astype(np.int32)
    local = extract_local(g, part, rank, reorder=True)
    owned_faces = np.unique(g.face_indices[part == rank].reshape(-1)).size
    return local, owned_faces / max(1, local.n_owned)


class DistributedEnvironment:
    """torch.distributed (NCCL) driver: one process per GPU, neighbour send/recv per iteration."""

    def __init__(self, local: LocalMesh, dynamics, scheme: str, dtype, device: int, n_global: int,
                 faces_per_cell: float = 1.5, native: bool = True):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        from_peers = gather_requests(local)
        self.engine = RankEngine(local, dynamics, scheme, dtype, device, from_peers)
        e = self.engine
        self.comm = HaloComm(e.peers_send, e.send_counts, e.peers_recv, e.recv_counts)
        self.env = self.engine.env
        self.n_owned, self.n_global, self.faces_per_cell = local.n_owned, n_global, faces_per_cell
        self.stream = torch.cuda.ExternalStream(self.env.stream_ptr, device=torch.device("cuda", device))
        self.native = native
        if native:
            # torch.distributed is only the plumbing: rank 0 mints the NCCL id, everybody gets it,
            # the engine then owns its communicator and exchanges halos without host code per step
            ids = [None]
            if self.rank == 0:
                import ctypes as C
                buf = C.create_string_buffer(_lib.COMM_ID_BYTES)
                _lib.check(_lib.load().fvdbm_comm_unique_id(buf))
                ids = [bytes(buf.raw)]
            dist.broadcast_object_list(ids, src=0)
            self.env.comm_attach(self.world, self.rank, ids[0], e.peers_send, e.send_counts, e.peers_recv, e.recv_counts)

    @classmethod
    def weak_scaling_square(cls, nx, scheme, dtype, rank, world, device, reorder="hilbert", native=True):
        """nx x nx quads per rank: the global mesh grows with the number of GPUs."""
        return cls.strips(nx, nx, scheme, dtype, rank, world, device, native=native)

    @classmethod
    def strips(cls, nx, ny_per_rank, scheme, dtype, rank, world, device, native=True):
        """Global nx x (ny_per_rank*world) x-periodic square, one horizontal strip per rank."""
        from .dynamics import D2Q9
        dyn = D2Q9(tau=0.8, delta_t=0.1)
        local, fpc = strip_local_mesh(nx, ny_per_rank, rank, world, dyn, scheme)
        return cls(local, dyn, scheme, dtype, device, n_global=2 * nx * ny_per_rank * world, faces_per_cell=fpc,
                   native=native)

    @classmethod
    def from_raw(cls, raw, dynamics, scheme, dtype, rank, world, device, boundary_conditions=None, initial_pdf=None,
                 owner=None, native=True):
        """General decomposition of an arbitrary raw mesh (points / elements / markers): Hilbert-chunk owners
        computed from the raw arrays, then only this rank's window is meshed (partition.local_from_raw)."""
        from .partition import local_from_raw
        local, fpc = local_from_raw(raw, rank, world, dynamics, scheme, boundary_conditions, initial_pdf, owner)
        return cls(local, dynamics, scheme, dtype, device, n_global=int(np.asarray(raw.elements).shape[0]),
                   faces_per_cell=fpc, native=native)

    def partition_stats(self):
        """Collective: peers / halo sizes / message bytes of every rank (for reports)."""
        e = self.engine
        mine = {"owned": int(self.n_owned), "halo": int(e.local.n_local - e.local.n_owned), "peers": len(e.peers_recv),
                "send_cells": int(e.n_send), "send_bytes_per_iteration": int(e.n_send) * e.Q * np.dtype(self.env.dtype).itemsize}
        allr = [None] * self.world
        self.dist.all_gather_object(allr, mine)
        return allr

    # ---- stepping ---------------------------------------------------------------------------------
    def _iterate(self):
        e = self.engine
        if e.n_send:
            e.env.halo_pack(e.send_buf.data_ptr())
        works = self.comm.start(e.send_buf, e.recv_buf)
        e.env.step_phase(0)
        self.comm.finish(works)
        if e.n_recv:
            e.env.halo_unpack(e.recv_buf.data_ptr())
        e.env.step_phase(1)

    def step(self, n: int = 1):
        if self.native:
            self.env.step(n)
            return self
        with self.torch.cuda.stream(self.stream):
            for _ in range(n):
                self._iterate()
        return self

    def step_timed(self, n: int) -> float:
        if self.native:
            return self.env.step_timed(n)
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            e0.record(self.stream)
            for _ in range(n):
                self._iterate()
            e1.record(self.stream)
        e1.synchronize()
        return float(e0.elapsed_time(e1))

    # ---- pass-throughs ----------------------------------------------------------------------------
    def set_option(self, option, value):
        self.env.set_option(option, value)
        return self

    def info(self, key):
        return self.env.info(key)

    def sync(self):
        self.env.sync()
        return self

    def get_into(self, name, out, wait=True):
        """Owned rows of a local cell field straight into ``out`` ((n_owned, width), e.g. pinned)."""
        return self.env.get_into(name, out, wait=wait)

    def set_cells_pdf(self, owned_pdf: np.ndarray, wait=True):
        self.env.set_cells_pdf(owned_pdf, wait=wait)
        return self

    def wait(self, ticket=None):
        self.env.wait(ticket)
        return self

    def close(self):
        self.env.close()
