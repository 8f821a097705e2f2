"""ctypes front-end of oracle/libstep_c.so (the C/OpenMP restatement). TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libstep_c.so")


class ODesc(C.Structure):
    _fields_ = [("scheme", C.c_int32), ("Q", C.c_int32), ("K", C.c_int32), ("M", C.c_int32),
                ("N", C.c_int64), ("F", C.c_int64), ("P", C.c_int64), ("tau", C.c_double), ("delta_t", C.c_double),
                ("lat_w", C.c_double * 16), ("cs2", C.c_double), ("two_cs4", C.c_double), ("two_cs2", C.c_double),
                ("two_cs6", C.c_double)] + [(n, C.c_void_p) for n in (
                    "cell_face_idx", "cell_face_sign", "face_cell_idx", "face_dists", "face_node_idx", "face_n",
                    "face_L", "node_type", "node_cell_idx", "node_cell_dist")]


def threads() -> int:
    return int(C.CDLL(LIB).fvdbm_oracle_threads())


def use_all_cores() -> int:
    """Ignore an inherited OMP_NUM_THREADS=1 (torchrun sets it) and use every host core."""
    lib = C.CDLL(LIB)
    lib.fvdbm_oracle_set_threads(C.c_int(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
    return int(lib.fvdbm_oracle_threads())


class COracle:
    """Same constructor contract as oracle.step_numpy.StepOracle."""

    def __init__(self, static, state, Q, tau, delta_t, scheme, dtype=np.float64):
        from .step_numpy import Lattice
        self.lib = C.CDLL(LIB)
        real = self.real = np.dtype(dtype)
        lat = Lattice(Q, real)
        keep = self.keep = {}

        def arr(name, a, dt, shape):
            keep[name] = np.ascontiguousarray(np.asarray(a), dtype=dt).reshape(shape)
            return keep[name]
        fi = np.asarray(static["cells.face_indices"])
        N, K = fi.shape
        F = np.asarray(static["faces.n"]).shape[0]
        ring = np.asarray(static["nodes.cells_index"])
        P, M = ring.shape
        arr("cell_face_idx", fi, np.int32, (N, K))
        arr("cell_face_sign", static["cells.face_normals"], np.int32, (N, K))
        arr("face_cell_idx", static["faces.stencil_cells_index"], np.int32, (F, 2))
        arr("face_dists", static["faces.stencil_dists"], real, (F, 2))
        arr("face_node_idx", static["faces.nodes_index"], np.int32, (F, 2))
        arr("face_n", static["faces.n"], real, (F, 2))
        L = np.asarray(static["faces.L"], dtype=real).reshape(F)
        if "faces.alpha" in static:      # CCStencilFaces: flux * cos(alpha) == a face length scaled by cos(alpha)
            L = (L * np.cos(np.asarray(static["faces.alpha"], dtype=real).reshape(F))).astype(real)
        arr("face_L", L, real, (F,))
        arr("node_type", static["nodes.type"], np.int32, (P,))
        arr("node_cell_idx", ring, np.int32, (P, M))
        arr("node_cell_dist", static["nodes.cell_dists"], real, (P, M))
        d = self.desc = ODesc()
        d.scheme = 0 if scheme == "upwind" else 1
        d.Q, d.K, d.M, d.N, d.F, d.P = Q, K, M, N, F, P
        d.tau, d.delta_t = tau, delta_t
        for q in range(Q):
            d.lat_w[q] = float(lat.w[q])
        two = real.type(2)
        d.cs2, d.two_cs4, d.two_cs2, d.two_cs6 = float(lat.c2), float(two * lat.c4), float(two * lat.c2), float(two * lat.c6)
        for n in keep:
            setattr(d, n, keep[n].ctypes.data)
        z = lambda k, shape: np.array(state.get(k, np.zeros(shape)), dtype=real).reshape(shape)
        self.pdf = z("cells.pdf", (N, Q)); self.rho = z("cells.rho", (N, 1)); self.vel = z("cells.vel", (N, 2))
        self.pdf_eq = z("cells.pdf_eq", (N, Q)); self.flux = z("faces.pdf", (F, Q))
        self.npdf = z("nodes.pdf", (P, Q)); self.nrho = z("nodes.rho", (P, 1)); self.nvel = z("nodes.vel", (P, 2))
        self.fn = getattr(self.lib, "fvdbm_oracle_step_f32" if real.itemsize == 4 else "fvdbm_oracle_step_f64")
        self.fn.restype = C.c_int

    def step(self, n=1):
        ptr = lambda a: C.c_void_p(a.ctypes.data)
        rc = self.fn(C.byref(self.desc), ptr(self.pdf), ptr(self.rho), ptr(self.vel), ptr(self.pdf_eq), ptr(self.flux),
                     ptr(self.npdf), ptr(self.nrho), ptr(self.nvel), int(n))
        assert rc == 0
        return self

    def state(self):
        return {"cells.pdf": self.pdf, "cells.rho": self.rho, "cells.vel": self.vel, "cells.pdf_eq": self.pdf_eq,
                "faces.pdf": self.flux, "nodes.pdf": self.npdf, "nodes.rho": self.nrho, "nodes.vel": self.nvel}
