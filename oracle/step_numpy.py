"""Vectorised NumPy restatement of the reference hot path ``Environment.step()``.

TEST INFRASTRUCTURE ONLY -- the checker for the CUDA path (tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline leg).  The product package never imports it.

Pinned (tests/test_oracle_golden.py) against tests/golden/*.npz, which were produced by executing
the reference's own sources under oracle/jaxshim (oracle/make_golden.py): agreement <= 1e-13 in
fp64 on all eight state arrays, both flux schemes, node types 0/1/2, K=3 and K=4, D2Q9 and D2Q13.

Each stage cites the reference lines it restates (paths relative to /root/reference):
  S1 moments        src/environment.py:60 -> src/containers.py:93-98  -> src/dynamics.py:35-47
  S2 equilibrium    src/environment.py:61 -> src/containers.py:100-105 -> src/dynamics.py:70-74,101-102
  S3 boundary nodes src/environment.py:62 -> src/containers.py:339-404, utils/utils.py:34-60
  S4 face flux      src/environment.py:63 -> src/containers.py:191-287, utils/utils.py:153-154
  S5 cell update    src/environment.py:64 -> src/containers.py:107-121
  cc stencil faces  src/faces.py:10-76 (CCStencilFaces: S4 flux times cos(alpha))
"""
from __future__ import annotations

import numpy as np

KSI9 = np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1], [1, 1], [-1, 1], [-1, -1], [1, -1]])
KSI13 = np.concatenate([KSI9, np.array([[2, 0], [0, 2], [-2, 0], [0, -2]])])
W9 = np.array([4 / 9, 1 / 9, 1 / 9, 1 / 9, 1 / 9, 1 / 36, 1 / 36, 1 / 36, 1 / 36])
W13 = np.array([3 / 8, 1 / 12, 1 / 12, 1 / 12, 1 / 12, 1 / 16, 1 / 16, 1 / 16, 1 / 16,
                1 / 96, 1 / 96, 1 / 96, 1 / 96])


class Lattice:
    """D2Q9 / D2Q13 constants evaluated in the working precision (src/dynamics.py:52-64,79-95).
    In fp32 this reproduces stock JAX: C = 1/sqrt(3) in fp32 so C**2 = 0.3333333f, not 1/3."""

    def __init__(self, Q: int, dtype):
        dt = np.dtype(dtype).type
        self.Q = Q
        self.ksi = (KSI9 if Q == 9 else KSI13).astype(dt)
        self.w = (W9 if Q == 9 else W13).astype(dt)
        c = dt(1) / np.sqrt(dt(3 if Q == 9 else 2))
        self.c2 = dt(c * c)
        self.c4 = dt(self.c2 * self.c2)
        self.c6 = dt(self.c4 * self.c2)
        self.dt = dt

    def eq(self, rho, vel):
        """src/dynamics.py:70-74 (D2Q9) / :101-102 (D2Q13); rho (...,1), vel (...,2) -> (...,Q)."""
        ku = vel @ self.ksi.T
        uu = np.sum(vel * vel, axis=-1, keepdims=True)
        two = self.dt(2)
        poly = 1 + ku / self.c2 + ku ** 2 / (two * self.c4) - uu / (two * self.c2)
        if self.Q == 13:
            poly = poly + ku ** 3 / (two * self.c6) - self.dt(3) * ku * uu / (two * self.c4)
        return self.w * rho * poly


class StepOracle:
    """State + statics in the reference's AoS layout; ``step()`` advances one FVDBM iteration."""

    def __init__(self, static: dict, state: dict, Q: int, tau: float, delta_t: float,
                 scheme: str, dtype=np.float64):
        if scheme not in ("upwind", "lax_wendroff"):
            raise ValueError(f"Unknown flux scheme: {scheme}")       # src/containers.py:203
        dt = np.dtype(dtype)
        self.dtype = dt
        self.lat = Lattice(Q, dt)
        self.tau, self.delta_t, self.scheme = tau, delta_t, scheme
        f = lambda a: np.array(a, dtype=dt)
        i = lambda a: np.array(a, dtype=np.int64)
        self.face_indices = i(static["cells.face_indices"])           # (N,K)
        self.face_signs = i(static["cells.face_normals"])             # (N,K) +-1
        self.nodes_index = i(static["faces.nodes_index"])             # (F,2)
        self.stencil = i(static["faces.stencil_cells_index"])         # (F,2), -1 = ghost
        self.dists = f(static["faces.stencil_dists"])                 # (F,2)
        self.n = f(static["faces.n"])                                 # (F,2)
        self.L = f(static["faces.L"]).reshape(-1, 1)                  # (F,1)
        # CCStencilFaces (src/faces.py:54-72): flux * cos(alpha); absent for the base Faces
        self.cos_alpha = np.cos(f(static["faces.alpha"]).reshape(-1, 1)) if "faces.alpha" in static else None
        self.type = i(static["nodes.type"]).reshape(-1, 1)            # (P,1)
        self.ring = i(static["nodes.cells_index"])                    # (P,M), -1 padded
        self.ring_d = f(static["nodes.cell_dists"])                   # (P,M), -1 padded
        self.pdf = f(state["cells.pdf"])
        self.inv_area = None if static.get("cells.inv_area") is None else f(static["cells.inv_area"]).reshape(-1)
        N, P = self.pdf.shape[0], self.type.shape[0]
        self.rho = f(state.get("cells.rho", np.zeros((N, 1))))
        self.vel = f(state.get("cells.vel", np.zeros((N, 2))))
        self.pdf_eq = f(state.get("cells.pdf_eq", np.zeros((N, Q))))
        self.flux = f(state.get("faces.pdf", np.zeros((self.stencil.shape[0], Q))))
        self.npdf = f(state["nodes.pdf"])
        self.nrho = f(state.get("nodes.rho", np.zeros((P, 1))))
        self.nvel = f(state.get("nodes.vel", np.zeros((P, 2))))
        # only nodes with a BC type need the ring gathers (reference evaluates all, then masks)
        self.active = np.nonzero(self.type[:, 0] != 0)[0]
        with np.errstate(divide="ignore"):
            w = dt.type(1.) / self.ring_d[self.active]                # utils/utils.py:58
        self.ring_w = np.where(w < 0, dt.type(0), w)                  # utils/utils.py:59
        self.ring_a = self.ring[self.active]

    # ---------------------------------------------------------------- stages
    def moments(self):                                                # S1
        self.rho = np.sum(self.pdf, axis=1, keepdims=True)
        self.vel = (self.pdf @ self.lat.ksi) / self.rho

    def equilibrium(self):                                            # S2
        self.pdf_eq = self.lat.eq(self.rho, self.vel)

    def nodes(self):                                                  # S3
        a = self.active
        if a.size == 0:
            return
        idx, w = self.ring_a, self.ring_w                             # (A,M)
        wsum = np.sum(w[..., None], axis=1)                           # (A,1)
        t = self.type[a]                                              # (A,1)
        # velocity nodes: density interpolated (containers.py:348-351,364-368)
        rho_i = np.sum(self.rho[idx] * w[..., None], axis=1) / wsum
        nrho = np.where(t == 1, rho_i, self.nrho[a])
        # density nodes: velocity interpolated (containers.py:343-346,370-374)
        vel_i = np.sum(self.vel[idx] * w[..., None], axis=1) / wsum
        nvel = np.where(t == 2, vel_i, self.nvel[a])
        # non-equilibrium extrapolation (containers.py:353-361,383-390)
        neq = np.where((idx == -1)[..., None], self.dtype.type(0), self.pdf[idx] - self.pdf_eq[idx])
        neq_i = np.sum(neq * w[..., None], axis=1) / wsum
        self.nrho[a] = nrho
        self.nvel[a] = nvel
        self.npdf[a] = self.lat.eq(nrho, nvel) + neq_i

    def _ghost(self, known, d_ghost, d_known):                        # containers.py:280-287
        gbar = np.mean(self.npdf[self.nodes_index], axis=1)
        fk = self.pdf[known]
        return gbar + (gbar - fk) * (d_ghost / d_known)[:, None]      # utils/utils.py:153-154

    def fluxes(self):                                                 # S4
        s0, s1 = self.stencil[:, 0], self.stencil[:, 1]
        d0, d1 = self.dists[:, 0], self.dists[:, 1]
        f0 = self.pdf[s0]
        f1 = self.pdf[s1]
        g0 = np.nonzero(s0 == -1)[0]
        g1 = np.nonzero(s1 == -1)[0]
        if g0.size:
            sub = _FaceSubset(self, g0)
            f0[g0] = sub.ghost(s1[g0], d0[g0], d1[g0])
        if g1.size:
            sub = _FaceSubset(self, g1)
            f1[g1] = sub.ghost(s0[g1], d1[g1], d0[g1])
        varpi = self.n @ self.lat.ksi.T                               # (F,Q)
        if self.scheme == "upwind":                                   # containers.py:234-238
            fs = np.where(varpi >= 0, f0, f1)
        else:                                                         # containers.py:266-275
            dd = (d0 + d1)[:, None]
            interp = (d0[:, None] / dd) - (varpi * self.delta_t) / (2 * dd)
            fs = f0 + (f1 - f0) * interp
        self.flux = fs * varpi * self.L
        if self.cos_alpha is not None:
            self.flux = self.flux * self.cos_alpha

    def cells(self):                                                  # S5
        fl = self.flux[self.face_indices] * self.face_signs[..., None].astype(self.dtype)
        total = np.sum(fl, axis=1)
        if self.inv_area is not None:      # optional physically consistent mode (NOT in the reference, containers.py:115-121)
            total = total * self.inv_area[:, None]
        self.pdf = self.pdf + self.delta_t * (1 / self.tau * (self.pdf_eq - self.pdf) - total)

    def step(self, n: int = 1):
        for _ in range(n):
            self.moments()
            self.equilibrium()
            self.nodes()
            self.fluxes()
            self.cells()
        return self

    def state(self) -> dict:
        return {"cells.pdf": self.pdf, "cells.rho": self.rho, "cells.vel": self.vel,
                "cells.pdf_eq": self.pdf_eq, "faces.pdf": self.flux, "nodes.pdf": self.npdf,
                "nodes.rho": self.nrho, "nodes.vel": self.nvel}


class _FaceSubset:
    """Ghost evaluation restricted to the boundary faces (the reference evaluates it for every
    face and discards it on interior ones, containers.py:227-232,258-263)."""

    def __init__(self, o: StepOracle, faces):
        self.o, self.faces = o, faces

    def ghost(self, known, d_ghost, d_known):
        o = self.o
        gbar = np.mean(o.npdf[o.nodes_index[self.faces]], axis=1)
        fk = o.pdf[known]
        return gbar + (gbar - fk) * (d_ghost / d_known)[:, None]
