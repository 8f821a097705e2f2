"""Lattice models: same names and attribute surface as /root/reference/src/dynamics.py:13-102.

Host-side NumPy; the device kernels take the lattice as template parameters (Q in {9,13}) and the
precision-dependent constants (W, C^2, 2C^4, 2C^2, 2C^6) from ``lattice_constants`` so that an
fp32 build uses the same fp32-rounded constants as stock JAX (C = 1/sqrt(3) evaluated in fp32:
C^2 = 0.3333333f), and an fp64 build the fp64 ones.
"""
from __future__ import annotations

import numpy as np

__all__ = ["Dynamics", "D2Q9", "D2Q13"]


class Dynamics:
    """Base class (reference dynamics.py:13-47)."""
    DIM: int
    NUM_QUIVERS: int
    KSI: np.ndarray
    W: np.ndarray
    C: float
    _SQRT_ARG: int
    tau: float
    delta_t: float

    def __init__(self):
        pass

    # -- precision-aware constants -------------------------------------------------------------
    @classmethod
    def lattice_constants(cls, dtype=np.float64):
        """(W, cs2, two_cs4, two_cs2, two_cs6) evaluated in ``dtype`` exactly like the reference
        expressions ``self.C**2``, ``2*self.C**4``, ``2*self.C**2``, ``2*self.C**6`` would be."""
        dt = np.dtype(dtype).type
        c = dt(1) / np.sqrt(dt(cls._SQRT_ARG))
        c2 = dt(c * c)
        c4 = dt(c2 * c2)
        c6 = dt(c4 * c2)
        w = np.asarray(cls.W, dtype=np.float64).astype(dt)
        return w, c2, dt(2) * c4, dt(2) * c2, dt(2) * c6

    def ones_pdf(self):
        return np.ones(self.NUM_QUIVERS)

    def density(self, pdf):
        return np.sum(pdf, keepdims=True)

    def velocity(self, pdf, rho):
        return np.dot(self.KSI.T, pdf) / rho

    def calc_macro(self, pdf):
        rho = self.density(pdf)
        return rho, self.velocity(pdf, rho)

    def calc_eq(self, rho, vel):
        """Equilibrium for one (rho, vel) pair or a batch (..., ) / (..., 2); dtype follows input."""
        vel = np.asarray(vel)
        dt = vel.dtype if vel.dtype.kind == "f" else np.dtype(np.float64)
        w, c2, tc4, tc2, tc6 = self.lattice_constants(dt)
        rho = np.asarray(rho, dtype=dt)
        ksi = self.KSI.astype(dt)
        ku = vel @ ksi.T                                           # (..., Q)
        uu = np.sum(vel * vel, axis=-1, keepdims=True)
        if rho.ndim == vel.ndim - 1:
            rho = rho[..., np.newaxis]
        poly = 1 + ku / c2 + ku ** 2 / tc4 - uu / tc2
        if self.NUM_QUIVERS == 13:                                 # reference dynamics.py:101-102
            poly = poly + ku ** 3 / tc6 - 3 * ku * uu / tc4
        return w * rho * poly

    def __repr__(self):
        return f"{type(self).__name__}(tau={self.tau}, delta_t={self.delta_t})"


class D2Q9(Dynamics):
    """reference dynamics.py:50-74"""
    DIM = 2
    NUM_QUIVERS = 9
    KSI = np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1],
                    [1, 1], [-1, 1], [-1, -1], [1, -1]], dtype=np.int32)
    W = np.array([4 / 9, 1 / 9, 1 / 9, 1 / 9, 1 / 9, 1 / 36, 1 / 36, 1 / 36, 1 / 36])
    _SQRT_ARG = 3
    C = 1 / np.sqrt(3)

    def __init__(self, tau, delta_t):
        super().__init__()
        self.tau = tau
        self.delta_t = delta_t


class D2Q13(Dynamics):
    """reference dynamics.py:77-102"""
    DIM = 2
    NUM_QUIVERS = 13
    KSI = np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1],
                    [1, 1], [-1, 1], [-1, -1], [1, -1],
                    [2, 0], [0, 2], [-2, 0], [0, -2]], dtype=np.int32)
    W = np.array([3 / 8, 1 / 12, 1 / 12, 1 / 12, 1 / 12, 1 / 16, 1 / 16, 1 / 16, 1 / 16,
                  1 / 96, 1 / 96, 1 / 96, 1 / 96])
    _SQRT_ARG = 2
    C = 1 / np.sqrt(2)

    def __init__(self, tau, delta_t):
        super().__init__()
        self.tau = tau
        self.delta_t = delta_t
