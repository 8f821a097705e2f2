"""ctypes binding of libfvdbm_b200.so (include/fvdbm_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no CPU
fallback: if the shared object is missing ``load()`` raises, and every compute entry point needs a
CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FVDBM_LIB", os.path.join(HERE, "libfvdbm_b200.so"))   # FVDBM_LIB: A/B builds

ABI_VERSION = 2
COMM_ID_BYTES = 128
OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4
SCHEME_UPWIND, SCHEME_LAX_WENDROFF = 0, 1
MODE_AUTO, MODE_STAGED, MODE_FUSED = 0, 1, 2
VARIANT_AUTO, VARIANT_DIRECT, VARIANT_TMA, VARIANT_PAIR, VARIANT_REC = 0, 1, 2, 3, 4
(CELL_PDF, CELL_RHO, CELL_VEL, CELL_PDF_EQ, FACE_FLUX, NODE_PDF, NODE_RHO, NODE_VEL,
 CELL_PDF_PREV) = range(9)
(INFO_MODE, INFO_STEPS, INFO_LAUNCHES, INFO_TRACKED_NODES, INFO_BOUNDARY_SIDES, INFO_DEVICE_BYTES,
 INFO_VARIANT, INFO_NPAD, INFO_FUSED_OK, INFO_HALO_CELLS, INFO_OWNED_CELLS, INFO_GRAPH_STEPS) = range(12)
(OPT_VARIANT, OPT_TILE_CELLS, OPT_STAGES, OPT_GRAPH_STEPS, OPT_CTAS_PER_SM, OPT_REVERSE_SWEEP,
 OPT_TEMPORAL, OPT_PREFETCH_DIST, OPT_PDL) = range(9)

EXPORTS = ("fvdbm_abi_version", "fvdbm_create", "fvdbm_destroy", "fvdbm_last_error", "fvdbm_step",
           "fvdbm_step_timed", "fvdbm_sync", "fvdbm_get", "fvdbm_set", "fvdbm_set_async", "fvdbm_get_async", "fvdbm_wait",
           "fvdbm_set_params",
           "fvdbm_set_option", "fvdbm_info", "fvdbm_check_finite", "fvdbm_halo_set_lists", "fvdbm_halo_pack",
           "fvdbm_halo_unpack", "fvdbm_step_phase", "fvdbm_stream", "fvdbm_comm_unique_id", "fvdbm_comm_init",
           "fvdbm_halo_set_peers", "fvdbm_sfc_keys", "fvdbm_mesh_ring_width", "fvdbm_mesh_properties", "fvdbm_mesh_unique_edges", "fvdbm_plan_create",
           "fvdbm_plan_destroy", "fvdbm_plan_array", "fvdbm_plan_scalar")


class MeshDesc(C.Structure):
    """fvdbm_mesh_desc (include/fvdbm_b200.h): raw triangle mesh in, Mesher attributes out."""
    _fields_ = [("N", C.c_int64), ("F", C.c_int64), ("P", C.c_int64), ("M", C.c_int32), ("reserved", C.c_int32)] + [
        (name, C.c_void_p) for name in (
            "points", "cells", "faces", "point_alias", "cell_centers", "cell_face_indices", "cell_face_normals",
            "cell_face_normal_signs", "faces_out", "face_centers", "face_normals", "face_lengths", "face_cell_indices",
            "face_cell_center_distances", "stencil_norms", "cc_stencil_dist", "face_stencil_angles",
            "point_cell_indices", "point_cell_center_distances")]


class Desc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device_id", C.c_int32), ("dtype", C.c_int32), ("scheme", C.c_int32),
        ("Q", C.c_int32), ("K", C.c_int32), ("M", C.c_int32), ("mode", C.c_int32),
        ("N", C.c_int64), ("F", C.c_int64), ("P", C.c_int64), ("N_owned", C.c_int64),
        ("tau", C.c_double), ("delta_t", C.c_double),
        ("lat_w", C.c_double * 16),
        ("cs2", C.c_double), ("two_cs4", C.c_double), ("two_cs2", C.c_double), ("two_cs6", C.c_double),
        ("cell_face_idx", C.c_void_p), ("cell_face_sign", C.c_void_p), ("face_cell_idx", C.c_void_p),
        ("face_dists", C.c_void_p), ("face_node_idx", C.c_void_p), ("face_n", C.c_void_p),
        ("face_L", C.c_void_p), ("node_type", C.c_void_p), ("node_cell_idx", C.c_void_p),
        ("node_cell_dist", C.c_void_p), ("cell_pdf", C.c_void_p), ("node_pdf", C.c_void_p),
        ("node_rho", C.c_void_p), ("node_vel", C.c_void_p), ("cell_perm", C.c_void_p),
        ("cell_inv_area", C.c_void_p),
    ]


_lib = None


def load():
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). fvdbm_jax_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    H, P = C.c_void_p, C.c_void_p
    lib.fvdbm_abi_version.restype = C.c_int
    lib.fvdbm_create.argtypes = [C.POINTER(Desc), C.POINTER(H)]
    lib.fvdbm_create.restype = C.c_int
    lib.fvdbm_destroy.argtypes = [H]
    lib.fvdbm_destroy.restype = None
    lib.fvdbm_last_error.argtypes = [H]
    lib.fvdbm_last_error.restype = C.c_char_p
    lib.fvdbm_step.argtypes = [H, C.c_int]
    lib.fvdbm_step_timed.argtypes = [H, C.c_int, C.POINTER(C.c_float)]
    lib.fvdbm_step_phase.argtypes = [H, C.c_int]
    lib.fvdbm_sync.argtypes = [H]
    lib.fvdbm_get.argtypes = [H, C.c_int, C.c_void_p, C.c_size_t]
    lib.fvdbm_set.argtypes = [H, C.c_int, C.c_void_p, C.c_size_t]
    lib.fvdbm_set_async.argtypes = [H, C.c_int, C.c_void_p, C.c_size_t]
    lib.fvdbm_get_async.argtypes = [H, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int64)]
    lib.fvdbm_wait.argtypes = [H, C.c_int64]
    lib.fvdbm_set_params.argtypes = [H, C.c_double, C.c_double]
    lib.fvdbm_set_option.argtypes = [H, C.c_int, C.c_int64]
    lib.fvdbm_info.argtypes = [H, C.c_int, C.POINTER(C.c_int64)]
    lib.fvdbm_check_finite.argtypes = [H, C.POINTER(C.c_int64)]
    lib.fvdbm_check_finite.restype = C.c_int
    lib.fvdbm_halo_set_lists.argtypes = [H, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    lib.fvdbm_halo_pack.argtypes = [H, C.c_void_p]
    lib.fvdbm_halo_unpack.argtypes = [H, C.c_void_p]
    lib.fvdbm_comm_unique_id.argtypes = [C.c_void_p]
    lib.fvdbm_comm_unique_id.restype = C.c_int
    lib.fvdbm_comm_init.argtypes = [H, C.c_int, C.c_int, C.c_void_p]
    lib.fvdbm_comm_init.restype = C.c_int
    lib.fvdbm_halo_set_peers.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    lib.fvdbm_halo_set_peers.restype = C.c_int
    lib.fvdbm_stream.argtypes = [H]
    lib.fvdbm_stream.restype = C.c_void_p
    lib.fvdbm_sfc_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
    lib.fvdbm_sfc_keys.restype = C.c_int
    lib.fvdbm_mesh_ring_width.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    lib.fvdbm_mesh_ring_width.restype = C.c_int64
    lib.fvdbm_mesh_unique_edges.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    lib.fvdbm_mesh_unique_edges.restype = C.c_int64
    lib.fvdbm_mesh_properties.argtypes = [C.POINTER(MeshDesc)]
    lib.fvdbm_mesh_properties.restype = C.c_int
    lib.fvdbm_plan_create.argtypes = [C.POINTER(Desc), C.POINTER(P)]
    lib.fvdbm_plan_destroy.argtypes = [P]
    lib.fvdbm_plan_destroy.restype = None
    lib.fvdbm_plan_array.argtypes = [P, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]
    lib.fvdbm_plan_array.restype = C.c_int64
    lib.fvdbm_plan_scalar.argtypes = [P, C.c_char_p]
    lib.fvdbm_plan_scalar.restype = C.c_int64
    for name in ("fvdbm_step", "fvdbm_step_timed", "fvdbm_step_phase", "fvdbm_sync", "fvdbm_get", "fvdbm_set",
                 "fvdbm_set_async", "fvdbm_get_async", "fvdbm_wait",
                 "fvdbm_set_params", "fvdbm_set_option", "fvdbm_info", "fvdbm_halo_set_lists",
                 "fvdbm_halo_pack", "fvdbm_halo_unpack", "fvdbm_plan_create"):
        getattr(lib, name).restype = C.c_int
    if lib.fvdbm_abi_version() != ABI_VERSION:
        raise RuntimeError("libfvdbm_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def last_error(handle=None) -> str:
    msg = load().fvdbm_last_error(handle)
    return msg.decode() if msg else ""


def check(rc: int, handle=None):
    """Map C status codes onto the reference's exception conventions."""
    if rc == OK:
        return
    msg = last_error(handle)
    if rc == ERR_ARG:
        raise ValueError(msg)
    raise RuntimeError(f"libfvdbm_b200: {msg} (status {rc})")


class DescArrays:
    """Builds a ``Desc`` from reference-layout arrays and keeps them alive."""

    def __init__(self, *, dtype, scheme, Q, K, tau, delta_t, lattice_constants,
                 cell_face_idx, cell_face_sign, face_cell_idx, face_dists, face_node_idx, face_n, face_L,
                 node_type, node_cell_idx, node_cell_dist, cell_pdf, node_pdf, node_rho, node_vel,
                 cell_perm=None, n_owned=0, device_id=0, mode=MODE_AUTO, cell_inv_area=None):
        real = np.dtype(dtype)
        if real not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise ValueError("dtype must be float32 or float64")
        if scheme not in ("upwind", "lax_wendroff"):
            raise ValueError(f"Unknown flux scheme: {scheme}")          # reference containers.py:203
        self.real = real
        keep = self.keep = {}

        def arr(name, a, dt, shape=None):
            a = np.ascontiguousarray(np.asarray(a), dtype=dt)
            if shape is not None:
                a = a.reshape(shape)
            keep[name] = a
            return a

        cfi = arr("cell_face_idx", cell_face_idx, np.int32)
        if cfi.ndim != 2:
            raise ValueError("cells.face_indices must be (N,K)")
        N, Kin = cfi.shape
        if Kin != K:
            raise ValueError("K does not match cells.face_indices")
        arr("cell_face_sign", cell_face_sign, np.int32, (N, K))
        fci = arr("face_cell_idx", face_cell_idx, np.int32)
        F = fci.shape[0]
        arr("face_cell_idx", fci, np.int32, (F, 2))
        arr("face_dists", face_dists, real, (F, 2))
        arr("face_node_idx", face_node_idx, np.int32, (F, 2))
        arr("face_n", face_n, real, (F, 2))
        arr("face_L", face_L, real, (F,))
        nt = arr("node_type", node_type, np.int32).reshape(-1)
        keep["node_type"] = nt
        Pn = nt.shape[0]
        nci = np.asarray(node_cell_idx)
        M = nci.shape[1] if nci.ndim == 2 else 1
        arr("node_cell_idx", nci, np.int32, (Pn, M))
        arr("node_cell_dist", node_cell_dist, real, (Pn, M))
        arr("cell_pdf", cell_pdf, real, (N, Q))
        arr("node_pdf", node_pdf, real, (Pn, Q))
        arr("node_rho", node_rho, real, (Pn,))
        arr("node_vel", node_vel, real, (Pn, 2))
        if cell_perm is not None:
            arr("cell_perm", cell_perm, np.int32, (N,))
        if cell_inv_area is not None:
            arr("cell_inv_area", cell_inv_area, real, (N,))
        self.N, self.F, self.P, self.M, self.Q, self.K = N, F, Pn, M, Q, K

        d = self.desc = Desc()
        d.abi_version = ABI_VERSION
        d.device_id = device_id
        d.dtype = real.itemsize * 8
        d.scheme = SCHEME_UPWIND if scheme == "upwind" else SCHEME_LAX_WENDROFF
        d.Q, d.K, d.M, d.mode = Q, K, M, mode
        d.N, d.F, d.P, d.N_owned = N, F, Pn, n_owned
        d.tau, d.delta_t = float(tau), float(delta_t)
        w, c2, tc4, tc2, tc6 = lattice_constants
        for q in range(Q):
            d.lat_w[q] = float(w[q])
        d.cs2, d.two_cs4, d.two_cs2, d.two_cs6 = float(c2), float(tc4), float(tc2), float(tc6)
        for name in ("cell_face_idx", "cell_face_sign", "face_cell_idx", "face_dists", "face_node_idx", "face_n",
                     "face_L", "node_type", "node_cell_idx", "node_cell_dist", "cell_pdf", "node_pdf", "node_rho",
                     "node_vel"):
            setattr(d, name, keep[name].ctypes.data)
        d.cell_perm = keep["cell_perm"].ctypes.data if cell_perm is not None else None
        d.cell_inv_area = keep["cell_inv_area"].ctypes.data if cell_inv_area is not None else None


class HostPlan:
    """Host-only view of the device layout (fvdbm_plan_*): unit-testable without a GPU."""

    def __init__(self, desc_arrays: DescArrays):
        self._lib = load()
        self._da = desc_arrays
        self._p = C.c_void_p()
        check(self._lib.fvdbm_plan_create(C.byref(desc_arrays.desc), C.byref(self._p)))

    def scalar(self, key: str) -> int:
        v = self._lib.fvdbm_plan_scalar(self._p, key.encode())
        if v < 0:
            raise KeyError(key)
        return int(v)

    def array(self, key: str) -> np.ndarray:
        ptr, eb = C.c_void_p(), C.c_int32()
        n = self._lib.fvdbm_plan_array(self._p, key.encode(), C.byref(ptr), C.byref(eb))
        if n < 0:
            raise KeyError(key)
        dt = np.int32 if key in _PLAN_I32 else self._da.real
        if n == 0:
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (n * eb.value)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dt).copy()

    def close(self):
        if self._p:
            self._lib.fvdbm_plan_destroy(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_PLAN_I32 = {"ring_fcell", "pos", "ipos", "ccode", "bf_na", "bf_nb", "tn_orig", "tn_type", "node_track", "ring_off",
             "ring_cell", "s_cface", "s_csign", "s_fcell", "s_fnode"}
