"""Drop-in replacement for the reference ``Environment`` (/root/reference/src/environment.py:10-68).

Same constructor / factories / ``init()`` / ``step()`` / ``__repr__`` and the same attribute surface
(``env.cells.{pdf,rho,vel,pdf_eq,face_indices,face_normals}``, ``env.faces.{pdf,...}``,
``env.nodes.{pdf,rho,vel,type,...}``) so the reference notebooks run unchanged, but ``step()``
enqueues hand-written sm_100a kernels through the C ABI of libfvdbm_b200.so instead of tracing a
jit.  Differences a caller can observe:

* ``step(n=1)``: optional step count (one C call, CUDA-graph batched); still returns an env
  (``self``), so ``env = env.step()`` keeps working.  The literal notebook loop
  ``for i in range(100000): env = env.step()`` (tests/flow_over_cyl.ipynb c17) is batched too: single
  steps only bump a pending counter that is flushed -- through the same CUDA-graph path as ``step(n)`` --
  every ``defer_batch`` steps and before anything observes or changes the state (any attribute read,
  ``sync``, assignment, pickling).  ``Environment.defer = False`` restores one C call per ``step()``.
* dynamic arrays are host NumPy arrays materialised on demand, in the ORIGINAL element numbering
  and the reference's shapes, with the reference's one-step lag for rho / vel / pdf_eq / flux
  (SURVEY.md A.2).
* precision is explicit: ``Environment.dtype`` / ``dtype=`` (float32 like stock JAX, or float64).
* optional ``cells.inv_area`` (N,) attribute: 1/area of every cell multiplying the flux divergence (the
  physically consistent finite-volume update on non-unit cells).  Absent = the reference's behaviour, which
  has no area anywhere (reference containers.py:115-121); ``Mesher.cell_inv_areas()`` computes it.
* no CPU fallback: without the CUDA library or a CUDA device ``step()`` raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .containers import Cells, Faces, Nodes
from .dynamics import Dynamics, D2Q9, D2Q13
from .reorder import choose_perm

__all__ = ["Environment"]

_LATTICES = {9: D2Q9, 13: D2Q13}


def _np(a, dtype=None):
    """np.asarray that also understands the reference's CustomArray / jax arrays."""
    if hasattr(a, "data") and hasattr(a, "default_value") and not isinstance(a, np.ndarray):
        a = a.data
    return np.asarray(a, dtype=dtype)


class _View:
    """Attribute view over one container: statics are plain arrays, dynamic fields are fetched
    from the device lazily (cached until the next step) and can be assigned to."""
    _dynamic: dict = {}
    _static: tuple = ()

    def __init__(self, env, src):
        object.__setattr__(self, "_env", env)
        object.__setattr__(self, "_src", src)

    def __getattr__(self, name):
        if name in type(self)._dynamic:
            return self._env._fetch(type(self)._dynamic[name])
        return getattr(self._src, name)

    def __setattr__(self, name, value):
        if name in type(self)._dynamic:
            self._env._store(type(self)._dynamic[name], value)
        else:
            setattr(self._src, name, value)

    def init(self):
        if hasattr(self._src, "init"):
            self._src.init()

    def __repr__(self):
        keys = list(type(self)._dynamic) + list(type(self)._static)
        return repr({k: getattr(self, k) for k in keys})


class _CellsView(_View):
    _dynamic = {"pdf": "cells.pdf", "rho": "cells.rho", "vel": "cells.vel", "pdf_eq": "cells.pdf_eq"}
    _static = ("face_indices", "face_normals")
    _optional = ("inv_area",)


class _FacesView(_View):
    _dynamic = {"pdf": "faces.pdf"}
    _static = ("nodes_index", "stencil_cells_index", "stencil_dists", "n", "L", "flux_scheme")
    _optional = ("alpha",)


class _NodesView(_View):
    _dynamic = {"pdf": "nodes.pdf", "rho": "nodes.rho", "vel": "nodes.vel"}
    _static = ("type", "cells_index", "cell_dists")


_FIELD = {"cells.pdf": _lib.CELL_PDF, "cells.rho": _lib.CELL_RHO, "cells.vel": _lib.CELL_VEL,
          "cells.pdf_eq": _lib.CELL_PDF_EQ, "faces.pdf": _lib.FACE_FLUX, "nodes.pdf": _lib.NODE_PDF,
          "nodes.rho": _lib.NODE_RHO, "nodes.vel": _lib.NODE_VEL}
_SETTABLE = ("cells.pdf", "nodes.pdf", "nodes.rho", "nodes.vel")


class Environment:
    """Environment(cells, faces, nodes).init(); env = env.step()"""
    # class-level knobs, mirroring the reference's class variable ``dynamics`` (environment.py:13)
    dynamics: Dynamics = None
    dtype = np.float32          # stock JAX runs this path in float32
    device = 0
    reorder = "auto"            # 'auto' | 'hilbert' | 'rcm' | 'none' | explicit permutation
    mode = "auto"               # 'auto' | 'fused' | 'staged'
    defer = True                # batch single step() calls (see module docstring)

    def __init__(self, cells, faces, nodes, dtype=None, device=None, reorder=None, mode=None,
                 n_owned=0):
        self._attach(cells, faces, nodes)
        if dtype is not None:
            self.dtype = np.dtype(dtype).type
        if device is not None:
            self.device = device
        if reorder is not None:
            self.reorder = reorder
        if mode is not None:
            self.mode = mode
        self._n_owned = n_owned

    # ------------------------------------------------------------------ factories (environment.py:20-36)
    @classmethod
    def create(cls, num_cells, num_faces, num_nodes):
        temp = cls.__new__(cls)
        temp._attach(Cells(num_cells, cls.dynamics), Faces(num_faces, cls.dynamics), Nodes(num_nodes, cls.dynamics))
        temp._n_owned = 0
        return temp

    @classmethod
    def define(cls, cells, faces, nodes):
        temp = cls.__new__(cls)
        temp._attach(cells, faces, nodes)
        temp._n_owned = 0
        return temp

    def _attach(self, cells, faces, nodes):
        self._handle = None
        self._lib = None
        self._closed = False
        self._steps = 0
        self._pending = 0
        self._batch = 32
        self._cache = {}
        self._host = {}
        self._options = {}
        self.cells = _CellsView(self, cells)
        self.faces = _FacesView(self, faces)
        self.nodes = _NodesView(self, nodes)

    def init(self):
        """reference environment.py:38-42: finalise the containers (CustomArray -> dense arrays)."""
        self.cells.init()
        self.faces.init()
        self.nodes.init()

    # ------------------------------------------------------------------ engine
    @property
    def real(self):
        return np.dtype(self.dtype)

    def _describe(self):
        c, f, n = self.cells._src, self.faces._src, self.nodes._src
        dyn = getattr(c, "dynamics", None) or type(self).dynamics
        if dyn is None:
            raise ValueError("no dynamics: pass containers built with a Dynamics or set Environment.dynamics")
        Q = int(dyn.NUM_QUIVERS)
        if Q not in _LATTICES or not np.array_equal(np.asarray(dyn.KSI), _LATTICES[Q].KSI):
            raise ValueError("unsupported lattice: D2Q9 and D2Q13 (reference dynamics.py) are compiled in")
        real = self.real
        fi = _np(c.face_indices)
        K = int(fi.shape[1])
        N = fi.shape[0]
        scheme = getattr(f, "flux_scheme", "upwind")
        stencil = _np(f.stencil_cells_index)
        face_L = np.asarray(_np(f.L), dtype=real).reshape(-1)
        if hasattr(f, "npq"):
            raise ValueError("CCStencilKsiFaces (cc_alt_upwind) is inoperable in the reference (divides by "
                             "KSI.n_PQ = 0 -> NaN) and is not supported")
        if hasattr(f, "alpha"):
            # CCStencilFaces (reference src/faces.py:54-72): flux * cos(alpha) == face length * cos(alpha)
            face_L = (face_L * np.cos(np.asarray(_np(f.alpha), dtype=real).reshape(-1))).astype(real)
        host = self._host
        host["cells.pdf"] = np.array(_np(c.pdf), dtype=real, order="C").reshape(N, Q)
        host["cells.rho"] = np.array(_np(c.rho), dtype=real, order="C").reshape(N, 1)
        host["cells.vel"] = np.array(_np(c.vel), dtype=real, order="C").reshape(N, 2)
        host["cells.pdf_eq"] = np.array(_np(c.pdf_eq), dtype=real, order="C").reshape(N, Q)
        host["faces.pdf"] = np.array(_np(f.pdf), dtype=real, order="C").reshape(stencil.shape[0], Q)
        Pn = _np(n.type).reshape(-1).shape[0]
        host["nodes.pdf"] = np.array(_np(n.pdf), dtype=real, order="C").reshape(Pn, Q)
        host["nodes.rho"] = np.array(_np(n.rho), dtype=real, order="C").reshape(Pn, 1)
        host["nodes.vel"] = np.array(_np(n.vel), dtype=real, order="C").reshape(Pn, 2)
        perm = choose_perm(self.reorder, N, getattr(c, "centers", None), stencil) if not self._n_owned else \
            (self.reorder if isinstance(self.reorder, np.ndarray) else None)
        mode = {"auto": _lib.MODE_AUTO, "fused": _lib.MODE_FUSED, "staged": _lib.MODE_STAGED}[self.mode]
        return _lib.DescArrays(
            dtype=real, scheme=scheme, Q=Q, K=K, tau=float(dyn.tau), delta_t=float(dyn.delta_t),
            lattice_constants=_LATTICES[Q].lattice_constants(real),
            cell_face_idx=fi, cell_face_sign=_np(c.face_normals), face_cell_idx=stencil,
            face_dists=_np(f.stencil_dists), face_node_idx=_np(f.nodes_index), face_n=_np(f.n), face_L=face_L,
            node_type=_np(n.type), node_cell_idx=_np(n.cells_index), node_cell_dist=_np(n.cell_dists),
            cell_pdf=host["cells.pdf"], node_pdf=host["nodes.pdf"], node_rho=host["nodes.rho"],
            node_vel=host["nodes.vel"], cell_perm=perm, n_owned=self._n_owned, device_id=self.device, mode=mode,
            cell_inv_area=None if getattr(c, "inv_area", None) is None else _np(c.inv_area).reshape(-1))

    def build(self):
        """Create the device engine (done implicitly by the first ``step``)."""
        if self._handle is not None:
            return self
        if self._closed:
            raise RuntimeError("this Environment was closed; its device state is gone (pickle it before close() to keep it)")
        lib = _lib.load()                      # raises if the CUDA library is not built
        da = self._describe()
        h = C.c_void_p()
        _lib.check(lib.fvdbm_create(C.byref(da.desc), C.byref(h)))
        self._lib, self._handle, self._desc_arrays = lib, h, da
        self._shape = {"cells.pdf": (da.N, da.Q), "cells.rho": (da.N, 1), "cells.vel": (da.N, 2),
                       "cells.pdf_eq": (da.N, da.Q), "faces.pdf": (da.F, da.Q), "nodes.pdf": (da.P, da.Q),
                       "nodes.rho": (da.P, 1), "nodes.vel": (da.P, 2)}
        for k, v in self._options.items():
            _lib.check(lib.fvdbm_set_option(h, k, v), h)
        self._set_batch()
        return self

    def _set_batch(self):
        g = C.c_int64()
        _lib.check(self._lib.fvdbm_info(self._handle, _lib.INFO_GRAPH_STEPS, C.byref(g)), self._handle)
        self._batch = int(g.value) if g.value > 0 else 32      # one CUDA-graph launch (or 32 plain iterations) per flush

    def _flush(self):
        """Hand the deferred single steps to the engine (one C call)."""
        if self._pending:
            n, self._pending = self._pending, 0
            _lib.check(self._lib.fvdbm_step(self._handle, n), self._handle)

    def step(self, n: int = 1):
        """n iterations of reference ``Environment.step`` (environment.py:55-65); asynchronous."""
        self.build()
        n = int(n)
        if n < 0:
            raise ValueError("nsteps must be >= 0")
        if n > 0:
            self._pending += n
            self._steps += n
            self._cache.clear()
            if n != 1 or not self.defer or self._pending >= self._batch:
                self._flush()
        return self

    def step_timed(self, n: int) -> float:
        """Like ``step(n)`` but blocks and returns the device time in milliseconds (CUDA events)."""
        self.build()
        self._flush()
        ms = C.c_float()
        _lib.check(self._lib.fvdbm_step_timed(self._handle, int(n), C.byref(ms)), self._handle)
        if n > 0:
            self._steps += int(n)
            self._cache.clear()
        return float(ms.value)

    def sync(self):
        if self._handle is not None:
            self._flush()
            _lib.check(self._lib.fvdbm_sync(self._handle), self._handle)
        return self

    def wait(self, ticket=None):
        """Block until the transfer with this ticket (from ``get_into(..., wait=False)``) has filled its host
        array; ``None`` waits for everything enqueued (= ``sync``)."""
        if ticket is None:
            return self.sync()
        _lib.check(self._lib.fvdbm_wait(self._handle, int(ticket)), self._handle)
        return self

    def set_option(self, option: int, value: int):
        self._options[option] = int(value)
        if self._handle is not None:
            self._flush()
            _lib.check(self._lib.fvdbm_set_option(self._handle, option, int(value)), self._handle)
            self._set_batch()
        return self

    def set_params(self, tau: float, delta_t: float):
        self.build()
        self._flush()
        _lib.check(self._lib.fvdbm_set_params(self._handle, float(tau), float(delta_t)), self._handle)
        return self

    def count_nonfinite(self) -> int:
        """Number of cells whose populations hold a NaN/Inf (the reference has no such check: a
        diverged run is only visible in the plots)."""
        self.build()
        self._flush()
        v = C.c_int64()
        _lib.check(self._lib.fvdbm_check_finite(self._handle, C.byref(v)), self._handle)
        return int(v.value)

    def info(self, key: int) -> int:
        self.build()
        self._flush()
        v = C.c_int64()
        _lib.check(self._lib.fvdbm_info(self._handle, key, C.byref(v)), self._handle)
        return int(v.value)

    # ------------------------------------------------------------------ dynamic fields
    def _fetch(self, name):
        if self._handle is None:
            src = {"cells": self.cells, "faces": self.faces, "nodes": self.nodes}[name.split(".")[0]]._src
            return getattr(src, name.split(".")[1])
        if name in self._cache:
            return self._cache[name]
        lagged = name in ("cells.rho", "cells.vel", "cells.pdf_eq", "faces.pdf")
        if lagged and self._steps == 0:
            return self._host[name]
        self._flush()
        shape = self._shape[name]
        if name.startswith("nodes."):
            out = np.array(self._host[name], copy=True, order="C")   # untracked rows keep their values
        else:
            out = np.empty(shape, dtype=self.real)
        _lib.check(self._lib.fvdbm_get(self._handle, _FIELD[name], out.ctypes.data, out.nbytes), self._handle)
        if name.startswith("nodes."):
            self._host[name] = out
        self._cache[name] = out
        return out

    def get_into(self, name: str, out: np.ndarray, wait: bool = True):
        """Download field ``name`` ("cells.rho", ...) into a caller-owned (e.g. pinned) array.  With
        ``wait=False`` the export + D2H copy are only enqueued (they overlap later iterations) and a ticket
        for ``wait(ticket)`` is returned instead of the array."""
        self.build()
        self._flush()
        full = int(np.prod(self._shape[name]))
        owned = self._n_owned * self._shape[name][1] if (self._n_owned and name.startswith("cells.")) else full
        if out.dtype != self.real or out.size not in (full, owned) or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous array of the engine dtype holding all (or all owned) rows")
        if wait:
            _lib.check(self._lib.fvdbm_get(self._handle, _FIELD[name], out.ctypes.data, out.nbytes), self._handle)
            return out
        t = C.c_int64()
        _lib.check(self._lib.fvdbm_get_async(self._handle, _FIELD[name], out.ctypes.data, out.nbytes, C.byref(t)), self._handle)
        return int(t.value)

    # ------------------------------------------------------------------ multi-GPU primitives
    def halo_set_lists(self, send_cells: np.ndarray, recv_cells: np.ndarray):
        """Local ids (original local numbering) of the cells packed for / unpacked from peers."""
        self.build()
        self._flush()
        s = np.ascontiguousarray(send_cells, dtype=np.int32)
        r = np.ascontiguousarray(recv_cells, dtype=np.int32)
        _lib.check(self._lib.fvdbm_halo_set_lists(self._handle, s.ctypes.data, s.size, r.ctypes.data, r.size), self._handle)
        return self

    def halo_pack(self, dev_ptr: int):
        self._flush()
        _lib.check(self._lib.fvdbm_halo_pack(self._handle, C.c_void_p(dev_ptr)), self._handle)

    def halo_unpack(self, dev_ptr: int):
        self._flush()
        _lib.check(self._lib.fvdbm_halo_unpack(self._handle, C.c_void_p(dev_ptr)), self._handle)

    def comm_attach(self, nranks: int, rank: int, unique_id: bytes, peers_send, send_counts, peers_recv, recv_counts):
        """Native exchange: give the engine its own NCCL communicator and the per-peer layout of the
        halo lists; afterwards ``step(n)`` runs complete distributed iterations inside the library."""
        self.build()
        self._flush()
        buf = C.create_string_buffer(bytes(unique_id), _lib.COMM_ID_BYTES)
        _lib.check(self._lib.fvdbm_comm_init(self._handle, int(nranks), int(rank), buf), self._handle)
        sp = np.ascontiguousarray(peers_send, dtype=np.int32); sc = np.ascontiguousarray(send_counts, dtype=np.int64)
        rp = np.ascontiguousarray(peers_recv, dtype=np.int32); rc = np.ascontiguousarray(recv_counts, dtype=np.int64)
        _lib.check(self._lib.fvdbm_halo_set_peers(self._handle, sp.ctypes.data, sc.ctypes.data, sp.size,
                                                  rp.ctypes.data, rc.ctypes.data, rp.size), self._handle)
        return self

    def step_phase(self, phase: int):
        """phase 0: interior cells (no halo / boundary dependence); phase 1: node kernel + border
        cells + buffer swap (completes the step)."""
        self.build()
        self._flush()
        _lib.check(self._lib.fvdbm_step_phase(self._handle, int(phase)), self._handle)
        if phase == 1:
            self._steps += 1
            self._cache.clear()
        return self

    @property
    def stream_ptr(self) -> int:
        self.build()
        return int(self._lib.fvdbm_stream(self._handle) or 0)

    def set_cells_pdf(self, arr: np.ndarray, wait: bool = True):
        """Upload populations from a caller-owned (e.g. pinned) (N,Q) array of the engine dtype.  With
        ``wait=False`` the copy is only enqueued on the engine's upload stream (it overlaps the iterations
        already enqueued); ``arr`` must then stay untouched until ``sync()`` / a later ``wait``."""
        self.build()
        self._flush()
        full = int(np.prod(self._shape["cells.pdf"]))
        owned = self._n_owned * self._shape["cells.pdf"][1] if self._n_owned else full
        if arr.dtype != self.real or arr.size not in (full, owned) or not arr.flags.c_contiguous:
            raise ValueError("arr must be a contiguous (N,Q) or (N_owned,Q) array of the engine dtype")
        fn = self._lib.fvdbm_set if wait else self._lib.fvdbm_set_async
        _lib.check(fn(self._handle, _lib.CELL_PDF, arr.ctypes.data, arr.nbytes), self._handle)
        self._cache.pop("cells.pdf", None)
        return self

    def _store(self, name, value):
        if self._handle is None:
            src = {"cells": self.cells, "faces": self.faces, "nodes": self.nodes}[name.split(".")[0]]._src
            setattr(src, name.split(".")[1], value)
            return
        if name not in _SETTABLE:
            raise AttributeError(f"{name} is derived state (recomputed every step) and cannot be assigned")
        self._flush()
        arr = np.ascontiguousarray(_np(value), dtype=self.real).reshape(self._shape[name])
        _lib.check(self._lib.fvdbm_set(self._handle, _FIELD[name], arr.ctypes.data, arr.nbytes), self._handle)
        if name.startswith("nodes."):
            self._host[name] = np.array(arr, copy=True)
        self._cache.pop(name, None)

    # ------------------------------------------------------------------ lifetime / pickling
    def close(self):
        if getattr(self, "_handle", None) is not None:
            self._lib.fvdbm_destroy(self._handle)
            self._handle = None
            self._closed = True          # a later step()/fetch raises instead of silently restarting from t = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __getstate__(self):
        """Pickle = download (reference Mesher.to_pickle dumps (env, mesher), mesher.py:600-602)."""
        c, f, n = self.cells._src, self.faces._src, self.nodes._src
        dyn = getattr(c, "dynamics", None) or type(self).dynamics
        dyn_state = (int(dyn.NUM_QUIVERS), float(dyn.tau), float(dyn.delta_t))
        st = {"dyn": dyn_state, "dtype": np.dtype(self.dtype).str, "device": self.device,
              "reorder": self.reorder if not isinstance(self.reorder, np.ndarray) else "auto",
              "mode": self.mode, "steps": self._steps, "n_owned": self._n_owned,
              "scheme": getattr(f, "flux_scheme", "upwind")}
        for view in (self.cells, self.faces, self.nodes):
            pre = {"_CellsView": "cells", "_FacesView": "faces", "_NodesView": "nodes"}[type(view).__name__]
            for k in type(view)._dynamic:
                st[f"{pre}.{k}"] = np.array(_np(getattr(view, k)), order="C")
            for k in type(view)._static + getattr(type(view), "_optional", ()):
                if k != "flux_scheme" and hasattr(view._src, k):
                    st[f"{pre}.{k}"] = np.array(_np(getattr(view._src, k)), order="C")
        if hasattr(c, "centers"):
            st["cells.centers"] = np.array(c.centers)
        return st

    def __setstate__(self, st):
        Q, tau, dt = st["dyn"]
        dyn = _LATTICES[Q](tau, dt)
        N, F, Pn = st["cells.pdf"].shape[0], st["faces.pdf"].shape[0], st["nodes.pdf"].shape[0]
        cells, faces, nodes = Cells(N, dyn), Faces(F, dyn, st["scheme"]), Nodes(Pn, dyn)
        for key, val in st.items():
            if "." in key:
                pre, attr = key.split(".")
                setattr({"cells": cells, "faces": faces, "nodes": nodes}[pre], attr, val)
        self._attach(cells, faces, nodes)
        self.dtype = np.dtype(st["dtype"]).type
        self.device, self.reorder, self.mode = st["device"], st["reorder"], st["mode"]
        self._n_owned = st["n_owned"]

    def __repr__(self):
        return f"Environment(cells={self.cells}, faces={self.faces}, nodes={self.nodes})"
