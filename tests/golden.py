"""Helpers shared by the tests: load tests/golden/*.npz (minted by oracle/make_golden.py from the
reference's own code) and build oracles / descriptors / environments from them."""
from __future__ import annotations

import glob
import os
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

STATE = ("cells.pdf", "cells.rho", "cells.vel", "cells.pdf_eq", "faces.pdf", "nodes.pdf", "nodes.rho", "nodes.vel")


def names(fp32=False):
    out = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))
    out = [n for n in out if n != "ldc_re100_centerlines"]          # validation data, not a step fixture
    return [n for n in out if n.endswith("_f32") == fp32]


class Case:
    def __init__(self, name):
        self.name = name
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.g = g
        self.static = {k[7:]: g[k] for k in g.files if k.startswith("static.")}
        self.init = {k[5:]: g[k] for k in g.files if k.startswith("init.")}
        self.Q, self.K = int(g["meta.Q"]), int(g["meta.K"])
        self.tau, self.delta_t = float(g["meta.tau"]), float(g["meta.delta_t"])
        self.scheme = str(g["meta.scheme"])
        self.steps = [int(s) for s in g["meta.steps"]]
        self.bits = int(g["meta.float_bits"])

    def expected(self, step, name):
        return self.g[f"s{step}.{name}"]

    def oracle(self, dtype=np.float64):
        from oracle.step_numpy import StepOracle
        return StepOracle(self.static, self.init, self.Q, self.tau, self.delta_t, self.scheme, dtype)

    def raw(self):
        g = self.g
        return types.SimpleNamespace(points=g["raw.points"], elements=g["raw.elements"], faces=g["raw.faces"],
                                     point_markers=g["raw.point_markers"])

    def containers(self):
        """Product-side containers (fvdbm_jax_b200.Cells/Faces/Nodes) filled from the fixture."""
        import fvdbm_jax_b200 as fb
        dyn = (fb.D2Q9 if self.Q == 9 else fb.D2Q13)(self.tau, self.delta_t)
        s, i = self.static, self.init
        cells = fb.Cells(s["cells.face_indices"].shape[0], dyn)
        cells.face_indices, cells.face_normals = s["cells.face_indices"], s["cells.face_normals"]
        cells.pdf, cells.rho, cells.vel, cells.pdf_eq = i["cells.pdf"], i["cells.rho"], i["cells.vel"], i["cells.pdf_eq"]
        faces = fb.Faces(s["faces.n"].shape[0], dyn, flux_scheme=self.scheme)
        faces.nodes_index, faces.stencil_cells_index = s["faces.nodes_index"], s["faces.stencil_cells_index"]
        faces.stencil_dists, faces.n, faces.L, faces.pdf = s["faces.stencil_dists"], s["faces.n"], s["faces.L"], i["faces.pdf"]
        if "faces.alpha" in s:
            faces.alpha = s["faces.alpha"]
        nodes = fb.Nodes(s["nodes.type"].shape[0], dyn)
        nodes.type, nodes.cells_index, nodes.cell_dists = s["nodes.type"], s["nodes.cells_index"], s["nodes.cell_dists"]
        nodes.pdf, nodes.rho, nodes.vel = i["nodes.pdf"], i["nodes.rho"], i["nodes.vel"]
        return cells, faces, nodes

    def desc_arrays(self, dtype, perm=None, mode=0):
        from fvdbm_jax_b200 import _lib, D2Q9, D2Q13
        s, i = self.static, self.init
        lat = (D2Q9 if self.Q == 9 else D2Q13).lattice_constants(dtype)
        face_L = np.asarray(s["faces.L"], dtype=dtype).reshape(-1)
        if "faces.alpha" in s:        # same folding as Environment._describe
            face_L = (face_L * np.cos(np.asarray(s["faces.alpha"], dtype=dtype).reshape(-1))).astype(dtype)
        return _lib.DescArrays(
            dtype=dtype, scheme=self.scheme, Q=self.Q, K=self.K, tau=self.tau, delta_t=self.delta_t,
            lattice_constants=lat, cell_face_idx=s["cells.face_indices"], cell_face_sign=s["cells.face_normals"],
            face_cell_idx=s["faces.stencil_cells_index"], face_dists=s["faces.stencil_dists"],
            face_node_idx=s["faces.nodes_index"], face_n=s["faces.n"], face_L=face_L,
            node_type=s["nodes.type"], node_cell_idx=s["nodes.cells_index"], node_cell_dist=s["nodes.cell_dists"],
            cell_pdf=i["cells.pdf"], node_pdf=i["nodes.pdf"], node_rho=i["nodes.rho"], node_vel=i["nodes.vel"],
            cell_perm=perm, mode=mode)


def rel_err(a, b):
    """max-norm error relative to the larger of max|b| and 1e-30 ... with an absolute floor so that
    fields that are identically ~0 (rest state velocities ~1e-18) do not blow up the ratio."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))), 1e-3)
    return float(np.max(np.abs(a - b))) / scale
