"""Kernel-configuration sweep on one GPU (development aid; prints one JSON line per config).
    python tools/sweep.py --nx 2236 --inner 50
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import fvdbm_jax_b200 as fb  # noqa: E402
from fvdbm_jax_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=2236)
    ap.add_argument("--inner", type=int, default=50)
    ap.add_argument("--scheme", default="lax_wendroff")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--reorders", default="hilbert,none,rcm")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    real = np.float32 if args.dtype == "f32" else np.float64
    m, dyn, cells, faces, nodes, t_mesh = bench.build_problem(args.nx, args.nx, args.scheme)
    n = cells.face_indices.shape[0]
    per_cell, per_face = bench.B_ALG[(args.dtype, args.scheme)]
    b_alg = per_cell + per_face * faces.n.shape[0] / n
    peak, _ = bench.measured_peak()
    print(json.dumps({"cells": n, "mesh_s": t_mesh, "b_alg": b_alg}), flush=True)
    cfgs = [dict(variant=3), dict(variant=1)]
    for tile in (128, 256, 512):
        for stages in (2, 3, 4):
            cfgs.append(dict(variant=2, tile=tile, stages=stages))
    cfgs += [dict(variant=2, tile=256, stages=3, reverse=1), dict(variant=1, reverse=1),
             dict(variant=2, tile=256, stages=3, ctas=1), dict(variant=2, tile=256, stages=3, ctas=2),
             dict(variant=2, tile=128, stages=4, ctas=4), dict(variant=2, tile=256, stages=3, graph=10),
             dict(variant=2, tile=256, stages=3, reverse=1, graph=10)]
    if args.quick:
        cfgs = [dict(variant=3), dict(variant=1), dict(variant=2, tile=256, stages=3)]
    for reorder in args.reorders.split(","):
        t0 = time.time()
        if reorder == "random":       # worst case: what an arbitrarily numbered unstructured mesh looks like
            reorder_arg = np.random.default_rng(0).permutation(n).astype(np.int32)
        else:
            reorder_arg = reorder
        env = fb.Environment(cells, faces, nodes, dtype=real, reorder=reorder_arg)
        env.init(); env.build()
        t_build = time.time() - t0
        for cfg in (cfgs if reorder == "hilbert" else cfgs[:2] + [c for c in cfgs if c.get("variant") == 2 and c.get("tile") == 256 and c.get("stages") == 3][:2]):
            if cfg["variant"] == 3 and real is np.float64:
                continue
            env.set_option(_lib.OPT_VARIANT, cfg["variant"])
            env.set_option(_lib.OPT_TILE_CELLS, cfg.get("tile", 256)).set_option(_lib.OPT_STAGES, cfg.get("stages", 3))
            env.set_option(_lib.OPT_REVERSE_SWEEP, cfg.get("reverse", 0)).set_option(_lib.OPT_GRAPH_STEPS, cfg.get("graph", 0))
            env.set_option(_lib.OPT_CTAS_PER_SM, cfg.get("ctas", 0))
            env.step(10); env.sync()
            best = min(env.step_timed(args.inner) for _ in range(3))
            it_ms = best / args.inner
            mcups = n / (it_ms * 1e-3) / 1e6
            print(json.dumps({"reorder": reorder, **cfg, "iter_ms": round(it_ms, 4), "MCUPS": round(mcups, 1),
                              "GBps_alg": round(mcups * 1e6 * b_alg / 1e9, 1), "frac": round(mcups * 1e6 * b_alg / 1e9 / peak, 4),
                              "build_s": round(t_build, 1)}), flush=True)
        env.close()


if __name__ == "__main__":
    main()
