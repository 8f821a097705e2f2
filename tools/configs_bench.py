"""MCUPS of the smaller BASELINE.json configs (latency-bound regime): LDC, cylinder, porous.
    python tools/configs_bench.py [--graphs 0,10,50]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fvdbm_jax_b200 as fb  # noqa: E402
from fvdbm_jax_b200 import _lib, meshgen  # noqa: E402


def quad_ldc():
    dyn = fb.D2Q13(tau=0.8, delta_t=0.1)
    return "ldc_quads_100x100_d2q13", meshgen.quad_cavity(100, 100, dyn, 0.1)


def problems():
    dyn = fb.D2Q9(tau=0.8, delta_t=0.1)
    raw = meshgen.triangulated_square(100, 100, seed=0)
    yield "ldc_tri_100x100", raw, dyn, "upwind", [("vel", 1, [0, 0]), ("vel", 2, [0, 0]), ("vel", 4, [0, 0]), ("vel", 3, [0.1, 0])]
    dyn2 = fb.D2Q9(tau=0.65, delta_t=0.1)
    cyl = [("vel", 4, [0.1, 0]), ("vel", 3, [0, 0]), ("vel", 1, [0, 0]), ("vel", 5, [0, 0]), ("rho", 2, 0.95)]
    yield "cylinder_scale9", meshgen.cylinder_channel(scale=9), dyn2, "lax_wendroff", cyl
    por = [("vel", 5, [0, 0]), ("vel", 1, [0, 0]), ("vel", 3, [0, 0]), ("rho", 4, 1.05), ("rho", 2, 0.95)]
    yield "porous_scale8.5", meshgen.porous_channel(scale=8.5), dyn2, "lax_wendroff", por


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graphs", default="0,10,50")
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--variants", default="0", help="comma list of FVDBM_VARIANT_* (0 = the engine's default)")
    ap.add_argument("--prefetch", default="-1", help="comma list of L2 prefetch distances (-1 = default)")
    ap.add_argument("--pdl", default="-1", help="comma list: 0 two-stream schedule, 1 PDL chain (-1 = default)")
    args = ap.parse_args()
    def built():
        name, (cells, faces, nodes) = quad_ldc()
        yield name, cells, faces, nodes
        for name, raw, dyn, scheme, bcs in problems():
            m = fb.Mesher(); m.import_meshpy(raw); m.calc_mesh_properties()
            cells, faces, nodes = m.to_env(dyn, flux_method=scheme)
            for kind, mk, val in bcs:
                nodes = m.set_vel_node(nodes, mk, np.array(val, dtype=float)) if kind == "vel" else m.set_rho_node(nodes, mk, val)
            yield name, cells, faces, nodes

    for name, cells, faces, nodes in built():
        n = np.asarray(cells.face_indices).shape[0]
        env = fb.Environment(cells, faces, nodes, dtype=np.float32)
        env.init(); env.build()
        combos = [(int(g), int(v), int(pf), int(pd)) for g in args.graphs.split(",") for v in args.variants.split(",")
                  for pf in args.prefetch.split(",") for pd in args.pdl.split(",")]
        for g, variant, pf, pd in combos:
            if pd >= 0:
                env.set_option(_lib.OPT_PDL, pd)
            env.set_option(_lib.OPT_GRAPH_STEPS, g)
            env.set_option(_lib.OPT_VARIANT, variant)
            if pf >= 0:
                env.set_option(_lib.OPT_PREFETCH_DIST, pf)
            env.step(200); env.sync()
            steps = args.steps if n < 1_000_000 else 400
            best = min(env.step_timed(steps) for _ in range(3))
            t0 = time.perf_counter(); env.step(steps); env.sync(); wall = (time.perf_counter() - t0) * 1e3
            # the literal notebook loop (tests/flow_over_cyl.ipynb c17): one step() per Python iteration
            t0 = time.perf_counter()
            for _ in range(steps):
                env = env.step()
            env.sync(); loop = (time.perf_counter() - t0) * 1e3
            fb.Environment.defer = False
            t0 = time.perf_counter()
            for _ in range(steps):
                env = env.step()
            env.sync(); loop_nodefer = (time.perf_counter() - t0) * 1e3
            fb.Environment.defer = True
            print(json.dumps({"config": name, "cells": n, "graph_steps": g, "variant": env.info(_lib.INFO_VARIANT), "prefetch": pf, "pdl": pd, "us_per_step": round(best / steps * 1e3, 2),
                              "wall_us_per_step": round(wall / steps * 1e3, 2),
                              "loop_us_per_step": round(loop / steps * 1e3, 2), "loop_no_defer_us_per_step": round(loop_nodefer / steps * 1e3, 2),
                              "MCUPS": round(n * steps / best / 1e3, 1), "finite": bool(np.isfinite(env.cells.rho).all())}), flush=True)
        env.close()


if __name__ == "__main__":
    main()
