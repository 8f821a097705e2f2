"""Where does the record-layout kernel (variant 4) beat the thread-per-cell AoSoA kernel (variant 1)?  Times both (and the
pair kernel, variant 3) on x-periodic triangulated squares from 50 k to 4 M cells with the engine's default schedule
(PDL chain + 50-iteration graphs below 1 M cells, two streams above), fp32 D2Q9 Lax-Wendroff: the data behind the
AUTO thresholds in csrc/api.cu (default_variant).
    python tools/variant_crossover.py [--nx 160,224,354,500,708,1000,1414] [--variants 1,4,3]"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fvdbm_jax_b200 as fb  # noqa: E402
from fvdbm_jax_b200 import _lib, meshgen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", default="160,224,354,500,708,1000,1414")
    ap.add_argument("--variants", default="1,4,3")
    ap.add_argument("--scheme", default="lax_wendroff")
    ap.add_argument("--schedules", default="", help="comma list of pdl:graph_steps pairs (e.g. 0:0,1:0,0:50,1:50) to time per "
                                                    "variant instead of the engine's default schedule")
    args = ap.parse_args()
    dyn = fb.D2Q9(tau=0.8, delta_t=0.1)
    for nx in [int(x) for x in args.nx.split(",")]:
        raw = meshgen.triangulated_square(nx, nx, seed=0, periodic_x=True)
        m = fb.Mesher(); m.import_meshpy(raw); m.calc_mesh_properties()
        cells, faces, nodes = m.to_env(dyn, flux_method=args.scheme)
        nodes = m.set_vel_node(nodes, meshgen.BOTTOM, np.array([0.0, 0.0]))
        nodes = m.set_vel_node(nodes, meshgen.TOP, np.array([0.1, 0.0]))
        c = m.cell_centers
        rho = 1 + 0.01 * np.sin(2 * np.pi * c[:, 0] / nx) * np.sin(2 * np.pi * c[:, 1] / nx)
        u = 0.05 * np.stack([np.sin(2 * np.pi * c[:, 1] / nx), np.sin(2 * np.pi * c[:, 0] / nx)], axis=1)
        cells.pdf = dyn.calc_eq(rho, u).astype(np.float32)
        n = cells.pdf.shape[0]
        env = fb.Environment(cells, faces, nodes, dtype=np.float32)
        env.init(); env.build()
        steps = int(max(200, min(4000, 4e8 // n)))
        steps -= steps % 100
        row = {"cells": n, "steps": steps, "auto_variant": env.info(_lib.INFO_VARIANT)}
        scheds = [tuple(int(y) for y in x.split(":")) for x in args.schedules.split(",") if x] or [(-1, -1)]
        for v in [int(x) for x in args.variants.split(",")]:
            env.set_option(_lib.OPT_VARIANT, v)
            for pdl, graph in scheds:
                tag = f"v{v}" if pdl < 0 else f"v{v}_pdl{pdl}_g{graph}"
                if pdl >= 0:
                    env.set_option(_lib.OPT_PDL, pdl)
                    env.set_option(_lib.OPT_GRAPH_STEPS, graph)
                env.step(200); env.sync()
                best = min(env.step_timed(steps) for _ in range(3))
                row[f"{tag}_us_per_step"] = round(best / steps * 1e3, 3)
                row[f"{tag}_MCUPS"] = round(n * steps / best / 1e3, 1)
        print(json.dumps(row), flush=True)
        env.close()


if __name__ == "__main__":
    main()
