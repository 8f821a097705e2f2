"""A/B harness for the fused cell kernels on one GPU (development aid): for one library build
(FVDBM_LIB=...) sweeps kernel variant x L2-prefetch distance on the bench.py workload and prints the
burst (best of 3 x 50 iterations) and sustained (2000 iterations after 500 of warm-up) time per iteration.
    FVDBM_LIB=gpurun_out/lib_m5.so python tools/ab_pair.py --dists 0,296,592,1184 --variants 3,1
"""
import argparse, json, os, pickle, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import fvdbm_jax_b200 as fb  # noqa: E402
from fvdbm_jax_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=2236)
    ap.add_argument("--dists", default="0")
    ap.add_argument("--variants", default="3")
    ap.add_argument("--scheme", default="lax_wendroff")
    ap.add_argument("--sustained", type=int, default=4000)
    ap.add_argument("--reorder", default="hilbert")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--tag", default=os.path.basename(os.environ.get("FVDBM_LIB", "default")))
    args = ap.parse_args()
    cache = f"/tmp/fvdbm_problem_{args.nx}_{args.scheme}.pkl"
    if os.path.exists(cache):
        cells, faces, nodes = pickle.load(open(cache, "rb"))
    else:
        m, dyn, cells, faces, nodes, _ = bench.build_problem(args.nx, args.nx, args.scheme)
        pickle.dump((cells, faces, nodes), open(cache, "wb"), protocol=4)
    n = cells.face_indices.shape[0]
    per_cell, per_face = bench.B_ALG[(args.dtype, args.scheme)]
    b_alg = per_cell + per_face * faces.n.shape[0] / n
    peak, _ = bench.measured_peak()
    env = fb.Environment(cells, faces, nodes, dtype=np.float32 if args.dtype == "f32" else np.float64, reorder=args.reorder)
    env.init(); env.build()
    for variant in [int(v) for v in args.variants.split(",")]:
        for dist in [int(d) for d in args.dists.split(",")]:
            if variant == 3 and args.dtype == "f64":
                continue
            env.set_option(_lib.OPT_VARIANT, variant).set_option(_lib.OPT_PREFETCH_DIST, dist)
            env.step(20); env.sync()
            burst = min(env.step_timed(50) for _ in range(3)) / 50
            env.step(500)
            sampler = bench.ClockSampler(0)
            sus = env.step_timed(args.sustained) / args.sustained
            clk = sampler.stop()
            f = lambda ms: round(n * b_alg / (ms * 1e-3) / 1e9 / peak, 4)
            print(json.dumps({"lib": args.tag, "dtype": args.dtype, "scheme": args.scheme, "reorder": args.reorder, "variant": variant, "prefetch_dist": dist, "burst_ms": round(burst, 4),
                              "sustained_ms": round(sus, 4), "burst_frac": f(burst), "sustained_frac": f(sus),
                              "sustained_MCUPS": round(n / sus / 1e3, 1), "sm_mhz": clk.get("sm_mhz"), "power_w": clk.get("power_w"),
                              "reasons": clk.get("reasons")}), flush=True)
            time.sleep(1.0)
    env.close()


if __name__ == "__main__":
    main()
