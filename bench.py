#!/usr/bin/env python
"""bench.py -- MCUPS (D2Q9 cell-updates/s) of the FVDBM time-step hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework (CUDA)
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm: oracle C/OpenMP port

Workload (BASELINE.json configs[3]): synthetic triangulated square, nx=ny=2236 quads
-> 9 999 392 triangles, jitter 0.2 (seed 0), alternating diagonals, x-periodic, y walls (bottom
vel 0, top lid vel (0.1,0)), D2Q9(tau=0.8, dt=0.1), Lax-Wendroff, fp32, perturbed equilibrium
start.  One bench "step" = one `Environment.step(inner)` call = `inner` FVDBM iterations of the
whole mesh (default 100).  `value` = cells * inner * K / device time (CUDA events on the engine's
stream, inputs resident in HBM).  `e2e` = same call through the public API with pinned HOST
buffers: populations uploaded and rho/vel downloaded inside the timed region, every step.
Inputs exceed L2 (1.3 GB touched per iteration vs 126 MB), so no explicit L2 flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG = {("f32", "lax_wendroff"): (108, 28), ("f32", "upwind"): (108, 20),
         ("f64", "lax_wendroff"): (192, 48), ("f64", "upwind"): (192, 36)}


def build_problem(nx, ny, scheme, seed=0, periodic=True):
    import fvdbm_jax_b200 as fb
    from fvdbm_jax_b200 import meshgen
    t0 = time.time()
    raw = meshgen.triangulated_square(nx, ny, jitter=0.2, seed=seed, periodic_x=periodic)
    m = fb.Mesher()
    m.import_meshpy(raw)
    m.calc_mesh_properties()
    dyn = fb.D2Q9(tau=0.8, delta_t=0.1)
    cells, faces, nodes = m.to_env(dyn, flux_method=scheme)
    nodes = m.set_vel_node(nodes, meshgen.BOTTOM, np.array([0.0, 0.0]))
    nodes = m.set_vel_node(nodes, meshgen.TOP, np.array([0.1, 0.0]))
    if not periodic:
        nodes = m.set_vel_node(nodes, meshgen.LEFT, np.array([0.0, 0.0]))
        nodes = m.set_vel_node(nodes, meshgen.RIGHT, np.array([0.0, 0.0]))
    c = m.cell_centers
    rho = 1 + 0.01 * np.sin(2 * np.pi * c[:, 0] / nx) * np.sin(2 * np.pi * c[:, 1] / ny)
    u = 0.05 * np.stack([np.sin(2 * np.pi * c[:, 1] / ny), np.sin(2 * np.pi * c[:, 0] / nx)], axis=1)
    cells.pdf = dyn.calc_eq(rho, u).astype(np.float32)
    return m, dyn, cells, faces, nodes, time.time() - t0


def static_state(cells, faces, nodes):
    static = {"cells.face_indices": cells.face_indices, "cells.face_normals": cells.face_normals,
              "faces.nodes_index": faces.nodes_index, "faces.stencil_cells_index": faces.stencil_cells_index,
              "faces.stencil_dists": faces.stencil_dists, "faces.n": faces.n, "faces.L": faces.L,
              "nodes.type": nodes.type, "nodes.cells_index": nodes.cells_index, "nodes.cell_dists": nodes.cell_dists}
    state = {"cells.pdf": cells.pdf, "nodes.pdf": nodes.pdf, "nodes.rho": nodes.rho, "nodes.vel": nodes.vel}
    return static, state


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        out, _ = self.proc.communicate(timeout=10)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w": statistics.median(pw) if pw else None}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(scheme, problem, seconds=12.0):
    """Oracle C/OpenMP port on the host cores: a bounded number of iterations of the SAME mesh the GPU
    arm ran (BASELINE.json configs[3] by default)."""
    from oracle.step_c import COracle, threads, use_all_cores
    use_all_cores()
    m, dyn, cells, faces, nodes = problem
    static, state = static_state(cells, faces, nodes)
    o = COracle(static, state, 9, dyn.tau, dyn.delta_t, scheme, np.float32)
    n = cells.face_indices.shape[0]
    o.step(2)
    t0 = time.perf_counter(); o.step(3); per = (time.perf_counter() - t0) / 3
    iters = max(5, int(seconds / per))
    t0 = time.perf_counter(); o.step(iters); dt = time.perf_counter() - t0
    return {"value": n * iters / dt / 1e6, "unit": "MCUPS", "cores": threads(), "kind": "port",
            "sample": f"the same {n}-cell mesh x {iters} iterations, oracle/step_c.c OpenMP fp32; "
                      "JAX is not installed on the box so the reference's jitted CPU step cannot be timed"}


def square_raw_and_callbacks(nx, ny, dyn):
    """Global raw mesh (points + elements only) of the benchmark square and the BC / initial-state callbacks the
    window-based decomposition applies per rank -- the same physics as build_problem()."""
    from fvdbm_jax_b200 import meshgen
    raw = meshgen.triangulated_square(nx, ny, jitter=0.2, seed=0, periodic_x=True, with_faces=False)

    def bcs(m, nodes):
        nodes = m.set_vel_node(nodes, meshgen.BOTTOM, np.array([0.0, 0.0]))
        return m.set_vel_node(nodes, meshgen.TOP, np.array([0.1, 0.0]))

    def init(m):
        c = m.cell_centers
        rho = 1 + 0.01 * np.sin(2 * np.pi * c[:, 0] / nx) * np.sin(2 * np.pi * c[:, 1] / ny)
        u = 0.05 * np.stack([np.sin(2 * np.pi * c[:, 1] / ny), np.sin(2 * np.pi * c[:, 0] / nx)], axis=1)
        return dyn.calc_eq(rho, u)
    return raw, bcs, init


def porous_raw_and_callbacks(scale, dyn):
    """BASELINE.json configs[2]: the porous obstacle field (reference outlines), BCs of tests/porous_flow.ipynb."""
    from fvdbm_jax_b200 import meshgen
    raw = meshgen.porous_channel(scale=scale)
    return raw, (lambda m, nodes: meshgen.porous_boundary_conditions(m, nodes)), None


def pinned_upload_buffer(shape, tdt, write_combined):
    """Pinned host buffer the producer only writes and the GPU only reads.  write_combined: cudaHostAlloc with
    cudaHostAllocWriteCombined (no CPU-cache snooping on the PCIe reads); falls back to torch's pin_memory()."""
    import torch
    if write_combined:
        try:
            import ctypes as C
            rt = C.CDLL("libcudart.so.12")
            ptr = C.c_void_p()
            nbytes = int(np.prod(shape)) * torch.empty((), dtype=tdt).element_size()
            if rt.cudaHostAlloc(C.byref(ptr), C.c_size_t(nbytes), C.c_uint(0x04)) == 0:      # cudaHostAllocWriteCombined
                buf = (C.c_char * nbytes).from_address(ptr.value)
                return torch.frombuffer(buf, dtype=tdt).reshape(shape)
        except OSError:
            pass
    return torch.empty(shape, dtype=tdt).pin_memory()


def bind_to_gpu_numa_node(index):
    """Pin this process (and therefore its first-touch pinned buffers) to the CPUs of the GPU's NUMA node."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def multi_gpu_check(rank, world, local_dev, scheme, real, nx=40, rows=24, iters=20, partition="strips"):
    """N>1 only, before the timed region: a small problem (nx x rows quads per rank) stepped through
    the SAME native path as the benchmark (engine-owned NCCL send/recv inside fvdbm_step, the same kind of
    partition) must equal, bit for bit, the whole mesh stepped by one handle on rank 0."""
    import torch.distributed as dist
    import fvdbm_jax_b200 as fb
    from fvdbm_jax_b200.distributed import DistributedEnvironment, strip_local_mesh, containers_from_mesh
    dyn = fb.D2Q9(tau=0.8, delta_t=0.1)
    if partition == "sfc":
        return multi_gpu_check_sfc(rank, world, local_dev, scheme, real, dyn, nx, rows * world, iters)
    lm, fpc = strip_local_mesh(nx, rows, rank, world, dyn, scheme)
    denv = DistributedEnvironment(lm, dyn, scheme, real, local_dev, 2 * nx * rows * world, fpc, native=True)
    denv.step(iters)
    denv.sync()
    mine = (lm.cell_gid[:lm.n_owned].copy(), np.array(denv.env.cells.pdf[:lm.n_owned]), len(denv.engine.peers_recv))
    denv.close()
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    out = None
    if rank == 0:
        whole, _ = strip_local_mesh(nx, rows * world, 0, 1, dyn, scheme)
        env = fb.Environment(*containers_from_mesh(whole.mesh, dyn, scheme), dtype=real, device=local_dev, reorder="hilbert")
        env.init()
        ref = np.empty((whole.n_owned, 9), dtype=real)
        ref[whole.cell_gid[:whole.n_owned]] = env.step(iters).cells.pdf[:whole.n_owned]
        env.close()
        got = np.zeros_like(ref)
        for gid, pdf, _ in parts:
            got[gid] = pdf
        out = {"bitwise_equal": bool(np.array_equal(got, ref)), "ranks": world, "cells": int(ref.shape[0]),
               "iterations": iters, "peers_per_rank": [int(p[2]) for p in parts],
               "what": f"{nx}x{rows}-quad strip per rank, native NCCL path vs one handle on rank 0"}
    dist.barrier()
    return out


def multi_gpu_check_sfc(rank, world, local_dev, scheme, real, dyn, nx, ny, iters):
    import torch.distributed as dist
    import fvdbm_jax_b200 as fb
    from fvdbm_jax_b200.distributed import DistributedEnvironment
    raw, bcs, init = square_raw_and_callbacks(nx, ny, dyn)
    denv = DistributedEnvironment.from_raw(raw, dyn, scheme, real, rank, world, local_dev, bcs, init)
    lm = denv.engine.local
    denv.step(iters)
    denv.sync()
    mine = (lm.cell_gid[:lm.n_owned].copy(), np.array(denv.env.cells.pdf[:lm.n_owned]), len(denv.engine.peers_recv))
    denv.close()
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    out = None
    if rank == 0:
        m, _, cells, faces, nodes, _ = build_problem(nx, ny, scheme)
        env = fb.Environment(cells, faces, nodes, dtype=real, device=local_dev, reorder="hilbert")
        env.init()
        ref = np.array(env.step(iters).cells.pdf)
        env.close()
        got = np.zeros_like(ref)
        for gid, pdf, _ in parts:
            got[gid] = pdf
        out = {"bitwise_equal": bool(np.array_equal(got, ref)), "ranks": world, "cells": int(ref.shape[0]),
               "iterations": iters, "peers_per_rank": [int(p[2]) for p in parts],
               "what": f"{nx}x{ny}-quad square cut into Hilbert chunks (window-based local meshes), native NCCL path vs one handle on rank 0"}
    dist.barrier()
    return out


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (oracle port; JAX absent)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.step_c import COracle, threads, use_all_cores
    use_all_cores()
    nx = args.ref_nx if args.ref_nx > 0 else args.nx            # default: the GPU arm's own mesh
    m, dyn, cells, faces, nodes, _ = build_problem(nx, nx, args.scheme)
    static, state = static_state(cells, faces, nodes)
    o = COracle(static, state, 9, dyn.tau, dyn.delta_t, args.scheme, np.float32)
    n = cells.face_indices.shape[0]
    t0 = time.perf_counter(); o.step(1); t1 = time.perf_counter() - t0
    # bounded sample: as many iterations per bench step as fit ~100 s for the whole --steps/--warmup run
    inner = max(1, min(args.ref_inner, int(100.0 / ((args.steps + args.warmup) * t1))))
    for _ in range(args.warmup):
        o.step(inner)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.step(inner)
    dt = time.perf_counter() - t0
    val = n * inner * args.steps / dt / 1e6
    sample = (f"{n} cells (nx=ny={nx}, {'the GPU arm mesh' if nx == args.nx else 'REDUCED mesh, same family'}) x {inner} "
              f"iterations per bench step (the GPU arm runs {args.inner}), oracle/step_c.c OpenMP fp32 on {threads()} threads")
    line = {"impl": "reference", "metric": "MCUPS", "value": val, "unit": "MCUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, nx, n, args.inner),
            "cpu_baseline": {"value": val, "unit": "MCUPS", "cores": threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "MCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args, nx, cells, inner):
    if getattr(args, "mesh", "square") == "porous" and args.impl == "b200":
        return {"workload": f"porous obstacle field (reference tests/test_bmp.mat outlines, scale {args.porous_scale}), {cells} cells on this GPU, "
                            f"velocity walls + density inlet/outlet, D2Q9 tau=0.65 dt=0.1, {args.scheme}",
                "inner_iterations_per_step": inner, "scheme": args.scheme, "l2": "inputs exceed L2 (no flush needed)",
                "reorder": args.reorder, "partition": args.partition}
    if getattr(args, "scaling", "weak") == "strong" and args.impl == "b200":
        shape = f"fixed global mesh nx={nx} x ny={args.ny_total} quads ({2 * nx * args.ny_total} cells) cut into one strip per GPU"
    else:
        shape = f"nx=ny={nx} ({cells if cells else 2 * nx * nx} cells) per GPU (weak scaling: strips of one global square)"
    return {"workload": f"synthetic triangulated square, {shape}, x-periodic + y walls (lid 0.1), D2Q9 tau=0.8 dt=0.1, {args.scheme}",
            "inner_iterations_per_step": inner, "scheme": args.scheme, "l2": "inputs exceed L2 (no flush needed)",
            "reorder": args.reorder, "variant": args.variant, "tile_cells": args.tile, "stages": args.stages,
            "reverse_sweep": args.reverse, "graph_steps": args.graph, "partition": getattr(args, "partition", "strips")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=2236, help="quads per side (cells = 2*nx*nx per GPU)")
    ap.add_argument("--inner", type=int, default=100, help="FVDBM iterations per bench step")
    ap.add_argument("--scheme", default="lax_wendroff", choices=["lax_wendroff", "upwind"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--reorder", default="hilbert")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--reverse", type=int, default=-1)
    ap.add_argument("--graph", type=int, default=-1)
    ap.add_argument("--ctas", type=int, default=-1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-multi-gpu-check", action="store_true")
    ap.add_argument("--wc-upload", action="store_true", help="e2e: write-combined pinned memory for the upload buffer")
    ap.add_argument("--ref-nx", type=int, default=0, help="reference arm mesh (0 = same as --nx)")
    ap.add_argument("--ref-inner", type=int, default=10)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: nx x nx quads per GPU (default); strong: a fixed nx x ny_total global mesh split into strips")
    ap.add_argument("--ny-total", type=int, default=8944, help="strong scaling: quad rows of the global mesh")
    ap.add_argument("--partition", default="strips", choices=["strips", "sfc"],
                    help="N>1: strips = 1-D strips, each rank meshes its own window of the analytic square; sfc = general path: "
                         "Hilbert-chunk owners from the raw mesh, window-based local meshes (4-6 peers per rank)")
    ap.add_argument("--mesh", default="square", choices=["square", "porous"],
                    help="porous: BASELINE configs[2] obstacle field (fixed size -> strong scaling over the GPUs, sfc partition)")
    ap.add_argument("--porous-scale", type=float, default=8.5)
    ap.add_argument("--config4", action="store_true",
                    help="BASELINE configs[4]: the fixed 99 993 920-cell square (2236 x 22360 quads) cut over the GPUs (= --scaling strong --ny-total 22360)")
    ap.add_argument("--torch-exchange", action="store_true",
                    help="N>1: drive the halo exchange from Python with torch.distributed P2P instead of the engine's own NCCL communicator")
    args = ap.parse_args()
    if args.config4:
        args.scaling, args.ny_total = "strong", 22360
    if args.mesh == "porous":
        args.scaling, args.partition = "strong", "sfc"
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import fvdbm_jax_b200 as fb
    from fvdbm_jax_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    out_fd = os.dup(1)                 # NCCL prints its version banner on stdout: keep stdout clean for
    os.dup2(2, 1)                      # the single JSON line by routing everything else to stderr
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local)
    real = np.float32 if args.dtype == "f32" else np.float64
    problem, t_mesh, t_plan, mg_check = None, None, None, None

    if args.scaling == "strong" and args.ny_total % world:
        raise SystemExit("--ny-total must be divisible by the number of GPUs")
    part_stats = None
    if world > 1:
        from fvdbm_jax_b200.distributed import DistributedEnvironment
        os.environ.setdefault("FVDBM_PLAN_THREADS", str(max(1, len(os.sched_getaffinity(0)) // world)))
        if not args.torch_exchange and not args.no_multi_gpu_check:
            mg_check = multi_gpu_check(rank, world, local, args.scheme, real, partition=args.partition)
        t0 = time.time()
        if args.mesh == "porous":
            dyn = fb.D2Q9(tau=0.65, delta_t=0.1)
            raw, bcs, init = porous_raw_and_callbacks(args.porous_scale, dyn)
            denv = DistributedEnvironment.from_raw(raw, dyn, args.scheme, real, rank, world, local, bcs, init,
                                                   native=not args.torch_exchange)
        elif args.partition == "sfc":
            dyn = fb.D2Q9(tau=0.8, delta_t=0.1)
            ny = args.nx * world if args.scaling == "weak" else args.ny_total
            raw, bcs, init = square_raw_and_callbacks(args.nx, ny, dyn)
            denv = DistributedEnvironment.from_raw(raw, dyn, args.scheme, real, rank, world, local, bcs, init,
                                                   native=not args.torch_exchange)
            del raw
        else:
            rows = args.nx if args.scaling == "weak" else args.ny_total // world
            denv = DistributedEnvironment.strips(args.nx, rows, args.scheme, real, rank, world, local,
                                                 native=not args.torch_exchange)
        t_mesh = time.time() - t0
        part_stats = denv.partition_stats()
        env, n_local, n_global = denv, denv.n_owned, denv.n_global
        stepper = denv
    elif args.mesh == "porous":
        dyn = fb.D2Q9(tau=0.65, delta_t=0.1)
        raw, bcs, init = porous_raw_and_callbacks(args.porous_scale, dyn)
        m = fb.Mesher()
        m.import_meshpy(raw)
        m.calc_mesh_properties()
        cells, faces, nodes = m.to_env(dyn, flux_method=args.scheme)
        nodes = bcs(m, nodes)
        env = fb.Environment(cells, faces, nodes, dtype=real, device=local, reorder=args.reorder)
        env.init()
        env.build()
        n_local = n_global = cells.face_indices.shape[0]
        stepper = env
    elif args.scaling == "strong":
        # same mesh family as the multi-GPU strips (hash jitter), whole mesh on one GPU
        from fvdbm_jax_b200.distributed import strip_local_mesh, containers_from_mesh
        dyn = fb.D2Q9(tau=0.8, delta_t=0.1)
        lm, fpc_strong = strip_local_mesh(args.nx, args.ny_total, 0, 1, dyn, args.scheme)
        cells, faces, nodes = containers_from_mesh(lm.mesh, dyn, args.scheme)
        env = fb.Environment(cells, faces, nodes, dtype=real, device=local, reorder=args.reorder)
        env.init()
        env.build()
        n_local = n_global = lm.n_owned
        stepper = env
    else:
        m, dyn, cells, faces, nodes, t_mesh = build_problem(args.nx, args.nx, args.scheme)
        problem = (m, dyn, cells, faces, nodes)
        env = fb.Environment(cells, faces, nodes, dtype=real, device=local, reorder=args.reorder)
        env.init()
        t0 = time.time()
        env.build()
        t_plan = time.time() - t0
        n_local = n_global = cells.face_indices.shape[0]
        stepper = env
    for opt, val in ((_lib.OPT_VARIANT, args.variant), (_lib.OPT_TILE_CELLS, args.tile), (_lib.OPT_STAGES, args.stages)):
        if val > 0:
            stepper.set_option(opt, val)
    for opt, val in ((_lib.OPT_REVERSE_SWEEP, args.reverse), (_lib.OPT_GRAPH_STEPS, args.graph), (_lib.OPT_CTAS_PER_SM, args.ctas)):
        if val >= 0:
            stepper.set_option(opt, val)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    inner = args.inner
    launches0 = stepper.info(_lib.INFO_LAUNCHES)
    for _ in range(args.warmup):
        stepper.step(inner)
    barrier()
    launches1 = stepper.info(_lib.INFO_LAUNCHES)
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    ms = stepper.step_timed(inner * args.steps)        # CUDA events on the engine's stream, K steps
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = stepper.info(_lib.INFO_LAUNCHES) - launches1
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = n_global * inner * args.steps / (ms * 1e-3) / 1e6

    # ---- e2e through the public API with pinned host buffers --------------------------------------
    # Every step: populations uploaded from pinned host memory, `inner` iterations, rho + vel downloaded to
    # pinned host memory and read by the host.  The calls are the asynchronous ones of the public API
    # (set_cells_pdf(wait=False) / get_into(wait=False) / wait(ticket)): step k+1's upload and step k-1's
    # download overlap step k's iterations on the engine's two copy streams; the host consumes result k-1
    # while step k runs.  All copies of all steps are inside the timed region.
    e2e = None
    if not args.no_e2e:
        tdt = torch.float32 if real is np.float32 else torch.float64
        host_pdf = pinned_upload_buffer((n_local, 9), tdt, args.wc_upload)
        rho_out = [torch.empty((n_local, 1), dtype=tdt).pin_memory() for _ in range(2)]
        vel_out = [torch.empty((n_local, 2), dtype=tdt).pin_memory() for _ in range(2)]
        stepper.get_into("cells.pdf", host_pdf.numpy())
        reps = max(3, min(args.steps, 10))
        seen = []

        def run(count):
            pending = None
            for k in range(count):
                stepper.set_cells_pdf(host_pdf.numpy(), wait=False)
                stepper.step(inner)
                stepper.get_into("cells.rho", rho_out[k & 1].numpy(), wait=False)
                ticket = stepper.get_into("cells.vel", vel_out[k & 1].numpy(), wait=False)
                if pending is not None:                       # consume the previous step's result on the host
                    stepper.wait(pending[0])
                    seen.append(float(rho_out[pending[1]][0, 0]) + float(vel_out[pending[1]][-1, 1]))
                pending = (ticket, k & 1)
            stepper.wait(pending[0])
            seen.append(float(rho_out[pending[1]][0, 0]) + float(vel_out[pending[1]][-1, 1]))
            stepper.sync()
        run(2)
        barrier()
        t0 = time.perf_counter()
        run(reps)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert all(np.isfinite(seen))
        e2e = {"value": n_global * inner * reps / dt / 1e6, "unit": "MCUPS",
               "h2d_bytes_per_step": int(host_pdf.numel() * host_pdf.element_size()),
               "d2h_bytes_per_step": int((rho_out[0].numel() + vel_out[0].numel()) * rho_out[0].element_size()),
               "steps": reps, "numa_node": numa_node, "upload_buffer": "pinned write-combined" if args.wc_upload else "pinned",
               "note": f"per GPU and step: cells.pdf <- pinned host (async upload stream); step({inner}); cells.rho, cells.vel -> "
                       "pinned host (one export pass, async download stream); host reads result k-1 while step k runs; wall clock"}

    if world > 1:
        # explicit teardown: engine communicator first, then a barrier so nobody tears NCCL down under
        # a peer; os._exit skips interpreter-exit destructors that could block on a departed peer
        import torch.distributed as dist
        n_launch_info = stepper.info(_lib.INFO_VARIANT)
        f_over_n_dist = denv.faces_per_cell
        stepper.close()
        dist.barrier()
        if rank != 0:
            sys.stderr.flush()
            os._exit(0)
    per_cell, per_face = B_ALG[(args.dtype, args.scheme)]
    if world == 1:
        f_over_n = faces.n.shape[0] / n_local
        n_launch_info = stepper.info(_lib.INFO_VARIANT)
    else:
        f_over_n = f_over_n_dist
    b_alg = per_cell + per_face * f_over_n
    peak, peak_src = measured_peak()
    iter_ms = ms / (inner * args.steps)
    achieved = n_local * b_alg / (iter_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": {1: "k_fused_direct", 2: "k_fused_tma", 3: "k_fused_pair", 4: "k_fused_rec"}.get(n_launch_info, "?"),
                "algorithmic_bytes_per_cell_update": b_alg, "cells_per_launch": n_local, "avg_launch_ms": iter_ms,
                "peak_source": peak_src,
                "note": "avg_launch_ms = event time / iterations (includes the O(sqrt N) node + border launches that run "
                        "concurrently); traffic = DRAM bytes of one launch from the committed ncu capture named in "
                        "traffic_source, null when that capture is of another kernel / size"}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):          # one ncu --set full capture of the dominant kernel; says which commit it was taken at
        try:
            t = json.load(open(prof))
            if t.get("kernel", "").startswith(roofline["kernel"]) and t.get("cells_per_launch") == n_local:
                roofline["traffic"] = t.get("dram_bytes_per_launch")
                roofline["traffic_source"] = f"{t.get('capture')} at commit {t.get('captured_at_commit')} (not re-measured in this run)"
        except Exception:
            pass
    line = {"metric": "MCUPS", "value": value, "unit": "MCUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, args.nx, n_local, inner), "roofline": roofline,
            "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e}
    if mg_check is not None:
        line["multi_gpu_check"] = mg_check
    if part_stats is not None:
        line["partition"] = {"method": args.partition, "peers_per_rank": [p["peers"] for p in part_stats],
                             "halo_cells_per_rank": [p["halo"] for p in part_stats],
                             "send_bytes_per_iteration_per_rank": [p["send_bytes_per_iteration"] for p in part_stats],
                             "owned_cells_per_rank": [p["owned"] for p in part_stats], "local_mesh_build_s": round(t_mesh, 2)}
    if t_mesh is not None and t_plan is not None:
        line["host_build_s"] = {"mesher": round(t_mesh, 2), "planner_and_upload": round(t_plan, 2)}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.scheme, problem) if problem is not None else None
    os.write(out_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
