"""k ranks == 1 rank, bit for bit (SURVEY.md 8e iii).  All ranks run in one process on one device
(InProcessCluster): every rank is a separate C-ABI handle with owned + halo cells, the exchange
goes through the real pack / unpack kernels and the two-phase (interior, border) step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

fb = pytest.importorskip("fvdbm_jax_b200")
from fvdbm_jax_b200 import _lib, meshgen  # noqa: E402
from fvdbm_jax_b200.distributed import InProcessCluster, RankEngine, containers_from_mesh, strip_local_mesh  # noqa: E402
from fvdbm_jax_b200.partition import (GlobalMesh, exchange_lists, halo_requests, partition_sfc, partition_strips,  # noqa: E402
                                      refine_partition)


def global_problem(raw, scheme, lid=0.1, walls=(1,), nx=None, ny=None):
    m = fb.Mesher()
    m.import_meshpy(raw)
    m.calc_mesh_properties()
    dyn = fb.D2Q9(0.8, 0.1)
    cells, faces, nodes = m.to_env(dyn, scheme)
    for mk in walls:
        nodes = m.set_vel_node(nodes, mk, np.array([0.0, 0.0]))
    nodes = m.set_vel_node(nodes, 3, np.array([lid, 0.0]))
    c = m.cell_centers
    nx = nx or c[:, 0].max()
    ny = ny or c[:, 1].max()
    rho = 1 + 0.01 * np.sin(2 * np.pi * c[:, 0] / nx) * np.sin(2 * np.pi * c[:, 1] / ny)
    u = 0.05 * np.stack([np.sin(2 * np.pi * c[:, 1] / ny), np.sin(2 * np.pi * c[:, 0] / nx)], axis=1)
    cells.pdf = dyn.calc_eq(rho, u)
    return dyn, cells, faces, nodes


@pytest.mark.parametrize("scheme", ["lax_wendroff", "upwind"])
@pytest.mark.parametrize("nparts,method", [(2, "strips"), (3, "sfc"), (4, "sfc+refine")])
def test_cluster_equals_single_handle(nparts, method, scheme):
    raw = meshgen.triangulated_square(48, 40, seed=5)
    dyn, cells, faces, nodes = global_problem(raw, scheme, walls=(1, 2, 4))
    single = fb.Environment(cells, faces, nodes, dtype=np.float32, reorder="hilbert")
    single.init()
    single = single.step(25)
    g = GlobalMesh.from_containers(cells, faces, nodes)
    if method == "strips":
        part = partition_strips(g.centers, nparts)
    else:
        part = partition_sfc(g.centers, nparts)
        if method.endswith("refine"):
            part = refine_partition(g.stencil, part, nparts)
    cl = InProcessCluster(g, part, nparts, dyn, scheme, np.float32)
    assert all(e.env.info(_lib.INFO_HALO_CELLS) > 0 for e in cl.engines)
    cl.step(25)
    for name in ("pdf", "rho", "vel"):
        np.testing.assert_array_equal(cl.gather_cells(name), getattr(single.cells, name), err_msg=name)
    cl.close()
    single.close()


def test_strip_windows_equal_whole_mesh():
    """bench.py's scalable path (each rank meshes only its own window of the global square) gives
    the same bits as meshing the whole square and running it on one handle."""
    nx, nyr, world = 40, 24, 3
    dyn = fb.D2Q9(0.8, 0.1)
    raw = meshgen.strip_window(nx, nyr * world, 0, nyr * world)
    _, cells, faces, nodes = global_problem(raw, "lax_wendroff", nx=nx, ny=nyr * world)
    single = fb.Environment(cells, faces, nodes, dtype=np.float32, reorder="hilbert")
    single.init()
    single = single.step(20)
    ref = single.cells.pdf
    locals_ = [strip_local_mesh(nx, nyr, r, world, dyn, "lax_wendroff")[0] for r in range(world)]
    reqs = [halo_requests(l) for l in locals_]
    engines = [RankEngine(l, dyn, "lax_wendroff", np.float32, 0, {s: reqs[s][r] for s in range(world) if r in reqs[s]})
               for r, l in enumerate(locals_)]
    cl = InProcessCluster.__new__(InProcessCluster)
    cl.locals, cl.engines, cl.n_global, cl.Q, cl.dtype = locals_, engines, 2 * nx * nyr * world, 9, np.dtype(np.float32)
    cl.step(20)
    got = cl.gather_cells("pdf")
    np.testing.assert_array_equal(got, ref)
    cl.close()
    single.close()


def _native_general_worker(rank, world, port, q):
    """General unstructured decomposition (obstacle mesh, Hilbert chunks + greedy refinement): every
    rank builds the small global mesh, extracts its own part, and exchanges with several peers."""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fvdbm_jax_b200.distributed import DistributedEnvironment
        from fvdbm_jax_b200.partition import extract_local
        dyn, g, part = _cylinder_global(world)
        local = extract_local(g, part, rank)
        denv = DistributedEnvironment(local, dyn, "lax_wendroff", np.float32, rank, g.num_cells, 1.5, native=True)
        denv.step(30)
        q.put((rank, local.cell_gid[:local.n_owned].copy(), np.array(denv.env.cells.pdf[:local.n_owned]),
               len(denv.engine.peers_recv)))
        denv.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _cylinder_global(world):
    raw = meshgen.masked_domain(72, 36, 24.0, 12.0, lambda x, y: (x - 7.0) ** 2 + (y - 6.0) ** 2 < 4.0, seed=6)
    m = fb.Mesher()
    m.import_meshpy(raw)
    m.calc_mesh_properties()
    dyn = fb.D2Q9(0.65, 0.1)
    cells, faces, nodes = m.to_env(dyn, "lax_wendroff")
    for mk, v in ((4, (0.1, 0.0)), (3, (0.0, 0.0)), (1, (0.0, 0.0)), (5, (0.0, 0.0))):
        nodes = m.set_vel_node(nodes, mk, np.array(v))
    nodes = m.set_rho_node(nodes, 2, 0.95)
    g = GlobalMesh.from_containers(cells, faces, nodes)
    part = refine_partition(g.stencil, partition_sfc(g.centers, world), world)
    return dyn, g, part


@pytest.mark.timeout(600)
def test_native_nccl_general_partition_equals_single_handle():
    import os
    import torch
    import torch.multiprocessing as mp
    from fvdbm_jax_b200.distributed import containers_from_mesh
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    dyn, g, part = _cylinder_global(world)
    single = fb.Environment(*containers_from_mesh(g, dyn, "lax_wendroff"), dtype=np.float32, reorder="hilbert")
    single.init()
    ref = single.step(30).cells.pdf
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 1000
    procs = [ctx.Process(target=_native_general_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = np.zeros_like(ref)
    for _ in procs:
        rank, gid, pdf, npeers = q.get(timeout=500)
        got[gid] = pdf
        assert npeers >= 1
    for p in procs:
        p.join(timeout=120)
    np.testing.assert_array_equal(got, ref)
    single.close()


def _native_worker(rank, world, port, q, native=True):
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    # native: torch.distributed is plumbing only (gloo), the data path is the engine's own NCCL
    # communicator; otherwise the halo buffers travel through torch.distributed P2P (NCCL backend)
    dist.init_process_group("gloo" if native else "cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    try:
        import fvdbm_jax_b200 as fb
        from fvdbm_jax_b200.distributed import DistributedEnvironment, strip_local_mesh
        dyn = fb.D2Q9(0.8, 0.1)
        nx, nyr = 40, 24
        local, fpc = strip_local_mesh(nx, nyr, rank, world, dyn, "lax_wendroff")
        denv = DistributedEnvironment(local, dyn, "lax_wendroff", np.float32, rank, 2 * nx * nyr * world, fpc, native=native)
        denv.step(20)
        denv.sync()
        pdf = np.array(denv.env.cells.pdf[:local.n_owned])
        q.put((rank, local.cell_gid[:local.n_owned].copy(), pdf))
        denv.close()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("native", [True, False])
def test_native_nccl_exchange_equals_single_handle(native):
    """Real multi-GPU path (engine-owned NCCL communicator, or torch.distributed P2P driven from
    Python; one process per GPU) == one handle."""
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    nx, nyr = 40, 24
    raw = meshgen.strip_window(nx, nyr * world, 0, nyr * world)
    _, cells, faces, nodes = global_problem(raw, "lax_wendroff", nx=nx, ny=nyr * world)
    single = fb.Environment(cells, faces, nodes, dtype=np.float32, reorder="hilbert")
    single.init()
    ref = single.step(20).cells.pdf
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import os
    port = 29700 + os.getpid() % 1000
    port += 0 if native else 17
    procs = [ctx.Process(target=_native_worker, args=(r, world, port, q, native)) for r in range(world)]
    for p in procs:
        p.start()
    got = np.zeros_like(ref)
    for _ in procs:
        rank, gid, pdf = q.get(timeout=500)
        got[gid] = pdf
    for p in procs:
        p.join(timeout=120)
    np.testing.assert_array_equal(got, ref)
    single.close()
