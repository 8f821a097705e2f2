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
