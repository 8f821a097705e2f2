"""N>1 host logic on CPU: world_size-2 gloo processes build their local meshes, agree on exchange
lists through the same collective the NCCL path uses, and run the neighbour send/recv pattern with
the global cell ids as payload -- every halo slot must receive exactly the id it stands for."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nx, ny_per_rank, q, general=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fvdbm_jax_b200 as fb
        from fvdbm_jax_b200 import meshgen
        from fvdbm_jax_b200.distributed import HaloComm, gather_requests, strip_local_mesh
        from fvdbm_jax_b200.partition import exchange_lists, local_from_raw
        dyn = fb.D2Q9(0.8, 0.1)
        if general:       # window-based local meshes of an obstacle mesh cut into Hilbert chunks
            raw = meshgen.masked_domain(nx, ny_per_rank * world, float(nx), float(ny_per_rank * world),
                                        lambda x, y: (x - nx / 3) ** 2 + (y - ny_per_rank * world / 2) ** 2 < 6.0, seed=2)
            raw.faces = np.zeros((0, 2), np.int32)            # the decomposition never needs global faces
            bcs = lambda m, nodes: m.set_rho_node(m.set_vel_node(nodes, 5, np.array([0.0, 0.0])), 2, 0.95)
            local, fpc = local_from_raw(raw, rank, world, dyn, "lax_wendroff", boundary_conditions=bcs)
        else:
            local, fpc = strip_local_mesh(nx, ny_per_rank, rank, world, dyn, "lax_wendroff")
        from_peers = gather_requests(local)
        ps, sc, scnt, pr, rc, rcnt = exchange_lists(local, from_peers)
        comm = HaloComm(ps, scnt, pr, rcnt)
        send = torch.from_numpy(local.cell_gid[sc].astype(np.float64)).reshape(-1, 1).repeat(1, 9).contiguous()
        recv = torch.full((rc.size, 9), -1.0, dtype=torch.float64)
        for _ in range(3):                                   # repeated exchanges must not deadlock
            comm.finish(comm.start(send, recv))
        ok = bool(np.array_equal(recv[:, 0].numpy().astype(np.int64), local.cell_gid[rc]))
        if general:
            ok &= rank not in pr and len(pr) >= 1 and 1.4 < fpc < 2.0
        else:
            ok &= local.n_owned == 2 * nx * ny_per_rank and 1.4 < fpc < 1.7
            ok &= set(pr) <= {(rank - 1) % world, (rank + 1) % world} and rank not in pr
        q.put((rank, ok, int(rc.size)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world,general", [(2, False), (3, False), (2, True), (4, True)])
def test_gloo_halo_exchange_pattern(world, general):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world + 7 * general
    procs = [ctx.Process(target=_worker, args=(r, world, port, 10, 6, q, general)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert all(n > 0 for _, _, n in res)
