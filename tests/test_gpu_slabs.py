"""Slab decomposition on the GPU: several slab ranks living on one device (in-process mailbox
instead of NCCL) must reproduce the single-handle kick!/drift!.  `-m gpu` only."""
import numpy as np
import pytest

import trixiparticles.jl_b200 as tp
from trixiparticles.jl_b200 import examples
from trixiparticles.jl_b200.slabs import LocalMailbox, SlabSemidiscretization

pytestmark = pytest.mark.gpu


def single_kick(fluid, wall, u, v):
    import torch
    semi = tp.Semidiscretization(fluid, wall, parallelization_backend=tp.B200Backend(ode_memory="device"))
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    dev = ode.u0.device
    u_d, v_d = torch.from_numpy(u.reshape(-1)).to(dev), torch.from_numpy(v.reshape(-1)).to(dev)
    dv_d = torch.full_like(v_d, float("nan"))
    ode.f1(dv_d, v_d, u_d, ode.p, 0.0)
    semi.synchronize()
    out = dv_d.cpu().numpy().reshape(v.shape)
    semi.close()
    return out


@pytest.mark.parametrize("config,world", [("dam_break_2d_f64", 2), ("dam_break_3d_f32", 3), ("dam_break_3d_f32", 1)])
def test_slab_ranks_reproduce_single_gpu_kick(config, world):
    import torch
    if config == "dam_break_2d_f64":
        fluid, wall, _ = examples.dam_break_2d(40)
        tol = 1e-12
    else:
        fluid, wall, _ = examples.dam_break_3d(0.05)
        tol = 1e-5
    nd = fluid.ndims
    u, v = examples.perturbed_state(fluid)
    ref = single_kick(fluid, wall, u, v)
    mb = LocalMailbox(world)
    slabs = [SlabSemidiscretization(fluid, wall, rank=r, world=world, device=0, transport=mb.transport(r))
             for r in range(world)]
    odes = [s.semidiscretize((0.0, 1.0), finish=False) for s in slabs]
    if world > 1:
        for s in slabs:
            s.setup_finish()
    assert sum(s.n_owned for s in slabs) == fluid.nparticles
    dev = odes[0].u0.device
    us = [torch.from_numpy(u[s.owned_index].reshape(-1)).to(dev) for s in slabs]
    vs = [torch.from_numpy(v[s.owned_index].reshape(-1)).to(dev) for s in slabs]
    dvs = [torch.full_like(x, float("nan")) for x in vs]
    dus = [torch.full_like(x, float("nan")) for x in us]
    for repeat in range(2):  # second pass: buffers reused, ghost counts re-set
        if world > 1:
            for s, uu, vv in zip(slabs, us, vs):
                s.kick_post(vv, uu)
            for s, dv in zip(slabs, dvs):
                s.kick_finish(dv)
        else:
            odes[0].f1(dvs[0], vs[0], us[0], odes[0].p, 0.0)
        for s, ode, du, vv, uu in zip(slabs, odes, dus, vs, us):
            ode.f2(du, vv, uu, ode.p, 0.0)
    got = np.full_like(ref, np.nan)
    for s, dv, du, vv in zip(slabs, dvs, dus, vs):
        s.semi.synchronize()
        got[s.owned_index] = dv.cpu().numpy().reshape(-1, v.shape[1])
        assert np.array_equal(du.cpu().numpy().reshape(-1, nd), v[s.owned_index][:, :nd].astype(u.dtype))
        if world > 1:
            assert s.n_ghost > 0 and not s.needs_rebalance(us[slabs.index(s)])
    assert np.isfinite(got).all()
    for block in (slice(0, nd), slice(nd, nd + 1)):
        err = np.abs(got[:, block] - ref[:, block]).max() / np.abs(ref[:, block]).max()
        assert err <= tol, (config, world, block, err)
    for s in slabs:
        s.close()


# ------------------------------------------------------------------ one process per GPU
def _peer_worker(rank, world, port, out):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fluid, wall, _ = examples.dam_break_3d(0.05)
        u, v = examples.perturbed_state(fluid)
        results = {}
        for mode in ("peer", "nccl"):
            os.environ["TPB_HALO"] = mode
            slab = SlabSemidiscretization(fluid, wall, rank=rank, world=world, device=rank)
            ode = slab.semidiscretize((0.0, 1.0))
            assert (slab.peer is not None) == (mode == "peer"), slab.halo_transport
            dev = ode.u0.device
            ode.u0.copy_(torch.from_numpy(u[slab.owned_index].reshape(-1)).to(dev))
            ode.v0.copy_(torch.from_numpy(v[slab.owned_index].reshape(-1)).to(dev))
            dv = torch.full_like(ode.v0, float("nan"))
            for _ in range(3):  # both parities of the receive area, flags reused
                ode.f1(dv, ode.v0, ode.u0, ode.p, 0.0)
            slab.semi.synchronize()
            results[mode] = (dv.cpu().numpy().reshape(-1, v.shape[1]), slab.owned_index.copy())
            dist.barrier()
            slab.close()
        out.put((rank, results["peer"][0], results["nccl"][0], results["peer"][1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(400)
def test_peer_memory_halo_matches_nccl_and_single_gpu():
    """Two processes on two GPUs: the pack-and-store kernel + flag protocol (tpb_halo.cuh)
    delivers exactly the rows NCCL send/recv delivers (dv bit-identical) and the slabs together
    reproduce the single-GPU kick."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    fluid, wall, _ = examples.dam_break_3d(0.05)
    u, v = examples.perturbed_state(fluid)
    ref = single_kick(fluid, wall, u, v)
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, out), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    got = np.full_like(ref, np.nan)
    for _ in range(world):
        rank, dv_peer, dv_nccl, owned = out.get(timeout=240)  # a crashed worker fails the test, not hangs it
        assert np.array_equal(dv_peer, dv_nccl), f"rank {rank}: peer-memory and NCCL exchanges differ"
        got[owned] = dv_peer
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.isfinite(got).all()
    for block in (slice(0, 3), slice(3, 4)):
        err = np.abs(got[:, block] - ref[:, block]).max() / np.abs(ref[:, block]).max()
        assert err <= 1e-5, (block, err)
