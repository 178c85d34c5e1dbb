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


def test_adaptive_cole_across_slab_ranks():
    """StateEquationAdaptiveCole on several slabs: `update_speed_of_sound!` (wcsph/system.jl:307-321)
    takes the maximum over ALL fluid particles, so every rank reduces its own max |v|^2 on the device,
    the words are combined (an integer MAX all-reduce between processes, the in-process mailbox here)
    and handed to the kick with tpb_set_max_speed2.  The fastest particle sits in ONE slab; all slabs
    must arrive at the single-handle speed of sound and dv."""
    import torch
    fluid, wall, _ = examples.dam_break_3d(0.05, adaptive_sound_speed=True)
    nd = fluid.ndims
    u, v = examples.perturbed_state(fluid)
    fastest = int(np.argmax(u[:, 0]))                  # a particle of the last slab
    v[fastest, :nd] = np.float32(4.0)
    semi = tp.Semidiscretization(fluid, wall, parallelization_backend=tp.B200Backend(ode_memory="device"))
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    dev = ode.u0.device
    dv_ref = torch.full((v.size,), float("nan"), dtype=torch.float32, device=dev)
    ode.f1(dv_ref, torch.from_numpy(v.reshape(-1)).to(dev), torch.from_numpy(u.reshape(-1)).to(dev), ode.p, 0.0)
    c_ref = semi.sound_speed()
    ref = dv_ref.cpu().numpy().reshape(v.shape)
    semi.close()
    assert 10.0 < c_ref < 100.0
    world = 3
    mb = LocalMailbox(world)
    slabs = [SlabSemidiscretization(fluid, wall, rank=r, world=world, device=0, transport=mb.transport(r))
             for r in range(world)]
    odes = [s.semidiscretize((0.0, 1.0), finish=False) for s in slabs]
    for s in slabs:
        s.setup_finish()
    us = [torch.from_numpy(u[s.owned_index].reshape(-1)).to(dev) for s in slabs]
    vs = [torch.from_numpy(v[s.owned_index].reshape(-1)).to(dev) for s in slabs]
    dvs = [torch.full_like(x, float("nan")) for x in vs]
    for s, uu, vv in zip(slabs, us, vs):
        s.kick_post(vv, uu)
    for s, dv in zip(slabs, dvs):
        s.kick_finish(dv)
    got = np.full_like(ref, np.nan)
    for s, dv in zip(slabs, dvs):
        assert s.semi.sound_speed() == c_ref, (s.rank, s.semi.sound_speed(), c_ref)
        got[s.owned_index] = dv.cpu().numpy().reshape(-1, v.shape[1])
    for block in (slice(0, nd), slice(nd, nd + 1)):
        err = np.abs(got[:, block] - ref[:, block]).max() / np.abs(ref[:, block]).max()
        assert err <= 1e-5, (block, err)
    # without the reduced maximum a handle with ghosts refuses to guess
    s0 = slabs[0]
    with pytest.raises(Exception):
        s0._lib.check(s0.semi._handle, s0._lib.load().tpb_kick(
            s0.semi._handle, dvs[0].data_ptr(), s0.v_ext.data_ptr(), s0.u_ext.data_ptr(), 0.0))
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


def _adaptive_worker(rank, world, port, out):
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fluid, wall, _ = examples.dam_break_3d(0.05, adaptive_sound_speed=True)
        u, v = examples.perturbed_state(fluid)
        v[int(np.argmax(u[:, 0])), :3] = np.float32(4.0)      # the fastest particle lives on the last rank
        slab = SlabSemidiscretization(fluid, wall, rank=rank, world=world, device=rank)
        ode = slab.semidiscretize((0.0, 1.0))
        dev = ode.u0.device
        ode.u0.copy_(torch.from_numpy(u[slab.owned_index].reshape(-1)).to(dev))
        ode.v0.copy_(torch.from_numpy(v[slab.owned_index].reshape(-1)).to(dev))
        dv = torch.full_like(ode.v0, float("nan"))
        for _ in range(2):
            ode.f1(dv, ode.v0, ode.u0, ode.p, 0.0)
        c = slab.semi.sound_speed()
        out.put((rank, dv.cpu().numpy().reshape(-1, v.shape[1]), slab.owned_index.copy(), c))
        dist.barrier()
        slab.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(400)
def test_two_gpu_adaptive_cole_matches_single_gpu():
    """StateEquationAdaptiveCole on two GPUs: max |v|^2 of every rank, an NCCL integer MAX all-reduce on
    the compute stream, tpb_set_max_speed2 -- both ranks arrive at the single-GPU speed of sound."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    fluid, wall, _ = examples.dam_break_3d(0.05, adaptive_sound_speed=True)
    u, v = examples.perturbed_state(fluid)
    v[int(np.argmax(u[:, 0])), :3] = np.float32(4.0)
    semi = tp.Semidiscretization(fluid, wall, parallelization_backend=tp.B200Backend(ode_memory="device"))
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    dev = ode.u0.device
    dv_d = torch.full((v.size,), float("nan"), dtype=torch.float32, device=dev)
    ode.f1(dv_d, torch.from_numpy(v.reshape(-1)).to(dev), torch.from_numpy(u.reshape(-1)).to(dev), ode.p, 0.0)
    c_ref = semi.sound_speed()
    ref = dv_d.cpu().numpy().reshape(v.shape)
    semi.close()
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_adaptive_worker, args=(r, world, port, out), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    got = np.full_like(ref, np.nan)
    for _ in range(world):
        rank, dv, owned, c = out.get(timeout=240)
        assert c == c_ref, (rank, c, c_ref)
        got[owned] = dv
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for block in (slice(0, 3), slice(3, 4)):
        err = np.abs(got[:, block] - ref[:, block]).max() / np.abs(ref[:, block]).max()
        assert err <= 1e-5, (block, err)


def _loop_worker(rank, world, port, out, n_steps, dt):
    import os
    import torch
    import torch.distributed as dist
    from trixiparticles.jl_b200.time_integration import CarpenterKennedy2N54
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fluid, wall, _ = examples.dam_break_3d(0.05)
        R = float(np.float32(2) * fluid.smoothing_length)
        # a skin of R / 20 forces several rebalances within the run
        slab = SlabSemidiscretization(fluid, wall, rank=rank, world=world, device=rank, skin=0.05 * R)
        slab.semidiscretize((0.0, n_steps * dt))
        t, v, u = slab.solve(CarpenterKennedy2N54(williamson_condition=False), dt=dt, n_steps=n_steps,
                             check_every=5)
        out.put((rank, slab.owned_index.copy(), u.cpu().numpy().reshape(-1, 3), v.cpu().numpy().reshape(-1, 4),
                 int(getattr(slab, "n_rebalances", 0)), slab.halo_transport))
        dist.barrier()
        slab.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_time_loop_with_rebalance_matches_single_gpu():
    """60 steps of CarpenterKennedy2N54 on the 3-D dam break (16 000 fluid particles), two
    processes on two GPUs with a tiny skin so that particles migrate between the slabs several
    times (`rebalance`: histogram planes, `migrate`, new handles), against the single-GPU loop.
    Stated tolerance: 1e-4 of the fluid height in position, 1e-3 of sqrt(g H) in velocity, 1e-4 of
    the reference density (Float32, different summation order per step)."""
    import socket
    import torch
    import torch.multiprocessing as mp
    from trixiparticles.jl_b200.time_integration import CarpenterKennedy2N54, solve
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n_steps = 60
    fluid, wall, _ = examples.dam_break_3d(0.05)
    semi = tp.Semidiscretization(fluid, wall, parallelization_backend=tp.B200Backend(ode_memory="device"))
    from trixiparticles.jl_b200.time_integration import StepsizeCallback
    dt = StepsizeCallback(cfl=0.9).dt(semi)
    ode = tp.semidiscretize(semi, (0.0, n_steps * dt))
    sol = solve(ode, CarpenterKennedy2N54(williamson_condition=False), dt=dt)
    assert sol.nsteps == n_steps
    ref_u = sol.u.cpu().numpy().reshape(-1, 3)
    ref_v = sol.v.cpu().numpy().reshape(-1, 4)
    semi.close()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_loop_worker, args=(r, world, port, out, n_steps, dt), daemon=True)
             for r in range(world)]
    for p in procs:
        p.start()
    got_u, got_v = np.full_like(ref_u, np.nan), np.full_like(ref_v, np.nan)
    for _ in range(world):
        rank, owned, u, v, n_reb, transport = out.get(timeout=400)
        assert n_reb >= 1, "the run was meant to rebalance"
        assert transport.startswith("peer memory")
        got_u[owned], got_v[owned] = u, v
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.isfinite(got_u).all() and np.isfinite(got_v).all()      # every particle owned exactly once
    errs = (np.abs(got_u - ref_u).max(), np.abs(got_v[:, :3] - ref_v[:, :3]).max() / np.sqrt(9.81 * 1.0),
            np.abs(got_v[:, 3] - ref_v[:, 3]).max() / 1000.0)
    assert errs[0] <= 1e-4 and errs[1] <= 1e-3 and errs[2] <= 1e-4, errs
