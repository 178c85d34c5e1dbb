"""Host logic of the slab decomposition (trixiparticles.jl_b200/slabs.py) on CPU:
layout, wall selection, and the ghost exchange over torch.distributed (gloo, world_size 2),
checked by running the CPU oracle on each rank's local particle set."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from trixiparticles.jl_b200 import examples
from trixiparticles.jl_b200.slabs import (DistTransport, HaloExchange, LocalMailbox, local_systems,
                                          make_layout)


def _radius(fluid):
    return float(fluid.eltype.type(2) * fluid.smoothing_length)


def test_layout_partitions_every_particle_once():
    fluid, wall, _ = examples.dam_break_2d(20)
    R = _radius(fluid)
    for world in (1, 2, 3):
        layout = make_layout(fluid.initial_condition.coordinates[:, 0], world, R, R)
        owned = [local_systems(fluid, wall, layout, r)[2] for r in range(world)]
        allidx = np.sort(np.concatenate(owned))
        assert np.array_equal(allidx, np.arange(fluid.nparticles))
        counts = [len(o) for o in owned]
        assert max(counts) - min(counts) <= 2 * 20 + 1  # at most one lattice column apart
        assert layout.halo == pytest.approx(2.25 * R)  # R_fluid + R_wall + skin, skin = R / 4


def test_layout_rejects_slabs_thinner_than_halo():
    fluid, _, _ = examples.dam_break_2d(20)
    R = _radius(fluid)
    with pytest.raises(ValueError):
        make_layout(fluid.initial_condition.coordinates[:, 0], 16, R, R)


def test_local_mailbox_exchange_matches_definition():
    fluid, wall, _ = examples.dam_break_2d(20)
    R = _radius(fluid)
    world = 3
    layout = make_layout(fluid.initial_condition.coordinates[:, 0], world, R, R)
    u, v = examples.perturbed_state(fluid)
    mb = LocalMailbox(world)
    parts, halos = [], []
    for r in range(world):
        _, _, owned, _ = local_systems(fluid, wall, layout, r)
        parts.append(owned)
        halos.append(HaloExchange(layout, r, mb.transport(r), wall.coordinates))
    tu = [torch.from_numpy(u[o]) for o in parts]
    tv = [torch.from_numpy(v[o]) for o in parts]
    tm = [torch.from_numpy(fluid.mass[o]) for o in parts]
    for r in range(world):
        halos[r].setup_post(tu[r], tm[r])
    for r in range(world):
        halos[r].setup_finish(tu[r], tv[r])
    for r in range(world):
        halos[r].post(tu[r], tv[r])
    for r in range(world):
        ug, vg = halos[r].collect(tu[r], tv[r])
        assert len(ug) == halos[r].n_ghost_slots == len(halos[r].ghost_mass)
        live = ~torch.isnan(ug[:, 0])
        lo, hi = layout.planes[r], layout.planes[r + 1]
        x = u[:, 0]
        depth = np.where(x < lo, lo - x, x - hi)        # distance behind the slab's faces
        outside = (x < lo) | (x >= hi)
        within_halo = np.nonzero(outside & (depth <= layout.halo))[0]
        # needed: everything within R_fluid + skin of a face, and deeper only what is within
        # R_wall of a wall particle (Adami pressure of the wall particles the slab interacts with)
        from scipy.spatial import cKDTree
        d_wall, _ = cKDTree(wall.coordinates.astype(np.float64)).query(u.astype(np.float64))
        needed = np.nonzero(outside & ((depth < layout.direct) | ((depth < layout.halo) & (d_wall <= R))))[0]
        got = {tuple(row) for row in ug[live].numpy()}
        assert got <= {tuple(row) for row in u[within_halo]}
        assert got >= {tuple(row) for row in u[needed]}
        assert len(got) < len(within_halo)  # the second radius is a shell along the walls only
        got_mass = dict(zip((tuple(row) for row in ug[live].numpy()), halos[r].ghost_mass[live].tolist()))
        for i in needed:
            assert got_mass[tuple(u[i])] == fluid.mass[i]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, no_slip=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import adapter
        fluid, wall, _ = examples.dam_break_2d(20)
        if no_slip:   # the wall velocity needs the same fluid as the Adami pressure: no extra exchange
            import trixiparticles.jl_b200 as tp
            wall.boundary_model.viscosity = tp.ViscosityAdami(nu=0.01)
        R = _radius(fluid)
        layout = make_layout(fluid.initial_condition.coordinates[:, 0], world, R, R)
        u, v = examples.perturbed_state(fluid)
        ref = adapter.kick(fluid, wall, u, v)["dv"]
        fluid_k, wall_k, owned, widx = local_systems(fluid, wall, layout, rank)
        halo = HaloExchange(layout, rank, DistTransport(rank, world), wall.coordinates)
        tu, tv, tm = torch.from_numpy(u[owned]), torch.from_numpy(v[owned]), torch.from_numpy(fluid.mass[owned])
        assert halo.check_drift(tu)
        halo.setup(tu, tv, tm)
        ug, vg = halo.exchange(tu, tv)
        live = ~torch.isnan(ug[:, 0])            # the library skips NaN slots; so does the checker
        ug, vg, mg = ug[live], vg[live], halo.ghost_mass[live]
        # the rank's local set: owned particles first, ghosts behind them
        import copy
        from trixiparticles.jl_b200.setups import InitialCondition
        lu = np.concatenate([u[owned], ug.numpy()])
        lv = np.concatenate([v[owned], vg.numpy()])
        local = copy.copy(fluid_k)
        local.mass = np.concatenate([fluid.mass[owned], mg.numpy()])
        local.initial_condition = InitialCondition(coordinates=lu, velocity=lv[:, :2], mass=local.mass,
                                                   density=lv[:, 2], pressure=np.zeros(len(lu)),
                                                   particle_spacing=fluid.initial_condition.particle_spacing)
        got = adapter.kick(local, wall_k, lu, lv)["dv"][: len(owned)]
        err = np.abs(got - ref[owned]).max() / np.abs(ref).max()
        out.put((rank, float(err), int(len(owned)), int(len(ug))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("no_slip", [False, True])
def test_gloo_two_rank_halo_exchange_reproduces_global_kick(oracle, no_slip):
    """world_size 2 over gloo: the oracle kick on (owned + received ghosts, local wall) equals
    the global oracle kick on the owned rows to rounding (summation order differs); also with a
    no-slip wall, whose wall velocity is interpolated from the same two-radii ghost layer."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out, no_slip)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = sorted(out.get() for _ in range(world))
    assert [r[0] for r in res] == [0, 1]
    for rank, err, n_owned, n_ghost in res:
        assert err < 1e-12, (rank, err)
        assert n_owned > 0 and n_ghost > 0


def _vmax_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        res = []
        for dtype, itype in ((np.float32, np.uint32), (np.float64, np.uint64)):
            v = rng.normal(size=(1000, 3)).astype(dtype)              # the same global state on every rank
            v[123] *= 7                                               # the fastest particle lives on rank 0
            mine = v[rank::world]
            s2 = (mine[:, 0] * mine[:, 0] + mine[:, 1] * mine[:, 1]) + mine[:, 2] * mine[:, 2]
            word = torch.tensor([int(s2.max().view(itype))], dtype=torch.int64)   # what tpb_max_speed2 writes
            dist.all_reduce(word, op=dist.ReduceOp.MAX)
            g2 = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]
            res.append(int(word) == int(g2.max().view(itype)))
        out.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_gloo_max_speed_word_all_reduce():
    """StateEquationAdaptiveCole across slabs: max |v|^2 travels as the zero-extended IEEE bit pattern and
    is combined with an integer MAX all-reduce (non-negative floats order like their bit patterns)."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_vmax_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, res in sorted(out.get() for _ in range(world)):
        assert res == [True, True], (rank, res)


def _rebalance_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from trixiparticles.jl_b200.slabs import balanced_planes, migrate
        g = torch.Generator().manual_seed(1234 + rank)
        n = 1000 + 700 * rank                                   # unbalanced on purpose
        x = torch.rand(n, generator=g, dtype=torch.float64) ** 2 * (1 + rank)   # skewed, overlapping ranges
        ids = torch.arange(n, dtype=torch.int64) + 100000 * rank
        planes = balanced_planes(x, world)
        dest = torch.bucketize(x, torch.as_tensor(planes[1:-1], dtype=torch.float64), right=True)
        ids2, x2, f2 = migrate([ids.view(-1, 1), x.view(-1, 1), x.to(torch.float32).view(-1, 1) * 2], dest, rank, world)
        assert len(ids2) == len(x2) == len(f2)
        assert bool(((x2[:, 0] >= planes[rank]) & (x2[:, 0] < planes[rank + 1])).all())
        assert torch.equal(f2[:, 0], x2[:, 0].to(torch.float32) * 2)              # rows stay together
        out.put((rank, planes.tolist(), ids2[:, 0].tolist(), n))
    finally:
        dist.destroy_process_group()


def test_gloo_rebalance_planes_and_migration_conserve_particles():
    """world_size 3 over gloo: the histogram planes give equal counts and `migrate` delivers
    every particle exactly once to the rank that owns its position."""
    world = 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rebalance_worker, args=(r, world, port, out), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    total = sum(r[3] for r in res)
    assert all(r[1] == res[0][1] for r in res)                 # every rank computed the same planes
    got = sorted(i for r in res for i in r[2])
    expect = sorted(i + 100000 * r for r in range(world) for i in range(1000 + 700 * r))
    assert got == expect
    counts = [len(r[2]) for r in res]
    assert max(counts) - min(counts) <= 0.02 * total, counts   # equal counts up to a few histogram bins
