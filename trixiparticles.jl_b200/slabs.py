"""1-D slab decomposition of the WCSPH right-hand side over the GPUs of one box.

The reference has no distributed path at all; this is the B200
design for it.  One process per GPU owns the fluid particles of one slab `[x_k, x_{k+1})`.  Per
`kick!`:

  1. every rank sends one fixed-size message `(x, v, rho)` per slab face (NCCL send/recv over
     NVLink): a row per *candidate* particle (owned, within `halo + skin` of the face when the
     partition was made); candidates currently outside the neighbour's ghost layer carry
     `x = NaN`.  Sizes and slot order are fixed between rebalances, masses travel once -- no
     host synchronisation, deterministic summation order;
  2. the received rows land behind the owned particles in the rank's extended `u`/`v` buffers;
     `tpb_set_fluid_count(n_owned + n_slots, n_owned)` (once) told the library that the tail is
     neighbours-only and that NaN rows there are empty slots;
  3. `tpb_kick` runs the usual rebuild + Adami + interact on the local set.

`halo = R_fluid + R_wall + skin`: a wall particle within `R_fluid` of an owned fluid particle needs
all fluid within `R_wall` of itself for its Adami pressure, so the fluid ghost layer reaches two
search radii deep and the wall pressure needs no second exchange -- but beyond `R_fluid + skin`
only the fluid within `R_wall` of a wall particle is needed, so the second radius of the layer is
a thin shell along the tank walls, not a full slice (`candidate_mask`).  `skin` is how far a
particle may move (and an owned particle drift out of its slab) before `rebalance` (re-partition
by position, a `SortingCallback`-like event: particle <-> ODE index may only change between time
steps) becomes necessary.  Wall particles are static: each rank keeps those within
`R_fluid + skin` of its slab.

The partition / selection / exchange logic is tensor-backend agnostic (CPU tensors + gloo in the
tests, CUDA tensors + NCCL on the box); only `SlabSemidiscretization` needs the CUDA library.
"""
from __future__ import annotations

import copy
import ctypes as C
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional

import numpy as np

from .setups import InitialCondition


# ------------------------------------------------------------------ layout (pure numpy)
@dataclass
class SlabLayout:
    planes: np.ndarray   # world + 1 slab faces along x; planes[0] = -inf, planes[-1] = +inf
    halo: float          # fluid ghost layer width (deepest ghost: R_fluid + R_wall + skin)
    wall_reach: float    # wall particles kept within this distance of the slab
    skin: float
    direct: float = 0.0  # R_fluid + skin: every particle this close to the face is a ghost
    near_wall: float = 0.0  # R_wall: deeper ghosts are only those this close to a wall particle

    @property
    def world(self) -> int:
        return len(self.planes) - 1

    def owner(self, x: np.ndarray) -> np.ndarray:
        """Rank owning coordinate(s) x: slab k is [planes[k], planes[k+1])."""
        return np.clip(np.searchsorted(self.planes, x, side="right") - 1, 0, self.world - 1)


def make_layout(x_fluid: np.ndarray, world: int, radius_fluid: float, radius_wall: float,
                skin: Optional[float] = None, per_column: Optional[int] = None) -> SlabLayout:
    """Faces between lattice columns such that every slab holds the same number of fluid
    particles (+-1 column): the dam-break column fills only part of the tank, equal-width slabs
    would leave most GPUs idle."""
    skin = 0.25 * float(radius_fluid) if skin is None else float(skin)
    xs = np.sort(np.asarray(x_fluid, dtype=np.float64))
    planes = [-np.inf]
    for k in range(1, world):
        if per_column is not None:
            # x_fluid = the distinct column coordinates of a lattice with `per_column` particles
            # each: the same faces as from the full coordinate array, without building it
            n = len(xs) * per_column
            i = min(max(int(round(n * k / world)), 1), n - 1)
            c = min(max(-(-i // per_column), 1), len(xs) - 1)
            planes.append(0.5 * (xs[c - 1] + xs[c]))
            continue
        i = int(round(len(xs) * k / world))
        i = min(max(i, 1), len(xs) - 1)
        # move to a column boundary: first index whose x differs from its predecessor
        while i < len(xs) - 1 and xs[i] == xs[i - 1]:
            i += 1
        planes.append(0.5 * (xs[i - 1] + xs[i]))
    planes.append(np.inf)
    planes = np.asarray(planes)
    halo = float(radius_fluid) + float(radius_wall) + skin
    widths = np.diff(planes[1:-1]) if world > 2 else np.array([np.inf])
    # an owner from rank k+2 may sit `skin` inside slab k+1 before the next rebalance
    if world > 1 and (np.any(np.diff(planes) <= 0) or np.any(widths < halo + skin)):
        raise ValueError("slabs are thinner than the ghost layer: use fewer ranks or a larger problem")
    return SlabLayout(planes=planes, halo=halo, wall_reach=float(radius_fluid) + skin, skin=skin,
                      direct=float(radius_fluid) + skin, near_wall=float(radius_wall))


def wall_tree(wall_coordinates):
    """KD-tree of the wall particles for `candidate_mask` (None without a wall)."""
    if wall_coordinates is None or hasattr(wall_coordinates, "query"):
        return wall_coordinates
    if len(wall_coordinates) == 0:
        return None
    from scipy.spatial import cKDTree
    return cKDTree(np.asarray(wall_coordinates, dtype=np.float64))


def candidate_mask(coords: np.ndarray, face: float, side: int, layout: SlabLayout, tree) -> np.ndarray:
    """Which particles (owned by the slab on the far side of `face`) are ghost *candidates* of
    the slab on the near side: side = -1: the owner lies to the right of the face (x >= face),
    side = +1: to the left (x < face).  A candidate is within `direct + skin` of the face, or
    within `halo + skin` of it and within `near_wall + skin` of a wall particle; the extra
    `skin` covers the motion until the next rebalance (HaloExchange.check_drift)."""
    coords = np.asarray(coords, dtype=np.float64)
    x = coords[:, 0]
    depth = (x - face) if side < 0 else (face - x)   # > 0 (>= 0) inside the owner's slab
    owned = depth >= 0 if side < 0 else depth > 0
    mask = owned & (depth < layout.direct + layout.skin)
    if tree is not None:
        band = np.nonzero(owned & ~mask & (depth < layout.halo + layout.skin))[0]
        if len(band):
            d, _ = tree.query(coords[band], k=1, distance_upper_bound=layout.near_wall + layout.skin)
            mask[band[np.isfinite(d)]] = True
    return mask


def subset_ic(ic: InitialCondition, idx: np.ndarray) -> InitialCondition:
    return InitialCondition(coordinates=np.ascontiguousarray(ic.coordinates[idx]),
                            velocity=np.ascontiguousarray(ic.velocity[idx]),
                            mass=np.ascontiguousarray(ic.mass[idx]),
                            density=np.ascontiguousarray(ic.density[idx]),
                            pressure=np.ascontiguousarray(ic.pressure[idx]),
                            particle_spacing=ic.particle_spacing)


def local_systems(fluid, wall, layout: SlabLayout, rank: int):
    """The rank's own systems: owned fluid particles, wall particles within reach of the slab.
    Returns (fluid_k, wall_k, owned_index, wall_index) with indices into the global systems."""
    x = fluid.initial_condition.coordinates[:, 0].astype(np.float64)
    owned = np.nonzero(layout.owner(x) == rank)[0]
    fluid_k = copy.copy(fluid)
    fluid_k.initial_condition = subset_ic(fluid.initial_condition, owned)
    fluid_k.mass = fluid.mass[owned].copy()
    fluid_k.pressure = np.zeros(len(owned), dtype=fluid.eltype)
    wall_k, widx = local_wall(wall, layout, rank)
    return fluid_k, wall_k, owned, widx


def local_wall(wall, layout: SlabLayout, rank: int):
    """The wall particles within `wall_reach` of slab `rank` (static: a subset of the global wall)."""
    if wall is None:
        return None, None
    xw = wall.coordinates[:, 0].astype(np.float64)
    lo, hi = layout.planes[rank] - layout.wall_reach, layout.planes[rank + 1] + layout.wall_reach
    widx = np.nonzero((xw >= lo) & (xw <= hi))[0]
    wall_k = copy.copy(wall)
    wall_k.initial_condition = subset_ic(wall.initial_condition, widx)
    wall_k.coordinates = wall_k.initial_condition.coordinates
    model = copy.copy(wall.boundary_model)
    model.initial_density = wall.boundary_model.initial_density[widx].copy()
    model.hydrodynamic_mass = wall.boundary_model.hydrodynamic_mass[widx].copy()
    model.pressure = np.zeros(len(widx), dtype=wall.eltype)
    model.cache = dict(density=model.initial_density.copy(), volume=np.zeros(len(widx), dtype=wall.eltype))
    wall_k.boundary_model = model
    return wall_k, widx


def dam_break_3d_slab(dx: float, rank: int, world: int, skin: Optional[float] = None, **kw):
    """Slab `rank` of `examples.dam_break_3d(dx)` cut into `world` slabs of equal fluid count,
    generated locally: the rank builds only its own fluid columns and the wall particles within
    reach of them (RectangularTank(x_window=...)); the rows are exactly those of the global lattice,
    in the same relative order.  Returns `(fluid_k, wall_k, local)` for
    `SlabSemidiscretization(fluid_k, wall_k, ..., local=local)` (BASELINE config 4: 10 M - 100 M
    particles)."""
    from . import examples
    from .model import compact_support
    # geometry only: an empty window generates no particles
    fluid0, wall0, tank0 = examples.dam_break_3d(dx, x_window=(0.0, 0.0), **kw)
    t = fluid0.eltype.type
    ct = np.dtype(fluid0.coordinates_eltype).type
    n_f = tank0.n_particles_per_dimension
    spacing = float(t(dx))
    # x coordinate of every fluid column, exactly as rectangular_shape_coords computes it
    xs = (np.float64(ct(spacing)) * (np.arange(1, n_f[0] + 1, dtype=np.float64) - 0.5)).astype(ct)
    R_f = float(compact_support(fluid0.smoothing_kernel, t(fluid0.smoothing_length)))
    R_w = float(compact_support(wall0.boundary_model.smoothing_kernel, t(wall0.boundary_model.smoothing_length)))
    layout = make_layout(xs, world, R_f, R_w, skin, per_column=int(np.prod(n_f[1:])))
    lo, hi = layout.planes[rank], layout.planes[rank + 1]
    reach = layout.wall_reach
    fluid_k, wall_k, _ = examples.dam_break_3d(
        dx, x_window=(lo, hi), boundary_x_window=(lo - reach, np.nextafter(hi + reach, np.inf)), **kw)
    fluid_k.pressure = np.zeros(fluid_k.nparticles, dtype=fluid_k.eltype)
    L = tank0.n_layers
    gmin = np.full(3, -(L - 0.5) * spacing)
    gmax = np.asarray(tank0.tank_size, dtype=np.float64) + (L - 0.5) * spacing
    local = dict(layout=layout, bounding_box=(gmin, gmax), n_fluid=int(np.prod(n_f)))
    return fluid_k, wall_k, local


# ------------------------------------------------------------------ rebalance: planes + migration
def balanced_planes(x, world: int, group=None, nbins: int = 1 << 14) -> np.ndarray:
    """Slab faces such that every slab holds the same number of particles (+- one histogram bin),
    from a global histogram of the owned x coordinates (torch tensor on any device; one
    all_reduce each for the range and the counts -- no rank ever sees another rank's particles)."""
    import torch
    import torch.distributed as dist
    x = x.to(torch.float64)
    big = torch.finfo(torch.float64).max
    rng = torch.stack([x.min() if x.numel() else torch.tensor(big, device=x.device, dtype=torch.float64),
                       -x.max() if x.numel() else torch.tensor(big, device=x.device, dtype=torch.float64)])
    if world > 1:
        dist.all_reduce(rng, op=dist.ReduceOp.MIN, group=group)
    xmin, xmax = float(rng[0]), float(-rng[1])
    width = max((xmax - xmin) / nbins, 1e-300) * (1 + 1e-12)
    idx = ((x - xmin) / width).to(torch.int64).clamp_(0, nbins - 1)
    hist = torch.bincount(idx, minlength=nbins)
    if world > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    cum = torch.cumsum(hist, 0).cpu().numpy()
    total = int(cum[-1])
    planes = [-np.inf]
    for k in range(1, world):
        b = int(np.searchsorted(cum, total * k / world, side="left"))
        planes.append(xmin + (b + 1) * width)
    planes.append(np.inf)
    return np.asarray(planes)


def migrate(payload, dest, rank: int, world: int, group=None):
    """Send row i of every tensor in `payload` (2-D, same number of rows) to rank `dest[i]`.
    Returns the rows this rank now owns, ordered by source rank (own rows in their place): a
    count exchange (all_gather of one row of the count matrix) followed by variable-size
    point-to-point messages -- NCCL send/recv over NVLink on the box, gloo in the tests."""
    import torch
    import torch.distributed as dist
    order = torch.argsort(dest, stable=True)
    counts = torch.bincount(dest, minlength=world)
    if world == 1:
        return [t[order] for t in payload]
    rows = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(rows, counts, group=group)
    matrix = torch.stack(rows).cpu()                     # matrix[src, dst]
    send_off = torch.cumsum(counts, 0).cpu().tolist()
    send_off = [0] + send_off
    out = []
    for t in payload:
        ts = t[order].contiguous()
        chunks, ops = [], []
        for src in range(world):
            n = int(matrix[src, rank])
            if src == rank:
                chunks.append(ts[send_off[rank]:send_off[rank + 1]])
                continue
            buf = torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            chunks.append(buf)
            if n:
                ops.append(dist.P2POp(dist.irecv, buf, src, group))
        for dst in range(world):
            if dst != rank and send_off[dst + 1] > send_off[dst]:
                ops.append(dist.P2POp(dist.isend, ts[send_off[dst]:send_off[dst + 1]], dst, group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        out.append(torch.cat(chunks, dim=0))
    return out


# ------------------------------------------------------------------ transports
class DistTransport:
    """Neighbour exchange over torch.distributed point-to-point ops (NCCL over NVLink on the
    box, gloo in the CPU tests)."""

    def __init__(self, rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.left = rank - 1 if rank > 0 else None
        self.right = rank + 1 if rank < world - 1 else None

    def _run(self, ops):
        import torch.distributed as dist
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()

    def exchange_var(self, to_left, to_right):
        """Setup-time exchange of variable-size messages (counts first; host sync)."""
        import torch
        import torch.distributed as dist
        dev, dt, k = to_left.device, to_left.dtype, to_left.shape[1]
        nsend = torch.tensor([to_left.shape[0], to_right.shape[0]], dtype=torch.int64, device=dev)
        nrecv = torch.zeros(2, dtype=torch.int64, device=dev)
        ops = []
        if self.left is not None:
            ops += [dist.P2POp(dist.isend, nsend[0:1], self.left, self.group),
                    dist.P2POp(dist.irecv, nrecv[0:1], self.left, self.group)]
        if self.right is not None:
            ops += [dist.P2POp(dist.isend, nsend[1:2], self.right, self.group),
                    dist.P2POp(dist.irecv, nrecv[1:2], self.right, self.group)]
        self._run(ops)
        n_l, n_r = (int(c) for c in nrecv.tolist())
        from_left = torch.empty((n_l, k), dtype=dt, device=dev)
        from_right = torch.empty((n_r, k), dtype=dt, device=dev)
        self.exchange_fixed(to_left, to_right, from_left, from_right)
        return from_left, from_right

    def exchange_fixed(self, to_left, to_right, from_left, from_right):
        """Per-kick exchange: all four message sizes are known, nothing touches the host."""
        import torch.distributed as dist
        ops = []
        if self.left is not None:
            if to_left.shape[0]:
                ops.append(dist.P2POp(dist.isend, to_left, self.left, self.group))
            if from_left.shape[0]:
                ops.append(dist.P2POp(dist.irecv, from_left, self.left, self.group))
        if self.right is not None:
            if to_right.shape[0]:
                ops.append(dist.P2POp(dist.isend, to_right, self.right, self.group))
            if from_right.shape[0]:
                ops.append(dist.P2POp(dist.irecv, from_right, self.right, self.group))
        self._run(ops)


class LocalMailbox:
    """In-process stand-in for the interconnect: lets several slab ranks live in one process
    (single-GPU tests).  All ranks `post`, then all ranks `collect`."""

    def __init__(self, world: int):
        self.world = world
        self.box = {}

    def transport(self, rank: int) -> "LocalTransport":
        return LocalTransport(self, rank)


class LocalTransport:
    two_phase = True

    def __init__(self, mailbox: LocalMailbox, rank: int):
        self.mb, self.rank, self.world = mailbox, rank, mailbox.world

    def post(self, to_left, to_right, tag=""):
        if self.rank > 0:
            self.mb.box[(tag, self.rank, self.rank - 1)] = to_left.clone()
        if self.rank < self.world - 1:
            self.mb.box[(tag, self.rank, self.rank + 1)] = to_right.clone()

    def collect(self, like, tag=""):
        empty = like[:0]
        from_left = self.mb.box.pop((tag, self.rank - 1, self.rank), empty)
        from_right = self.mb.box.pop((tag, self.rank + 1, self.rank), empty)
        return from_left, from_right


# ------------------------------------------------------------------ halo selection + packing
class HaloExchange:
    """Fixed-slot ghost exchange of one rank (no host synchronisation per kick).

    At `setup` (semidiscretize / rebalance) every rank fixes its *candidate* lists: the owned
    particles within `halo + skin` of a slab face -- a superset of whatever can be inside the
    neighbour's ghost layer until the next rebalance.  Every kick it sends one row
    `(x[ND], v[NV])` per candidate; candidates currently outside the ghost layer get `x = NaN`,
    which the library treats as an empty ghost slot.  Message sizes, slot order (and therefore
    the summation order on the receiving rank) are fixed; masses travel once, at setup.
    Works on torch tensors of any device: u (n, ND) coordinates, v (n, NV) velocity + density."""

    def __init__(self, layout: SlabLayout, rank: int, transport, wall_coordinates=None):
        self.layout, self.rank, self.transport = layout, rank, transport
        self.lo, self.hi = float(layout.planes[rank]), float(layout.planes[rank + 1])
        self.has_left, self.has_right = rank > 0, rank < layout.world - 1
        self.tree = wall_tree(wall_coordinates)
        self.ready = False

    # -- setup -----------------------------------------------------------------------------
    def candidates(self, u):
        """Setup-time selection on the host (`candidate_mask`); returns index tensors on u's device."""
        import torch
        coords = u.detach().cpu().numpy()
        empty = torch.empty(0, dtype=torch.int64, device=u.device)

        def pick(face, side):
            idx = np.nonzero(candidate_mask(coords, face, side, self.layout, self.tree))[0]
            return torch.from_numpy(idx).to(u.device)

        cand_l = pick(self.lo, -1) if self.has_left else empty
        cand_r = pick(self.hi, +1) if self.has_right else empty
        return cand_l, cand_r

    def setup_post(self, u, mass):
        self.cand_l, self.cand_r = self.candidates(u)
        self.u_setup = u.clone()
        self._m_l = mass.index_select(0, self.cand_l).unsqueeze(1).contiguous()
        self._m_r = mass.index_select(0, self.cand_r).unsqueeze(1).contiguous()
        if getattr(self.transport, "two_phase", False):
            self.transport.post(self._m_l, self._m_r, tag="mass")

    def setup_finish(self, u, v):
        import torch
        if getattr(self.transport, "two_phase", False):
            m_l, m_r = self.transport.collect(self._m_l, tag="mass")
        else:
            m_l, m_r = self.transport.exchange_var(self._m_l, self._m_r)
        self.ghost_mass = torch.cat([m_l, m_r], dim=0).squeeze(1).contiguous()
        self.n_from_l, self.n_from_r = m_l.shape[0], m_r.shape[0]
        k = u.shape[1] + v.shape[1]
        self.recv_l = torch.empty((self.n_from_l, k), dtype=u.dtype, device=u.device)
        self.recv_r = torch.empty((self.n_from_r, k), dtype=u.dtype, device=u.device)
        self.ready = True

    def setup(self, u, v, mass):
        self.setup_post(u, mass)
        self.setup_finish(u, v)

    @property
    def n_ghost_slots(self) -> int:
        return self.n_from_l + self.n_from_r

    # -- per kick --------------------------------------------------------------------------
    def pack(self, u, v):
        """Rows (x, v) of the candidates; x[0] = NaN marks a candidate outside the ghost layer."""
        import torch
        out = []
        for cand, keep in ((self.cand_l, lambda x: x < self.lo + self.layout.halo),
                           (self.cand_r, lambda x: x >= self.hi - self.layout.halo)):
            rows = torch.cat([u.index_select(0, cand), v.index_select(0, cand).to(u.dtype)], dim=1)
            x = rows[:, 0]
            rows[:, 0] = torch.where(keep(x), x, torch.full_like(x, float("nan")))
            out.append(rows)
        return out

    def check_drift(self, u, margin: float = 0.0) -> bool:
        """True while every owned particle is within `skin - margin` of its slab and, once the
        candidate lists are fixed, has moved less than `skin - margin` since then (else: rebalance).
        `margin` = how far a particle may still travel until the next check (the caller's
        `steps between checks * dt * max|v|`), so that the budget is never exceeded in between."""
        x = u[:, 0]
        ok = True
        skin = max(self.layout.skin - float(margin), 0.0)
        if self.ready:
            moved2 = ((u - self.u_setup) ** 2).sum(dim=1).max() if len(u) else 0.0
            ok = ok and float(moved2) < skin ** 2
        if self.has_left:
            ok = ok and bool((x >= self.lo - skin).all())
        if self.has_right:
            ok = ok and bool((x < self.hi + skin).all())
        return ok

    def exchange(self, u, v):
        """Returns (u_g, v_g): one row per ghost slot (left neighbour's first)."""
        to_l, to_r = self.pack(u, v)
        self.transport.exchange_fixed(to_l, to_r, self.recv_l, self.recv_r)
        return self._unpack(u, v)

    def post(self, u, v):
        to_l, to_r = self.pack(u, v)
        self.transport.post(to_l, to_r, tag="halo")

    def collect(self, u, v):
        fl, fr = self.transport.collect(self.recv_l, tag="halo")
        self.recv_l, self.recv_r = fl, fr
        return self._unpack(u, v)

    def _unpack(self, u, v):
        import torch
        nd, nv = u.shape[1], v.shape[1]
        both = torch.cat([self.recv_l, self.recv_r], dim=0)
        return both[:, :nd], both[:, nd:nd + nv].to(v.dtype)


# ------------------------------------------------------------------ exchange over peer memory
class PeerHalo:
    """Per-kick ghost exchange through the neighbours' device memory (csrc/tpb_halo.cuh): one
    pack-and-store kernel per neighbour, one wait-and-install kernel, no NCCL call and no torch
    op on the path.  Set up after `HaloExchange.setup_finish` (candidate lists and slot counts
    are known); needs one process per GPU with CUDA IPC between them (one box)."""

    ALIGN = 256

    def __init__(self, slab, group=None, timeout_s: float = 20.0):
        import torch
        import torch.distributed as dist
        self.slab, self.halo = slab, slab.halo
        self.L, self.lib = slab._lib.load(), slab._lib
        self.rank, self.world = slab.rank, slab.world
        self.timeout_s = float(timeout_s)
        nd, nv = slab.nd, slab.nv
        csize = slab.u_ext.element_size()
        tsize = slab.v_ext.element_size()
        n_g = self.halo.n_ghost_slots
        up = lambda b: (b + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        # receive area: [stage_u p0 | stage_u p1 | stage_v p0 | stage_v p1 | flags | local words]
        self.off_u = [0, up(n_g * nd * csize)]
        self.off_v = [self.off_u[1] + up(n_g * nd * csize)]
        self.off_v.append(self.off_v[0] + up(n_g * nv * tsize))
        self.off_flag = self.off_v[1] + up(n_g * nv * tsize)      # +0: from left, +4: from right
        self.off_local = self.off_flag + self.ALIGN               # +0/+8: done counters, +16: timed_out
        total = self.off_local + self.ALIGN
        base = C.c_void_p()
        if self.L.tpb_peer_alloc(total, C.byref(base)) != 0:
            raise RuntimeError("tpb_peer_alloc failed")
        self.base = base.value
        handle = C.create_string_buffer(64)
        if self.L.tpb_peer_export(C.c_void_p(self.base), handle) != 0:
            raise RuntimeError("tpb_peer_export failed (CUDA IPC unavailable)")
        mine = dict(handle=bytes(handle.raw), off_u=self.off_u, off_v=self.off_v, off_flag=self.off_flag,
                    n_from_l=self.halo.n_from_l, n_from_r=self.halo.n_from_r)
        infos = [None] * self.world
        dist.all_gather_object(infos, mine, group=group)
        self.peers = {}
        ok = 1
        for side, nb in ((-1, self.rank - 1), (+1, self.rank + 1)):
            if nb < 0 or nb >= self.world:
                continue
            ptr = C.c_void_p()
            if self.L.tpb_peer_import(C.create_string_buffer(infos[nb]["handle"], 64), C.byref(ptr)) != 0:
                ok = 0
                continue
            self.peers[side] = (ptr.value, infos[nb])
        flag = torch.tensor([ok], dtype=torch.int32, device=slab.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag) == 0:
            self.close()
            raise RuntimeError("tpb_peer_import failed on some rank (no peer access between the GPUs?)")
        self.epoch = 0
        self.blocks_done = {-1: 0, +1: 0}
        self.launches = 0

    def exchange(self):
        """Queue this kick's exchange on the handle's stream (after the owned rows of
        u_ext / v_ext are final, before tpb_kick)."""
        s, h, L = self.slab, self.slab.semi._handle, self.L
        halo, lay = self.halo, self.halo.layout
        s.semi._bind_stream()
        self.epoch += 1
        e, par = self.epoch & 0xFFFFFFFF, self.epoch & 1
        nd, nv = s.nd, s.nv
        csize, tsize = s.u_ext.element_size(), s.v_ext.element_size()
        u_ptr, v_ptr = s.u_ext.data_ptr(), s.v_ext.data_ptr()
        for side, cand, thr in ((-1, halo.cand_l, halo.lo + lay.halo), (+1, halo.cand_r, halo.hi - lay.halo)):
            if side not in self.peers:
                continue
            pbase, info = self.peers[side]
            # my rows land behind the rows of the neighbour's left neighbour when I am its right one
            row0 = info["n_from_l"] if side < 0 else 0
            peer_u = pbase + info["off_u"][par] + row0 * nd * csize
            peer_v = pbase + info["off_v"][par] + row0 * nv * tsize
            peer_flag = pbase + info["off_flag"] + (4 if side < 0 else 0)
            done = self.base + self.off_local + (0 if side < 0 else 8)
            blocks = C.c_int32(0)
            self.lib.check(h, L.tpb_halo_pack(h, side, float(thr), C.c_void_p(u_ptr), C.c_void_p(v_ptr),
                                              C.c_void_p(cand.data_ptr()), int(cand.numel()),
                                              C.c_void_p(peer_u), C.c_void_p(peer_v), C.c_void_p(done),
                                              self.blocks_done[side], C.c_void_p(peer_flag), e, C.byref(blocks)))
            self.blocks_done[side] += blocks.value
            self.launches += 1
        n0, n_g = s.n_owned, halo.n_ghost_slots
        fl = C.c_void_p(self.base + self.off_flag) if -1 in self.peers else C.c_void_p(None)
        fr = C.c_void_p(self.base + self.off_flag + 4) if +1 in self.peers else C.c_void_p(None)
        self.lib.check(h, L.tpb_halo_install(h, n_g, C.c_void_p(self.base + self.off_u[par]),
                                             C.c_void_p(self.base + self.off_v[par]),
                                             C.c_void_p(u_ptr + n0 * nd * csize), C.c_void_p(v_ptr + n0 * nv * tsize),
                                             fl, fr, e, self.timeout_s, C.c_void_p(None)))
        self.launches += 1

    def close(self):
        for ptr, _ in getattr(self, "peers", {}).values():
            self.L.tpb_peer_close(C.c_void_p(ptr))
        self.peers = {}
        if getattr(self, "base", None):
            self.L.tpb_peer_free(C.c_void_p(self.base))
            self.base = None


# ------------------------------------------------------------------ the per-rank GPU object
class SlabSemidiscretization:
    """One slab of `Semidiscretization(fluid, wall)` on one GPU.

    `ode = slab.semidiscretize(tspan)` returns the rank's part of the `DynamicalODEProblem`:
    `ode.u0` / `ode.v0` are the owned particles (torch CUDA tensors, views of the extended
    buffers that also hold the ghosts), `ode.f1` = kick!, `ode.f2` = drift!."""

    def __init__(self, fluid, wall, *, rank: int, world: int, device: int = 0, transport=None,
                 skin: Optional[float] = None, ghost_capacity: Optional[int] = None,
                 interact_variant: int = 0, local: Optional[dict] = None):
        """`fluid`, `wall`: the GLOBAL systems (every rank builds the whole lattice and keeps its
        slab), or -- `local=dict(layout=SlabLayout, bounding_box=(min, max), n_fluid=N)` -- the rank's
        OWN systems already (fluid particles of slab `rank`, wall particles within `wall_reach` of
        it; see `dam_break_3d_slab`): nothing of global size is ever built, the ghost-slot counts
        come from the neighbours (one all_gather)."""
        import torch
        from . import _lib
        from .semidiscretization import (B200Backend, FullGridCellList, GridNeighborhoodSearch,
                                         Semidiscretization)
        self.rank, self.world = rank, world
        t = fluid.eltype.type
        from .model import compact_support
        R_f = float(compact_support(fluid.smoothing_kernel, t(fluid.smoothing_length)))
        R_w = (float(compact_support(wall.boundary_model.smoothing_kernel, t(wall.boundary_model.smoothing_length)))
               if wall is not None else R_f)
        self._radii = (R_f, R_w)
        self._device_index, self._interact_variant, self._transport_arg = device, interact_variant, transport
        self._lib = _lib
        if local is not None:
            self.global_n_fluid = int(local["n_fluid"])
            self._wall_global = None            # rebalance needs the global wall: not available
            self.layout = local["layout"]
            self.fluid, self.wall, self.owned_index, self.wall_index = fluid, wall, None, None
            lo, hi = self.layout.planes[rank], self.layout.planes[rank + 1]
            # the wall particles that can matter for the candidate selection: near the two faces
            tree = None
            if wall is not None and world > 1:
                xw = wall.coordinates[:, 0].astype(np.float64)
                reach = self.layout.halo + 2 * self.layout.skin + R_w
                near = np.zeros(len(xw), dtype=bool)
                for face in (lo, hi):
                    if np.isfinite(face):
                        near |= np.abs(xw - face) <= reach
                tree = wall_tree(wall.coordinates[near])
            self._tree = tree
            gmin, gmax = local["bounding_box"]
            self._gbox = (np.asarray(gmin, dtype=np.float64) - 2 * max(R_f, R_w),
                          np.asarray(gmax, dtype=np.float64) + 2 * max(R_f, R_w))
            n_slots = ghost_capacity
            if n_slots is None:
                n_slots = self._slots_from_neighbours(fluid.initial_condition.coordinates, self.layout,
                                                      torch.device("cuda", device) if transport is None else None)
            self._build(int(n_slots))
            return
        self.global_n_fluid = fluid.nparticles
        self._wall_global = wall
        self.layout = make_layout(fluid.initial_condition.coordinates[:, 0], world, R_f, R_w, skin)
        self.fluid, self.wall, self.owned_index, self.wall_index = local_systems(fluid, wall, self.layout, rank)
        # ghost slots = the neighbours' candidate counts (particles within halo + skin of the
        # faces): computed from the global lattice here, confirmed by the setup exchange
        lo, hi = self.layout.planes[rank], self.layout.planes[rank + 1]
        gcoords = fluid.initial_condition.coordinates
        tree = wall_tree(wall.coordinates if wall is not None else None)
        n_slots = 0
        if rank > 0:
            n_slots += int(candidate_mask(gcoords, lo, +1, self.layout, tree).sum())
        if rank < world - 1:
            n_slots += int(candidate_mask(gcoords, hi, -1, self.layout, tree).sum())
        # global bounding box: the whole tank (every slab shares it in y, z)
        allc = [fluid.initial_condition.coordinates.astype(np.float64)]
        if wall is not None:
            allc.append(wall.coordinates.astype(np.float64))
        allc = np.concatenate(allc)
        self._gbox = (allc.min(axis=0) - 2 * max(R_f, R_w), allc.max(axis=0) + 2 * max(R_f, R_w))
        self._tree = tree
        self._build(int(ghost_capacity if ghost_capacity is not None else n_slots))

    def _slots_from_neighbours(self, coords, layout, device=None, group=None) -> int:
        """Ghost slots of this rank = what its neighbours will send: every rank counts its own
        candidates per face, one all_gather distributes the counts."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return 0
        lo, hi = float(layout.planes[self.rank]), float(layout.planes[self.rank + 1])
        mine = torch.zeros(2, dtype=torch.int64, device=device)
        if self.rank > 0:
            mine[0] = int(candidate_mask(coords, lo, -1, layout, self._tree).sum())
        if self.rank < self.world - 1:
            mine[1] = int(candidate_mask(coords, hi, +1, layout, self._tree).sum())
        allc = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allc, mine, group=group)
        n_slots = 0
        if self.rank > 0:
            n_slots += int(allc[self.rank - 1][1])
        if self.rank < self.world - 1:
            n_slots += int(allc[self.rank + 1][0])
        return n_slots

    def _build(self, n_slots: int):
        """Everything that depends on the partition: the rank's Semidiscretization (bounding box =
        slab + ghost layer in x), transport and halo objects.  `self.fluid`, `self.wall`,
        `self.layout`, `self.owned_index` are set."""
        import torch
        from .semidiscretization import (B200Backend, FullGridCellList, GridNeighborhoodSearch,
                                         Semidiscretization)
        rank, world, device = self.rank, self.world, self._device_index
        R_f, R_w = self._radii
        nd = self.fluid.ndims
        tree = self._tree
        self.n_owned = self.fluid.nparticles
        lo, hi = self.layout.planes[rank], self.layout.planes[rank + 1]
        self.ghost_capacity = 0 if world == 1 else int(n_slots)
        # bounding box of the rank: slab + ghost layer in x, the whole tank in y, z
        gmin, gmax = self._gbox
        pad = self.layout.halo + self.layout.skin + max(R_f, R_w)
        mn, mx = gmin.copy(), gmax.copy()
        mn[0] = max(gmin[0], lo - pad) if np.isfinite(lo) else gmin[0]
        mx[0] = min(gmax[0], hi + pad) if np.isfinite(hi) else gmax[0]
        nhs = GridNeighborhoodSearch(nd, cell_list=FullGridCellList(min_corner=mn, max_corner=mx))
        backend = B200Backend(device=device, ode_memory="device", interact_variant=self._interact_variant)
        backend.ghost_capacity = self.ghost_capacity
        systems = (self.fluid,) if self.wall is None else (self.fluid, self.wall)
        self.semi = Semidiscretization(*systems, neighborhood_search=nhs, parallelization_backend=backend)
        self.device = torch.device("cuda", device)
        from .model import StateEquationAdaptiveCole
        self.adaptive_sound_speed = isinstance(self.fluid.state_equation, StateEquationAdaptiveCole)
        self._vmax_bits = None
        self.peer = None
        self.transport = self._transport_arg if self._transport_arg is not None else DistTransport(rank, world)
        self.halo = HaloExchange(self.layout, rank, self.transport, tree)
        self.nd, self.nv = nd, self.fluid.v_nvariables
        self.n_ghost = 0

    # -- setup ---------------------------------------------------------------------------
    def semidiscretize(self, tspan, finish=True):
        import torch
        from .semidiscretization import DynamicalODEProblem, semidiscretize
        ode0 = semidiscretize(self.semi, tspan)        # creates the handle with capacity
        cap = self.n_owned + self.ghost_capacity
        cdt, vdt = ode0.u0.dtype, ode0.v0.dtype
        self.u_ext = torch.zeros((cap, self.nd), dtype=cdt, device=self.device)
        self.v_ext = torch.zeros((cap, self.nv), dtype=vdt, device=self.device)
        self.u_ext[: self.n_owned] = ode0.u0.view(self.n_owned, self.nd)
        self.v_ext[: self.n_owned] = ode0.v0.view(self.n_owned, self.nv)
        self.u_ext[self.n_owned:, 0] = float("nan")    # empty ghost slots
        self.mass = torch.from_numpy(np.ascontiguousarray(self.fluid.mass)).to(self.device)
        u0 = self.u_ext[: self.n_owned].view(-1)
        v0 = self.v_ext[: self.n_owned].view(-1)
        self.ode = DynamicalODEProblem(self.kick_, self.drift_, v0, u0, tuple(tspan), SimpleNamespace(semi=self))
        if self.world > 1:
            self.halo.setup_post(self.u_ext[: self.n_owned], self.mass)
            if finish:
                self.setup_finish()
        return self.ode

    def setup_finish(self):
        """Second half of the setup exchange (separate only for the in-process transport)."""
        n0 = self.n_owned
        self.halo.setup_finish(self.u_ext[:n0], self.v_ext[:n0])
        n_g = self.halo.n_ghost_slots
        if n_g > self.ghost_capacity:
            raise RuntimeError(f"rank {self.rank}: {n_g} ghost slots exceed the capacity {self.ghost_capacity}")
        L, h = self._lib.load(), self.semi._handle
        self.semi._bind_stream()
        if n_g:
            self._lib.check(h, L.tpb_set_fluid_mass(h, n0, n_g, C.c_void_p(self.halo.ghost_mass.data_ptr())))
        self._lib.check(h, L.tpb_set_fluid_count(h, n0 + n_g, n0))
        self.n_ghost = n_g
        # exchange over peer memory when every rank is its own process on a GPU of this box
        self.peer = None
        import os
        if isinstance(self.transport, DistTransport) and os.environ.get("TPB_HALO", "peer") == "peer":
            try:
                self.peer = PeerHalo(self, group=self.transport.group)
            except RuntimeError as err:
                import sys
                print(f"[slabs] rank {self.rank}: peer-memory halo unavailable ({err}); using NCCL send/recv",
                      file=sys.stderr)
                self.peer = None

    def close(self):
        if getattr(self, "peer", None) is not None:
            self.peer.close()
            self.peer = None
        self.semi.close()

    @property
    def halo_transport(self) -> str:
        return "peer memory (pack-and-store kernel + flag)" if getattr(self, "peer", None) is not None \
            else "NCCL send/recv"

    # -- per kick ------------------------------------------------------------------------
    def _stage_owned(self, v_ode, u_ode):
        """Owned rows into the extended buffers (no copy when the caller works on the views)."""
        if u_ode.data_ptr() != self.u_ext.data_ptr():
            self.u_ext[: self.n_owned].copy_(u_ode.view(self.n_owned, self.nd))
        if v_ode.data_ptr() != self.v_ext.data_ptr():
            self.v_ext[: self.n_owned].copy_(v_ode.view(self.n_owned, self.nv))

    def _install_ghosts(self, ghosts):
        u_g, v_g = ghosts
        n0, n_g = self.n_owned, self.n_ghost
        if n_g:
            self.u_ext[n0:n0 + n_g] = u_g
            self.v_ext[n0:n0 + n_g] = v_g

    def _local_max_speed(self):
        import torch
        L, h = self._lib.load(), self.semi._handle
        if getattr(self, "_vmax_bits", None) is None:
            self._vmax_bits = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.semi._bind_stream()
        self._lib.check(h, L.tpb_max_speed2(h, C.c_void_p(self.v_ext.data_ptr()), C.c_void_p(self._vmax_bits.data_ptr())))
        return self._vmax_bits

    def _reduce_max_speed(self):
        """StateEquationAdaptiveCole: `update_speed_of_sound!` (wcsph/system.jl:307-321) takes the maximum
        over ALL fluid particles -- the rank's max |v|^2 over its owned rows (one kernel), an integer MAX
        all-reduce of the 8-byte bit pattern on the same stream, and the result is handed to the kick;
        nothing waits for the host."""
        import torch
        import torch.distributed as dist
        L, h = self._lib.load(), self.semi._handle
        if isinstance(self.transport, LocalTransport) and self.world > 1:
            # in-process ranks (two-phase kick): every rank has posted its word in kick_post
            words = [self.transport.mb.box[("vmax", r)] for r in range(self.world)]
            bits = torch.stack(words).max(dim=0).values
        else:
            bits = self._local_max_speed()
            if self.world > 1:
                dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=self.transport.group)
        self._lib.check(h, L.tpb_set_max_speed2(h, C.c_void_p(bits.data_ptr())))
        self._vmax_last = bits

    def _compute(self, dv_ode, t):
        L, h = self._lib.load(), self.semi._handle
        self.semi._bind_stream()
        if self.adaptive_sound_speed:
            self._reduce_max_speed()
        self._lib.check(h, L.tpb_kick(h, C.c_void_p(dv_ode.data_ptr()), C.c_void_p(self.v_ext.data_ptr()),
                                      C.c_void_p(self.u_ext.data_ptr()), float(t)))

    def kick_(self, dv_ode, v_ode, u_ode, p, t):
        self._stage_owned(v_ode, u_ode)
        if self.world > 1:
            if self.peer is not None:
                self.peer.exchange()
            else:
                n0 = self.n_owned
                self._install_ghosts(self.halo.exchange(self.u_ext[:n0], self.v_ext[:n0]))
        self._compute(dv_ode, t)
        return dv_ode

    # two-phase form (LocalTransport): all ranks post, then all ranks finish
    def kick_post(self, v_ode, u_ode):
        self._stage_owned(v_ode, u_ode)
        self.halo.post(self.u_ext[: self.n_owned], self.v_ext[: self.n_owned])
        if self.adaptive_sound_speed:
            self.transport.mb.box[("vmax", self.rank)] = self._local_max_speed().clone()

    def kick_finish(self, dv_ode, t=0.0):
        self._install_ghosts(self.halo.collect(self.u_ext[: self.n_owned], self.v_ext[: self.n_owned]))
        self._compute(dv_ode, t)
        return dv_ode

    def drift_(self, du_ode, v_ode, u_ode, p, t):
        L, h = self._lib.load(), self.semi._handle
        self.semi._bind_stream()
        self._lib.check(h, L.tpb_drift(h, C.c_void_p(du_ode.data_ptr()), C.c_void_p(v_ode.data_ptr()),
                                       C.c_void_p(u_ode.data_ptr()), float(t)))
        return du_ode

    def needs_rebalance(self, u_ode, v_ode=None, lookahead: float = 0.0) -> bool:
        """`lookahead` = time until the next check: with `v_ode` given, the motion `lookahead *
        max|v|` still to come is taken off the skin."""
        margin = 0.0
        if v_ode is not None and lookahead > 0.0 and self.n_owned:
            vel = v_ode.view(self.n_owned, self.nv)[:, : self.nd]
            margin = float(lookahead) * float((vel.double() ** 2).sum(dim=1).max().sqrt())
        return not self.halo.check_drift(u_ode.view(self.n_owned, self.nd), margin)

    # -- rebalance: new partition by position, particles migrate to their new owners ----------
    def rebalance(self, tspan=None):
        """Re-partition by the current positions (the state in `u_ext` / `v_ext`, i.e. in the
        views `ode.u0` / `ode.v0`): new slab faces of equal fluid count from a global histogram,
        every particle (x, v, rho, m, global index) moves to its new owner (`migrate`), the rank's
        Semidiscretization, ghost slots and exchange areas are rebuilt.  Collective: every rank
        calls it between two time steps (a `SortingCallback`-like event, sorting.jl:94-114 --
        particle <-> ODE index changes).  Returns the new `DynamicalODEProblem`; the owned rows
        are ordered by global particle index (`self.owned_index`)."""
        import torch
        import torch.distributed as dist
        from .model import ContinuityDensity
        if self.nv != self.nd + 1:
            raise ValueError("rebalance needs ContinuityDensity (the density travels in v_ode)")
        if self._wall_global is None and self.wall is not None:
            raise ValueError("rebalance needs the global wall system (this slab was built from local systems)")
        group = getattr(self.transport, "group", None)
        n0, nd = self.n_owned, self.nd
        u, v = self.u_ext[:n0], self.v_ext[:n0]
        planes = balanced_planes(u[:, 0], self.world, group)
        old = self.layout
        layout = SlabLayout(planes=planes, halo=old.halo, wall_reach=old.wall_reach, skin=old.skin,
                            direct=old.direct, near_wall=old.near_wall)
        if self.world > 2 and np.any(np.diff(planes[1:-1]) < layout.halo + layout.skin):
            raise ValueError("slabs are thinner than the ghost layer: use fewer ranks or a larger problem")
        dest = torch.bucketize(u[:, 0].to(torch.float64).contiguous(),
                               torch.as_tensor(planes[1:-1], dtype=torch.float64, device=u.device), right=True)
        ids = torch.from_numpy(np.ascontiguousarray(self.owned_index, dtype=np.int64)).to(u.device)
        ids, u, v, mass = migrate([ids.view(-1, 1), u, v, self.mass.view(-1, 1)], dest, self.rank, self.world, group)
        order = torch.argsort(ids[:, 0])
        ids, u, v, mass = ids[order, 0], u[order], v[order], mass[order, 0]
        # the rank's new systems
        un, vn, mn = u.cpu().numpy(), v.cpu().numpy(), mass.cpu().numpy()
        tspan = tuple(tspan) if tspan is not None else self.ode.tspan
        template = self.fluid
        fluid_k = copy.copy(template)
        fluid_k.initial_condition = InitialCondition(
            coordinates=np.ascontiguousarray(un), velocity=np.ascontiguousarray(vn[:, :nd]),
            mass=np.ascontiguousarray(mn), density=np.ascontiguousarray(vn[:, nd]),
            pressure=np.zeros(len(un), dtype=template.eltype),
            particle_spacing=template.initial_condition.particle_spacing)
        fluid_k.mass = np.ascontiguousarray(mn)
        fluid_k.pressure = np.zeros(len(un), dtype=template.eltype)
        wall_k, widx = local_wall(self._wall_global, layout, self.rank)
        # ghost slots: the neighbours' candidate counts under the new partition
        n_slots = self._slots_from_neighbours(un, layout, u.device, group)
        # swap the handle
        self.close()
        self.fluid, self.wall, self.wall_index = fluid_k, wall_k, widx
        self.owned_index = ids.cpu().numpy()
        self.layout = layout
        self.n_rebalances = getattr(self, "n_rebalances", 0) + 1
        self._build(n_slots)
        return self.semidiscretize(tspan)

    # -- time loop ---------------------------------------------------------------------------
    def solve(self, alg, *, dt: float, n_steps: int, check_every: int = 10, t0: float = 0.0):
        """`n_steps` fixed steps of the 2N-storage Runge-Kutta scheme `alg`
        (time_integration.CarpenterKennedy2N54) on the rank's particles: per stage one kick!
        (ghost exchange + RHS) and one drift!, stage updates fused on the device.  Every
        `check_every` steps all ranks agree (one all_reduce) whether some particle has moved
        further than the skin allows and `rebalance` if so.  Returns (t, v, u): the owned rows
        of the final state, matching `self.owned_index`."""
        import torch
        import torch.distributed as dist
        from . import _lib
        L = _lib.load()
        group = getattr(self.transport, "group", None)

        def stage(A, B, rhs, tmp, state):
            eltype = _lib.F32 if state.dtype == torch.float32 else _lib.F64
            self.semi._bind_stream()
            _lib.check(self.semi._handle, L.tpb_vec_rk2n_stage(
                self.semi._handle, state.numel(), eltype, float(A), float(B), float(dt),
                C.c_void_p(rhs.data_ptr()), C.c_void_p(tmp.data_ptr()), C.c_void_p(state.data_ptr())))

        def buffers():
            v, u = self.ode.v0, self.ode.u0   # views of the extended buffers: no staging copies
            return v, u, torch.zeros_like(v), torch.zeros_like(u), torch.zeros_like(v), torch.zeros_like(u)

        v, u, dv, du, tmp_v, tmp_u = buffers()
        t = float(t0)
        for step in range(int(n_steps)):
            if self.world > 1 and step % check_every == 0:
                flag = torch.tensor([1 if self.needs_rebalance(u, v, check_every * dt) else 0],
                                    dtype=torch.int32, device=u.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
                if int(flag):
                    self.rebalance()
                    v, u, dv, du, tmp_v, tmp_u = buffers()
            for A, B, c in zip(alg.A, alg.B, alg.c):
                self.kick_(dv, v, u, self.ode.p, t + c * dt)
                self.drift_(du, v, u, self.ode.p, t + c * dt)
                stage(A, B, dv, tmp_v, v)
                stage(A, B, du, tmp_u, u)
            t += dt
        self.semi.synchronize()
        return t, v, u
