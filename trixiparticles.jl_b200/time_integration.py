"""Time loop around the B200 right-hand side: the minimum of OrdinaryDiffEq + TrixiParticles
callbacks that the reference's dam-break validation run uses, kept GPU-resident.

Mirrors (names, arguments, semantics):
  `CarpenterKennedy2N54(williamson_condition=false)`  OrdinaryDiffEqLowStorageRK (external), used by
        /root/reference/examples/fluid/dam_break_2d.jl:127-130
  `StepsizeCallback(cfl=...)`      /root/reference/src/callbacks/stepsize.jl:46-79
  `calculate_dt`                   /root/reference/src/schemes/fluid/fluid.jl:199-239
  `PostprocessCallback(; dt, funcs...)`  /root/reference/src/callbacks/post_process.jl (time series)
  `max_x_coord`                    /root/reference/src/general/custom_quantities.jl
  `solve(ode, alg; dt, save_everystep, callback)`

With `B200Backend(ode_memory="device")` the state lives in torch CUDA tensors and every stage
update is one fused kernel of the library (`tpb_vec_rk2n_stage`); nothing but scalars crosses
PCIe.  This is SURVEY.md section 8(f) rank 1-2 ("time-loop residency", "probes").
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np

from . import _lib
from .model import WeaklyCompressibleSPHSystem


@dataclass(frozen=True)
class CarpenterKennedy2N54:
    """Carpenter & Kennedy (1994) five-stage fourth-order 2N-storage scheme."""
    williamson_condition: bool = False
    A = (0.0, -567301805773 / 1357537059087, -2404267990393 / 2016746695238,
         -3550918686646 / 2091501179385, -1275806237668 / 842570457699)
    B = (1432997174477 / 9575080441755, 5161836677717 / 13612068292357, 1720146321549 / 2090206949498,
         3134564353537 / 4481467310338, 2277821191437 / 14882151754819)
    c = (0.0, 1432997174477 / 9575080441755, 2526269341429 / 6820363962896,
         2006345519317 / 3224310063776, 2802321613138 / 2924317926251)


def calculate_dt(system: WeaklyCompressibleSPHSystem, cfl_number: float) -> float:
    """fluid.jl:199-239 with `ArtificialViscosityMonaghan` (viscosity.jl:82-87)."""
    h = float(system.smoothing_length)
    c = float(system.state_equation.sound_speed)
    dt_viscosity = math.inf
    if system.viscosity is not None:
        if hasattr(system.viscosity, "nu"):   # ViscosityMorris / ViscosityAdami: kinematic_viscosity = nu
            nu = float(system.viscosity.nu)
        else:
            nu = float(system.viscosity.alpha) * h * c / (2 * system.ndims + 4)
        dt_viscosity = 0.125 * h ** 2 / nu
    # sqrt(h / 0) = Inf in the reference: without gravity the other two limits decide
    a = float(np.linalg.norm(system.acceleration))
    dt_acceleration = math.inf if a == 0.0 else 0.25 * math.sqrt(h / a)
    dt_sound_speed = cfl_number * h / c
    return min(dt_viscosity, dt_acceleration, dt_sound_speed)


def calculate_dt_structure(system, cfl_number: float) -> float:
    """total_lagrangian_sph/system.jl:701-717: cfl h / sqrt(K / rho_min), bulk modulus
    K = E / (ND (1 - 2 nu))."""
    E, nu = float(system.young_modulus), float(system.poisson_ratio)
    K = E / (system.ndims * (1 - 2 * nu))
    sound_speed = math.sqrt(K / float(np.min(system.material_density)))
    return cfl_number * float(system.smoothing_length) / sound_speed


@dataclass
class StepsizeCallback:
    cfl: float

    def dt(self, semi) -> float:
        # minimum over all systems (stepsize.jl:63-79); walls contribute Inf
        from .model import TotalLagrangianSPHSystem
        dts = [calculate_dt(s, self.cfl) for s in semi.systems if isinstance(s, WeaklyCompressibleSPHSystem)]
        dts += [calculate_dt_structure(s, self.cfl) for s in semi.systems if isinstance(s, TotalLagrangianSPHSystem)]
        return min(dts)


def max_x_coord(system, v_ode, u_ode, semi, t) -> float:
    """Largest x coordinate of the system's particles (the dam-break surge front)."""
    if system.n_integrated_particles == 0:
        return float(np.max(system.coordinates[:, 0]))
    a, b = semi.ranges_u[semi.system_index(system)]
    nd = system.ndims
    if isinstance(u_ode, np.ndarray):
        return float(u_ode[a:b].reshape(-1, nd)[:, 0].max())
    out = C.c_double(0.0)
    eltype = _lib.F32 if semi.coordinates_eltype == np.float32 else _lib.F64
    ptr = u_ode.data_ptr() + a * u_ode.element_size()
    semi._bind_stream()
    _lib.check(semi._handle, _lib.load().tpb_vec_strided_max(
        semi._handle, (b - a) // nd, eltype, nd, 0, C.c_void_p(ptr), C.byref(out)))
    return out.value


class PostprocessCallback:
    """Records `name -> f(system, v_ode, u_ode, semi, t)` for the fluid system every `dt`."""

    def __init__(self, dt: float, **funcs: Callable):
        self.dt = float(dt)
        self.funcs = funcs
        self.times: List[float] = []
        self.values: Dict[str, List[float]] = {k: [] for k in funcs}

    def __call__(self, t, v_ode, u_ode, semi):
        self.times.append(t)
        for name, f in self.funcs.items():
            self.values[name].append(f(semi.fluid, v_ode, u_ode, semi, t))


@dataclass
class Solution:
    t: float
    v: object
    u: object
    nsteps: int
    nf: int                      # RHS evaluations (kick! + drift! pairs)
    retcode: str = "Success"
    dts: List[float] = field(default_factory=list)


class _VecOps:
    """Stage update on the ODE vectors: fused library kernel on the device, numpy on the host."""

    def __init__(self, semi):
        self.semi = semi
        self.device = semi.parallelization_backend.ode_memory == "device"

    def zeros_like(self, x):
        if self.device:
            import torch
            return torch.zeros_like(x)
        return np.zeros_like(x)

    def rk2n_stage(self, A, B, dt, rhs, tmp, state):
        if not self.device:
            if A == 0.0:
                tmp[:] = dt * rhs
            else:
                tmp *= A
                tmp += dt * rhs
            state += B * tmp
            return
        import torch
        eltype = _lib.F32 if state.dtype == torch.float32 else _lib.F64
        self.semi._bind_stream()
        _lib.check(self.semi._handle, _lib.load().tpb_vec_rk2n_stage(
            self.semi._handle, state.numel(), eltype, float(A), float(B), float(dt),
            C.c_void_p(rhs.data_ptr()), C.c_void_p(tmp.data_ptr()), C.c_void_p(state.data_ptr())))


def solve(ode, alg: CarpenterKennedy2N54, *, dt: Optional[float] = None, save_everystep: bool = False,
          callback=(), maxiters: int = 10 ** 7, cuda_graph: bool = False) -> Solution:
    """`solve(ode, CarpenterKennedy2N54(); dt, callback)` for the `DynamicalODEProblem` returned by
    `semidiscretize`: per stage dv = kick!(v, u), du = drift!(v, u), then the 2N-storage update of
    both partitions.  Fixed step from `dt` or a `StepsizeCallback`; `PostprocessCallback`s add
    tstops at multiples of their `dt`, which the step is shortened to hit (OrdinaryDiffEq's
    `modify_dt_for_tstops!`).  `cuda_graph=True` (device-resident vectors): one full step of the
    regular size -- all stages, about 15 kernel launches each -- is captured once into a CUDA graph
    and replayed, so that small problems are not bound by launch latency; the shortened steps
    before an output time run eagerly."""
    semi = ode.p.semi
    if callback is None:
        callback = ()
    callbacks = list(callback) if isinstance(callback, (list, tuple)) else [callback]
    stepsize = next((c for c in callbacks if isinstance(c, StepsizeCallback)), None)
    posts = [c for c in callbacks if isinstance(c, PostprocessCallback)]
    if stepsize is not None:
        dt = stepsize.dt(semi)
    if dt is None or not dt > 0:
        raise ValueError("a positive `dt` or a `StepsizeCallback` is required (fixed-step scheme)")
    adaptive_eos = [s_.state_equation for s_ in semi.systems
                    if isinstance(s_, WeaklyCompressibleSPHSystem) and hasattr(s_.state_equation, "update_speed_of_sound")]
    ops = _VecOps(semi)
    v = ode.v0.clone() if ops.device else ode.v0.copy()
    u = ode.u0.clone() if ops.device else ode.u0.copy()
    dv, du = ops.zeros_like(v), ops.zeros_like(u)
    tmp_v, tmp_u = ops.zeros_like(v), ops.zeros_like(u)
    t, t_end = float(ode.tspan[0]), float(ode.tspan[1])
    next_stop = [t + p.dt for p in posts]
    for p in posts:
        p(t, v, u, semi)
    nsteps = nf = 0
    dts: List[float] = []
    graph, warmed, graph_ok = None, False, bool(cuda_graph and ops.device)

    def run_stages(step_):
        for A, B, c in zip(alg.A, alg.B, alg.c):
            ode.f1(dv, v, u, ode.p, t + c * step_)
            ode.f2(du, v, u, ode.p, t + c * step_)
            ops.rk2n_stage(A, B, step_, dv, tmp_v, v)
            ops.rk2n_stage(A, B, step_, du, tmp_u, u)

    while t < t_end and nsteps < maxiters:
        stop = min([t_end] + next_stop)
        step = dt
        if stop - t <= step * (1 + 1e-10):
            step = stop - t
        if graph_ok and step == dt and not adaptive_eos:
            if graph is None:
                if not warmed:
                    run_stages(step)       # first regular step eagerly (one-time kernel attributes)
                    warmed = True
                else:
                    import torch
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        run_stages(step)       # recorded, not executed
                    semi._bind_stream()        # back from the capture's side stream
                    graph.replay()
            else:
                graph.replay()
        else:
            run_stages(step)
        nf += len(alg.A)
        t = stop if step != dt else t + step
        nsteps += 1
        if stepsize is not None and adaptive_eos:
            # StateEquationAdaptiveCole: the StepsizeCallback sees the speed of sound of the last
            # right-hand-side evaluation (stepsize.jl:63-79, fluid.jl:199-239)
            for se_ in adaptive_eos:
                se_.sound_speed = se_.param_eltype.type(semi.sound_speed())
            dt = stepsize.dt(semi)
        if save_everystep:
            dts.append(step)
        for i, p in enumerate(posts):
            if t >= next_stop[i] - 1e-12 * max(1.0, abs(t)):
                p(t, v, u, semi)
                next_stop[i] = (len(p.times)) * p.dt + float(ode.tspan[0])
    if ops.device:
        semi.synchronize()   # surfaces a deferred out-of-bounds error of an asynchronous kick
    return Solution(t=t, v=v, u=u, nsteps=nsteps, nf=nf, dts=dts,
                    retcode="Success" if t >= t_end else "MaxIters")
