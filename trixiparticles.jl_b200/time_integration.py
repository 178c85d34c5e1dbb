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
import sys
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


def _rdpk3spfsal35_embedded_weights(c_stages, b_main):
    """Embedded (second-order) weights of RDPK3SpFSAL35 over its five stages and the FSAL stage.

    The scheme lives in OrdinaryDiffEqLowStorageRK.jl (third-party, not under /root/reference; no
    network here).  The main method's coefficients below are verified: they reproduce the abscissae
    c_i and satisfy all four third-order conditions to 1e-16 (tests/test_host_logic.py).  The
    embedded weights could not be verified the same way -- the values on record violate
    sum(bhat) = 1 and sum(bhat c) = 1/2 in the fourth digit -- so the weights used here are the
    closest ones (least squares) that satisfy both second-order conditions exactly.  The step-size
    sequence therefore follows the published controller on a second-order error estimate of the same
    family, not bit for bit OrdinaryDiffEq's."""
    recorded = np.array([1.046363371354093758897668305991705199e-01, 9.520431574956758809511173383346476348e-02,
                         4.482446645568668405072421350300379357e-01, 2.449030295461310135957132640369862245e-01,
                         7.962176910593929356839751650674380730e-02, 2.718753614680220546367613932962426040e-02])
    c_all = np.append(np.asarray(c_stages, dtype=np.float64), 1.0)
    M = np.stack([np.ones(6), c_all])
    bhat = recorded - M.T @ np.linalg.solve(M @ M.T, M @ recorded - np.array([1.0, 0.5]))
    return tuple(np.append(np.asarray(b_main, dtype=np.float64), 0.0) - bhat)   # error weights b - bhat


@dataclass(frozen=True)
class RDPK3SpFSAL35:
    """Ranocha, Dalcin, Parsani, Ketcheson (2021): five-stage third-order 3S*+ low-storage scheme with
    FSAL and an embedded second-order error estimate -- the integrator of the reference's shipped
    examples (examples/fluid/dam_break_3d.jl:85, hydrostatic_water_column_2d.jl:86).  Per stage
        tmp += delta_i u;  u = gamma1_i u + gamma2_i tmp + gamma3_i uprev + beta_i dt f(u)
    Step-size control: PID controller (0.70, -0.23, 0.00) with the limiter 1 + atan(x - 1) and the
    accept threshold 0.81, as OrdinaryDiffEq selects for this scheme."""
    gamma1 = (0.0, 2.587771979725733308135192812685323706e-01, -1.324380360140723382965420909764953437e-01,
              5.056033948190826045833606441415585735e-02, 5.670532000739313812633197158607642990e-01)
    gamma2 = (1.0, 5.528354909301389892439698870483746541e-01, 6.731871608203061824849561782794643600e-01,
              2.803103963297672407841316576323901761e-01, 5.521525447020610386070346724931300367e-01)
    gamma3 = (0.0, 0.0, 0.0, 2.752563273304676380891217287572780582e-01, -8.950526174674033822276061734289327568e-01)
    delta = (1.0, 3.407655879334525365094815965895763636e-01, 3.414382655003386206551709871126405331e-01,
             7.229275366787987419692007421895451953e-01, 0.0)
    beta = (2.300298624518076223899418286314123354e-01, 3.021434166948288809034402119555380003e-01,
            8.025606185416310937583009085873554681e-01, 4.362158943603440930655148245148766471e-01,
            1.129272530455059129782111662594436580e-01)
    c = (0.0, 2.300298624518076223899418286314123354e-01, 4.050046072094990912268498160116125481e-01,
         8.947822893693433545220710894560512805e-01, 7.235136928826589010272834603680114769e-01)
    pid = (0.70, -0.23, 0.00)
    order = 3
    adaptive_order = 2

    @classmethod
    def butcher(cls):
        """Effective Butcher tableau (A, b) of the low-storage recurrence."""
        S = 5
        up = np.zeros(S + 1); up[0] = 1.0
        u, tmp = up.copy(), up.copy()
        u[1] += cls.beta[0]
        rows = [up.copy()]
        for i in range(1, S):
            rows.append(u.copy())
            tmp = tmp + cls.delta[i] * u
            u = cls.gamma1[i] * u + cls.gamma2[i] * tmp + cls.gamma3[i] * up
            u[i + 1] += cls.beta[i]
        return np.array([r[1:] for r in rows]), u[1:].copy()

    @classmethod
    def error_weights(cls):
        A, b = cls.butcher()
        return _rdpk3spfsal35_embedded_weights(A.sum(axis=1), b)


@dataclass(frozen=True)
class SymplecticPositionVerlet:
    """DualSPHysics' symplectic position Verlet scheme as the reference defines it
    (ext/TrixiParticlesOrdinaryDiffEqSymplecticRKExt.jl:89-171): drift half, kick half, kick at the
    half step, then v = v_prev + dt dv except for an integrated density, which is advanced by
    rho = rho_prev (2 - eps) / (2 + eps), eps = -(drho / rho_half) dt; drift half."""


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
    # (vectors: the stiffest particle decides)
    E, nu = np.asarray(system.young_modulus, dtype=np.float64), np.asarray(system.poisson_ratio, dtype=np.float64)
    K = float(np.max(E / (system.ndims * (1 - 2 * nu))))
    sound_speed = math.sqrt(K / float(np.min(system.material_density)))
    return cfl_number * float(system.smoothing_length) / sound_speed


@dataclass
class StepsizeCallback:
    cfl: float

    def dt(self, semi) -> float:
        # minimum over all systems (stepsize.jl:63-79); walls contribute Inf
        from .model import TotalLagrangianSPHSystem
        dts = [calculate_dt(s, self.cfl) for s in semi.systems if isinstance(s, WeaklyCompressibleSPHSystem)]
        if getattr(semi, "integrate_tlsph", True):   # calculate_dt skips TLSPH systems that are not integrated (:514-517)
            dts += [calculate_dt_structure(s, self.cfl) for s in semi.systems if isinstance(s, TotalLagrangianSPHSystem)]
        return min(dts)


class SplitIntegrationCallback:
    """`SplitIntegrationCallback(alg; stage_coupling=false, predict_positions=true, dt, callback=StepsizeCallback(...))`
    (callbacks/split_integration.jl:1-110): the `TotalLagrangianSPHSystem` is taken out of the main integrator
    (`semi.integrate_tlsph[] = false`) and advanced by its own CarpenterKennedy2N54 sub-integrator with smaller
    steps -- after every step of the main integrator, or to every stage time with `stage_coupling=true`.  During
    one sub-integration the force of the fluid on the structure is constant; it is computed for the new fluid
    state and the structure positions predicted by an Euler step (`predict_positions`).  Device-resident vectors."""

    def __init__(self, alg, *, stage_coupling: bool = False, predict_positions: bool = True, dt: float = None,
                 callback=None):
        if not isinstance(alg, CarpenterKennedy2N54):
            raise ValueError("SplitIntegrationCallback: CarpenterKennedy2N54 is the sub-integrator on this path")
        self.alg, self.stage_coupling, self.predict_positions = alg, bool(stage_coupling), bool(predict_positions)
        self.dt_sub, self.stepsize = dt, callback
        self.n_substeps = self.n_calls = 0

    # initialize_split_integration! (split_integration.jl:112-170)
    def initialize(self, semi, v_ode, u_ode, t):
        from .model import TotalLagrangianSPHSystem
        st = next((s_ for s_ in semi.systems if isinstance(s_, TotalLagrangianSPHSystem)), None)
        if st is None:
            raise ValueError("`SplitIntegrationCallback` must be used with a `TotalLagrangianSPHSystem`")
        if semi.parallelization_backend.ode_memory != "device":
            raise ValueError("SplitIntegrationCallback needs B200Backend(ode_memory='device')")
        if getattr(semi, "_merged", None) is not None:
            raise ValueError("SplitIntegrationCallback with several structure systems is outside the accelerated path")
        self.semi, self.system = semi, st
        i = semi.system_index(st)
        self.rv, self.ru = semi.ranges_v[i], semi.ranges_u[i]
        semi.set_integrate_structure(False)
        if isinstance(self.stepsize, StepsizeCallback):
            self.dt_sub = calculate_dt_structure(st, self.stepsize.cfl)
        if self.dt_sub is None or not self.dt_sub > 0:
            raise ValueError("SplitIntegrationCallback: `dt` or `callback=StepsizeCallback(...)` is required")
        # copy_to_split!
        self.v = v_ode[self.rv[0]:self.rv[1]].clone()
        self.u = u_ode[self.ru[0]:self.ru[1]].clone()
        self.dv, self.du = self.v.clone(), self.u.clone()
        self.tmp_v, self.tmp_u = self.v.clone(), self.u.clone()
        self.force = self.v.clone()
        self.t = float(t)
        self.ops = _VecOps(semi)

    def _copy_from_split(self, v_ode, u_ode, dt_predict):
        # copy_from_split! (split_integration.jl:490-510): u += v (t_new - t_previous) with PREDICT
        v_ode[self.rv[0]:self.rv[1]].copy_(self.v)
        us = u_ode[self.ru[0]:self.ru[1]]
        us.copy_(self.u)
        if dt_predict != 0.0:
            us.add_(self.v.to(us.dtype), alpha=dt_predict)

    # split_integrate! (split_integration.jl:237-350)
    def integrate_to(self, v_ode, u_ode, t_new):
        t_prev = self.t
        if t_new < t_prev - 1e-12 * max(1.0, abs(t_prev)):
            raise ValueError("stage-level coupling with `SplitIntegrationCallback` requires monotonically increasing "
                             "stage times")
        span = t_new - t_prev
        self._copy_from_split(v_ode, u_ode, span if self.predict_positions else 0.0)
        if abs(span) <= 1e-14 * max(1.0, abs(t_new)):
            return
        semi, L, h = self.semi, _lib.load(), self.semi._handle
        semi._bind_stream()
        # update_systems_and_nhs + other_interaction_split!: the fluid's force for this sub-integration
        _lib.check(h, L.tpb_structure_fluid_force(h, C.c_void_p(self.force.data_ptr()), C.c_void_p(v_ode.data_ptr()),
                                                  C.c_void_p(u_ode.data_ptr())))
        n = max(1, int(math.ceil(span / self.dt_sub - 1e-9)))
        dt = span / n
        alg = self.alg
        for _ in range(n):
            for A, B, c in zip(alg.A, alg.B, alg.c):
                _lib.check(h, L.tpb_kick_structure(h, C.c_void_p(self.dv.data_ptr()), C.c_void_p(self.v.data_ptr()),
                                                   C.c_void_p(self.u.data_ptr()), C.c_void_p(self.force.data_ptr())))
                # drift_split!: du = v; both partitions are updated from the state of the stage's start
                self.du.copy_(self.v.to(self.du.dtype))
                self.ops.rk2n_stage(A, B, dt, self.dv, self.tmp_v, self.v)
                self.ops.rk2n_stage(A, B, dt, self.du, self.tmp_u, self.u)
        self.n_substeps += n
        self.n_calls += 1
        self.t = float(t_new)
        self._copy_from_split(v_ode, u_ode, 0.0)


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


class SortingCallback:
    """callbacks/sorting.jl:9-57: `SortingCallback(; interval=-1, dt=0.0, initial_sort=true)` -- reorders the fluid
    particles in the ODE vectors by neighbourhood-search cell, every `interval` steps or at the first step after
    each `dt` of simulation time.  The reference needs it against a 3-4x slowdown of its per-particle kernels on
    shuffled particles; here the library sorts its own records at every kick and only the gather / scatter between
    the ODE vectors and those records profits (5-9 % for a random order)."""

    def __init__(self, *, interval: int = -1, dt: float = 0.0, initial_sort: bool = True):
        if dt > 0 and interval > 0:
            raise ValueError("setting both `interval` and `dt` is not supported")
        if not dt > 0 and interval <= 0:
            raise ValueError("either `interval` or `dt` must be set to a positive value")
        self.interval = float(dt) if dt > 0 else int(interval)
        self.initial_sort = bool(initial_sort)
        self.last_t = 0.0
        self.n_sorts = 0

    def initialize(self, semi, v, u, t) -> bool:
        self.last_t = float(t)
        if self.initial_sort:
            self._sort(semi, v, u, t)
        return self.initial_sort

    def _sort(self, semi, v, u, t):
        semi.sort_particles(v, u)
        self.last_t = float(t)
        self.n_sorts += 1

    def __call__(self, semi, v, u, t, nsteps) -> bool:
        """condition + affect! (sorting.jl:76-112); True when the vectors were reordered (the integrator then has a
        derivative discontinuity: FSAL right-hand sides are stale)."""
        due = (nsteps % self.interval == 0) if isinstance(self.interval, int) else (t - self.last_t >= self.interval)
        if due:
            self._sort(semi, v, u, t)
        return bool(due)


class DensityReinitializationCallback:
    """callbacks/density_reinit.jl:33-121: `DensityReinitializationCallback(system, semi; interval=0, dt=0.0,
    reinit_initial_solution=true)` -- every `interval` steps (or at the first step after each `dt`) the integrated
    density of the fluid is replaced by the Shepard-corrected summation density (Panizzo 2007)."""

    def __init__(self, system=None, semi=None, *, interval: int = 0, dt: float = 0.0, reinit_initial_solution: bool = True):
        if dt > 0 and interval > 0:
            raise ValueError("Setting both interval and dt is not supported!")
        if system is not None and not isinstance(getattr(system, "density_calculator", None), ContinuityDensityType()):
            raise ValueError("DensityReinitializationCallback: the system must integrate its density (ContinuityDensity)")
        self.interval = float(dt) if dt > 0 else int(interval)
        self.reinit_initial_solution = bool(reinit_initial_solution)
        self.last_t = -math.inf
        self.n_reinits = 0

    def initialize(self, semi, v, u, t):
        if self.reinit_initial_solution:
            self._apply(semi, v, u, t)
        self.last_t = float(t)

    def _apply(self, semi, v, u, t):
        semi.reinit_density(v, u)
        self.last_t = float(t)
        self.n_reinits += 1

    def __call__(self, semi, v, u, t, nsteps) -> bool:
        if isinstance(self.interval, int):
            due = self.interval > 0 and nsteps % self.interval == 0
        else:
            due = (t - self.last_t) > self.interval
        if due:
            self._apply(semi, v, u, t)
        return bool(due)


def ContinuityDensityType():
    from .model import ContinuityDensity
    return ContinuityDensity


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

    def _eltype(self, x):
        import torch
        return _lib.F32 if x.dtype == torch.float32 else _lib.F64

    def lincomb(self, y, *terms):
        """y = sum_k a_k x_k for up to four (a, x) terms (an x may be y itself)."""
        terms = [(float(a), x) for a, x in terms if x is not None]
        assert 1 <= len(terms) <= 4
        if not self.device:
            acc = sum(a * x.astype(np.float64) for a, x in terms)
            y[:] = acc.astype(y.dtype)
            return
        terms += [(0.0, None)] * (4 - len(terms))
        self.semi._bind_stream()
        args = []
        for a, x in terms:
            args += [a, C.c_void_p(x.data_ptr()) if x is not None else None]
        _lib.check(self.semi._handle, _lib.load().tpb_vec_lincomb4(
            self.semi._handle, y.numel(), self._eltype(y), *args, C.c_void_p(y.data_ptr())))

    def wrms_sumsq(self, err, a, b, abstol, reltol):
        """sum_i (err_i / (abstol + reltol max(|a_i|, |b_i|)))^2 and the number of entries."""
        if not self.device:
            e = err.astype(np.float64) / (abstol + reltol * np.maximum(np.abs(a), np.abs(b)).astype(np.float64))
            return float((e * e).sum()), err.size
        n = err.numel()
        if n == 0:
            return 0.0, 0
        out = C.c_double(0.0)
        self.semi._bind_stream()
        _lib.check(self.semi._handle, _lib.load().tpb_vec_wrms_norm(
            self.semi._handle, n, self._eltype(err), C.c_void_p(err.data_ptr()), C.c_void_p(a.data_ptr()),
            C.c_void_p(b.data_ptr()), float(abstol), float(reltol), C.byref(out)))
        return out.value ** 2 * n, n

    def verlet_update(self, system, semi, dt, kdu, duprev, du):
        """update_velocity! / update_density! of one system on its rows of the v vectors."""
        a, b = semi.ranges_v[semi.system_index(system)]
        if b == a:
            return
        nd, nv = system.ndims, system.v_nvariables
        n = (b - a) // nv
        if not self.device:
            k, p, d = (x[a:b].reshape(n, nv) for x in (kdu, duprev, du))
            t = d.dtype.type
            if nv > nd:
                eps = -k[:, nd] / d[:, nd] * t(dt)
                rho = p[:, nd] * (t(2) - eps) / (t(2) + eps)
            d[:, :nd] = p[:, :nd] + t(dt) * k[:, :nd]
            if nv > nd:
                d[:, nd] = rho
            return
        self.semi._bind_stream()
        off = a * du.element_size()
        _lib.check(self.semi._handle, _lib.load().tpb_vec_verlet_update(
            self.semi._handle, n, nd, nv, self._eltype(du), float(dt), C.c_void_p(kdu.data_ptr() + off),
            C.c_void_p(duprev.data_ptr() + off), C.c_void_p(du.data_ptr() + off)))

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


def solve(ode, alg, *, dt: Optional[float] = None, save_everystep: bool = False,
          callback=(), maxiters: int = 10 ** 7, cuda_graph: bool = False, abstol: float = 1e-6,
          reltol: float = 1e-3, dtmax: Optional[float] = None) -> Solution:
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
    if isinstance(alg, RDPK3SpFSAL35):
        return _solve_rdpk3(ode, alg, callbacks, dt=dt, abstol=abstol, reltol=reltol, dtmax=dtmax,
                            maxiters=maxiters, save_everystep=save_everystep)
    if isinstance(alg, SymplecticPositionVerlet):
        return _solve_verlet(ode, callbacks, dt=dt, maxiters=maxiters, save_everystep=save_everystep)
    stepsize = next((c for c in callbacks if isinstance(c, StepsizeCallback)), None)
    posts = [c for c in callbacks if isinstance(c, PostprocessCallback)]
    split = next((c for c in callbacks if isinstance(c, SplitIntegrationCallback)), None)
    sorting = next((c for c in callbacks if isinstance(c, SortingCallback)), None)   # (fluid rows only, sorting.jl:5)
    reinit = next((c for c in callbacks if isinstance(c, DensityReinitializationCallback)), None)
    if split is not None:
        split.initialize(semi, ode.v0, ode.u0, ode.tspan[0])   # (also: the StepsizeCallback skips the structure)
        cuda_graph = False                                      # the number of sub-steps varies from stage to stage
    if getattr(semi, "_motion_system", None) is not None:
        cuda_graph = False                                      # the movement function runs on the host, per kick
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
    if sorting is not None:
        sorting.initialize(semi, v, u, t)
    if reinit is not None:
        reinit.initialize(semi, v, u, t)
    next_stop = [t + p.dt for p in posts]
    for p in posts:
        p(t, v, u, semi)
    nsteps = nf = 0
    dts: List[float] = []
    graph, warmed, graph_ok = None, False, bool(cuda_graph and ops.device)

    def run_stages(step_):
        for A, B, c in zip(alg.A, alg.B, alg.c):
            if split is not None and split.stage_coupling:
                split.integrate_to(v, u, t + c * step_)    # split_integrate_stage! (semidiscretization.jl:594)
            ode.f1(dv, v, u, ode.p, t + c * step_)
            ode.f2(du, v, u, ode.p, t + c * step_)
            ops.rk2n_stage(A, B, step_, dv, tmp_v, v)
            ops.rk2n_stage(A, B, step_, du, tmp_u, u)

    while t < t_end and nsteps < maxiters:
        stop = min([t_end] + next_stop)
        step = dt
        if stop - t <= step * (1 + 1e-10):
            step = stop - t
        # StateEquationAdaptiveCole: replayable when the speed of sound stays on the device and dt is fixed
        if graph_ok and step == dt and (not adaptive_eos or
                                        (stepsize is None and semi.adaptive_sound_speed_on_device())):
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
        if split is not None:
            split.integrate_to(v, u, t)                    # the callback's affect! at the end of every step
        if sorting is not None:
            sorting(semi, v, u, t, nsteps)                 # in place: a captured graph keeps replaying on the same vectors
        if reinit is not None:
            reinit(semi, v, u, t, nsteps)
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


# ------------------------------------------------------------------ adaptive low-storage scheme
def _rhs(ode, dv, du, v, u, t):
    ode.f1(dv, v, u, ode.p, t)
    ode.f2(du, v, u, ode.p, t)


def _solve_rdpk3(ode, alg: RDPK3SpFSAL35, callbacks, *, dt, abstol, reltol, dtmax, maxiters, save_everystep):
    """`solve(ode, RDPK3SpFSAL35(); abstol, reltol, dtmax)` on the (v, u) partition of the
    `DynamicalODEProblem`: OrdinaryDiffEq's 3S*+ FSAL step (`LowStorageRK3SpFSAL`), its automatic
    initial step (Hairer & Wanner) and its PID step-size controller; the error norm is the RMS of
    err / (abstol + reltol max(|u_prev|, |u|)) over both partitions (`ODE_DEFAULT_NORM`)."""
    semi = ode.p.semi
    ops = _VecOps(semi)
    posts = [c for c in callbacks if isinstance(c, PostprocessCallback)]
    sorting = next((c for c in callbacks if isinstance(c, SortingCallback)), None)
    t, t_end = float(ode.tspan[0]), float(ode.tspan[1])
    dtmax = float(dtmax) if dtmax is not None else t_end - t
    v = ode.v0.clone() if ops.device else ode.v0.copy()
    u = ode.u0.clone() if ops.device else ode.u0.copy()
    reinit = next((c for c in callbacks if isinstance(c, DensityReinitializationCallback)), None)
    if sorting is not None:
        sorting.initialize(semi, v, u, t)
    if reinit is not None:
        reinit.initialize(semi, v, u, t)
    Z = ops.zeros_like
    kv, ku, k0v, k0u = Z(v), Z(u), Z(v), Z(u)           # stage rhs / FSAL rhs
    tv, tu, pv, pu, ev, eu = Z(v), Z(u), Z(v), Z(u), Z(v), Z(u)   # tmp register, uprev, error estimate
    ew = alg.error_weights()
    nf = 0

    def norm(ev_, eu_, av, au, bv, bu):
        s1, n1 = ops.wrms_sumsq(ev_, av, bv, abstol, reltol)
        s2, n2 = ops.wrms_sumsq(eu_, au, bu, abstol, reltol)
        return math.sqrt((s1 + s2) / max(n1 + n2, 1))

    _rhs(ode, k0v, k0u, v, u, t)
    nf += 1
    if dt is None or not dt > 0:
        # ode_determine_initdt (Hairer, Norsett, Wanner I, II.4): d0 = |u0|, d1 = |f0|, an explicit Euler
        # probe for the second derivative
        zv, zu = Z(v), Z(u)
        d0 = norm(v, u, v, u, v, u)
        d1 = norm(k0v, k0u, v, u, v, u)
        dt0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
        dt0 = min(dt0, dtmax)
        ops.lincomb(tv, (1.0, v), (dt0, k0v))
        ops.lincomb(tu, (1.0, u), (dt0, k0u))
        _rhs(ode, kv, ku, tv, tu, t + dt0)
        nf += 1
        ops.lincomb(zv, (1.0, kv), (-1.0, k0v))
        ops.lincomb(zu, (1.0, ku), (-1.0, k0u))
        d2 = norm(zv, zu, v, u, v, u) / dt0
        dmax = max(d1, d2)
        dt1 = max(1e-6, dt0 * 1e-3) if dmax <= 1e-15 else 10.0 ** (-(2.0 + math.log10(dmax)) / (alg.order + 1))
        dt = min(100 * dt0, dt1, dtmax)
        del zv, zu
    dt = min(float(dt), dtmax)
    err_hist = [1.0, 1.0, 1.0]
    kexp = min(alg.order, alg.adaptive_order) + 1
    nsteps = nrejected = 0
    dts: List[float] = []
    next_stop = [t + p.dt for p in posts]
    for p in posts:
        p(t, v, u, semi)
    retcode = "Success"
    while t < t_end:
        if nsteps >= maxiters:
            retcode = "MaxIters"
            break
        stop = min([t_end] + next_stop)
        step = min(dt, stop - t)
        ops.lincomb(pv, (1.0, v))
        ops.lincomb(pu, (1.0, u))
        ops.lincomb(tv, (1.0, v))
        ops.lincomb(tu, (1.0, u))
        ops.lincomb(v, (1.0, tv), (alg.beta[0] * step, k0v))
        ops.lincomb(u, (1.0, tu), (alg.beta[0] * step, k0u))
        ops.lincomb(ev, (ew[0] * step, k0v))
        ops.lincomb(eu, (ew[0] * step, k0u))
        for i in range(1, 5):
            _rhs(ode, kv, ku, v, u, t + alg.c[i] * step)
            ops.lincomb(tv, (1.0, tv), (alg.delta[i], v))
            ops.lincomb(tu, (1.0, tu), (alg.delta[i], u))
            ops.lincomb(v, (alg.gamma1[i], v), (alg.gamma2[i], tv), (alg.gamma3[i], pv), (alg.beta[i] * step, kv))
            ops.lincomb(u, (alg.gamma1[i], u), (alg.gamma2[i], tu), (alg.gamma3[i], pu), (alg.beta[i] * step, ku))
            ops.lincomb(ev, (1.0, ev), (ew[i] * step, kv))
            ops.lincomb(eu, (1.0, eu), (ew[i] * step, ku))
        _rhs(ode, kv, ku, v, u, t + step)       # FSAL evaluation
        nf += 5
        ops.lincomb(ev, (1.0, ev), (ew[5] * step, kv))
        ops.lincomb(eu, (1.0, eu), (ew[5] * step, ku))
        est = norm(ev, eu, pv, pu, v, u)
        # PIDController (Soederlind; OrdinaryDiffEqCore controllers)
        est = max(est, 1.0 / sys.float_info.max) if est == est else float("inf")
        err_hist = [1.0 / est, err_hist[0], err_hist[1]]
        factor = (err_hist[0] ** (alg.pid[0] / kexp) * err_hist[1] ** (alg.pid[1] / kexp)
                  * err_hist[2] ** (alg.pid[2] / kexp))
        factor = 1.0 + math.atan(factor - 1.0)
        if factor >= 0.81:
            t = stop if step != dt else t + step
            nsteps += 1
            k0v, kv = kv, k0v
            k0u, ku = ku, k0u
            if save_everystep:
                dts.append(step)
            if step == dt:                       # a step shortened for an output time keeps the controller's dt
                dt = min(dt * factor, dtmax)
            changed = sorting is not None and sorting(semi, v, u, t, nsteps)
            changed = (reinit is not None and reinit(semi, v, u, t, nsteps)) or changed
            if changed:
                _rhs(ode, k0v, k0u, v, u, t)     # derivative_discontinuity!(integrator, true): the FSAL value is stale
                nf += 1
            for i, p in enumerate(posts):
                if t >= next_stop[i] - 1e-12 * max(1.0, abs(t)):
                    p(t, v, u, semi)
                    next_stop[i] = len(p.times) * p.dt + float(ode.tspan[0])
        else:
            nrejected += 1
            ops.lincomb(v, (1.0, pv))
            ops.lincomb(u, (1.0, pu))
            dt = step * factor
            if not dt > 1e-14 * max(1.0, abs(t)):
                retcode = "DtLessThanMin"
                break
    if ops.device:
        semi.synchronize()
    sol = Solution(t=t, v=v, u=u, nsteps=nsteps, nf=nf, dts=dts, retcode=retcode)
    sol.nrejected = nrejected
    return sol


def _solve_verlet(ode, callbacks, *, dt, maxiters, save_everystep):
    """`solve(ode, SymplecticPositionVerlet(); dt, callback=StepsizeCallback(cfl))`."""
    semi = ode.p.semi
    ops = _VecOps(semi)
    stepsize = next((c for c in callbacks if isinstance(c, StepsizeCallback)), None)
    posts = [c for c in callbacks if isinstance(c, PostprocessCallback)]
    if stepsize is not None:
        dt = stepsize.dt(semi)
    if dt is None or not dt > 0:
        raise ValueError("a positive `dt` or a `StepsizeCallback` is required (fixed-step scheme)")
    v = ode.v0.clone() if ops.device else ode.v0.copy()
    u = ode.u0.clone() if ops.device else ode.u0.copy()
    Z = ops.zeros_like
    pv, pu, kdu, ku = Z(v), Z(u), Z(v), Z(u)
    t, t_end = float(ode.tspan[0]), float(ode.tspan[1])
    sorting = next((c for c in callbacks if isinstance(c, SortingCallback)), None)
    if sorting is not None:
        sorting.initialize(semi, v, u, t)
    next_stop = [t + p.dt for p in posts]
    for p in posts:
        p(t, v, u, semi)
    nsteps = nf = 0
    dts: List[float] = []
    while t < t_end and nsteps < maxiters:
        stop = min([t_end] + next_stop)
        step = dt if stop - t > dt * (1 + 1e-10) else stop - t
        ops.lincomb(pv, (1.0, v))
        ops.lincomb(pu, (1.0, u))
        ode.f2(ku, pv, pu, ode.p, t)                       # position half step
        ops.lincomb(u, (1.0, pu), (0.5 * step, ku))
        ode.f1(kdu, pv, pu, ode.p, t)                      # velocity half step
        ops.lincomb(v, (1.0, pv), (0.5 * step, kdu))
        ode.f1(kdu, v, u, ode.p, t + 0.5 * step)           # full kick from the half-step state
        for system in semi.systems:
            ops.verlet_update(system, semi, step, kdu, pv, v)
        ode.f2(ku, v, u, ode.p, t + step)                  # second position half step
        ops.lincomb(u, (1.0, u), (0.5 * step, ku))
        nf += 2
        t = stop if step != dt else t + step
        nsteps += 1
        if sorting is not None:
            sorting(semi, v, u, t, nsteps)                 # (no state but v, u survives a step)
        if save_everystep:
            dts.append(step)
        for i, p in enumerate(posts):
            if t >= next_stop[i] - 1e-12 * max(1.0, abs(t)):
                p(t, v, u, semi)
                next_stop[i] = len(p.times) * p.dt + float(ode.tspan[0])
    if ops.device:
        semi.synchronize()
    return Solution(t=t, v=v, u=u, nsteps=nsteps, nf=nf, dts=dts, retcode="Success" if t >= t_end else "MaxIters")
