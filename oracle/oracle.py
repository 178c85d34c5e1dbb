"""ctypes binding of the CPU oracle (oracle/liboracle_sph.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under trixiparticles.jl_b200/
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_sph.so")

KERNEL_WENDLAND_C2 = 0
KERNEL_SCHOENBERG_CUBIC = 1
KERNEL_WENDLAND_C4 = 2
KERNEL_WENDLAND_C6 = 3
KERNEL_SCHOENBERG_QUARTIC = 4
KERNEL_SCHOENBERG_QUINTIC = 5
DENSITY_CONTINUITY = 0
DENSITY_SUMMATION = 1


class FluidParams(C.Structure):
    _fields_ = [
        ("ndims", C.c_int32), ("kernel", C.c_int32), ("density_calculator", C.c_int32),
        ("clip_negative_pressure", C.c_int32), ("has_viscosity", C.c_int32),
        ("has_diffusion", C.c_int32), ("reserved0", C.c_int32), ("reserved1", C.c_int32),
        ("smoothing_length", C.c_double), ("sound_speed", C.c_double),
        ("exponent", C.c_double), ("reference_density", C.c_double),
        ("background_pressure", C.c_double), ("alpha", C.c_double), ("beta", C.c_double),
        ("epsilon", C.c_double), ("delta", C.c_double), ("acceleration", C.c_double * 3),
        ("damping_coefficient", C.c_double),
    ]


class WallParams(C.Structure):
    _fields_ = [
        ("kernel", C.c_int32), ("clip_negative_pressure", C.c_int32),
        ("eos_clip_negative_pressure", C.c_int32), ("reserved0", C.c_int32),
        ("smoothing_length", C.c_double), ("sound_speed", C.c_double),
        ("exponent", C.c_double), ("reference_density", C.c_double),
        ("background_pressure", C.c_double), ("pressure_offset", C.c_double),
        ("has_viscosity", C.c_int32), ("density_calculator", C.c_int32),
        ("visc_alpha", C.c_double), ("visc_beta", C.c_double), ("visc_epsilon", C.c_double),
    ]


WALL_ADAMI, WALL_CONTINUITY = 0, 1


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("sph_oracle.c", "sph_oracle_impl.inc", "tlsph_oracle_impl.inc", "sph_oracle.h")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle_sph.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


_SUFFIX = {("float64", "float64"): "f64", ("float32", "float32"): "f32",
           ("float32", "float64"): "f32c64"}


BOUNDARY_NONE = 0
BOUNDARY_MONAGHAN_KAJTAR = 1
BOUNDARY_DUMMY_PARTICLES = 2


class TlsphParams(C.Structure):
    _fields_ = [
        ("ndims", C.c_int32), ("kernel", C.c_int32), ("has_penalty", C.c_int32),
        ("boundary_model", C.c_int32), ("smoothing_length", C.c_double),
        ("young_modulus", C.c_double), ("poisson_ratio", C.c_double), ("penalty_alpha", C.c_double),
        ("acceleration", C.c_double * 3), ("mk_K", C.c_double), ("mk_beta", C.c_double),
        ("mk_spacing", C.c_double),
        ("bm_kernel", C.c_int32), ("bm_clip_negative_pressure", C.c_int32),
        ("bm_smoothing_length", C.c_double), ("bm_sound_speed", C.c_double), ("bm_exponent", C.c_double),
        ("bm_reference_density", C.c_double), ("bm_background_pressure", C.c_double),
        ("bm_pressure_offset", C.c_double),
        ("bm_wall_semantics", C.c_int32), ("bm_reserved", C.c_int32),
        ("bm_bernoulli_factor", C.c_double),
    ]


def suffix(dtype, coords_dtype=None) -> str:
    dtype = np.dtype(dtype)
    coords_dtype = np.dtype(coords_dtype) if coords_dtype is not None else dtype
    return _SUFFIX[(dtype.name, coords_dtype.name)]


def _declare(L):
    d, i, i64, p = C.c_double, C.c_int, C.c_int64, C.c_void_p
    for s in ("f64", "f32", "f32c64"):
        for name in ("orc_kernel", "orc_kernel_unsafe", "orc_kernel_deriv_div_r"):
            f = getattr(L, f"{name}_{s}"); f.restype = d; f.argtypes = [i, i, d, d]
        f = getattr(L, f"orc_eos_{s}"); f.restype = d; f.argtypes = [d, d, d, d, i, d]
        f = getattr(L, f"orc_inverse_eos_{s}"); f.restype = d; f.argtypes = [d, d, d, d, d]
        f = getattr(L, f"orc_viscosity_pair_{s}"); f.restype = None
        f.argtypes = [i, i, d, d, d, d, d, d, d, d, p, p, p]
        f = getattr(L, f"orc_viscosity_pair_nu_{s}"); f.restype = None
        f.argtypes = [i, i, i, d, d, d, d, d, d, d, p, p, p]
        f = getattr(L, f"orc_interact_pair_{s}"); f.restype = None
        f.argtypes = [C.POINTER(FluidParams), i, d, d, d, d, d, p, p, p, p, p]
        for name in ("orc_pairs_bruteforce", "orc_pairs_grid"):
            f = getattr(L, f"{name}_{s}"); f.restype = i64
            f.argtypes = [i, i64, p, i64, p, d, i64, p, p]
        f = getattr(L, f"orc_kick_{s}"); f.restype = i
        f.argtypes = [C.POINTER(FluidParams), C.POINTER(WallParams), i64, p, i64, p, p, p, p, p,
                      p, p, p, p, p, i, i]
        f = getattr(L, f"orc_kick_noslip_{s}"); f.restype = i
        f.argtypes = [C.POINTER(FluidParams), C.POINTER(WallParams), i64, p, i64, p, p, p, p, p,
                      p, p, p, p, p, p, i, i]
        f = getattr(L, f"orc_drift_{s}"); f.restype = None
        f.argtypes = [i, i, i64, p, p]
        f = getattr(L, f"orc_reinit_density_{s}"); f.restype = i
        f.argtypes = [C.POINTER(FluidParams), C.POINTER(WallParams), i64, p, i64, p, p, p, p, p]
        TP = C.POINTER(TlsphParams)
        f = getattr(L, f"orc_tlsph_correction_matrix_{s}"); f.restype = i
        f.argtypes = [TP, i64, p, p, p, p]
        f = getattr(L, f"orc_tlsph_update_{s}"); f.restype = i
        f.argtypes = [TP, i64, p, p, p, p, p, p, p]
        f = getattr(L, f"orc_tlsph_interact_{s}"); f.restype = i
        f.argtypes = [TP, i64, i64, p, p, p, p, p, p, p]
        f = getattr(L, f"orc_kick_fsi_{s}"); f.restype = i
        f.argtypes = [C.POINTER(FluidParams), C.POINTER(WallParams), TP, i64, p, i64, p, p, i64, i64,
                      p, p, p, p, p, p, p, p, p, p, i]
        f = getattr(L, f"orc_kick_fsi2_{s}"); f.restype = i
        f.argtypes = [C.POINTER(FluidParams), C.POINTER(WallParams), TP, i64, p, i64, p, p, i64, i64,
                      p, p, p, p, p, p, p, p, p, p, p, p, i]
        f = getattr(L, f"orc_kick_fsi3_{s}"); f.restype = i
        f.argtypes = [C.POINTER(FluidParams), C.POINTER(WallParams), TP, i64, p, i64, p, p, i64, i64,
                      p, p, p, p, p, p, p, p, p, p, p, p, p, p, p, i]
    L.orc_max_threads.restype = i
    L.orc_max_threads.argtypes = []


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _vec3(x):
    out = np.zeros(3, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64).ravel()
    out[: x.size] = x
    return out


# ---------------------------------------------------------------- scalar helpers

def kernel(kernel_id, ndims, r, h, dtype=np.float64):
    return getattr(lib(), f"orc_kernel_{suffix(dtype)}")(kernel_id, ndims, float(r), float(h))


def kernel_unsafe(kernel_id, ndims, r, h, dtype=np.float64):
    return getattr(lib(), f"orc_kernel_unsafe_{suffix(dtype)}")(kernel_id, ndims, float(r), float(h))


def kernel_deriv_div_r(kernel_id, ndims, r, h, dtype=np.float64):
    return getattr(lib(), f"orc_kernel_deriv_div_r_{suffix(dtype)}")(kernel_id, ndims, float(r), float(h))


def eos(c, gamma, rho0, p_bg, clip, density, dtype=np.float64):
    return getattr(lib(), f"orc_eos_{suffix(dtype)}")(float(c), float(gamma), float(rho0),
                                                      float(p_bg), int(clip), float(density))


def inverse_eos(c, gamma, rho0, p_bg, pressure, dtype=np.float64):
    return getattr(lib(), f"orc_inverse_eos_{suffix(dtype)}")(float(c), float(gamma), float(rho0),
                                                              float(p_bg), float(pressure))


def viscosity_pair(kernel_id, ndims, h, alpha, beta, epsilon, c, m_b, rho_a, rho_b, v_diff,
                   pos_diff, dtype=np.float64):
    vd, pd, out = _vec3(v_diff), _vec3(pos_diff), np.zeros(3)
    getattr(lib(), f"orc_viscosity_pair_{suffix(dtype)}")(
        kernel_id, ndims, float(h), float(alpha), float(beta), float(epsilon), float(c),
        float(m_b), float(rho_a), float(rho_b), _ptr(vd), _ptr(pd), _ptr(out))
    return out[:ndims]


def viscosity_pair_nu(model, kernel_id, ndims, h, nu, epsilon, m_a, m_b, rho_a, rho_b, v_diff, pos_diff,
                      dtype=np.float64):
    """ViscosityMorris (model 2) / ViscosityAdami (model 3) functor on one pair."""
    vd, pd, out = _vec3(v_diff), _vec3(pos_diff), np.zeros(3)
    getattr(lib(), f"orc_viscosity_pair_nu_{suffix(dtype)}")(
        int(model), kernel_id, ndims, float(h), float(nu), float(epsilon), float(m_a), float(m_b),
        float(rho_a), float(rho_b), _ptr(vd), _ptr(pd), _ptr(out))
    return out[:ndims]


def interact_pair(fp: FluidParams, neighbor_is_wall, m_b, rho_a, rho_b, p_a, p_b, v_a, v_b,
                  pos_diff, dtype=np.float64):
    va, vb, pd, dv, drho = _vec3(v_a), _vec3(v_b), _vec3(pos_diff), np.zeros(3), np.zeros(1)
    getattr(lib(), f"orc_interact_pair_{suffix(dtype)}")(
        C.byref(fp), int(neighbor_is_wall), float(m_b), float(rho_a), float(rho_b), float(p_a),
        float(p_b), _ptr(va), _ptr(vb), _ptr(pd), _ptr(dv), _ptr(drho))
    return dv[: fp.ndims], float(drho[0])


# ---------------------------------------------------------------- neighbour sets

def neighbor_pairs(x, y, radius, dtype=None, grid=False):
    """All (i, j) with |x_i - y_j|^2 <= R^2 (PointNeighbors predicate), sorted (i, j).

    x, y: (n, ND) arrays (particle-major == Julia's ND x n column-major) of the coordinate
    dtype; dtype = eltype(system) (defaults to the coordinate dtype)."""
    x = np.ascontiguousarray(x)
    y = np.ascontiguousarray(y, dtype=x.dtype)
    dtype = np.dtype(dtype) if dtype is not None else x.dtype
    s = suffix(dtype, x.dtype)
    f = getattr(lib(), f"orc_pairs_{'grid' if grid else 'bruteforce'}_{s}")
    nd = x.shape[1]
    cap = max(1024, 64 * x.shape[0])
    while True:
        oi = np.empty(cap, dtype=np.int32)
        oj = np.empty(cap, dtype=np.int32)
        n = f(nd, x.shape[0], _ptr(x), y.shape[0], _ptr(y), float(radius), cap, _ptr(oi), _ptr(oj))
        if n < 0:
            raise MemoryError("oracle grid allocation failed")
        if n <= cap:
            return oi[:n].copy(), oj[:n].copy()
        cap = int(n)


def neighbor_pair_count(x, y, radius, dtype=None) -> int:
    """Number of (i, j) with |x_i - y_j|^2 <= R^2 (cell-list search; nothing is stored)."""
    x = np.ascontiguousarray(x)
    y = np.ascontiguousarray(y, dtype=x.dtype)
    dtype = np.dtype(dtype) if dtype is not None else x.dtype
    f = getattr(lib(), f"orc_pairs_grid_{suffix(dtype, x.dtype)}")
    dummy = np.empty(1, dtype=np.int32)
    n = f(x.shape[1], x.shape[0], _ptr(x), y.shape[0], _ptr(y), float(radius), 0, _ptr(dummy), _ptr(dummy))
    if n < 0:
        raise MemoryError("oracle grid allocation failed")
    return int(n)


# ---------------------------------------------------------------- kick! / drift!

def kick(fp: FluidParams, wp, mass_f, coords_w, mass_w, v_ode, u_ode, dtype, use_grid=True,
         nthreads=0, v_wall=None):
    """One `kick!` of Semidiscretization(fluid[, wall]).  Arrays are particle-major:
    u_ode (n_f, ND) coordinate dtype; v_ode (n_f, nv) dtype.  Returns a dict.
    use_grid: False = all pairs, True = cell lists built per call, 2 = timing mode: the static
    wall's cell list is built once and reused across calls (as the reference does)."""
    dtype = np.dtype(dtype)
    u_ode = np.ascontiguousarray(u_ode)
    cdt = u_ode.dtype
    s = suffix(dtype, cdt)
    v_ode = np.ascontiguousarray(v_ode, dtype=dtype)
    mass_f = np.ascontiguousarray(mass_f, dtype=dtype)
    n_f = u_ode.shape[0]
    nd = fp.ndims
    nv = nd if fp.density_calculator == DENSITY_SUMMATION else nd + 1
    assert u_ode.shape == (n_f, nd) and v_ode.shape == (n_f, nv), (u_ode.shape, v_ode.shape)
    if wp is not None and coords_w is not None and len(coords_w) > 0:
        coords_w = np.ascontiguousarray(coords_w, dtype=cdt)
        mass_w = np.ascontiguousarray(mass_w, dtype=dtype)
        n_w = coords_w.shape[0]
        wpp = C.byref(wp)
    else:
        coords_w = np.zeros((0, nd), dtype=cdt)
        mass_w = np.zeros(0, dtype=dtype)
        n_w = 0
        wpp = None
    # BoundaryModelDummyParticles{ContinuityDensity}: the wall's density rows sit behind the fluid's rows
    # (`v_wall`, n_w values); the result gains `dv_wall`
    wall_cont = wpp is not None and wp.density_calculator == WALL_CONTINUITY
    v_in, dv_flat = v_ode, None
    if wall_cont:
        v_wall = np.ascontiguousarray(v_wall, dtype=dtype)
        assert v_wall.shape == (n_w,)
        v_in = np.concatenate([v_ode.reshape(-1), v_wall])
        dv_flat = np.full(v_in.shape, np.nan, dtype=dtype)
    out = dict(
        dv=np.zeros((n_f, nv), dtype=dtype), pressure=np.zeros(n_f, dtype=dtype),
        density=np.zeros(n_f, dtype=dtype), wall_pressure=np.zeros(n_w, dtype=dtype),
        wall_density=np.zeros(n_w, dtype=dtype), wall_volume=np.zeros(n_w, dtype=dtype),
        wall_velocity=np.zeros((n_w, nd), dtype=dtype))
    rc = getattr(lib(), f"orc_kick_noslip_{s}")(
        C.byref(fp), wpp, n_f, _ptr(mass_f), n_w, _ptr(coords_w), _ptr(mass_w), _ptr(v_in),
        _ptr(u_ode), _ptr(dv_flat if wall_cont else out["dv"]), _ptr(out["pressure"]), _ptr(out["density"]),
        _ptr(out["wall_pressure"]), _ptr(out["wall_density"]), _ptr(out["wall_volume"]),
        _ptr(out["wall_velocity"]), (2 if use_grid == 2 else int(bool(use_grid))), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"orc_kick failed: {rc}")
    if wall_cont:
        out["dv"] = dv_flat[: n_f * nv].reshape(n_f, nv)
        out["dv_wall"] = dv_flat[n_f * nv:]
    return out


def drift(v_ode, ndims, coords_dtype):
    v_ode = np.ascontiguousarray(v_ode)
    s = suffix(v_ode.dtype, coords_dtype)
    du = np.zeros((v_ode.shape[0], ndims), dtype=coords_dtype)
    getattr(lib(), f"orc_drift_{s}")(ndims, v_ode.shape[1], v_ode.shape[0], _ptr(v_ode), _ptr(du))
    return du


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ---------------------------------------------------------------- TLSPH structure + FSI

def tlsph_correction_matrix(sp: TlsphParams, x0, mass, rho, dtype):
    """`initialize!` of a TotalLagrangianSPHSystem: (n, ND, ND) correction matrices, [p, j, i] =
    Julia's L[i, j, p] (the memory layout of the reference's ND x ND x n array)."""
    dtype = np.dtype(dtype)
    x0 = np.ascontiguousarray(x0)
    s = suffix(dtype, x0.dtype)
    n, nd = x0.shape
    mass, rho = np.ascontiguousarray(mass, dtype=dtype), np.ascontiguousarray(rho, dtype=dtype)
    L = np.zeros((n, nd, nd), dtype=dtype)
    rc = getattr(lib(), f"orc_tlsph_correction_matrix_{s}")(C.byref(sp), n, _ptr(x0), _ptr(mass), _ptr(rho), _ptr(L))
    assert rc == 0
    return L


def tlsph_update(sp: TlsphParams, x0, x_cur, mass, rho, L, dtype):
    """Deformation gradient and PK1 / rho^2 of every particle: two (n, ND, ND) arrays in the
    reference's memory layout ([p, j, i] = M[i, j, p])."""
    dtype = np.dtype(dtype)
    x0, x_cur = np.ascontiguousarray(x0), np.ascontiguousarray(x_cur, dtype=np.asarray(x0).dtype)
    s = suffix(dtype, x0.dtype)
    n, nd = x0.shape
    mass, rho = np.ascontiguousarray(mass, dtype=dtype), np.ascontiguousarray(rho, dtype=dtype)
    L = np.ascontiguousarray(L, dtype=dtype)
    F, P = np.zeros((n, nd, nd), dtype=dtype), np.zeros((n, nd, nd), dtype=dtype)
    rc = getattr(lib(), f"orc_tlsph_update_{s}")(C.byref(sp), n, _ptr(x0), _ptr(x_cur), _ptr(mass), _ptr(rho),
                                                  _ptr(L), _ptr(F), _ptr(P))
    assert rc == 0
    return F, P


def tlsph_interact(sp: TlsphParams, n_int, x0, x_cur, mass, rho, F, pk1_rho2, dtype):
    """`interact_structure_structure!`: dv (n_int, ND) of the integrated particles."""
    dtype = np.dtype(dtype)
    x0, x_cur = np.ascontiguousarray(x0), np.ascontiguousarray(x_cur, dtype=np.asarray(x0).dtype)
    s = suffix(dtype, x0.dtype)
    n, nd = x0.shape
    mass, rho = np.ascontiguousarray(mass, dtype=dtype), np.ascontiguousarray(rho, dtype=dtype)
    F, P = np.ascontiguousarray(F, dtype=dtype), np.ascontiguousarray(pk1_rho2, dtype=dtype)
    dv = np.zeros((n_int, nd), dtype=dtype)
    rc = getattr(lib(), f"orc_tlsph_interact_{s}")(C.byref(sp), n, n_int, _ptr(x0), _ptr(x_cur), _ptr(mass),
                                                    _ptr(rho), _ptr(F), _ptr(P), _ptr(dv))
    assert rc == 0
    return dv


def reinit_density(fp: FluidParams, wp, mass_f, coords_w, mass_w, v_ode, u_ode, dtype):
    """`reinit_density!` of the DensityReinitializationCallback (orc_reinit_density): the new fluid densities (n_f,)."""
    dtype = np.dtype(dtype)
    u_ode = np.ascontiguousarray(u_ode)
    cdt = u_ode.dtype
    s = suffix(dtype, cdt)
    v_ode = np.ascontiguousarray(v_ode, dtype=dtype)
    mass_f = np.ascontiguousarray(mass_f, dtype=dtype)
    n_f = mass_f.size
    if wp is not None and coords_w is not None and len(coords_w) > 0:
        coords_w = np.ascontiguousarray(coords_w, dtype=cdt)
        mass_w = np.ascontiguousarray(mass_w, dtype=dtype)
        n_w, wpp = coords_w.shape[0], C.byref(wp)
    else:
        coords_w, mass_w, n_w, wpp = np.zeros((0, fp.ndims), dtype=cdt), np.zeros(0, dtype=dtype), 0, None
    out = np.zeros(n_f, dtype=dtype)
    rc = getattr(lib(), f"orc_reinit_density_{s}")(C.byref(fp), wpp, n_f, _ptr(mass_f), n_w, _ptr(coords_w),
                                                    _ptr(mass_w), _ptr(v_ode), _ptr(u_ode), _ptr(out))
    if rc != 0:
        raise RuntimeError(f"orc_reinit_density failed: {rc}")
    return out


def kick_fsi(fp: FluidParams, wp, sp: TlsphParams, mass_f, coords_w, mass_w, n_s_int, x0_s, mass_s, rho_s,
             hydro_mass_s, L, v_ode, u_ode, dtype, nthreads=0, clamped_coords=None, clamped_velocity=None,
             clamped_acceleration=None):
    """One `kick!` of Semidiscretization(fluid, wall, structure) on flat ODE vectors [fluid | structure]
    (see orc_kick_fsi).  Returns dict(dv=flat dv_ode, F=..., pk1_rho2=...).  clamped_*: the prescribed
    motion of the clamped particles (orc_kick_fsi3), (n_s - n_s_int, ND) each or None."""
    dtype = np.dtype(dtype)
    u_ode = np.ascontiguousarray(u_ode)
    cdt = u_ode.dtype
    s = suffix(dtype, cdt)
    v_ode = np.ascontiguousarray(v_ode, dtype=dtype)
    mass_f = np.ascontiguousarray(mass_f, dtype=dtype)
    n_f, nd = mass_f.size, fp.ndims
    x0_s = np.ascontiguousarray(x0_s, dtype=cdt)
    n_s = x0_s.shape[0]
    assert u_ode.size == nd * (n_f + n_s_int) and v_ode.size == (nd + 1) * n_f + nd * n_s_int
    if wp is not None and coords_w is not None and len(coords_w) > 0:
        coords_w = np.ascontiguousarray(coords_w, dtype=cdt)
        mass_w = np.ascontiguousarray(mass_w, dtype=dtype)
        n_w, wpp = coords_w.shape[0], C.byref(wp)
    else:
        coords_w, mass_w, n_w, wpp = np.zeros((0, nd), dtype=cdt), np.zeros(0, dtype=dtype), 0, None
    mass_s, rho_s = np.ascontiguousarray(mass_s, dtype=dtype), np.ascontiguousarray(rho_s, dtype=dtype)
    hydro = np.ascontiguousarray(hydro_mass_s, dtype=dtype)
    L = np.ascontiguousarray(L, dtype=dtype)
    dv = np.zeros_like(v_ode)
    F, P = np.zeros((n_s, nd, nd), dtype=dtype), np.zeros((n_s, nd, nd), dtype=dtype)
    ps, ds = np.zeros(n_s, dtype=dtype), np.zeros(n_s, dtype=dtype)
    n_cl = n_s - n_s_int
    xc = None if clamped_coords is None else np.ascontiguousarray(clamped_coords, dtype=cdt).reshape(n_cl, nd)
    vc = None if clamped_velocity is None else np.ascontiguousarray(clamped_velocity, dtype=dtype).reshape(n_cl, nd)
    ac = None if clamped_acceleration is None else \
        np.ascontiguousarray(clamped_acceleration, dtype=dtype).reshape(n_cl, nd)
    rc = getattr(lib(), f"orc_kick_fsi3_{s}")(C.byref(fp), wpp, C.byref(sp), n_f, _ptr(mass_f), n_w, _ptr(coords_w),
                                               _ptr(mass_w), n_s, n_s_int, _ptr(x0_s), _ptr(mass_s), _ptr(rho_s),
                                               _ptr(hydro), _ptr(L), _ptr(v_ode), _ptr(u_ode), _ptr(dv), _ptr(F),
                                               _ptr(P), _ptr(ps), _ptr(ds), _ptr(xc), _ptr(vc), _ptr(ac),
                                               int(nthreads))
    if rc != 0:
        raise RuntimeError(f"orc_kick_fsi failed: {rc}")
    return dict(dv=dv, F=F, pk1_rho2=P, structure_pressure=ps, structure_density=ds)
