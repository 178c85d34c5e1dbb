"""ctypes binding of libtpb200.so -- exactly the entry points declared in include/tpb200.h.

There is no CPU fallback: if the CUDA library is missing this module raises on first use.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libtpb200.so")

TPB_OK = 0
ERR_NAMES = {1: "INVALID_ARGUMENT", 2: "UNSUPPORTED", 3: "CUDA", 4: "OUT_OF_BOUNDS", 5: "STATE", 6: "CAPACITY"}
F32, F64 = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
FIELD_PRESSURE, FIELD_DENSITY, FIELD_VOLUME, FIELD_WALL_VELOCITY = 0, 1, 2, 3
FIELD_DEFORMATION_GRADIENT, FIELD_PK1_RHO2, FIELD_CORRECTION_MATRIX = 4, 5, 6
BOUNDARY_NONE, BOUNDARY_MONAGHAN_KAJTAR, BOUNDARY_DUMMY_PARTICLES = 0, 1, 2

EXPORTS = [
    "tpb_version", "tpb_last_error", "tpb_create", "tpb_destroy", "tpb_add_fluid_system",
    "tpb_add_wall_system", "tpb_add_structure_system", "tpb_set_interaction", "tpb_semidiscretize", "tpb_ode_sizes",
    "tpb_system_range", "tpb_kick", "tpb_drift", "tpb_get_system_field", "tpb_neighbor_pairs",
    "tpb_synchronize", "tpb_set_stream", "tpb_get_stats", "tpb_get_sound_speed", "tpb_max_speed2",
    "tpb_set_max_speed2", "tpb_set_integrate_structure", "tpb_structure_fluid_force", "tpb_kick_structure",
    "tpb_set_clamped_motion", "tpb_sort_system", "tpb_set_structure_material", "tpb_reinit_density",
    "tpb_host_register",
    "tpb_host_unregister", "tpb_set_profiling", "tpb_get_phase_times",
    "tpb_set_fluid_count", "tpb_set_fluid_mass",
    "tpb_vec_axpby", "tpb_vec_rk2n_stage", "tpb_vec_fill", "tpb_vec_lincomb4", "tpb_vec_verlet_update",
    "tpb_vec_div_fast", "tpb_vec_strided_max", "tpb_vec_wrms_norm",
    "tpb_peer_alloc", "tpb_peer_free", "tpb_peer_export", "tpb_peer_import", "tpb_peer_close",
    "tpb_halo_pack", "tpb_halo_install",
]
PHASES = ("rebuild", "density", "boundary", "interact")


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("ndims", C.c_int32), ("eltype", C.c_int32),
        ("coords_eltype", C.c_int32), ("device", C.c_int32), ("ode_memory", C.c_int32),
        ("has_bounds", C.c_int32), ("max_points_per_cell", C.c_int32),
        ("deterministic", C.c_int32), ("interact_variant", C.c_int32),
        ("min_corner", C.c_double * 3), ("max_corner", C.c_double * 3),
    ]


class FluidParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("kernel", C.c_int32), ("density_calculator", C.c_int32),
        ("clip_negative_pressure", C.c_int32), ("has_viscosity", C.c_int32),
        ("has_diffusion", C.c_int32), ("smoothing_length", C.c_double),
        ("sound_speed", C.c_double), ("exponent", C.c_double), ("reference_density", C.c_double),
        ("background_pressure", C.c_double), ("alpha", C.c_double), ("beta", C.c_double),
        ("epsilon", C.c_double), ("delta", C.c_double), ("acceleration", C.c_double * 3),
        ("damping_coefficient", C.c_double),
        ("adaptive_sound_speed", C.c_int32), ("adaptive_params_f32", C.c_int32),
        ("mach_number_target", C.c_double), ("min_sound_speed", C.c_double), ("max_sound_speed", C.c_double),
    ]


class WallParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("kernel", C.c_int32), ("clip_negative_pressure", C.c_int32),
        ("sound_speed_from_fluid", C.c_int32), ("smoothing_length", C.c_double), ("sound_speed", C.c_double),
        ("exponent", C.c_double), ("reference_density", C.c_double),
        ("background_pressure", C.c_double), ("pressure_offset", C.c_double),
        ("has_viscosity", C.c_int32), ("density_calculator", C.c_int32),
        ("alpha", C.c_double), ("beta", C.c_double), ("epsilon", C.c_double),
        ("eos_clip_negative_pressure", C.c_int32), ("reserved", C.c_int32),
    ]


WALL_DENSITY_ADAMI, WALL_DENSITY_CONTINUITY = 0, 1


class StructureParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("kernel", C.c_int32), ("has_penalty_force", C.c_int32),
        ("boundary_model", C.c_int32), ("smoothing_length", C.c_double), ("young_modulus", C.c_double),
        ("poisson_ratio", C.c_double), ("penalty_alpha", C.c_double), ("acceleration", C.c_double * 3),
        ("mk_K", C.c_double), ("mk_beta", C.c_double), ("mk_spacing", C.c_double),
        ("bm_kernel", C.c_int32), ("bm_clip_negative_pressure", C.c_int32), ("bm_smoothing_length", C.c_double),
        ("bm_sound_speed", C.c_double), ("bm_exponent", C.c_double), ("bm_reference_density", C.c_double),
        ("bm_background_pressure", C.c_double), ("bm_pressure_offset", C.c_double),
        ("bm_wall_semantics", C.c_int32), ("bm_reserved", C.c_int32), ("bm_bernoulli_factor", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("kernel_launches_total", C.c_int64), ("kicks", C.c_int64), ("drifts", C.c_int64),
        ("n_cells", C.c_int64), ("launches_last_kick", C.c_int32),
        ("launches_last_drift", C.c_int32), ("interact_variant_used", C.c_int32),
        ("reserved", C.c_int32),
    ]


class TpbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"tpb200 error {code} ({ERR_NAMES.get(code, '?')}): {message}")
        self.code = code


_lib = None


def load():
    """Load libtpb200.so; raises loudly when it has not been built (no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise RuntimeError(
            f"{SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the accelerated path.")
    L = C.CDLL(SO)
    i32, i64, p, d = C.c_int32, C.c_int64, C.c_void_p, C.c_double
    L.tpb_version.restype = C.c_char_p; L.tpb_version.argtypes = []
    L.tpb_last_error.restype = C.c_char_p; L.tpb_last_error.argtypes = [p]
    L.tpb_create.restype = i32; L.tpb_create.argtypes = [C.POINTER(Config), C.POINTER(p)]
    L.tpb_destroy.restype = i32; L.tpb_destroy.argtypes = [p]
    L.tpb_add_fluid_system.restype = i32
    L.tpb_add_fluid_system.argtypes = [p, C.POINTER(FluidParams), i64, p, C.POINTER(i32)]
    L.tpb_add_wall_system.restype = i32
    L.tpb_add_wall_system.argtypes = [p, C.POINTER(WallParams), i64, p, p, p, C.POINTER(i32)]
    L.tpb_add_structure_system.restype = i32
    L.tpb_add_structure_system.argtypes = [p, C.POINTER(StructureParams), i64, i64, p, p, p, p, C.POINTER(i32)]
    L.tpb_set_interaction.restype = i32; L.tpb_set_interaction.argtypes = [p, i32, i32, i32]
    L.tpb_semidiscretize.restype = i32; L.tpb_semidiscretize.argtypes = [p, p]
    L.tpb_ode_sizes.restype = i32; L.tpb_ode_sizes.argtypes = [p, C.POINTER(i64), C.POINTER(i64)]
    L.tpb_system_range.restype = i32
    L.tpb_system_range.argtypes = [p, i32] + [C.POINTER(i64)] * 4
    L.tpb_kick.restype = i32; L.tpb_kick.argtypes = [p, p, p, p, d]
    L.tpb_drift.restype = i32; L.tpb_drift.argtypes = [p, p, p, p, d]
    L.tpb_get_system_field.restype = i32; L.tpb_get_system_field.argtypes = [p, i32, i32, p, i64]
    L.tpb_neighbor_pairs.restype = i32
    L.tpb_neighbor_pairs.argtypes = [p, i32, i32, p, i64, p, p, C.POINTER(i64)]
    L.tpb_synchronize.restype = i32; L.tpb_synchronize.argtypes = [p]
    L.tpb_set_stream.restype = i32; L.tpb_set_stream.argtypes = [p, p]
    L.tpb_get_stats.restype = i32; L.tpb_get_stats.argtypes = [p, C.POINTER(Stats)]
    L.tpb_host_register.restype = i32; L.tpb_host_register.argtypes = [p, i64]
    L.tpb_get_sound_speed.restype = i32; L.tpb_get_sound_speed.argtypes = [p, C.POINTER(d)]
    L.tpb_max_speed2.restype = i32; L.tpb_max_speed2.argtypes = [p, p, p]
    L.tpb_set_max_speed2.restype = i32; L.tpb_set_max_speed2.argtypes = [p, p]
    L.tpb_set_integrate_structure.restype = i32; L.tpb_set_integrate_structure.argtypes = [p, i32]
    L.tpb_structure_fluid_force.restype = i32; L.tpb_structure_fluid_force.argtypes = [p, p, p, p]
    L.tpb_kick_structure.restype = i32; L.tpb_kick_structure.argtypes = [p, p, p, p, p]
    L.tpb_set_clamped_motion.restype = i32; L.tpb_set_clamped_motion.argtypes = [p, p, p, p, i32]
    L.tpb_sort_system.restype = i32; L.tpb_sort_system.argtypes = [p, i32, p, p]
    L.tpb_set_structure_material.restype = i32; L.tpb_set_structure_material.argtypes = [p, p, p]
    L.tpb_reinit_density.restype = i32; L.tpb_reinit_density.argtypes = [p, p, p]
    u32 = C.c_uint32
    L.tpb_peer_alloc.restype = i32; L.tpb_peer_alloc.argtypes = [i64, C.POINTER(p)]
    L.tpb_peer_free.restype = i32; L.tpb_peer_free.argtypes = [p]
    L.tpb_peer_export.restype = i32; L.tpb_peer_export.argtypes = [p, p]
    L.tpb_peer_import.restype = i32; L.tpb_peer_import.argtypes = [p, C.POINTER(p)]
    L.tpb_peer_close.restype = i32; L.tpb_peer_close.argtypes = [p]
    L.tpb_halo_pack.restype = i32
    L.tpb_halo_pack.argtypes = [p, i32, d, p, p, p, i64, p, p, p, i64, p, u32, C.POINTER(i32)]
    L.tpb_halo_install.restype = i32
    L.tpb_halo_install.argtypes = [p, i64, p, p, p, p, p, p, u32, d, p]
    L.tpb_host_unregister.restype = i32; L.tpb_host_unregister.argtypes = [p]
    L.tpb_set_profiling.restype = i32; L.tpb_set_profiling.argtypes = [p, i32]
    L.tpb_get_phase_times.restype = i32
    L.tpb_get_phase_times.argtypes = [p, C.POINTER(d), C.POINTER(i32)]
    L.tpb_vec_axpby.restype = i32; L.tpb_vec_axpby.argtypes = [p, i64, i32, d, p, d, p]
    L.tpb_vec_rk2n_stage.restype = i32; L.tpb_vec_rk2n_stage.argtypes = [p, i64, i32, d, d, d, p, p, p]
    L.tpb_vec_fill.restype = i32; L.tpb_vec_fill.argtypes = [p, i64, i32, d, p]
    L.tpb_vec_lincomb4.restype = i32; L.tpb_vec_lincomb4.argtypes = [p, i64, i32, d, p, d, p, d, p, d, p, p]
    L.tpb_vec_verlet_update.restype = i32; L.tpb_vec_verlet_update.argtypes = [p, i64, i32, i32, i32, d, p, p, p]
    L.tpb_vec_div_fast.restype = i32; L.tpb_vec_div_fast.argtypes = [p, i64, i32, d, p, p]
    L.tpb_vec_strided_max.restype = i32
    L.tpb_vec_strided_max.argtypes = [p, i64, i32, i32, i32, p, C.POINTER(d)]
    L.tpb_vec_wrms_norm.restype = i32
    L.tpb_vec_wrms_norm.argtypes = [p, i64, i32, p, p, p, d, d, C.POINTER(d)]
    L.tpb_set_fluid_count.restype = i32; L.tpb_set_fluid_count.argtypes = [p, i64, i64]
    L.tpb_set_fluid_mass.restype = i32; L.tpb_set_fluid_mass.argtypes = [p, i64, i64, p]
    _lib = L
    return L


def check(handle, code):
    if code != TPB_OK:
        msg = load().tpb_last_error(handle)
        raise TpbError(code, msg.decode() if msg else "")
