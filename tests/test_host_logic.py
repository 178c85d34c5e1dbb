"""CPU-only checks of the host side: particle generators, ODE layout, argument validation and
that the C-ABI library loads and exports every symbol include/tpb200.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

import trixiparticles.jl_b200 as tp
from trixiparticles.jl_b200 import _lib, examples

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config_sizes_match_survey():
    """SURVEY.md section 8 size table (computed from the example scripts)."""
    f, w, t = examples.dam_break_2d(40)
    assert (f.nparticles, w.nparticles) == (3200, 3912)
    assert t.tank_size == pytest.approx((3.21, 4.005))
    f, w, _ = examples.hydrostatic_water_column_2d()
    assert (f.nparticles, w.nparticles) == (360, 276)
    assert f.eltype == np.float32 and f.coordinates_eltype == np.float32
    f, w, _ = examples.dam_break_3d(0.08)
    assert (f.nparticles, w.nparticles) == (3600, 46800)


def test_rectangular_shape_ordering():
    """rectangular_shape.jl:168-224: default order has x fastest, :x_first has y fastest."""
    c = tp.setups.rectangular_shape_coords(0.1, (3, 2), (0.0, 0.0))
    assert np.allclose(c[:3, 0], [0.05, 0.15, 0.25]) and np.allclose(c[:3, 1], 0.05)
    c = tp.setups.rectangular_shape_coords(0.1, (3, 2), (0.0, 0.0), loop_order="x_first")
    assert np.allclose(c[:2, 1], [0.05, 0.15]) and np.allclose(c[:2, 0], 0.05)
    c = tp.setups.rectangular_shape_coords(1.0, (2, 2, 2), (0, 0, 0), loop_order="x_first")
    assert np.allclose(c[1], [0.5, 0.5, 1.5])


def test_hydrostatic_initialisation():
    """rectangular_shape.jl:226-267: explicit-Euler column from the free surface down."""
    f, _, tank = examples.dam_break_2d(40)
    p = tank.fluid.pressure.reshape(40, 80)  # rows = y
    assert np.all(np.diff(p[:, 0]) < 0)  # pressure decreases upwards
    assert p[-1, 0] == pytest.approx(0.5 * 0.015 * 9.81 * 1000.0)
    assert np.allclose(tank.fluid.mass, tank.fluid.density * 0.015**2)


def test_ode_layout():
    """test/general/semidiscretization.jl:53-54: ranges for (wall-like, fluid-like) systems and
    wcsph write_v0! layout (test/systems/wcsph_system.jl:231-294)."""
    f, w, _ = examples.hydrostatic_water_column_2d()
    semi = tp.Semidiscretization(f, w)
    assert semi.ranges_u == ((0, 720), (720, 720))
    assert semi.ranges_v == ((0, 1080), (1080, 1080))
    semi2 = tp.Semidiscretization(w, f)
    assert semi2.ranges_u == ((0, 0), (0, 720))
    f_s, w_s, _ = examples.hydrostatic_water_column_2d(density_calculator=tp.SummationDensity())
    assert tp.Semidiscretization(f_s, w_s).ranges_v == ((0, 720), (720, 720))


def test_constructor_validation():
    f, w, _ = examples.hydrostatic_water_column_2d()
    f64, w64, _ = examples.hydrostatic_water_column_2d(eltype=np.float64, coordinates_eltype=np.float64)
    with pytest.raises(ValueError):
        tp.Semidiscretization(f, w64)  # mixed eltypes (semidiscretization.jl:303-317)
    with pytest.raises(ValueError):
        tp.Semidiscretization(f, w, interaction_matrix=np.ones((3, 3), bool))
    with pytest.raises(ValueError):
        tp.WeaklyCompressibleSPHSystem(f.initial_condition, smoothing_kernel=tp.WendlandC2Kernel(3),
                                       smoothing_length=0.1, density_calculator=tp.ContinuityDensity(),
                                       state_equation=f.state_equation)
    with pytest.raises(ValueError):
        tp.WeaklyCompressibleSPHSystem(f.initial_condition, smoothing_kernel=tp.WendlandC2Kernel(2),
                                       smoothing_length=0.1, density_calculator=tp.ContinuityDensity(),
                                       state_equation=f.state_equation, acceleration=(0.0, 0.0, 1.0))
    # wall viscosity (no-slip wall): the three models of the accelerated path, nothing else
    m = w.boundary_model
    for visc in (tp.ArtificialViscosityMonaghan(alpha=0.02), tp.ViscosityMorris(nu=1e-3), tp.ViscosityAdami(nu=1e-3)):
        model = tp.BoundaryModelDummyParticles(m.initial_density, m.hydrodynamic_mass, m.density_calculator,
                                               m.smoothing_kernel, m.smoothing_length,
                                               state_equation=m.state_equation, viscosity=visc)
        assert model.viscosity is visc
    with pytest.raises(ValueError):
        tp.BoundaryModelDummyParticles(m.initial_density, m.hydrodynamic_mass, m.density_calculator,
                                       m.smoothing_kernel, m.smoothing_length, state_equation=m.state_equation,
                                       viscosity="no-slip")


def test_state_equation_host_mirror(oracle):
    se = tp.StateEquationCole(sound_speed=1484.0, reference_density=998.34, exponent=7.15,
                              background_pressure=101_325.0)
    assert se.inverse(100 * 101_325.0) == 1002.8323123356663
    for rho in [990.0, 1000.0, 1010.0]:
        assert se(rho) == oracle.eos(1484.0, 7.15, 998.34, 101_325.0, 0, rho)


def test_abi_library_exports_header_symbols():
    """The shared library loads without a GPU and exports exactly what include/tpb200.h declares."""
    header = open(os.path.join(ROOT, "include", "tpb200.h")).read()
    declared = set(re.findall(r"\b(tpb_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    L = _lib.load()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.tpb_version()
    # struct sizes agree between the ctypes mirror and the header (compile-time check in C)
    assert ctypes.sizeof(_lib.Config) == 4 * 10 + 8 * 6
    assert ctypes.sizeof(_lib.FluidParams) == 4 * 6 + 8 * 13 + 4 * 2 + 8 * 3   # + StateEquationAdaptiveCole fields
    assert ctypes.sizeof(_lib.WallParams) == 4 * 4 + 8 * 6 + 4 * 2 + 8 * 3 + 4 * 2   # + wall viscosity (no-slip), + ContinuityDensity


def test_abi_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on CPU."""
    L = _lib.load()
    h = ctypes.c_void_p()
    cfg = _lib.Config()
    cfg.struct_size = 1
    assert L.tpb_create(ctypes.byref(cfg), ctypes.byref(h)) == 1
    assert b"struct_size" in L.tpb_last_error(None)
    cfg.struct_size = ctypes.sizeof(_lib.Config)
    cfg.ndims = 4
    assert L.tpb_create(ctypes.byref(cfg), ctypes.byref(h)) == 1
    cfg.ndims = 3
    cfg.eltype, cfg.coords_eltype = _lib.F64, _lib.F32
    assert L.tpb_create(ctypes.byref(cfg), ctypes.byref(h)) == 1
    assert L.tpb_kick(None, None, None, None, 0.0) == 1
    assert L.tpb_destroy(None) == 0


def test_product_path_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the package may import or load it."""
    pkg = os.path.join(ROOT, "trixiparticles.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.lower() or fn == "tpb_tiles.cuh" and False, fn


def test_moving_wall_next_to_a_structure_shares_the_slot_only_when_exact():
    """A moving dummy-particle wall and a structure with the same boundary model share the library's structure slot
    (dam_break_gate_2d.jl); a different boundary model or a gate inside the structure's kernel support is refused."""
    import numpy as np
    import pytest
    import trixiparticles.jl_b200 as tp
    from trixiparticles.jl_b200 import examples
    fluid, tank_w, gate_w, plate, _ = examples.dam_break_gate_2d(0.02)
    semi = tp.Semidiscretization(fluid, tank_w, gate_w, plate)
    assert semi._gate is gate_w and semi.lib_index(gate_w) == semi.lib_index(plate) == 2
    assert semi.system_index(gate_w) == 2 and semi.system_index(plate) == 3
    assert semi.ranges_u[2] == (400, 400)          # the gate has no rows in the ODE vectors
    # another pressure offset on the gate: not the same boundary model any more
    m = gate_w.boundary_model
    other = tp.BoundaryModelDummyParticles(m.initial_density, m.hydrodynamic_mass,
                                           tp.AdamiPressureExtrapolation(pressure_offset=10.0), m.smoothing_kernel,
                                           m.smoothing_length, state_equation=m.state_equation,
                                           clip_negative_pressure=True)
    gate2 = tp.WallBoundarySystem(gate_w.initial_condition, other, prescribed_motion=gate_w.prescribed_motion)
    with pytest.raises(ValueError, match="share one slot"):
        tp.Semidiscretization(fluid, tank_w, gate2, plate)
    # a gate that starts inside the plate's kernel support
    ic = gate_w.initial_condition
    near = tp.InitialCondition(ic.coordinates + [plate.initial_coordinates[:, 0].min() - ic.coordinates[:, 0].max() - 0.005, 0.0],
                               ic.velocity, ic.mass, ic.density, ic.pressure, ic.particle_spacing)
    m2 = tp.BoundaryModelDummyParticles(m.initial_density, m.hydrodynamic_mass, tp.AdamiPressureExtrapolation(),
                                        m.smoothing_kernel, m.smoothing_length, state_equation=m.state_equation,
                                        clip_negative_pressure=True)
    gate3 = tp.WallBoundarySystem(near, m2, prescribed_motion=gate_w.prescribed_motion)
    with pytest.raises(ValueError, match="kernel support"):
        tp.Semidiscretization(fluid, tank_w, gate3, plate)
