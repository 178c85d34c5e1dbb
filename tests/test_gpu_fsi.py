"""BASELINE config 5 on the GPU: WCSPH fluid + dummy-particle tank + TotalLagrangianSPHSystem plate
(examples/fsi/dam_break_plate_2d.jl with the dimensions of the reference's GPU tests,
test/examples/gpu.jl:664-726) through the C ABI against the CPU oracle.  `-m gpu` only."""
import numpy as np
import pytest

import trixiparticles.jl_b200 as tp
from trixiparticles.jl_b200 import examples
from oracle import adapter

pytestmark = pytest.mark.gpu


def fsi_state(fluid, structure, seed=7, deform=0.05):
    """ODE vectors [fluid | structure]: the fluid as `perturbed_state`, the plate bent and moving."""
    u_f, v_f = examples.perturbed_state(fluid, seed=seed, position_jitter=0.05)
    rng = np.random.default_rng(seed + 1)
    n_int = structure.n_integrated_particles
    x0 = structure.initial_coordinates[:n_int].astype(np.float64)
    ds = structure.initial_condition.particle_spacing
    height = x0[:, 1] - x0[:, 1].min()
    u_s = x0.copy()
    u_s[:, 0] += 0.5 * height ** 2 / max(height.max(), 1e-12)           # bending
    u_s += rng.uniform(-deform * ds, deform * ds, u_s.shape)              # + noise (non-affine)
    v_s = rng.uniform(-0.5, 0.5, u_s.shape)
    u = np.concatenate([u_f.reshape(-1), u_s.astype(structure.coordinates_eltype).reshape(-1)])
    v = np.concatenate([v_f.reshape(-1), v_s.astype(structure.eltype).reshape(-1)])
    return np.ascontiguousarray(u), np.ascontiguousarray(v)


@pytest.mark.parametrize("eltype,coords,tol", [(np.float64, np.float64, 1e-11), (np.float32, np.float32, 2e-5),
                                                (np.float32, np.float64, 2e-5)])
@pytest.mark.parametrize("memory", ["host", "device"])
@pytest.mark.parametrize("boundary_model", ["monaghan_kajtar", "dummy_particles"])
def test_fsi_kick_matches_oracle(eltype, coords, tol, memory, boundary_model):
    """`boundary_model`: BoundaryModelMonaghanKajtar as the example ships it, or the BoundaryModelDummyParticles
    (Adami) alternative of dam_break_plate_2d.jl:120-133 / hydrostatic_water_column_2d.jl:109-124."""
    # the plate next to the water column, so that the coupling is active
    fluid, wall, structure, _ = examples.dam_break_plate_2d(
        0.01, eltype=eltype, coordinates_eltype=coords, initial_fluid_size=(0.15, 0.29), plate_position=(0.165, 0.0),
        structure_boundary_model=boundary_model)
    u, v = fsi_state(fluid, structure)
    ref = adapter.kick_fsi(fluid, wall, structure, u, v)
    semi = tp.Semidiscretization(fluid, wall, structure,
                                 parallelization_backend=tp.B200Backend(device=0, ode_memory=memory))
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    assert semi.ranges_u[-1][1] == u.size and semi.ranges_v[-1][1] == v.size
    if memory == "device":
        import torch
        dev = ode.u0.device
        u_d, v_d = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
        dv_d, du_d = torch.full_like(v_d, float("nan")), torch.full_like(u_d, float("nan"))
        ode.f1(dv_d, v_d, u_d, ode.p, 0.0)
        ode.f2(du_d, v_d, u_d, ode.p, 0.0)
        semi.synchronize()
        dv, du = dv_d.cpu().numpy(), du_d.cpu().numpy()
    else:
        dv, du = np.full_like(v, np.nan), np.full_like(u, np.nan)
        ode.f1(dv, v, u, ode.p, 0.0)
        ode.f2(du, v, u, ode.p, 0.0)
    nd, n_f, n_int = 2, fluid.nparticles, structure.n_integrated_particles
    # correction matrix, deformation gradient, PK1 / rho^2
    for name, key in (("correction_matrix", "L"), ("deformation_grad", "F"), ("pk1_rho2", "pk1_rho2")):
        got = semi.system_field(structure, name)
        scale = np.abs(ref[key]).max()
        assert np.abs(got - ref[key]).max() <= tol * scale, (name, np.abs(got - ref[key]).max() / scale)
    dv_f, ref_f = dv[: 3 * n_f].reshape(n_f, 3), ref["dv"][: 3 * n_f].reshape(n_f, 3)
    dv_s, ref_s = dv[3 * n_f:].reshape(n_int, 2), ref["dv"][3 * n_f:].reshape(n_int, 2)
    assert np.isfinite(dv).all()
    # the coupling is active on both sides
    no_plate = adapter.kick(fluid, wall, u[: 2 * n_f].reshape(n_f, 2), v[: 3 * n_f].reshape(n_f, 3))["dv"]
    assert np.abs(ref_f - no_plate).max() > 1.0
    if boundary_model == "dummy_particles":
        for name, key in (("pressure", "structure_pressure"), ("density", "structure_density")):
            got = semi.system_field(structure, name)
            assert np.abs(got - ref[key]).max() <= 10 * tol * np.abs(ref[key]).max(), name
    for name, a, b in (("fluid acceleration", dv_f[:, :2], ref_f[:, :2]), ("fluid drho", dv_f[:, 2], ref_f[:, 2]),
                       ("structure acceleration", dv_s, ref_s)):
        err = np.abs(a - b).max() / np.abs(b).max()
        assert err <= tol, (name, err)
    # drift!: du = v for both systems
    assert np.array_equal(du[: 2 * n_f].reshape(n_f, 2), v[: 3 * n_f].reshape(n_f, 3)[:, :2].astype(u.dtype))
    assert np.array_equal(du[2 * n_f:], v[3 * n_f:].astype(u.dtype))
    semi.close()


def test_fsi_structure_first_ordering():
    """Semidiscretization(structure, fluid, wall) -- the order of examples/fsi/hydrostatic_water_column_2d.jl:
    the ODE vectors are laid out [structure | fluid]; same physics."""
    fluid, wall, structure, _ = examples.dam_break_plate_2d(
        0.01, initial_fluid_size=(0.15, 0.29), plate_position=(0.165, 0.0))
    u, v = fsi_state(fluid, structure)
    ref = adapter.kick_fsi(fluid, wall, structure, u, v)["dv"]
    n_f, n_int = fluid.nparticles, structure.n_integrated_particles
    u2 = np.concatenate([u[2 * n_f:], u[: 2 * n_f]])
    v2 = np.concatenate([v[3 * n_f:], v[: 3 * n_f]])
    semi = tp.Semidiscretization(structure, fluid, wall, parallelization_backend=tp.B200Backend(device=0))
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    assert semi.ranges_u == ((0, 2 * n_int), (2 * n_int, 2 * n_int + 2 * n_f), (2 * n_int + 2 * n_f,) * 2)
    dv = np.full_like(v2, np.nan)
    ode.f1(dv, v2, u2, ode.p, 0.0)
    got = np.concatenate([dv[2 * n_int:], dv[: 2 * n_int]])
    assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max()
    semi.close()


def test_dam_break_plate_2d_time_loop():
    """test/examples/gpu.jl:696-726: Float32, initial_fluid_size = (0.15, 0.29), CarpenterKennedy2N54 with
    StepsizeCallback(cfl = 1.2) to t = 0.05: the run completes, nothing leaves the tank, the clamped
    base holds the plate and the plate has not moved by more than a fraction of its thickness before
    the water arrives."""
    from trixiparticles.jl_b200.time_integration import CarpenterKennedy2N54, StepsizeCallback, solve
    fluid, wall, structure, _ = examples.dam_break_plate_2d(
        0.01, eltype=np.float32, coordinates_eltype=np.float32, initial_fluid_size=(0.15, 0.29))
    semi = tp.Semidiscretization(fluid, wall, structure,
                                 parallelization_backend=tp.B200Backend(device=0, ode_memory="device"))
    ode = tp.semidiscretize(semi, (0.0, 0.05))
    cb = StepsizeCallback(cfl=1.2)
    sol = solve(ode, CarpenterKennedy2N54(williamson_condition=False), dt=cb.dt(semi), callback=cb)
    assert sol.retcode == "Success"
    u, v = sol.u.cpu().numpy(), sol.v.cpu().numpy()
    assert np.isfinite(u).all() and np.isfinite(v).all()
    n_f, n_int = fluid.nparticles, structure.n_integrated_particles
    u_s = u[2 * n_f:].reshape(n_int, 2)
    disp = np.abs(u_s - structure.initial_coordinates[:n_int]).max()
    assert disp < 0.012, disp          # plate thickness
    u_f = u[: 2 * n_f].reshape(n_f, 2)
    assert u_f[:, 0].max() > 0.16       # the column has started to collapse
    semi.close()


def test_hydrostatic_water_column_fsi_validation_trace():
    """The reference's FSI validation run (validation/hydrostatic_water_column_2d/validation.jl, n_particles_plate_y
    = 3; test/validation/validation.jl:95-109): a 2 m water column on an elastic aluminium plate clamped at both
    ends, TLSPH with dummy-particle (Adami) coupling.  One CarpenterKennedy2N54 loop with the plate's CFL step
    (35 000 steps to t = 0.3, CUDA-graph replay) reproduces the reference's own mid-plate deflection trace
    (121 samples, tests/golden/fsi_hydrostatic_wcsph_3_trace.json) and meets the reference's bar against the
    analytical deflection (relative error of the average over t >= 0.25: <= 0.045)."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import run_fsi_hydrostatic_validation as V
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                      "fsi_hydrostatic_wcsph_3_trace.json")))
    r = V.run(t_end=0.3)
    assert r["finite"]
    rt, ry = np.array(ref["time"]), np.array(ref["y_deflection_structure_1"])
    assert len(r["times"]) == len(rt) and np.allclose(r["times"], rt, atol=1e-9)
    a = abs(ref["analytical_value"])
    err = np.abs(r["y"] - ry).max() / a
    print(f"max |deflection - reference trace| = {err:.2e} x analytical value; average over t >= 0.25: "
          f"{r['avg']:.4e} (relative error {r['rel_error']:.4f})")
    assert err <= 5e-3, err                      # measured: a few 1e-5 -- the traces agree to four digits
    assert r["rel_error"] <= 0.045               # the reference's own bar
    ref_avg = ry[rt >= 0.25 - 1e-12].mean()
    assert abs(r["avg"] - ref_avg) <= 1e-3 * a


def test_split_integration_pieces_match_oracle():
    """The three entry points of the SplitIntegrationCallback (callbacks/split_integration.jl): with
    `integrate_tlsph = false` kick! / drift! leave the structure's rows zero while the fluid still feels the plate;
    `tpb_structure_fluid_force` (other_interaction_split!) + `tpb_kick_structure` (kick_split!) together give the
    structure rows of the full oracle kick."""
    import ctypes as C
    import torch
    from trixiparticles.jl_b200 import _lib
    for bm in ("monaghan_kajtar", "dummy_particles"):
        fluid, wall, structure, _ = examples.dam_break_plate_2d(
            0.01, initial_fluid_size=(0.15, 0.29), plate_position=(0.165, 0.0), structure_boundary_model=bm)
        u, v = fsi_state(fluid, structure)
        ref = adapter.kick_fsi(fluid, wall, structure, u, v)["dv"]
        semi = tp.Semidiscretization(fluid, wall, structure,
                                     parallelization_backend=tp.B200Backend(device=0, ode_memory="device"))
        ode = tp.semidiscretize(semi, (0.0, 1.0))
        dev = ode.u0.device
        u_d, v_d = torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev)
        n_f, n_int = fluid.nparticles, structure.n_integrated_particles
        semi.set_integrate_structure(False)
        dv_d, du_d = torch.full_like(v_d, float("nan")), torch.full_like(u_d, float("nan"))
        ode.f1(dv_d, v_d, u_d, ode.p, 0.0)
        ode.f2(du_d, v_d, u_d, ode.p, 0.0)
        semi.synchronize()
        dv, du = dv_d.cpu().numpy(), du_d.cpu().numpy()
        assert np.all(dv[3 * n_f:] == 0) and np.all(du[2 * n_f:] == 0)
        assert np.abs(dv[: 3 * n_f] - ref[: 3 * n_f]).max() <= 1e-11 * np.abs(ref[: 3 * n_f]).max()
        L, h = _lib.load(), semi._handle
        force = torch.full((2 * n_int,), float("nan"), dtype=v_d.dtype, device=dev)
        dv_s = torch.full_like(force, float("nan"))
        _lib.check(h, L.tpb_structure_fluid_force(h, C.c_void_p(force.data_ptr()), C.c_void_p(v_d.data_ptr()),
                                                  C.c_void_p(u_d.data_ptr())))
        v_s, u_s = v_d[3 * n_f:].clone(), u_d[2 * n_f:].clone()
        _lib.check(h, L.tpb_kick_structure(h, C.c_void_p(dv_s.data_ptr()), C.c_void_p(v_s.data_ptr()),
                                           C.c_void_p(u_s.data_ptr()), C.c_void_p(force.data_ptr())))
        semi.synchronize()
        got, want = dv_s.cpu().numpy(), ref[3 * n_f:]
        assert np.abs(force.cpu().numpy()).max() > 0
        assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max(), bm
        semi.close()


def test_dam_break_plate_2d_split_integration():
    """test/examples/gpu.jl:662-705: `fsi/dam_break_plate_2d.jl` with a SplitIntegrationCallback (CarpenterKennedy2N54
    sub-integrator, dt = 5e-5, stage_coupling = true) in Float32, stiffer plate (E = 1e7) moved next to the water
    column, tspan = (0, 0.2): the reference verifies that fewer than 400 iterations of the main integrator are
    needed (its step is set by the fluid alone).  Here also: the plate's tip follows the monolithic run, which
    integrates everything with the plate's step."""
    from trixiparticles.jl_b200.time_integration import (CarpenterKennedy2N54, SplitIntegrationCallback,
                                                         StepsizeCallback, solve)

    def run(split):
        fluid, wall, structure, _ = examples.dam_break_plate_2d(
            0.01, eltype=np.float32, coordinates_eltype=np.float32, initial_fluid_size=(0.15, 0.29),
            plate_position=(0.2, 0.0), E=1e7)
        semi = tp.Semidiscretization(fluid, wall, structure,
                                     parallelization_backend=tp.B200Backend(device=0, ode_memory="device"))
        ode = tp.semidiscretize(semi, (0.0, 0.2))
        cbs = [StepsizeCallback(cfl=1.2)]
        if split:
            cbs.append(SplitIntegrationCallback(CarpenterKennedy2N54(williamson_condition=False), stage_coupling=True,
                                                dt=5e-5))
        sol = solve(ode, CarpenterKennedy2N54(williamson_condition=False), callback=cbs,
                    maxiters=400 if split else 10 ** 6, cuda_graph=not split)
        u = sol.u.cpu().numpy()
        n_f, n_int = fluid.nparticles, structure.n_integrated_particles
        tip = u[2 * n_f:].reshape(n_int, 2)[-1] - structure.initial_coordinates[n_int - 1]
        out = dict(nsteps=sol.nsteps, t=sol.t, tip=tip, finite=bool(np.isfinite(u).all()),
                   substeps=cbs[-1].n_substeps if split else 0)
        semi.close()
        return out

    a = run(split=True)
    assert a["finite"] and abs(a["t"] - 0.2) < 1e-9 and a["nsteps"] < 400, a
    b = run(split=False)
    assert b["nsteps"] > 3 * a["nsteps"]                  # the plate's CFL step is far smaller than the fluid's
    print(f"split: {a['nsteps']} steps + {a['substeps']} sub-steps, tip displacement {a['tip']}; "
          f"monolithic: {b['nsteps']} steps, tip displacement {b['tip']}")
    assert a["tip"][0] > 1e-4 and b["tip"][0] > 1e-4     # the water has bent the plate
    assert abs(a["tip"][0] - b["tip"][0]) <= 0.25 * abs(b["tip"][0]) + 2e-4


@pytest.mark.parametrize("eltype,coords,tol", [(np.float64, np.float64, 1e-11), (np.float32, np.float32, 3e-5),
                                                (np.float32, np.float64, 3e-5)])
@pytest.mark.parametrize("boundary_model", ["monaghan_kajtar", "dummy_particles"])
def test_fsi_3d_kick_matches_oracle(eltype, coords, tol, boundary_model):
    """BASELINE config 5 names "2D/3D": the plate set-up extruded along z (examples.dam_break_plate_3d; 3 x 3
    matrices, WendlandC2{3}, 27-cell neighbourhoods) -- correction matrix, deformation gradient, PK1 and every
    coupling term against the oracle, whose 3-D TLSPH part is pinned by affine-deformation known answers and the
    fluid/plate momentum balance (tests/test_oracle_tlsph.py)."""
    fluid, wall, structure, _ = examples.dam_break_plate_3d(
        0.02, eltype=eltype, coordinates_eltype=coords, plate_position=(0.175, 0.0, 0.0075),
        structure_boundary_model=boundary_model)
    u, v = fsi_state(fluid, structure)
    ref = adapter.kick_fsi(fluid, wall, structure, u, v)
    semi = tp.Semidiscretization(fluid, wall, structure, parallelization_backend=tp.B200Backend(device=0))
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    assert semi.ranges_u[-1][1] == u.size and semi.ranges_v[-1][1] == v.size
    dv, du = np.full_like(v, np.nan), np.full_like(u, np.nan)
    ode.f1(dv, v, u, ode.p, 0.0)
    ode.f2(du, v, u, ode.p, 0.0)
    n_f, n_int = fluid.nparticles, structure.n_integrated_particles
    for name, key in (("correction_matrix", "L"), ("deformation_grad", "F"), ("pk1_rho2", "pk1_rho2")):
        got = semi.system_field(structure, name)
        assert got.shape == (structure.nparticles, 3, 3)
        scale = np.abs(ref[key]).max()
        assert np.abs(got - ref[key]).max() <= tol * scale, (name, np.abs(got - ref[key]).max() / scale)
    dv_f, ref_f = dv[: 4 * n_f].reshape(n_f, 4), ref["dv"][: 4 * n_f].reshape(n_f, 4)
    dv_s, ref_s = dv[4 * n_f:].reshape(n_int, 3), ref["dv"][4 * n_f:].reshape(n_int, 3)
    assert np.isfinite(dv).all()
    no_plate = adapter.kick(fluid, wall, u[: 3 * n_f].reshape(n_f, 3), v[: 4 * n_f].reshape(n_f, 4))["dv"]
    assert np.abs(ref_f - no_plate).max() > 1.0        # the coupling is active
    for name, a, b in (("fluid acceleration", dv_f[:, :3], ref_f[:, :3]), ("fluid drho", dv_f[:, 3], ref_f[:, 3]),
                       ("structure acceleration", dv_s, ref_s)):
        err = np.abs(a - b).max() / np.abs(b).max()
        assert err <= tol, (name, err)
    # drift!: du = v for both systems
    assert np.array_equal(du[: 3 * n_f].reshape(n_f, 3), v[: 4 * n_f].reshape(n_f, 4)[:, :3].astype(du.dtype))
    assert np.array_equal(du[3 * n_f:], v[4 * n_f:].astype(du.dtype))
    semi.close()


def test_fsi_3d_time_loop_plate_bends_away_from_the_column():
    """A short run of the 3-D set-up (CarpenterKennedy2N54, CUDA-graph replay): the column collapses onto the plate,
    whose free edge is pushed in +x while the clamped base stays; the state stays finite and uniform along z to the
    extent the set-up is (the tank's front and back walls break the symmetry only slightly)."""
    from trixiparticles.jl_b200.time_integration import CarpenterKennedy2N54, solve
    fluid, wall, structure, _ = examples.dam_break_plate_3d(0.02, eltype=np.float32, coordinates_eltype=np.float32,
                                                            plate_position=(0.175, 0.0, 0.0075))
    semi = tp.Semidiscretization(fluid, wall, structure, parallelization_backend=tp.B200Backend(ode_memory="device"))
    ode = tp.semidiscretize(semi, (0.0, 0.08))
    sol = solve(ode, CarpenterKennedy2N54(), dt=2e-5, cuda_graph=True)
    assert sol.retcode == "Success"
    u = sol.u.cpu().numpy()
    assert np.isfinite(u).all()
    n_f, n_int = fluid.nparticles, structure.n_integrated_particles
    x_s = u[3 * n_f:].reshape(n_int, 3)
    x0 = structure.initial_coordinates[:n_int]
    top = x0[:, 1] > x0[:, 1].max() - 1e-6
    deflection = (x_s[top, 0] - x0[top, 0])
    assert deflection.mean() > 1e-4                      # pushed away from the column
    assert deflection.std() < 0.5 * abs(deflection.mean())   # roughly uniform along the depth
    semi.close()


def test_oscillating_beam_2d_validation_trace():
    """validation/oscillating_beam_2d (test/validation/validation.jl:13-29): a clamped elastic beam swinging under
    gravity, TLSPH only -- a structure-only semidiscretization (the large-deformation counterpart of the hydrostatic
    plate above: the tip moves by a third of the beam's length).  The reference integrates with RDPK3SpFSAL49 at
    abstol 1e-8 / reltol 1e-6, i.e. to time-step convergence; CarpenterKennedy2N54 with dt = 2e-5 is converged as well,
    and the deflection of the mid particle of the free end follows the reference's own trace
    (tests/golden/oscillating_beam_2d_5_trace.json, 101 samples to t = 1)."""
    import json
    import os
    from trixiparticles.jl_b200.model import PenaltyForceGanzenmueller
    from trixiparticles.jl_b200.time_integration import CarpenterKennedy2N54, PostprocessCallback, solve
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                      "oscillating_beam_2d_5_trace.json")))
    beam, info = examples.oscillating_beam_2d(5, penalty_force=PenaltyForceGanzenmueller(alpha=0.01))
    semi = tp.Semidiscretization(beam, parallelization_backend=tp.B200Backend(ode_memory="device"))
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    assert semi.ranges_u[-1][1] == 2 * beam.n_integrated_particles
    mid, x0 = info["mid_particle"], info["start_position"]
    pp = PostprocessCallback(0.01,
                             deflection_x=lambda sys_, v, u, s, t: float(u[2 * mid].item()) - x0[0],
                             deflection_y=lambda sys_, v, u, s, t: float(u[2 * mid + 1].item()) - x0[1])
    sol = solve(ode, CarpenterKennedy2N54(), dt=2e-5, callback=(pp,), cuda_graph=True)
    assert sol.retcode == "Success"
    n = len(pp.times)
    rt = np.array(ref["time"][:n])
    assert n == 101 and np.allclose(pp.times, rt, atol=1e-9)
    dx, dy = np.array(pp.values["deflection_x"]), np.array(pp.values["deflection_y"])
    rx, ry = np.array(ref["deflection_x_structure_1"][:n]), np.array(ref["deflection_y_structure_1"][:n])
    amp = np.abs(ry).max()
    ex, ey = np.abs(dx - rx).max() / amp, np.abs(dy - ry).max() / amp
    print(f"oscillating beam: max |deflection - reference| / amplitude: x {ex:.2e}, y {ey:.2e}; "
          f"amplitude {amp:.4f} m; mse x {np.mean((dx - rx) ** 2):.2e}, y {np.mean((dy - ry) ** 2):.2e}")
    assert amp > 0.1                                     # a third of the beam's length
    assert ex <= 1e-5 and ey <= 1e-5, (ex, ey)         # measured: 1.1e-7 / 2.6e-7
    assert np.mean((dx - rx) ** 2) <= 1e-14 and np.mean((dy - ry) ** 2) <= 1e-14   # the reference's own check: MSE ~ 0
    semi.close()


@pytest.mark.parametrize("eltype,tol", [(np.float64, 1e-11), (np.float32, 3e-5)])
def test_falling_sphere_2d_kick_matches_oracle(eltype, tol):
    """examples/fsi/falling_sphere_2d.jl: an elastic TLSPH sphere without clamped particles whose dummy particles use
    `BernoulliPressureExtrapolation` -- on a structure the dynamic pressure term factor rho_f (v_rel . n)^2 / 2 always
    applies (dummy_particles.jl:696-707), on the static tank it never does.  The sphere half way into the water,
    moving down at 2 m/s."""
    fluid, wall, sphere, _ = examples.falling_sphere_2d(0.04, eltype=eltype, coordinates_eltype=eltype)
    assert sphere.n_integrated_particles == sphere.nparticles
    u_f, v_f = examples.perturbed_state(fluid, seed=3, position_jitter=0.05)
    rng = np.random.default_rng(4)
    x_s = sphere.initial_coordinates.astype(np.float64) + [0.0, -0.7]          # centre at y = 0.9: the free surface
    x_s = x_s * [1.02, 0.98] + rng.uniform(-2e-4, 2e-4, x_s.shape)             # squeezed a little
    v_s = np.tile([0.1, -2.0], (sphere.nparticles, 1)) + rng.uniform(-0.05, 0.05, x_s.shape)
    # the water the sphere has displaced is gone
    keep = np.linalg.norm(u_f.astype(np.float64) - [0.5 * 1.02, 0.9 * 0.98], axis=1) > 0.3 + 0.03
    u_f[~keep] += [0.0, 0.45]                                                  # parked above the surface, out of reach
    u = np.concatenate([u_f.reshape(-1), x_s.astype(eltype).reshape(-1)])
    v = np.concatenate([v_f.reshape(-1), v_s.astype(eltype).reshape(-1)])
    ref = adapter.kick_fsi(fluid, wall, sphere, u, v)
    # the Bernoulli term matters: with the factor switched off the sphere's pressure differs
    sphere.boundary_model.density_calculator = tp.AdamiPressureExtrapolation()
    ref_adami = adapter.kick_fsi(fluid, wall, sphere, u, v)
    sphere.boundary_model.density_calculator = tp.BernoulliPressureExtrapolation()
    assert np.abs(ref["structure_pressure"] - ref_adami["structure_pressure"]).max() > 100.0
    semi = tp.Semidiscretization(fluid, wall, sphere, parallelization_backend=tp.B200Backend())
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    dv = np.full_like(v, np.nan)
    ode.f1(dv, v, u, ode.p, 0.0)
    n_f = fluid.nparticles
    assert np.isfinite(dv).all()
    for name, a, b in (("fluid", dv[: 3 * n_f], ref["dv"][: 3 * n_f]), ("sphere", dv[3 * n_f:], ref["dv"][3 * n_f:])):
        assert np.abs(a - b).max() <= tol * np.abs(b).max(), (name, np.abs(a - b).max() / np.abs(b).max())
    p_err = np.abs(semi.system_field(sphere, "pressure") - ref["structure_pressure"]).max()
    assert p_err <= 10 * tol * np.abs(ref["structure_pressure"]).max()
    semi.close()


def test_falling_sphere_2d_time_loop():
    """The example for 0.5 s (RDPK3SpFSAL35 with the example's tolerances): free fall until the sphere meets the water
    at t = 0.29, then it is slowed down; the lighter-than-water sphere (density 500) does not sink to the bottom."""
    from trixiparticles.jl_b200.time_integration import RDPK3SpFSAL35, solve
    fluid, wall, sphere, _ = examples.falling_sphere_2d(0.04, eltype=np.float64)
    semi = tp.Semidiscretization(fluid, wall, sphere, parallelization_backend=tp.B200Backend(ode_memory="device"))
    ode = tp.semidiscretize(semi, (0.0, 0.5))
    sol = solve(ode, RDPK3SpFSAL35(), abstol=1e-6, reltol=1e-3)
    assert sol.retcode == "Success"
    u, v = sol.u.cpu().numpy(), sol.v.cpu().numpy()
    assert np.isfinite(u).all() and np.isfinite(v).all()
    n_f, n_s = fluid.nparticles, sphere.nparticles
    y = u[2 * n_f:].reshape(n_s, 2)[:, 1].mean()
    vy = v[3 * n_f:].reshape(n_s, 2)[:, 1].mean()
    free_fall_v = -9.81 * 0.5
    assert 0.6 < y < 1.2                        # in the water, not at the bottom (centre starts at 1.6, surface at 0.9)
    assert vy > 0.5 * free_fall_v               # decelerated by the water
    semi.close()


def test_falling_water_column_fsi_2d():
    """examples/fsi/falling_water_column_2d.jl: water dropped onto the clamped beam of oscillating_beam_2d.jl -- fluid +
    structure without any wall, cubic-spline fluid kernel, Monaghan-Kajtar coupling with three structure particles per
    fluid particle spacing.  Kick against the oracle with the water lowered onto the beam, then the example for
    0.25 s: the beam is bent down by the water."""
    from trixiparticles.jl_b200.time_integration import RDPK3SpFSAL35, solve
    fluid, beam, info = examples.falling_water_column_fsi_2d()
    box = tp.GridNeighborhoodSearch(2, cell_list=tp.FullGridCellList((-0.5, -3.0), (1.5, 1.5)))
    u, v = fsi_state(fluid, beam, seed=21, deform=0.02)
    n_f, n_int = fluid.nparticles, beam.n_integrated_particles
    u[: 2 * n_f].reshape(n_f, 2)[:, 1] -= 0.13            # the water touches the beam (top at y = 0.05)
    ref = adapter.kick_fsi(fluid, None, beam, u, v)
    semi = tp.Semidiscretization(fluid, beam, neighborhood_search=box, parallelization_backend=tp.B200Backend())
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    dv = np.full_like(v, np.nan)
    ode.f1(dv, v, u, ode.p, 0.0)
    free = adapter.kick(fluid, None, u[: 2 * n_f].reshape(n_f, 2), v[: 3 * n_f].reshape(n_f, 3))["dv"]
    assert np.abs(ref["dv"][: 3 * n_f].reshape(n_f, 3) - free).max() > 1.0      # the beam pushes back
    for name, a, b in (("fluid", dv[: 3 * n_f], ref["dv"][: 3 * n_f]), ("beam", dv[3 * n_f:], ref["dv"][3 * n_f:])):
        assert np.abs(a - b).max() <= 1e-11 * np.abs(b).max(), (name, np.abs(a - b).max() / np.abs(b).max())
    semi.close()
    semi = tp.Semidiscretization(fluid, beam, neighborhood_search=box,
                                 parallelization_backend=tp.B200Backend(ode_memory="device"))
    ode = tp.semidiscretize(semi, (0.0, 0.25))
    sol = solve(ode, RDPK3SpFSAL35(), abstol=1e-6, reltol=1e-4, dtmax=1e-3)
    assert sol.retcode == "Success"
    uu = sol.u.cpu().numpy()
    assert np.isfinite(uu).all()
    tip_y = uu[2 * n_f:].reshape(n_int, 2)[info["mid_particle"], 1] - info["start_position"][1]
    assert tip_y < -5e-3                                    # bent down (gravity alone: about -2e-3 by then)
    semi.close()


def test_falling_spheres_2d_two_structure_systems():
    """examples/fsi/falling_spheres_2d.jl: two `TotalLagrangianSPHSystem`s with different radius, density and Young's
    modulus.  Inside the library they share the structure slot (per-particle material constants,
    tpb_set_structure_material).  Independent check by superposition against three oracle runs with ONE sphere each:
    fluid dv = (with sphere 1) + (with sphere 2) - (with neither); each sphere's dv, stress and pressure as alone."""
    fluid, wall, s1, s2, _ = examples.falling_spheres_2d(0.04)
    n_f, n1, n2 = fluid.nparticles, s1.nparticles, s2.nparticles
    u_f, v_f = examples.perturbed_state(fluid, seed=5, position_jitter=0.05)
    rng = np.random.default_rng(6)

    def sphere_state(sph, drop, squeeze):
        x = sph.initial_coordinates.astype(np.float64) + [0.0, -drop]
        c = x.mean(axis=0)
        x = c + (x - c) * squeeze + rng.uniform(-2e-4, 2e-4, x.shape)
        vel = np.tile([0.05, -1.5], (sph.nparticles, 1)) + rng.uniform(-0.05, 0.05, x.shape)
        return x, vel
    x1, w1 = sphere_state(s1, 0.75, [1.03, 0.97])
    x2, w2 = sphere_state(s2, 0.65, [0.98, 1.02])
    for c, r in ((x1.mean(axis=0), 0.3), (x2.mean(axis=0), 0.2)):       # the displaced water is parked out of reach
        inside = np.linalg.norm(u_f - c, axis=1) < r + 0.04
        u_f[inside] += [0.0, 0.5]
    u = np.concatenate([u_f.reshape(-1), x1.reshape(-1), x2.reshape(-1)])
    v = np.concatenate([v_f.reshape(-1), w1.reshape(-1), w2.reshape(-1)])
    semi = tp.Semidiscretization(fluid, wall, s1, s2, parallelization_backend=tp.B200Backend())
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    assert semi.lib_index(s1) == semi.lib_index(s2) == 2 and semi.ranges_u[-1][1] == u.size
    dv = np.full_like(v, np.nan)
    ode.f1(dv, v, u, ode.p, 0.0)
    assert np.isfinite(dv).all()
    cat = lambda *a: np.concatenate([np.asarray(q).reshape(-1) for q in a])
    r1 = adapter.kick_fsi(fluid, wall, s1, cat(u_f, x1), cat(v_f, w1))
    r2 = adapter.kick_fsi(fluid, wall, s2, cat(u_f, x2), cat(v_f, w2))
    r0 = adapter.kick(fluid, wall, u_f, v_f)["dv"].reshape(-1)
    want_f = r1["dv"][: 3 * n_f] + r2["dv"][: 3 * n_f] - r0
    assert np.abs(r1["dv"][: 3 * n_f] - r0).max() > 1.0 and np.abs(r2["dv"][: 3 * n_f] - r0).max() > 1.0
    tol = 1e-11
    got_f, got_1, got_2 = dv[: 3 * n_f], dv[3 * n_f: 3 * n_f + 2 * n1], dv[3 * n_f + 2 * n1:]
    assert np.abs(got_f - want_f).max() <= tol * np.abs(want_f).max()
    assert np.abs(got_1 - r1["dv"][3 * n_f:]).max() <= tol * np.abs(r1["dv"][3 * n_f:]).max()
    assert np.abs(got_2 - r2["dv"][3 * n_f:]).max() <= tol * np.abs(r2["dv"][3 * n_f:]).max()
    # the two materials differ: swapping the spheres' moduli would change sphere 2's acceleration
    for sph, r in ((s1, r1), (s2, r2)):
        for name, key in (("pressure", "structure_pressure"), ("deformation_grad", "F"), ("pk1_rho2", "pk1_rho2")):
            got = semi.system_field(sph, name)
            assert np.abs(got - r[key]).max() <= 10 * tol * max(np.abs(r[key]).max(), 1e-300), (name, sph.nparticles)
    assert np.abs(r2["pk1_rho2"]).max() > 0
    semi.close()


def test_young_modulus_per_particle():
    """`young_modulus` / `poisson_ratio` as vectors (system.jl:108-161): constant vectors give the scalar result bit
    for bit; with the upper part of the plate ten times softer the stress there is a tenth, elsewhere unchanged."""
    fluid, wall, plate, _ = examples.dam_break_plate_2d(0.01, initial_fluid_size=(0.15, 0.29), plate_position=(0.165, 0.0))
    u, v = fsi_state(fluid, plate)
    n = plate.nparticles
    out = {}
    for name, E, nu in (("scalar", 1e6, 0.0), ("vector", np.full(n, 1e6), np.zeros(n)),
                        ("soft top", np.where(plate.initial_coordinates[:, 1] > 0.04, 1e5, 1e6), np.zeros(n))):
        st = tp.TotalLagrangianSPHSystem(
            examples.dam_break_plate_2d(0.01, initial_fluid_size=(0.15, 0.29), plate_position=(0.165, 0.0))[2].initial_condition,
            smoothing_kernel=plate.smoothing_kernel, smoothing_length=plate.smoothing_length, young_modulus=E,
            poisson_ratio=nu, boundary_model=plate.boundary_model, acceleration=plate.acceleration,
            penalty_force=plate.penalty_force, clamped_particles=range(n - plate.n_clamped_particles, n))
        semi = tp.Semidiscretization(fluid, wall, st, parallelization_backend=tp.B200Backend())
        ode = tp.semidiscretize(semi, (0.0, 1.0))
        dv = np.full_like(v, np.nan)
        ode.f1(dv, v, u, ode.p, 0.0)
        out[name] = (dv.copy(), semi.system_field(st, "pk1_rho2"))
        semi.close()
    assert np.array_equal(out["scalar"][0], out["vector"][0]) and np.array_equal(out["scalar"][1], out["vector"][1])
    y = plate.initial_coordinates[:, 1]
    soft, stiff = y > 0.04, y <= 0.04
    assert soft.sum() > 10 and np.abs(out["scalar"][1][soft]).max() > 0
    # nu = 0: S = E * strain, so the stress scales with E where E was lowered ...
    assert np.allclose(out["soft top"][1][soft], 0.1 * out["scalar"][1][soft], rtol=1e-13, atol=0)
    assert np.array_equal(out["soft top"][1][stiff], out["scalar"][1][stiff])      # ... and only there
