#!/usr/bin/env python
"""Runs the reference's 2-D dam-break validation case (validation/dam_break_2d, resolution 40,
WCSPH, CarpenterKennedy2N54, dt from StepsizeCallback(cfl=0.9), samples every 0.01/sqrt(g/H))
on the B200 library and compares the surge front and the four pressure sensors P1..P4
(validation/dam_break_2d/sensors.jl, setup_marrone_2011.jl:46-74) with the reference trace fixture."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trixiparticles.jl_b200 as tp
from trixiparticles.jl_b200 import examples
from trixiparticles.jl_b200.interpolation import interpolated_pressure
from trixiparticles.jl_b200.time_integration import (CarpenterKennedy2N54, PostprocessCallback,
                                                     StepsizeCallback, max_x_coord, solve)


def pressure_sensors(H, tank_right_wall_x):
    """Sensor positions used by De Courcy et al. (2024), setup_marrone_2011.jl:46-74."""
    sensor_size = 0.0084
    out = {}
    for name, offset in (("pressure_P1", 6), ("pressure_P2", 30), ("pressure_P3", 60), ("pressure_P4", 160)):
        y_top = (offset + 4.2) / 600 * H
        out[name] = interpolated_pressure([tank_right_wall_x, y_top], [tank_right_wall_x, y_top - sensor_size])
    return out


def run(t_end=None, eltype=np.float64, memory="device", sensors=False, cuda_graph=False):
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "dam_break_2d_wcsph_40_trace.json")))
    H, g = 0.6, 9.81
    fluid, wall, tank = examples.dam_break_2d(40, alpha=fx["fluid"]["viscosity_model"]["alpha"],
                                              delta=fx["fluid"]["density_diffusion"]["delta"],
                                              sound_speed_factor=20.0, eltype=eltype, coordinates_eltype=eltype)
    assert abs(float(fluid.state_equation.sound_speed) - fx["sound_speed"]) < 1e-4 * fx["sound_speed"]
    tspan = (0.0, t_end if t_end is not None else fx["time_integrator"]["final_time"])
    semi = tp.Semidiscretization(fluid, wall, parallelization_backend=tp.B200Backend(ode_memory=memory))
    ode = tp.semidiscretize(semi, tspan)
    step_cb = StepsizeCallback(cfl=0.9)
    funcs = dict(max_x_coord=max_x_coord)
    if sensors:
        funcs.update(pressure_sensors(H, float(np.floor(5.366 * H / 0.015) * 0.015)))
    post = PostprocessCallback(dt=0.01 / np.sqrt(g / H), **funcs)
    t0 = time.perf_counter()
    sol = solve(ode, CarpenterKennedy2N54(williamson_condition=False), callback=[step_cb, post],
                cuda_graph=cuda_graph)
    wall_s = time.perf_counter() - t0
    n = len(post.times)
    ref_t = np.array(fx["time"][:n])
    ref_x = np.array(fx["max_x_coord_fluid_1"][:n])
    got_x = np.array(post.values["max_x_coord"])
    assert np.allclose(ref_t, post.times, rtol=0, atol=1e-9)
    semi.close()
    out = dict(sol=sol, times=np.array(post.times), got=got_x, ref=ref_x, dt=step_cb.dt(semi), wall_s=wall_s,
               dt_max_ref=fx["time_integrator"]["dt_max"])
    if sensors:
        out["pressure"] = {k: (np.array(post.values[k]), np.array(fx[k + "_fluid_1"][:n]))
                           for k in ("pressure_P1", "pressure_P2", "pressure_P3", "pressure_P4")}
    return out


if __name__ == "__main__":
    t_end = float(sys.argv[1]) if len(sys.argv) > 1 else None
    r = run(t_end, sensors=True)
    err = np.abs(r["got"] - r["ref"])
    print(f"dt = {r['dt']:.16g} (reference dt_max {r['dt_max_ref']:.16g}); {r['sol'].nsteps} steps, "
          f"{r['sol'].nf} RHS evaluations in {r['wall_s']:.1f} s")
    for k in range(0, len(err), max(1, len(err) // 28)):
        print(f"t = {r['times'][k]:.4f}  front = {r['got'][k]:.6f}  reference = {r['ref'][k]:.6f}  |diff| = {err[k]:.2e}")
    for w in (0.25, 0.5, 0.75, 1.0, 1.25, 1.5, 1.75):
        m = r["times"] <= w
        if m.any():
            print(f"max |diff| for t <= {w}: {err[m].max():.3e}")
    # pressure sensors: pointwise deviation and the deviation of a moving average over 0.05 s
    # (WCSPH sensor pressures are dominated by acoustic noise after the impact at t ~ 0.6 s)
    scale = 1000.0 * 9.81 * 0.6
    for name, (got, ref) in r["pressure"].items():
        d = np.abs(got - ref)
        k = np.ones(21) / 21
        ds = np.abs(np.convolve(got, k, mode="same") - np.convolve(ref, k, mode="same"))
        line = f"{name}: first non-zero sample t = {r['times'][np.nonzero(got)[0][0]] if got.any() else None}"
        for w in (0.57, 0.65, 0.8, 1.0, 1.75):
            m = r["times"] <= w
            if m.any():
                line += f" | t<={w}: {d[m].max() / scale:.2e} (avg {ds[m].max() / scale:.2e})"
        print(line + "   [units of rho g H = 5886 Pa]")
