#!/usr/bin/env python
"""Runs the reference's 2-D dam-break validation case (validation/dam_break_2d, resolution 40,
WCSPH, CarpenterKennedy2N54, dt from StepsizeCallback(cfl=0.9), samples every 0.01/sqrt(g/H))
on the B200 library and compares the surge front with the reference trace fixture."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trixiparticles.jl_b200 as tp
from trixiparticles.jl_b200 import examples
from trixiparticles.jl_b200.time_integration import (CarpenterKennedy2N54, PostprocessCallback,
                                                     StepsizeCallback, max_x_coord, solve)


def run(t_end=None, eltype=np.float64, memory="device"):
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
    post = PostprocessCallback(dt=0.01 / np.sqrt(g / H), max_x_coord=max_x_coord)
    t0 = time.perf_counter()
    sol = solve(ode, CarpenterKennedy2N54(williamson_condition=False), callback=[step_cb, post])
    wall_s = time.perf_counter() - t0
    n = len(post.times)
    ref_t = np.array(fx["time"][:n])
    ref_x = np.array(fx["max_x_coord_fluid_1"][:n])
    got_x = np.array(post.values["max_x_coord"])
    assert np.allclose(ref_t, post.times, rtol=0, atol=1e-9)
    semi.close()
    return dict(sol=sol, times=np.array(post.times), got=got_x, ref=ref_x, dt=step_cb.dt(semi), wall_s=wall_s,
                dt_max_ref=fx["time_integrator"]["dt_max"])


if __name__ == "__main__":
    t_end = float(sys.argv[1]) if len(sys.argv) > 1 else None
    r = run(t_end)
    err = np.abs(r["got"] - r["ref"])
    print(f"dt = {r['dt']:.16g} (reference dt_max {r['dt_max_ref']:.16g}); {r['sol'].nsteps} steps, "
          f"{r['sol'].nf} RHS evaluations in {r['wall_s']:.1f} s")
    for k in range(0, len(err), max(1, len(err) // 28)):
        print(f"t = {r['times'][k]:.4f}  front = {r['got'][k]:.6f}  reference = {r['ref'][k]:.6f}  |diff| = {err[k]:.2e}")
    for w in (0.25, 0.5, 0.75, 1.0, 1.25, 1.5, 1.75):
        m = r["times"] <= w
        if m.any():
            print(f"max |diff| for t <= {w}: {err[m].max():.3e}")
