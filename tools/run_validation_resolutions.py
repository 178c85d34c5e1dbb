#!/usr/bin/env python
"""Second resolutions of the two structure validation cases, against the reference's own traces
(tests/golden/*_trace.json):  the oscillating beam with 9 particles across (validation_reference_9.json) and the
hydrostatic water column on an elastic plate with 5 particles across the plate (validation_reference_wcsph_5.json).
One-off runs on the GPU box; output kept in profiles/.   usage: python tools/run_validation_resolutions.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import trixiparticles.jl_b200 as tp  # noqa: E402
from trixiparticles.jl_b200 import examples  # noqa: E402
from trixiparticles.jl_b200.model import PenaltyForceGanzenmueller  # noqa: E402
from trixiparticles.jl_b200.time_integration import CarpenterKennedy2N54, PostprocessCallback, solve  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def beam(n_y, dt):
    ref = json.load(open(os.path.join(GOLD, f"oscillating_beam_2d_{n_y}_trace.json")))
    system, info = examples.oscillating_beam_2d(n_y, penalty_force=PenaltyForceGanzenmueller(alpha=0.01))
    semi = tp.Semidiscretization(system, parallelization_backend=tp.B200Backend(ode_memory="device"))
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    mid, x0 = info["mid_particle"], info["start_position"]
    pp = PostprocessCallback(0.01, dx=lambda s_, v, u, s, t: float(u[2 * mid].item()) - x0[0],
                             dy=lambda s_, v, u, s, t: float(u[2 * mid + 1].item()) - x0[1])
    t0 = time.time()
    sol = solve(ode, CarpenterKennedy2N54(), dt=dt, callback=(pp,), cuda_graph=True)
    wall = time.time() - t0
    n = len(pp.times)
    rx, ry = np.array(ref["deflection_x_structure_1"][:n]), np.array(ref["deflection_y_structure_1"][:n])
    dx, dy = np.array(pp.values["dx"]), np.array(pp.values["dy"])
    amp = np.abs(ry).max()
    print(f"oscillating beam, {n_y} particles across ({system.nparticles} particles, {sol.nsteps} steps, {wall:.1f} s): "
          f"max |deflection - reference| / amplitude  x {np.abs(dx - rx).max() / amp:.2e}  y {np.abs(dy - ry).max() / amp:.2e}; "
          f"MSE x {np.mean((dx - rx) ** 2):.2e}  y {np.mean((dy - ry) ** 2):.2e}; amplitude {amp:.4f} m", flush=True)
    semi.close()


def plate(n_y):
    import run_fsi_hydrostatic_validation as V
    ref = json.load(open(os.path.join(GOLD, f"fsi_hydrostatic_wcsph_{n_y}_trace.json")))
    rt, ry = np.array(ref["time"]), np.array(ref["y_deflection_structure_1"])
    r = V.run(t_end=0.3, n_particles_plate_y=n_y)
    n = min(len(r["times"]), len(rt))
    a = abs(ref["analytical_value"])
    print(f"hydrostatic water column on a plate, {n_y} particles across ({r['nsteps']} steps, {r['wall_s']:.1f} s): "
          f"max |deflection - reference trace| = {np.abs(r['y'][:n] - ry[:n]).max() / a:.2e} x analytical value; "
          f"average over t >= 0.25: {r['avg']:.4e} (reference trace {ry[rt >= 0.25 - 1e-12].mean():.4e}, analytical "
          f"{ref['analytical_value']:.4e}, relative error {r['rel_error']:.4f})", flush=True)


if __name__ == "__main__":
    beam(5, 2e-5)
    beam(9, 1e-5)
    plate(3)
    plate(5)
