#!/usr/bin/env python
"""FSI validation run of the reference (validation/hydrostatic_water_column_2d/validation.jl with
n_particles_plate_y = 3, the case of test/validation/validation.jl:95-109): a 2 m water column on an elastic
aluminium plate clamped at both ends, WCSPH + TLSPH with dummy-particle coupling, entirely on the GPU.

The reference integrates the fluid with RDPK3SpFSAL49 and sub-cycles the plate with CarpenterKennedy2N54
(SplitIntegrationCallback); here ONE CarpenterKennedy2N54 loop advances both systems with the plate's CFL step
(about 1e-5 s, 35 000 steps of 5 stages to t = 0.35, replayed from a CUDA graph) -- the same semi-discretisation,
a finer time discretisation of the fluid.  Checked: the mid-plate deflection averaged over t >= 0.25 against the
analytical value (reference bar: relative error <= 0.045) and against the reference's own trace
(tests/golden/fsi_hydrostatic_wcsph_3_trace.json).

usage: python tools/run_fsi_hydrostatic_validation.py [t_end]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import trixiparticles.jl_b200 as tp
from trixiparticles.jl_b200 import examples
from trixiparticles.jl_b200.time_integration import CarpenterKennedy2N54, PostprocessCallback, StepsizeCallback, solve


def run(t_end=0.35, n_particles_plate_y=3, cfl=1.6, cuda_graph=True, eltype=np.float64):
    structure, fluid, wall, info = examples.hydrostatic_water_column_fsi_2d(n_particles_plate_y, eltype=eltype)
    semi = tp.Semidiscretization(structure, fluid, wall,
                                 parallelization_backend=tp.B200Backend(ode_memory="device"))
    ode = tp.semidiscretize(semi, (0.0, t_end))
    a, _ = semi.ranges_u[semi.system_index(structure)]
    mid, half = info["mid_particle"], info["plate_size"][1] / 2

    def y_deflection(system, v_ode, u_ode, semi_, t):      # validation.jl:15-21
        return float(u_ode[a + 2 * mid + 1]) + half

    pp = PostprocessCallback(dt=0.0025, y_deflection=y_deflection)
    dt = StepsizeCallback(cfl=cfl).dt(semi)
    t0 = time.time()
    sol = solve(ode, CarpenterKennedy2N54(williamson_condition=False), dt=dt, callback=[pp], cuda_graph=cuda_graph,
                maxiters=10 ** 7)
    wall_s = time.time() - t0
    times, y = np.array(pp.times), np.array(pp.values["y_deflection"])
    late = times >= 0.25 - 1e-12
    avg = float(y[late].mean()) if late.any() else float("nan")
    u = sol.u.cpu().numpy()
    semi.close()
    return dict(times=times, y=y, avg=avg, analytical=info["analytical_value"], dt=dt, nsteps=sol.nsteps,
                wall_s=wall_s, finite=bool(np.isfinite(u).all()),
                rel_error=abs(avg - info["analytical_value"]) / abs(info["analytical_value"]))


if __name__ == "__main__":
    t_end = float(sys.argv[1]) if len(sys.argv) > 1 else 0.35
    r = run(t_end)
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "fsi_hydrostatic_wcsph_3_trace.json")))
    rt, ry = np.array(ref["time"]), np.array(ref["y_deflection_structure_1"])
    ref_avg = ry[rt >= 0.25 - 1e-12].mean()
    print(f"dt = {r['dt']:.3e} s, {r['nsteps']} steps in {r['wall_s']:.1f} s, state finite: {r['finite']}")
    print(f"mid-plate deflection averaged over t in [0.25, {t_end}]: {r['avg']:.4e} m "
          f"(analytical {r['analytical']:.4e}, relative error {r['rel_error']:.4f}; reference run, t in [0.25, 0.3]: "
          f"{ref_avg:.4e}, relative error {abs(ref_avg - r['analytical']) / abs(r['analytical']):.4f})")
    for tt in (0.05, 0.1, 0.15, 0.2, 0.25, 0.3):
        if tt <= t_end:
            i, j = int(np.argmin(np.abs(r["times"] - tt))), int(np.argmin(np.abs(rt - tt)))
            print(f"  t = {tt:4.2f}: deflection {r['y'][i]: .4e} m   (reference trace {ry[j]: .4e})")
