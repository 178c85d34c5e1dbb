"""Step times of the widened rows: BASELINE config 5 (dam_break_plate_2d, WCSPH + tank + TLSPH
plate) at the example's resolution and refined, config 2 (hydrostatic water column 2D) and the moving-wall example --
one kick!+drift! on device-resident vectors (CUDA events, L2 not flushed: the sets are far smaller than L2 anyway),
launches per kick, the time loop's steps per second with CUDA-graph replay, and the CPU oracle beside it.
Usage (GPU box):  python tools/bench_fsi.py > gpurun_out/fsi_bench.txt"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trixiparticles.jl_b200 as tp  # noqa: E402
from trixiparticles.jl_b200 import examples  # noqa: E402
from trixiparticles.jl_b200.time_integration import CarpenterKennedy2N54, solve  # noqa: E402
from oracle import adapter, oracle as O  # noqa: E402


def time_rhs(semi, ode, v, u, reps=200, t=0.0):
    dv, du = torch.zeros_like(v), torch.zeros_like(u)
    for _ in range(5):
        ode.f1(dv, v, u, ode.p, t); ode.f2(du, v, u, ode.p, t)
    semi.synchronize()
    st = semi.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ode.f1(dv, v, u, ode.p, t); ode.f2(du, v, u, ode.p, t)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, st.launches_last_kick + st.launches_last_drift


def case(name, systems, oracle_kick, steps_dt=None, nhs=None):
    semi = tp.Semidiscretization(*systems, neighborhood_search=nhs, parallelization_backend=tp.B200Backend(ode_memory="device"))
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    ms, launches = time_rhs(semi, ode, ode.v0, ode.u0)
    out = {"case": name, "particles": {type(s).__name__: s.nparticles for s in systems}, "rhs_ms": round(ms, 4),
           "launches_per_rhs": launches}
    if steps_dt is not None:
        n, dt = steps_dt
        semi2 = tp.Semidiscretization(*systems, neighborhood_search=nhs, parallelization_backend=tp.B200Backend(ode_memory="device"))
        ode2 = tp.semidiscretize(semi2, (0.0, n * dt))
        solve(ode2, CarpenterKennedy2N54(), dt=dt, cuda_graph=True, maxiters=20)   # warm-up (graph capture)
        semi2.synchronize()
        t0 = time.perf_counter()
        sol = solve(ode2, CarpenterKennedy2N54(), dt=dt, cuda_graph=True)
        semi2.synchronize()
        wall_s = time.perf_counter() - t0
        out["time_loop"] = {"steps": sol.nsteps, "rhs_evaluations": sol.nf, "ms_per_rhs": round(1e3 * wall_s / sol.nf, 4),
                            "steps_per_s": round(sol.nsteps / wall_s, 1), "cuda_graph": semi2._motion_system is None}
        semi2.close()
    u, v = ode.u0.cpu().numpy(), ode.v0.cpu().numpy()
    oracle_kick(u, v)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        oracle_kick(u, v)
    out["cpu_oracle_kick_ms"] = round(1e3 * (time.perf_counter() - t0) / reps, 3)
    out["cpu_threads"] = O.max_threads()
    semi.close()
    print(json.dumps(out), flush=True)


def main():
    O.build()
    for dx, npx in ((0.01, 5), (0.005, 9), (0.0025, 17)):
        fluid, wall, plate, _ = examples.dam_break_plate_2d(dx, n_particles_x=npx, eltype=np.float32, coordinates_eltype=np.float32,
                                                            initial_fluid_size=(0.15, 0.29))
        case(f"config 5: dam_break_plate_2d dx={dx} (Float32, Monaghan-Kajtar coupling)", (fluid, wall, plate),
             lambda u, v, f=fluid, w=wall, p=plate: adapter.kick_fsi(f, w, p, u, v), steps_dt=(400, 2e-5 * dx / 0.01))
    fluid, wall, plate, _ = examples.dam_break_plate_2d(0.005, n_particles_x=9, eltype=np.float32, coordinates_eltype=np.float32,
                                                        initial_fluid_size=(0.15, 0.29), structure_boundary_model="dummy_particles")
    case("config 5 with BoundaryModelDummyParticles on the plate, dx=0.005", (fluid, wall, plate),
         lambda u, v, f=fluid, w=wall, p=plate: adapter.kick_fsi(f, w, p, u, v), steps_dt=(400, 1e-5))
    for dx in (0.02, 0.01):
        fluid, wall, plate, _ = examples.dam_break_plate_3d(dx, n_particles_x=3 if dx == 0.02 else 5, eltype=np.float32,
                                                            coordinates_eltype=np.float32, plate_position=(0.175, 0.0, 0.0075))
        case(f"config 5 in 3-D: dam_break_plate_3d dx={dx} (Float32, Monaghan-Kajtar coupling)", (fluid, wall, plate),
             lambda u, v, f=fluid, w=wall, p=plate: adapter.kick_fsi(f, w, p, u, v), steps_dt=(400, 2e-5 * dx / 0.02))
    for dx in (0.05, 0.005):
        fluid, wall, _ = examples.hydrostatic_water_column_2d(dx)
        case(f"config 2: hydrostatic_water_column_2d dx={dx} (Float32)", (fluid, wall),
             lambda u, v, f=fluid, w=wall: adapter.kick(f, w, u.reshape(-1, 2), v.reshape(-1, 3)), steps_dt=(400, 2e-4 * dx / 0.05))
    fluid, wall, _ = examples.moving_wall_2d(0.005, eltype=np.float32, coordinates_eltype=np.float32)
    wall.apply_prescribed_motion(0.0)
    case("moving_wall_2d dx=0.005 (Float32; the moving wall on the structure path)", (fluid, wall),
         lambda u, v, f=fluid, w=wall: adapter.kick_moving_wall(f, w, u.reshape(-1, 2), v.reshape(-1, 3)), steps_dt=(200, 2e-5))


if __name__ == "__main__":
    main()
