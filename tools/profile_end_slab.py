#!/usr/bin/env python
"""The last slab of the 8-GPU config-4 run (12.5 M fluid particles, 22 M wall particles: the rank whose box
runs on to the far end of the tank) as a single handle on one GPU, without ghosts: where does its rebuild
phase (1.07 ms on 8 GPUs against 0.55 ms for a whole 12.5 M dam break) go?
usage: python tools/profile_end_slab.py [world] [dx] [kicks]   (run under ncu for the launch list)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trixiparticles.jl_b200 as tp
from trixiparticles.jl_b200 import _lib
from trixiparticles.jl_b200.slabs import dam_break_3d_slab

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dx = float(sys.argv[2]) if len(sys.argv) > 2 else 0.00271442
kicks = int(sys.argv[3]) if len(sys.argv) > 3 else 10
t0 = time.time()
fluid, wall, local = dam_break_3d_slab(dx, world - 1, world)
lay = local["layout"]
lo = lay.planes[world - 1]
gmin, gmax = (np.array(b, dtype=np.float64) for b in local["bounding_box"])
mn = gmin.copy()
mn[0] = max(gmin[0], lo - lay.halo - lay.skin - 3 * dx * 3)
print(f"rank {world - 1} of {world}: {fluid.nparticles} fluid + {wall.nparticles} wall particles, box {mn} .. {gmax}, "
      f"set-up {time.time() - t0:.1f} s", flush=True)
nhs = tp.GridNeighborhoodSearch(3, cell_list=tp.FullGridCellList(min_corner=mn, max_corner=gmax))
semi = tp.Semidiscretization(fluid, wall, neighborhood_search=nhs,
                             parallelization_backend=tp.B200Backend(ode_memory="device"))
ode = tp.semidiscretize(semi, (0.0, 1.0))
dv = torch.zeros_like(ode.v0)
semi.set_profiling(kicks)
for _ in range(3):
    ode.f1(dv, ode.v0, ode.u0, ode.p, 0.0)
semi.synchronize()
semi.set_profiling(kicks)
for _ in range(kicks):
    ode.f1(dv, ode.v0, ode.u0, ode.p, 0.0)
semi.synchronize()
print("cells", semi.stats().n_cells, "phases_ms", semi.phase_times())
semi.close()
