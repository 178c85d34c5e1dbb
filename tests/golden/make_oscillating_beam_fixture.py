#!/usr/bin/env python
"""Extracts the first two seconds of the reference's own oscillating-beam validation trace (TLSPH only, 5 particles
across the beam) into a small fixture.  Source (read-only reference tree):
/root/reference/validation/oscillating_beam_2d/validation_reference_5.json -- the run of
validation/oscillating_beam_2d/validation_oscillating_beam_2d.jl (RDPK3SpFSAL49, abstol 1e-8, reltol 1e-6,
PenaltyForceGanzenmueller(alpha = 0.01)), which test/validation/validation.jl:13-29 re-runs to t = 1 and compares
with an interpolated mean squared error of zero (atol = eps()).  Run in the build container (the reference tree does
not exist on the GPU box):

    python tests/golden/make_oscillating_beam_fixture.py
"""
import json
import os

SRC = "/root/reference/validation/oscillating_beam_2d/validation_reference_5.json"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oscillating_beam_2d_5_trace.json")

d = json.load(open(SRC))
t = d["deflection_x_structure_1"]["time"]
n = sum(1 for tt in t if tt <= 2.0 + 1e-12)
out = {
    "source": SRC,
    "n_particles_y": 5,
    "time": t[:n],
    "deflection_x_structure_1": d["deflection_x_structure_1"]["values"][:n],
    "deflection_y_structure_1": d["deflection_y_structure_1"]["values"][:n],
}
json.dump(out, open(DST, "w"))
# second resolution (9 particles across the beam), to t = 1
d9 = json.load(open(SRC.replace("_5.json", "_9.json")))
t9 = d9["deflection_x_structure_1"]["time"]
n9 = sum(1 for tt in t9 if tt <= 1.0 + 1e-12)
json.dump({"source": SRC.replace("_5.json", "_9.json"), "n_particles_y": 9, "time": t9[:n9],
           "deflection_x_structure_1": d9["deflection_x_structure_1"]["values"][:n9],
           "deflection_y_structure_1": d9["deflection_y_structure_1"]["values"][:n9]},
          open(DST.replace("_5_trace", "_9_trace"), "w"))
print(f"{n} samples to t = {out['time'][-1]}; min deflection_y = {min(out['deflection_y_structure_1']):.6e}, "
      f"min deflection_x = {min(out['deflection_x_structure_1']):.6e}")
