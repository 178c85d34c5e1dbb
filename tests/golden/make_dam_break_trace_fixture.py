#!/usr/bin/env python
"""Extracts the surge-front and pressure-sensor traces (and the run parameters that pin them) of the reference's
2-D dam-break validation run into a small fixture.  Source (read-only reference tree):
/root/reference/validation/dam_break_2d/validation_reference_wcsph_40.json, the file the
reference's own test/validation/validation.jl:48-68 compares against.  Run in the build
container (the reference tree does not exist on the GPU box):

    python tests/golden/make_dam_break_trace_fixture.py
"""
import json
import os

SRC = "/root/reference/validation/dam_break_2d/validation_reference_wcsph_40.json"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dam_break_2d_wcsph_40_trace.json")

d = json.load(open(SRC))
fl = d["meta"]["system_data"]["fluid_1"]
out = {
    "source": SRC,
    "time_integrator": d["meta"]["simulation_info"]["time_integrator"],
    "fluid": {k: fl[k] for k in ("viscosity_model", "density_diffusion", "smoothing_length",
                                 "particle_spacing", "acceleration", "state_equation",
                                 "density_calculator", "smoothing_kernel") if k in fl},
    "sound_speed": fl.get("sound_speed", fl.get("state_equation", {}).get("sound_speed")),
    "time": d["max_x_coord_fluid_1"]["time"],
    "max_x_coord_fluid_1": d["max_x_coord_fluid_1"]["values"],
    "pressure_P1_fluid_1": d["pressure_P1_fluid_1"]["values"],
    "pressure_P2_fluid_1": d["pressure_P2_fluid_1"]["values"],
    "pressure_P3_fluid_1": d["pressure_P3_fluid_1"]["values"],
    "pressure_P4_fluid_1": d["pressure_P4_fluid_1"]["values"],
}
json.dump(out, open(DST, "w"))
print(DST, os.path.getsize(DST), "bytes; sound_speed =", out["sound_speed"])
