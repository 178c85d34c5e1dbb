#!/usr/bin/env python
"""Extracts the mid-plate deflection trace of the reference's own FSI validation run (hydrostatic water column on
an elastic plate, WCSPH, n_particles_plate_y = 3) into a small fixture.  Source (read-only reference tree):
/root/reference/validation/hydrostatic_water_column_2d/validation_reference_wcsph_3.json -- the run behind
test/validation/validation.jl:95-109 (relative error of the deflection averaged over t >= 0.25 against the
analytical value: <= 0.045).  Run in the build container (the reference tree does not exist on the GPU box):

    python tests/golden/make_fsi_hydrostatic_fixture.py
"""
import json
import os

SRC = "/root/reference/validation/hydrostatic_water_column_2d/validation_reference_wcsph_3.json"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fsi_hydrostatic_wcsph_3_trace.json")

d = json.load(open(SRC))
out = {
    "source": SRC,
    "n_particles_plate_y": 3,
    "analytical_value": d["analytical_solution"]["values"][0],
    "time": d["y_deflection_structure_1"]["time"],
    "y_deflection_structure_1": d["y_deflection_structure_1"]["values"],
    "kinetic_energy_structure_1": d["kinetic_energy_structure_1"]["values"],
    "kinetic_energy_fluid_1": d["kinetic_energy_fluid_1"]["values"],
}
json.dump(out, open(DST, "w"))
# second resolution (5 particles across the plate)
d5 = json.load(open(SRC.replace("_3.json", "_5.json")))
json.dump({"source": SRC.replace("_3.json", "_5.json"), "n_particles_plate_y": 5,
           "analytical_value": d5["analytical_solution"]["values"][0], "time": d5["y_deflection_structure_1"]["time"],
           "y_deflection_structure_1": d5["y_deflection_structure_1"]["values"]},
          open(DST.replace("_3_trace", "_5_trace"), "w"))
t, y = out["time"], out["y_deflection_structure_1"]
late = [v for tt, v in zip(t, y) if tt >= 0.25]
print(f"{len(t)} samples to t = {t[-1]}; mean deflection over t >= 0.25: {sum(late) / len(late):.6e}; "
      f"analytical {out['analytical_value']:.6e}")
