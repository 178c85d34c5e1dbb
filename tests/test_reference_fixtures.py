"""Consumes the fixtures `tools/dump_reference_kick.jl` writes with the reference's own Julia `kick!` /
`drift!` (semidiscretization.jl:522-612): one directory per example under
`tests/golden/reference_kick/<name>/` (or `baseline/_ref/reference_kick/<name>/`), raw little-endian
`u_ode.bin v_ode.bin dv_ode.bin du_ode.bin [fluid_pressure.bin wall_pressure.bin wall_density.bin]`
+ `meta.txt`.  Julia is not installed in the build image, so no such directory is committed: the GPU
test skips when none is found, and a Julia box closes the "reference's own kick! within 1e-12 / 1e-5"
claim with

    julia --project -t auto tools/dump_reference_kick.jl examples/fluid/dam_break_2d.jl \
        tests/golden/reference_kick/dam_break_2d
    python -m pytest tests/test_reference_fixtures.py -m gpu

The loader itself is covered without Julia: the CPU test writes a fixture in the same format from
the oracle and reads it back.
"""
import os

import numpy as np
import pytest

import trixiparticles.jl_b200 as tp
from trixiparticles.jl_b200 import examples

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE_ROOTS = [os.path.join(ROOT, "tests", "golden", "reference_kick"),
                 os.path.join(ROOT, "baseline", "_ref", "reference_kick")]

# example name (directory) -> builder of the same systems with the example's defaults
# (examples/fluid/dam_break_2d.jl, hydrostatic_water_column_2d.jl, dam_break_3d.jl)
BUILDERS = {
    "dam_break_2d": lambda et, ct: examples.dam_break_2d(eltype=et, coordinates_eltype=ct),
    "hydrostatic_water_column_2d": lambda et, ct: examples.hydrostatic_water_column_2d(eltype=et),
    "dam_break_3d": lambda et, ct: examples.dam_break_3d(eltype=et, coordinates_eltype=ct),
}
DTYPES = {"Float64": np.float64, "Float32": np.float32}
TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


def read_meta(path):
    meta, systems = {}, []
    with open(path) as f:
        for line in f:
            key, _, val = line.partition("=")
            key, val = key.strip(), val.strip()
            if key.startswith("system["):
                name, _, n = val.partition(" n = ")
                systems.append((name.strip(), int(n)))
            elif key:
                meta[key] = val
    meta["systems"] = systems
    return meta


def load_fixture(directory):
    """-> dict(u, v, dv, du [, fluid_pressure, wall_pressure, wall_density], eltype_u, eltype_v, systems)"""
    meta = read_meta(os.path.join(directory, "meta.txt"))
    tu, tv = DTYPES[meta["eltype_u"]], DTYPES[meta["eltype_v"]]
    out = {"eltype_u": np.dtype(tu), "eltype_v": np.dtype(tv), "systems": meta["systems"]}
    for name, dt, length in (("u", tu, "length_u"), ("du", tu, "length_u"), ("v", tv, "length_v"),
                             ("dv", tv, "length_v")):
        a = np.fromfile(os.path.join(directory, f"{name}_ode.bin"), dtype=np.dtype(dt).newbyteorder("<"))
        if a.size != int(meta[length]):
            raise ValueError(f"{directory}: {name}_ode.bin holds {a.size} values, meta.txt says {meta[length]}")
        out[name] = a.astype(dt)
    for name in ("fluid_pressure", "wall_pressure", "wall_density"):
        p = os.path.join(directory, name + ".bin")
        if os.path.exists(p):
            out[name] = np.fromfile(p, dtype=np.dtype(tv).newbyteorder("<")).astype(tv)
    return out


def write_fixture(directory, u, v, dv, du, systems, extra=None):
    """The format of tools/dump_reference_kick.jl (used by the loader's own test)."""
    os.makedirs(directory, exist_ok=True)
    names = {np.dtype(np.float64): "Float64", np.dtype(np.float32): "Float32"}
    for name, a in (("u_ode", u), ("v_ode", v), ("dv_ode", dv), ("du_ode", du), *((extra or {}).items())):
        np.ascontiguousarray(a).astype(a.dtype.newbyteorder("<")).tofile(os.path.join(directory, name + ".bin"))
    with open(os.path.join(directory, "meta.txt"), "w") as f:
        f.write(f"eltype_v = {names[v.dtype]}\neltype_u = {names[u.dtype]}\n")
        f.write(f"length_v = {v.size}\nlength_u = {u.size}\n")
        for i, (name, n) in enumerate(systems):
            f.write(f"system[{i + 1}] = {name} n = {n}\n")


def fixture_dirs():
    found = []
    for root in FIXTURE_ROOTS:
        if os.path.isdir(root):
            for name in sorted(os.listdir(root)):
                if os.path.exists(os.path.join(root, name, "meta.txt")):
                    found.append((name, os.path.join(root, name)))
    return found


def rel_inf(a, b):
    scale = np.abs(b).max()
    return np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / (scale if scale > 0 else 1.0)


def test_fixture_format_round_trip(tmp_path):
    """A fixture written in the Julia script's format from the oracle's kick reads back exactly, with
    the ODE layout of semidiscretization.jl:128-135 (fluid rows only; the wall has no ODE variables)."""
    from oracle import adapter
    fluid, wall, _ = examples.dam_break_2d(10, eltype=np.float64, coordinates_eltype=np.float64)
    u, v = examples.perturbed_state(fluid, seed=3)
    ref = adapter.kick(fluid, wall, u, v)
    ref["du"] = np.ascontiguousarray(v[:, :fluid.ndims]).astype(u.dtype)  # drift!: du = v (semidiscretization.jl:522-536)
    d = str(tmp_path / "dam_break_2d")
    write_fixture(d, u.reshape(-1), v.reshape(-1), ref["dv"].reshape(-1), ref["du"].reshape(-1),
                  [("WeaklyCompressibleSPHSystem", fluid.nparticles), ("WallBoundarySystem", wall.nparticles)],
                  extra={"fluid_pressure": ref["pressure"], "wall_pressure": ref["wall_pressure"]})
    fx = load_fixture(d)
    assert fx["systems"] == [("WeaklyCompressibleSPHSystem", fluid.nparticles),
                             ("WallBoundarySystem", wall.nparticles)]
    assert fx["eltype_u"] == np.float64 and fx["u"].size == fluid.nparticles * 2
    np.testing.assert_array_equal(fx["dv"], ref["dv"].reshape(-1))
    np.testing.assert_array_equal(fx["du"], ref["du"].reshape(-1))
    np.testing.assert_array_equal(fx["fluid_pressure"], ref["pressure"])
    # a truncated file is an error, not a silent partial comparison
    with open(os.path.join(d, "dv_ode.bin"), "r+b") as f:
        f.truncate(16)
    with pytest.raises(ValueError):
        load_fixture(d)


@pytest.mark.gpu
@pytest.mark.parametrize("name,directory", fixture_dirs() or [pytest.param(None, None, marks=pytest.mark.skip(
    reason="no reference fixtures (tools/dump_reference_kick.jl needs Julia; see the module docstring)"))])
def test_cuda_path_matches_reference_julia_kick(name, directory):
    if name not in BUILDERS:
        pytest.skip(f"no builder for the example '{name}'")
    fx = load_fixture(directory)
    fluid, wall, _ = BUILDERS[name](fx["eltype_v"].type, fx["eltype_u"].type)
    nd = fluid.ndims
    assert [n for _, n in fx["systems"]] == [fluid.nparticles, wall.nparticles], "particle counts differ"
    # the example's initial condition itself (setups: RectangularTank, state equation, hydrostatic pressure)
    u0 = np.asarray(fluid.coordinates).reshape(-1)
    assert rel_inf(u0, fx["u"]) <= 4 * np.finfo(fx["eltype_u"]).eps, "initial coordinates differ from the reference's"
    semi = tp.Semidiscretization(fluid, wall, parallelization_backend=tp.B200Backend())
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    dv, du = np.full_like(fx["v"], np.nan), np.full_like(fx["u"], np.nan)
    tp.kick_(dv, fx["v"], fx["u"], ode.p, 0.0)
    tp.drift_(du, fx["v"], fx["u"], ode.p, 0.0)
    tol = TOL[fx["eltype_v"]]
    nv = fx["v"].size // fluid.nparticles
    got, ref = dv.reshape(-1, nv), fx["dv"].reshape(-1, nv)
    assert rel_inf(got[:, :nd], ref[:, :nd]) <= tol
    if nv > nd:
        assert rel_inf(got[:, nd], ref[:, nd]) <= tol
    np.testing.assert_array_equal(du, fx["du"])
    if "fluid_pressure" in fx:
        assert rel_inf(semi.system_field(fluid, "pressure"), fx["fluid_pressure"]) <= tol
    if "wall_pressure" in fx:
        assert rel_inf(semi.system_field(wall, "pressure"), fx["wall_pressure"]) <= 10 * tol
    semi.close()
