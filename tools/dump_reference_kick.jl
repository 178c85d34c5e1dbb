# Produces golden fixtures of the reference's own CPU `kick!` / `drift!` for the parity tests.
# Needs Julia >= 1.10 with TrixiParticles.jl and OrdinaryDiffEq installed (not available in the
# build image, so this script has not been executed there).  Usage:
#
#   julia --project -t auto tools/dump_reference_kick.jl examples/fluid/dam_break_2d.jl out_dir
#
# Writes raw little-endian binaries (Julia column-major == particle-major C order):
#   u_ode.bin v_ode.bin dv_ode.bin du_ode.bin   ODE vectors (semidiscretization.jl:128-135)
#   fluid_pressure.bin wall_pressure.bin wall_density.bin
#   meta.txt                                     sizes, eltypes, system order
# tests/test_reference_fixtures.py picks them up from tests/golden/reference_kick/<name>/ and
# compares the CUDA path at 1e-12 (Float64) / 1e-5 (Float32).
using TrixiParticles
using OrdinaryDiffEq

example, outdir = ARGS[1], ARGS[2]
mkpath(outdir)
# build everything the example builds, but do not solve (reference idiom: sol = nothing)
trixi_include(@__MODULE__, joinpath(TrixiParticles.examples_dir(), relpath(example, "examples")),
              sol = nothing)
v_ode, u_ode = ode.u0.x
dv_ode, du_ode = similar(v_ode), similar(u_ode)
TrixiParticles.kick!(dv_ode, v_ode, u_ode, ode.p, 0.0)
TrixiParticles.drift!(du_ode, v_ode, u_ode, ode.p, 0.0)
dump(name, a) = write(joinpath(outdir, name), Array(a))
dump("u_ode.bin", u_ode); dump("v_ode.bin", v_ode)
dump("dv_ode.bin", dv_ode); dump("du_ode.bin", du_ode)
semi = ode.p.semi
open(joinpath(outdir, "meta.txt"), "w") do io
    println(io, "eltype_v = ", eltype(v_ode), "\neltype_u = ", eltype(u_ode))
    println(io, "length_v = ", length(v_ode), "\nlength_u = ", length(u_ode))
    for (i, system) in enumerate(semi.systems)
        println(io, "system[$i] = ", nameof(typeof(system)), " n = ", TrixiParticles.nparticles(system))
        if system isa TrixiParticles.WeaklyCompressibleSPHSystem
            dump("fluid_pressure.bin", system.pressure)
        elseif system isa TrixiParticles.WallBoundarySystem
            dump("wall_pressure.bin", system.boundary_model.pressure)
            dump("wall_density.bin", system.boundary_model.cache.density)
        end
    end
end
