# Run on any machine with Julia 1.9 + the reference's pinned Manifest (Oceananigans v0.90.1) to turn the
# "parity unpinned" caveat of DESIGN.md into a measured number: dumps the prognostic fields of the REAL reference
# after `nsteps` steps of the reference's own test configuration (test/runtests.jl:11-19: resolution 1/5 ->
# 72 x 30 x 10, dt = 600 s, DoubleDrake) so that tests/cases.py::ref_test can be diffed against it.
#
#   julia --project=/path/to/OceanScalingTests.jl julia/compare_with_oceananigans.jl out.jld2 10
using OceanScalingTests, Oceananigans, Oceananigans.Units, JLD2, MPI
MPI.Init()
out, nsteps = ARGS[1], parse(Int, ARGS[2])
sim = OceanScalingTests.scaling_test_simulation(1/5, (1, 1, 1), 10minutes, 365days;
                                                child_arch = CPU(), Nz = 10, experiment = :DoubleDrake)
for _ in 1:nsteps
    time_step!(sim.model, 600.0)
end
m = sim.model
jldsave(out; u = Array(interior(m.velocities.u)), v = Array(interior(m.velocities.v)), w = Array(interior(m.velocities.w)),
        η = Array(interior(m.free_surface.η)), T = Array(interior(m.tracers.T)), S = Array(interior(m.tracers.S)))
