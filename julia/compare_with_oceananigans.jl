# Run on any machine with Julia 1.9 + the reference's pinned Manifest (Oceananigans v0.90.1) to turn the "parity unpinned"
# caveat of DESIGN.md into a measured number.  Dumps the prognostic fields of the REAL reference after `nsteps` steps of
# the case tests/cases.py calls `ref_test`: the reference's own test configuration (test/runtests.jl:11-19: resolution
# 1/5 -> 72 x 30 x 10, dt = 600 s, DoubleDrake, CPU) started from `initialize_model!`
# (src/boundary_and_initial_conditions.jl:9-18) PLUS the closed-form perturbation of tests/cases.py (Case.initialize,
# perturb = True) so that every term of the step is exercised.  Output: an .npz (NPZ.jl) that tests/compare_dump.py
# diffs against the oracle's reference-arithmetic mode and, on a GPU box, against libocean_b200.so.
#
#   julia --project=/path/to/OceanScalingTests.jl julia/compare_with_oceananigans.jl ref_test_oceananigans.npz 100
#   python tests/compare_dump.py ref_test_oceananigans.npz
using OceanScalingTests, Oceananigans, Oceananigans.Units, NPZ, MPI
MPI.Init()
out, nsteps = ARGS[1], parse(Int, ARGS[2])
sim = OceanScalingTests.scaling_test_simulation(1/5, (1, 1, 1), 10minutes, 365days;
                                                child_arch = CPU(), Nz = 10, experiment = :DoubleDrake)
m = sim.model
# DESIGN.md spec decision 10: does this Oceananigans seed the averaged barotropic transports from u, v when a run is initialised?
HFSM = Oceananigans.Models.HydrostaticFreeSurfaceModels
@info "initialize_free_surface! defined in the pinned Oceananigans" isdefined(HFSM, :initialize_free_surface!)
# tests/cases.py :: Case.initialize (angles in degrees -> radians with r = pi / 180; D = 3000 m)
const D, r = 3000.0, π / 180
const T_reference = OceanScalingTests.T_reference                   # src/bathymetry_and_forcings.jl:154
const exponential_profile = OceanScalingTests.exponential_profile   # src/bathymetry_and_forcings.jl:45
Tᵢ(λ, φ, z) = exponential_profile(z; Lz = D, h = D / 15) * max(T_reference(φ), 2.0) + 0.5 * sin(3λ * r) * cos(2φ * r) * exp(z / 500)
Sᵢ(λ, φ, z) = 35.0 + 0.2 * cos(2λ * r + 0.3) * sin(3φ * r) * exp(z / 800)
uᵢ(λ, φ, z) = 0.1 * sin(2λ * r) * cos(φ * r) * (1 + z / D)
vᵢ(λ, φ, z) = 0.05 * cos(3λ * r) * sin(2φ * r) * exp(z / 1000)
ηᵢ(λ, φ, z) = 0.1 * cos(φ * r) * sin(2λ * r)
set!(m, T = Tᵢ, S = Sᵢ, u = uᵢ, v = vᵢ, η = ηᵢ)
for _ in 1:nsteps
    time_step!(m, 600.0)
end
# numpy reads the bytes of a Julia (i, j, k) array as (k, j, i): permute so that the .npz holds (k, j, i) C-ordered arrays
pd(a) = permutedims(Array(a), (3, 2, 1))
npzwrite(out, Dict("steps" => nsteps, "dt" => 600.0,
                   "U" => pd(interior(m.velocities.u)), "V" => pd(interior(m.velocities.v)), "W" => pd(interior(m.velocities.w)),
                   "ETA" => pd(interior(m.free_surface.η))[1, :, :], "T" => pd(interior(m.tracers.T)), "S" => pd(interior(m.tracers.S))))
