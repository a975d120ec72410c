# OceanB200.jl -- Julia glue that makes libocean_b200.so a drop-in for Oceananigans' `time_step!` on the model
# OceanScalingTests.jl builds (src/near_global_simulation.jl:105-113).  NOT executed in this repository's image
# (no Julia here); kept small and literal.  See INTEGRATION.md.
#
# After `OceanB200.attach(model; ...)` the reference's own call sites -- `time_step!(model, Δt)` at
# src/near_global_simulation.jl:181,196 and Oceananigans' `run!` (experiments/run.jl:60) -- dispatch to the library
# through the method defined at the bottom of this file; a model that was not attached falls through to Oceananigans.
module OceanB200

using Oceananigans
using Oceananigans.Grids: halo_size
using Oceananigans.Models.HydrostaticFreeSurfaceModels: HydrostaticFreeSurfaceModel
using Oceananigans.TimeSteppers: QuasiAdamsBashforth2TimeStepper
using Oceananigans: AbstractModel
using CUDA
import Oceananigans.TimeSteppers: time_step!

# The reference takes its precision from ENV["PRECISION"] (experiments/run.jl:38); so does the choice of the library: the
# Float32 build has the same entry points and Float32 field buffers (include/ocean_b200.h `ob200_real`).
const FT = get(ENV, "PRECISION", "Float64") == "Float32" ? Float32 : Float64
const LIB = get(ENV, "OCEAN_B200_LIB", FT === Float32 ? "libocean_b200_f32.so" : "libocean_b200.so")
const HALO = 7

# include/ocean_b200.h :: ob200_config (field order and types must match the header)
Base.@kwdef mutable struct Config
    struct_bytes::Int32 = 0
    Nx::Int32; Ny::Int32; Nz::Int32
    Nx_global::Int32; i_offset::Int32; rank::Int32; nranks::Int32
    lon_west::Float64 = -180; lon_east::Float64 = 180; lat_south::Float64 = -75; lat_north::Float64 = 75
    radius::Float64 = 6371.0e3; gravity::Float64 = 9.80665; rotation_rate::Float64 = 7.292115e-5
    eos_kind::Int32 = 0; eos_alpha::Float64 = 1.67e-4; eos_beta::Float64 = 7.80e-4; eos_rho0::Float64 = 1020
    nu_z::Float64 = 5e-4; kappa_z::Float64 = 3e-5
    ri_enabled::Int32 = 1
    ri_nu0::Float64 = 0.7; ri_kappa0::Float64 = 0.5; ri_kappa_ca::Float64 = 1.7; ri_Cen::Float64 = 0.1
    ri_Cav::Float64 = 0.6; ri_Ri0::Float64 = 0.1; ri_Ridelta::Float64 = 0.4
    order_vorticity::Int32 = 9; order_divergence::Int32 = 5; order_tracer::Int32 = 7
    ab2_chi::Float64 = 0.1
    n_substeps::Int32; frac_dtau::Float64
    bc_kind::Int32 = 0
    wind_coeffs::NTuple{24, Float64} = ntuple(_ -> 0.0, 24)
    salt_coeffs::NTuple{16, Float64} = ntuple(_ -> 0.0, 16)
    T_restoring_lambda::Float64 = 0; drag_mu::Float64 = 0
    qgleith_enabled::Int32 = 0; leith_C::Float64 = 2; leith_min_N2::Float64 = 1e-20; leith_Vscale::Float64 = 1
    bottom_halo_columns::Int32 = 7
    barotropic_exchange::Int32 = 0      # OB200_BAROX_AUTO: wide halo when the slab allows it, per-substep exchange otherwise
    z_faces::Ptr{Float64} = C_NULL; bottom_height::Ptr{Float64} = C_NULL; averaging_weights::Ptr{Float64} = C_NULL
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    sync_every::Int          # refresh the Julia-owned arrays every n iterations (0 = only on request)
end

"models that were attached: `time_step!(model, Δt)` looks its handle up here"
const ATTACHED = IdDict{Any, Handle}()

check(h, rc) = rc == 0 || error("libocean_b200: " * unsafe_string(ccall((:ob200_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

# field ids: include/ocean_b200.h :: enum ob200_field
const F_U, F_V, F_W, F_T, F_S = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
const F_GU, F_GV, F_GT, F_GS = Int32(8), Int32(9), Int32(10), Int32(11)
const F_GU_PREV, F_GV_PREV, F_GT_PREV, F_GS_PREV = Int32(12), Int32(13), Int32(14), Int32(15)
const F_ETA, F_UBAR, F_VBAR = Int32(16), Int32(17), Int32(18)

# x-slab geometry from the Distributed architecture: equal `Partition(Rx)` or `Partition(x = Sizes(local_Nx...))`
# (src/grid_load_balance.jl:62 with LOADBALANCE=1) -- the offset is the sum of the slabs to the west, not rank * Nx
function slab_geometry(arch, Nx)
    arch isa Oceananigans.DistributedComputations.Distributed || return (Nx, 0, 0, 1)
    R, r = arch.ranks[1], arch.local_index[1] - 1
    px = arch.partition.x
    sizes = px isa Oceananigans.DistributedComputations.Sizes ? collect(Int, px.sizes) : fill(Nx, R)
    return (sum(sizes), sum(sizes[1:r]), r, R)
end

"""
    attach(model; bottom_height, bc_kind, wind_coeffs, salt_coeffs, λ, μ, comm = nothing, MPI = nothing, sync_every = 10)

Create the native model for `model`, register the Julia-owned arrays the reference's callbacks and writers read
(`ob200_bind`: u, v, w, η, T, S, Gⁿ, G⁻ -- src/near_global_simulation.jl:138-162, src/simulation_outputs.jl:14-37,47-64),
copy the state in and remember the handle, so that `time_step!(model, Δt)` runs on the library from now on.

`bottom_height` is the rank's host `Array{Float64}(Nx + 2 h, Ny)` (i fastest) INCLUDING `h = bottom_halo_columns(model)` x-halo columns
(`partition_global_array` plus periodic / neighbour columns; h = max(7, substeps + 1) on x-slabs, where the barotropic
solver carries a wide halo).  `sync_every` = 10 matches the reference's callback schedules (wizard every 10 iterations,
progress every 100, near_global_simulation.jl:160-162); use 0 in profile mode and call `sync_outputs!` yourself.
"""
function attach(model; device = CUDA.deviceid(), bottom_height, bc_kind, wind_coeffs, salt_coeffs, λ, μ,
                comm = nothing, MPI = nothing, sync_every = 10)
    grid = model.grid
    eltype(grid) === FT || error("model precision $(eltype(grid)) but ENV[\"PRECISION\"] selected the $(FT) library ($LIB)")
    ccall((:ob200_real_bytes, LIB), Int32, ()) == sizeof(FT) || error("$LIB is not the $(FT) build")
    arch = Oceananigans.architecture(grid)
    Nx, Ny, Nz = size(grid)
    fs = model.free_surface
    w = collect(Float64, fs.settings.substepping.averaging_weights)
    ug = hasproperty(grid, :underlying_grid) ? grid.underlying_grid : grid      # Quiescent has no immersed boundary
    zf = collect(Float64, Array(ug.zᵃᵃᶠ[1:Nz+1]))
    Nx_global, i_offset, rank, nranks = slab_geometry(arch, Nx)
    hb = bottom_halo_columns(model)
    size(bottom_height) == (Nx + 2hb, Ny) || error("bottom_height must be (Nx + 2*$hb, Ny) = ($(Nx + 2hb), $Ny), got $(size(bottom_height))")
    cfg = Config(; Nx, Ny, Nz, Nx_global, i_offset, rank, nranks,
                 eos_kind = model.buoyancy.model.equation_of_state isa LinearEquationOfState ? 0 : 1,
                 n_substeps = length(w), frac_dtau = fs.settings.substepping.fractional_step_size,
                 bc_kind, wind_coeffs = NTuple{24, Float64}(Tuple(wind_coeffs)), salt_coeffs = NTuple{16, Float64}(Tuple(salt_coeffs)),
                 T_restoring_lambda = λ, drag_mu = μ, bottom_halo_columns = hb)
    cfg.struct_bytes = sizeof(Config)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve w zf bottom_height begin
        cfg.z_faces = pointer(zf); cfg.bottom_height = pointer(bottom_height); cfg.averaging_weights = pointer(w)
        rc = ccall((:ob200_create, LIB), Cint, (Ref{Config}, Int32, Ref{Ptr{Cvoid}}), cfg, device, h)
        rc == 0 || error("ob200_create failed: " * unsafe_string(ccall((:ob200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
    handle = Handle(h[], sync_every)
    nranks > 1 && comm_init!(handle, comm, MPI)
    bind_model!(handle, model)
    sync_inputs!(handle)                 # the state `set!` / `initialize_model!` left in the Julia arrays, + update_state!
    seed_barotropic_averages!(handle, model)
    ATTACHED[model] = handle
    return handle
end

"x-halo columns the bottom-height array must carry: 7, or substeps + 1 on x-slabs (wide barotropic halo, SURVEY F9)"
function bottom_halo_columns(model)
    arch = Oceananigans.architecture(model.grid)
    M = length(model.free_surface.settings.substepping.averaging_weights)
    return arch isa Oceananigans.DistributedComputations.Distributed && arch.ranks[1] > 1 ? max(HALO, M + 1) : HALO
end

"ob200_bind: the library borrows the parent arrays for the lifetime of the handle (the model must outlive it)"
function bind!(h::Handle, id, field)
    Hx, Hy, Hz = halo_size(field.grid)
    check(h.ptr, ccall((:ob200_bind, LIB), Cint, (Ptr{Cvoid}, Int32, CuPtr{FT}, Int32, Int32, Int32),
                       h.ptr, id, pointer(parent(field)), Hx, Hy, size(parent(field), 3) == 1 ? 0 : Hz))
end
function bind_model!(h::Handle, model)
    u, v, w = model.velocities
    bind!(h, F_U, u); bind!(h, F_V, v); bind!(h, F_W, w)
    bind!(h, F_T, model.tracers.T); bind!(h, F_S, model.tracers.S); bind!(h, F_ETA, model.free_surface.η)
    G⁻, Gⁿ = model.timestepper.G⁻, model.timestepper.Gⁿ
    bind!(h, F_GU, Gⁿ.u); bind!(h, F_GV, Gⁿ.v); bind!(h, F_GT, Gⁿ.T); bind!(h, F_GS, Gⁿ.S)
    bind!(h, F_GU_PREV, G⁻.u); bind!(h, F_GV_PREV, G⁻.v); bind!(h, F_GT_PREV, G⁻.T); bind!(h, F_GS_PREV, G⁻.S)
end
"Refresh every bound Julia array from the library (what callbacks / writers / the checkpointer read)."
sync_outputs!(h::Handle) = check(h.ptr, ccall((:ob200_sync_outputs, LIB), Cint, (Ptr{Cvoid},), h.ptr))
sync_outputs!(model) = sync_outputs!(ATTACHED[model])
"Take the bound Julia arrays into the library (after `set!(model, ...)` or a restore) and run update_state!."
function sync_inputs!(h::Handle)
    CUDA.synchronize()
    check(h.ptr, ccall((:ob200_sync_inputs, LIB), Cint, (Ptr{Cvoid},), h.ptr))
end
function sync_inputs!(model)
    h = ATTACHED[model]
    sync_inputs!(h)
    seed_barotropic_averages!(h, model)
end

"Every step starts its substeps from the previous step's averaged transports (U ← U̅, V ← V̅).  Whatever Oceananigans holds in
`free_surface.state.U̅, V̅` -- zeros for a fresh model, `barotropic_mode!` of the velocities where its `initialize_free_surface!`
ran -- seeds the library, so the first native step continues exactly where Oceananigans' own step would have."
function seed_barotropic_averages!(h::Handle, model)
    st = model.free_surface.state
    set_field!(h, F_UBAR, st.U̅); set_field!(h, F_VBAR, st.V̅)
end

"Copy a Julia-owned field (parent array, any halo) into / out of the library: device-to-device."
function set_field!(h::Handle, id, field)
    Hx, Hy, Hz = halo_size(field.grid)
    CUDA.synchronize()
    check(h.ptr, ccall((:ob200_set_field, LIB), Cint, (Ptr{Cvoid}, Int32, CuPtr{FT}, Int32, Int32, Int32, Int32),
                       h.ptr, id, pointer(parent(field)), 1, Hx, Hy, size(parent(field), 3) == 1 ? 0 : Hz))
end
function get_field!(field, h::Handle, id)
    Hx, Hy, Hz = halo_size(field.grid)
    check(h.ptr, ccall((:ob200_get_field, LIB), Cint, (Ptr{Cvoid}, Int32, CuPtr{FT}, Int32, Int32, Int32, Int32),
                       h.ptr, id, pointer(parent(field)), 1, Hx, Hy, size(parent(field), 3) == 1 ? 0 : Hz))
end

function push_state!(h::Handle, model)
    u, v, _ = model.velocities
    set_field!(h, F_U, u); set_field!(h, F_V, v)
    set_field!(h, F_T, model.tracers.T); set_field!(h, F_S, model.tracers.S)
    set_field!(h, F_ETA, model.free_surface.η)
    check(h.ptr, ccall((:ob200_update_state, LIB), Cint, (Ptr{Cvoid},), h.ptr))
end

"One horizontal level of a 3-D field into a Julia-owned 2-D device array (the surface writer's `indices = (:, :, Nz)` views,
src/simulation_outputs.jl:17-25) without moving the whole field; `k` is 1-based like the reference's index."
function get_level!(dst::CuArray{FT, 2}, h::Handle, id, k; halo = (0, 0))
    check(h.ptr, ccall((:ob200_get_level, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, CuPtr{FT}, Int32, Int32, Int32),
                       h.ptr, id, k - 1, pointer(dst), 1, halo[1], halo[2]))
end

"Refresh the Julia-owned arrays the output writers / checkpointer read (src/simulation_outputs.jl:14-37,47-64)."
function pull_state!(model, h::Handle)
    u, v, w = model.velocities
    get_field!(u, h, F_U); get_field!(v, h, F_V); get_field!(w, h, F_W)
    get_field!(model.tracers.T, h, F_T); get_field!(model.tracers.S, h, F_S)
    get_field!(model.free_surface.η, h, F_ETA)
    G⁻, Gⁿ = model.timestepper.G⁻, model.timestepper.Gⁿ
    get_field!(Gⁿ.u, h, F_GU); get_field!(Gⁿ.v, h, F_GV); get_field!(Gⁿ.T, h, F_GT); get_field!(Gⁿ.S, h, F_GS)
    get_field!(G⁻.u, h, F_GU_PREV); get_field!(G⁻.v, h, F_GV_PREV); get_field!(G⁻.T, h, F_GT_PREV); get_field!(G⁻.S, h, F_GS_PREV)
    check(h.ptr, ccall((:ob200_sync, LIB), Cint, (Ptr{Cvoid},), h.ptr))
end

"The native step: ob200_time_step, then the Julia clock (the reference's callbacks and schedules read `model.clock`)."
function time_step!(h::Handle, model, Δt)
    check(h.ptr, ccall((:ob200_time_step, LIB), Cint, (Ptr{Cvoid}, Float64), h.ptr, Δt))
    Oceananigans.TimeSteppers.tick!(model.clock, Δt)
    hasproperty(model.clock, :last_Δt) && (model.clock.last_Δt = Δt)
    h.sync_every > 0 && model.clock.iteration % h.sync_every == 0 && sync_outputs!(h)
    return nothing
end

# THE DROP-IN: a method of Oceananigans' own `time_step!`, more specific than its
# `time_step!(::AbstractModel{<:QuasiAdamsBashforth2TimeStepper}, Δt; ...)`, so `time_step!(model, Δt)` at
# src/near_global_simulation.jl:181,196 and inside `run!` (experiments/run.jl:60) needs NO edit.  Attached models step on
# the library (the `euler` decision is the library's: a Δt that differs from the previous one takes a forward-Euler step,
# exactly Oceananigans' rule); any other model is handed to the method this one shadows.
function time_step!(model::HydrostaticFreeSurfaceModel{<:QuasiAdamsBashforth2TimeStepper}, Δt; callbacks = [], kw...)
    h = get(ATTACHED, model, nothing)
    h === nothing && return invoke(time_step!, Tuple{AbstractModel{<:QuasiAdamsBashforth2TimeStepper}, Any}, model, Δt; callbacks, kw...)
    time_step!(h, model, Δt)
    isempty(callbacks) || (sync_outputs!(h); foreach(cb -> cb(model), callbacks))    # state callbacks of Oceananigans' run!
    return nothing
end

"Release the native model (the Julia arrays stay with Julia)."
function detach!(model)
    h = pop!(ATTACHED, model)
    ccall((:ob200_destroy, LIB), Cvoid, (Ptr{Cvoid},), h.ptr)
end

"NCCL ring bootstrap over the MPI communicator the reference already has (experiments/run.jl:9-27)."
function comm_init!(h::Handle, comm, MPI)
    id = zeros(UInt8, 128)
    MPI.Comm_rank(comm) == 0 && ccall((:ob200_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id)
    MPI.Bcast!(id, 0, comm)
    check(h.ptr, ccall((:ob200_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), h.ptr, id))
end

# ---- checkpoint containers ------------------------------------------------------------------------------------
# The Python host mirror (simulation_outputs.py) stores the reference's JLD2 groups under the same names in an
# uncompressed zip of .npy members.  With NPZ.jl (not in the reference's Manifest; `] add NPZ`):
#
#   checkpoint_to_npz("RealisticOcean_checkpoint_0_iteration1000.jld2", "RealisticOcean_checkpoint_0_iteration1000.npz")
#
# (the inverse reads the same group names back with `NPZ.npzread` and writes them with `JLD2.jldopen(path, "w")`).
# Julia arrays are (i, j, k) column-major; numpy reads the same bytes as (k, j, i) C-ordered, so `permutedims` with
# (3, 2, 1) converts between what NPZ.jl writes and what the Python side expects.
const CHECKPOINT_GROUPS = vcat(["$n/data" for n in ("u", "v", "T", "S", "η")],
                               ["timestepper/$G/$n/data" for G in ("Gⁿ", "G⁻") for n in ("u", "v", "T", "S")])

function checkpoint_to_npz(jld2path, npzpath; JLD2, NPZ)
    JLD2.jldopen(jld2path, "r") do file
        out = Dict{String, Any}()
        for g in CHECKPOINT_GROUPS
            haskey(file, g) && (out[g] = permutedims(Array(file[g]), (3, 2, 1)))
        end
        clock = file["clock"]
        out["clock/time"] = Float64(clock.time)
        out["clock/iteration"] = Int64(clock.iteration)
        NPZ.npzwrite(npzpath, out)
    end
end

end # module
