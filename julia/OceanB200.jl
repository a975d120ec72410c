# OceanB200.jl -- Julia glue that makes libocean_b200.so a drop-in for Oceananigans' `time_step!` on the model
# OceanScalingTests.jl builds (src/near_global_simulation.jl:105-113).  NOT executed in this repository's image
# (no Julia here); kept small and literal.  See INTEGRATION.md.
module OceanB200

using Oceananigans
using Oceananigans.Grids: halo_size
using CUDA

const LIB = get(ENV, "OCEAN_B200_LIB", "libocean_b200.so")
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
    z_faces::Ptr{Float64} = C_NULL; bottom_height::Ptr{Float64} = C_NULL; averaging_weights::Ptr{Float64} = C_NULL
end

mutable struct Handle
    ptr::Ptr{Cvoid}
end

check(h, rc) = rc == 0 || error("libocean_b200: " * unsafe_string(ccall((:ob200_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

# field ids: include/ocean_b200.h :: enum ob200_field
const F_U, F_V, F_W, F_T, F_S = Int32(0), Int32(1), Int32(2), Int32(3), Int32(4)
const F_GU, F_GV, F_GT, F_GS = Int32(8), Int32(9), Int32(10), Int32(11)
const F_GU_PREV, F_GV_PREV, F_GT_PREV, F_GS_PREV = Int32(12), Int32(13), Int32(14), Int32(15)
const F_ETA, F_UBAR, F_VBAR = Int32(16), Int32(17), Int32(18)

"Create the native model from an Oceananigans model and copy its state in (replaces device allocation)."
function attach(model; device = CUDA.deviceid(), bottom_height, bc_kind, wind_coeffs, salt_coeffs, λ, μ)
    grid = model.grid
    arch = Oceananigans.architecture(grid)
    Nx, Ny, Nz = size(grid)
    fs = model.free_surface
    w = collect(Float64, fs.settings.substepping.averaging_weights)
    zf = collect(Float64, Array(grid.underlying_grid.zᵃᵃᶠ[1:Nz+1]))
    cfg = Config(; Nx, Ny, Nz, Nx_global = Nx * arch.ranks[1], i_offset = (arch.local_index[1] - 1) * Nx,
                 rank = arch.local_rank, nranks = arch.ranks[1],
                 eos_kind = model.buoyancy.model.equation_of_state isa LinearEquationOfState ? 0 : 1,
                 n_substeps = length(w), frac_dtau = fs.settings.substepping.fractional_step_size,
                 bc_kind, wind_coeffs, salt_coeffs, T_restoring_lambda = λ, drag_mu = μ)
    cfg.struct_bytes = sizeof(Config)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve w zf bottom_height begin
        cfg.z_faces = pointer(zf); cfg.bottom_height = pointer(bottom_height); cfg.averaging_weights = pointer(w)
        rc = ccall((:ob200_create, LIB), Cint, (Ref{Config}, Int32, Ref{Ptr{Cvoid}}), cfg, device, h)
        rc == 0 || error("ob200_create failed: " * unsafe_string(ccall((:ob200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
    end
    handle = Handle(h[])
    push_state!(handle, model)
    return handle
end

"Copy a Julia-owned field (parent array, any halo) into / out of the library: device-to-device."
function set_field!(h::Handle, id, field)
    Hx, Hy, Hz = halo_size(field.grid)
    CUDA.synchronize()
    check(h.ptr, ccall((:ob200_set_field, LIB), Cint, (Ptr{Cvoid}, Int32, CuPtr{Float64}, Int32, Int32, Int32, Int32),
                       h.ptr, id, pointer(parent(field)), 1, Hx, Hy, size(parent(field), 3) == 1 ? 0 : Hz))
end
function get_field!(field, h::Handle, id)
    Hx, Hy, Hz = halo_size(field.grid)
    check(h.ptr, ccall((:ob200_get_field, LIB), Cint, (Ptr{Cvoid}, Int32, CuPtr{Float64}, Int32, Int32, Int32, Int32),
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
function get_level!(dst::CuArray{Float64, 2}, h::Handle, id, k; halo = (0, 0))
    check(h.ptr, ccall((:ob200_get_level, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, CuPtr{Float64}, Int32, Int32, Int32),
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

"The replacement for Oceananigans.TimeSteppers.time_step!(model, Δt) (near_global_simulation.jl:181,196)."
function time_step!(h::Handle, model, Δt)
    check(h.ptr, ccall((:ob200_time_step, LIB), Cint, (Ptr{Cvoid}, Float64), h.ptr, Δt))
    Oceananigans.TimeSteppers.tick!(model.clock, Δt)
    model.clock.last_Δt = Δt
    return nothing
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
