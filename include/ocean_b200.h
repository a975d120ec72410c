/*
 * libocean_b200.so -- C ABI of the B200-native hydrostatic free-surface time step.
 *
 * This is the drop-in boundary for the ONE hot path OceanScalingTests.jl
 * benchmarks: Oceananigans' `time_step!(model::HydrostaticFreeSurfaceModel, dt)`
 * as configured by /root/reference/src/near_global_simulation.jl:36-113 and
 * called at src/near_global_simulation.jl:181,196 (profile loop) and through
 * `run!` at experiments/run.jl:60.  A Julia maintainer binds these entry points
 * with `ccall` (see INTEGRATION.md and julia/OceanB200.jl); in this repository
 * the Python host mirror (oceanscalingtests.jl_b200/) binds the same symbols
 * with ctypes.  Plain pointers and sizes only -- no torch / CUDA.jl types.
 *
 * Conventions
 *   - every entry point returns int: 0 = OB200_OK, negative = error; the message
 *     is retrievable with ob200_last_error(); nothing throws across the ABI.
 *   - one host thread per handle (the reference is 1 MPI rank : 1 GPU).
 *   - arithmetic and field buffers are Float64 in libocean_b200.so (experiments/run.jl:38 PRECISION default) and
 *     Float32 in libocean_b200_f32.so (PRECISION="Float32", launch_strong_scaling.sh:7): the same sources built with
 *     -DOB200_F32, the same entry points; `ob200_real` below is the element type of every FIELD buffer, and
 *     ob200_real_bytes() tells a binding which of the two libraries it loaded.  Configuration values, the model
 *     clock, dt and the diagnostics are double in both.
 *   - indices are 0-based here; face i lies between cells i-1 and i
 *     (Oceananigans' convention shifted by one).
 */
#ifndef OCEAN_B200_H
#define OCEAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef OB200_F32
typedef float ob200_real;
#else
typedef double ob200_real;
#endif

#define OB200_OK            0
#define OB200_EINVAL       -1   /* bad argument / inconsistent config        */
#define OB200_ECUDA        -2   /* CUDA runtime error (message has details)  */
#define OB200_ENOMEM       -3
#define OB200_ENCCL        -4   /* NCCL error                                */
#define OB200_ESTATE       -5   /* call order violated (e.g. comm not ready) */

#define OB200_HALO          7   /* src/grid_load_balance.jl:25 halo = (7,7,7) */
#define OB200_MAX_SUBSTEPS  512

/* Equation of state: src/near_global_simulation.jl:23-24 */
#define OB200_EOS_LINEAR    0   /* DoubleDrake: LinearEquationOfState        */
#define OB200_EOS_TEOS10    1   /* Quiescent / RealisticOcean                */

/* Boundary-condition sets: src/boundary_and_initial_conditions.jl:48-137 */
#define OB200_BC_QUIESCENT    0 /* zero top fluxes (:48-56)                  */
#define OB200_BC_DOUBLEDRAKE  1 /* cubic wind/salt, T restoring, linear drag (:58-81) */
#define OB200_BC_REALISTIC    2 /* time-interpolated flux arrays, quadratic drag (:83-137) */

#define OB200_BAROX_AUTO      0
#define OB200_BAROX_WIDE      1
#define OB200_BAROX_SUBSTEP   2

/* Field identifiers for ob200_set_field / ob200_get_field.
 * Interior extents (i fastest, then j, then k):
 *   U,T,S,P,GU,GT,GS,..  Nx x Ny     x Nz
 *   V,GV,..              Nx x (Ny+1) x Nz        (Face in Bounded y)
 *   W,KAPPA_C,KAPPA_U    Nx x Ny     x (Nz+1)    (Face in z)
 *   ETA,UBAR(TRANS)      Nx x Ny ;  VBAR(TRANS)  Nx x (Ny+1)
 */
enum ob200_field {
    OB200_F_U = 0, OB200_F_V, OB200_F_W, OB200_F_T, OB200_F_S, OB200_F_P,
    OB200_F_KAPPA_C, OB200_F_KAPPA_U,
    OB200_F_GU, OB200_F_GV, OB200_F_GT, OB200_F_GS,          /* timestepper.G^n */
    OB200_F_GU_PREV, OB200_F_GV_PREV, OB200_F_GT_PREV, OB200_F_GS_PREV, /* G^-  */
    OB200_F_ETA, OB200_F_UBAR, OB200_F_VBAR,                 /* eta, averaged transports */
    OB200_F_UTRANS, OB200_F_VTRANS,                          /* instantaneous U, V */
    OB200_F_GUBT, OB200_F_GVBT,                              /* barotropic forcing G^U, G^V */
    OB200_F_NU_LEITH, OB200_F_QX, OB200_F_QY, OB200_F_LD,    /* QG-Leith diagnostics */
    OB200_F_COUNT
};

/* Stages of one time step, exposed so that tests (and a host that wants to do
 * its own halo exchange, e.g. Julia's CUDA-aware MPI) can sequence them.
 * ob200_time_step() runs them in this order.                 */
enum ob200_stage {
    OB200_ST_BAROTROPIC_FORCING = 0, /* G^U,G^V = sum_k dz AB2(Gu,Gv)           */
    OB200_ST_AB2_IMPLICIT,           /* ab2_step_field! + implicit_step! x4     */
    OB200_ST_BAROTROPIC_INIT,        /* U<-Ubar, V<-Vbar, averages zeroed       */
    OB200_ST_BAROTROPIC_ETA,         /* substep m: eta update (arg = m)         */
    OB200_ST_BAROTROPIC_VEL,         /* substep m: U,V update + averaging       */
    OB200_ST_BAROTROPIC_FINISH,      /* eta <- eta_bar                          */
    OB200_ST_CORRECT_AND_STORE,      /* barotropic corrector + G^- <- G^n       */
    OB200_ST_MASK_AND_LOCAL_HALOS,   /* mask fields, fill y/z (and periodic x)  */
    OB200_ST_AUXILIARIES,            /* w, Ri/kappa, pHY', (QG-Leith)           */
    OB200_ST_TENDENCIES,             /* Gu,Gv,GT,GS incl. flux BCs              */
    OB200_ST_BAROTROPIC_LOOP,        /* all substeps incl. internal exchanges   */
    OB200_ST_EXCHANGE_HALOS,         /* x-halo exchange of u,v,T,S,eta (NCCL)   */
    OB200_ST_COUNT
};

typedef struct ob200_config {
    int32_t struct_bytes;      /* = sizeof(ob200_config), ABI check            */
    /* grid: src/near_global_simulation.jl:59-64, src/grid_load_balance.jl:21-26 */
    int32_t Nx, Ny, Nz;        /* LOCAL interior size of this rank's x-slab    */
    int32_t Nx_global;         /* 360*resolution                               */
    int32_t i_offset;          /* global column index of local column 0        */
    int32_t rank, nranks;      /* ring position along the periodic x axis      */
    double  lon_west, lon_east;    /* (-180, 180)                              */
    double  lat_south, lat_north;  /* (-75, 75)                                */
    double  radius;            /* 6371.0e3                                     */
    double  gravity;           /* 9.80665                                      */
    double  rotation_rate;     /* 7.292115e-5                                  */
    /* buoyancy: src/near_global_simulation.jl:86 */
    int32_t eos_kind;
    double  eos_alpha, eos_beta;   /* linear: 1.67e-4, 7.80e-4                 */
    double  eos_rho0;              /* TEOS-10 reference density (1020)         */
    /* closures: src/near_global_simulation.jl:70-73,49,87 */
    double  nu_z, kappa_z;     /* 5e-4, 3e-5                                   */
    int32_t ri_enabled;
    double  ri_nu0, ri_kappa0, ri_kappa_ca, ri_Cen, ri_Cav, ri_Ri0, ri_Ridelta;
    /* advection orders (2N-1): src/near_global_simulation.jl:30-32,75-78 */
    int32_t order_vorticity, order_divergence, order_tracer;  /* 9, 5, 7       */
    double  ab2_chi;           /* 0.1                                          */
    /* split-explicit free surface: src/near_global_simulation.jl:82,11-19 */
    int32_t n_substeps;        /* M = number of averaging weights              */
    double  frac_dtau;         /* barotropic step as a fraction of dt (2/N)    */
    /* boundary conditions: src/boundary_and_initial_conditions.jl:58-81 */
    int32_t bc_kind;
    double  wind_coeffs[24];   /* 6 pieces x (a3,a2,a1,a0), already /1000      */
    double  salt_coeffs[16];   /* 4 pieces x 4, already x S_ref                */
    double  T_restoring_lambda;/* 5 m / 7 days                                 */
    double  drag_mu;           /* 0.003 linear (DoubleDrake), 0.001 quadratic  */
    /* optional QG-Leith closure: src/quasi_geostrophic_leith.jl:37-48 */
    int32_t qgleith_enabled;
    double  leith_C, leith_min_N2, leith_Vscale;
    /* host arrays, copied during ob200_create */
    int32_t bottom_halo_columns;     /* x-halo columns carried by bottom_height: >= OB200_HALO; the wide barotropic
                                        halo (below) needs >= n_substeps + 1                                      */
    /* x-slabs only: how the barotropic substeps see their neighbours (BASELINE configs[4]).
     * OB200_BAROX_WIDE: halo of n_substeps + 1 columns exchanged ONCE per step, no communication inside the loop
     *   (what Oceananigans' distributed solver does; src/simulation_outputs.jl:70-80); needs Nx and
     *   bottom_halo_columns >= n_substeps + 1.
     * OB200_BAROX_SUBSTEP: one-column exchanges of eta / U every half substep; works on slabs as narrow as the
     *   reference's 20-column clamp (src/grid_load_balance.jl:132-145).
     * OB200_BAROX_AUTO: wide when the slab allows it, per-substep otherwise.                                      */
    int32_t barotropic_exchange;
    const double* z_faces;           /* Nz+1, bottom to top (bathymetry_and_forcings.jl:47-58) */
    const double* bottom_height;     /* Ny rows x (Nx + 2*bottom_halo_columns) columns incl. x-halos */
    const double* averaging_weights; /* n_substeps                                             */
} ob200_config;

typedef struct ob200_handle ob200_handle;

/* Stage groups timed with CUDA events on the handle's stream when option "timing" = 1:
 * 0 barotropic forcing + AB2/implicit x4, 1 barotropic substeps, 2 corrector, 3 halos/exchange,
 * 4 w + Ri sweep, 5 pHY' + kappa sweep (+ QG-Leith), 6 momentum tendencies, 7 tracer tendencies,
 * 8 the T,S part of group 0.  With option "overlap" = 1 (default 0) group 8 runs on a second stream beside
 * group 1 and the groups no longer add up to the step time.                                               */
#define OB200_TIMED_GROUPS 9
typedef struct ob200_timing {
    int32_t n;                 /* number of timed stage groups (OB200_TIMED_GROUPS) */
    float   ms[16];            /* last step: per stage group, CUDA-event ms          */
    int64_t launches;          /* kernels launched by this library since create      */
} ob200_timing;

/* --- lifecycle ---------------------------------------------------------- */
/* Replaces HydrostaticFreeSurfaceModel(...) device allocation
 * (src/near_global_simulation.jl:105-113). device = CUDA ordinal.           */
int  ob200_create(const ob200_config* cfg, int32_t device, ob200_handle** out);
void ob200_destroy(ob200_handle* h);
const char* ob200_last_error(const ob200_handle* h);   /* h may be NULL      */
const char* ob200_version(void);
int32_t ob200_real_bytes(void);                        /* 8 (libocean_b200.so) or 4 (libocean_b200_f32.so) */

/* --- field transfer ------------------------------------------------------
 * `buf` is laid out like an Oceananigans parent array with halos (hx,hy,hz)
 * (0,0,0 = interior(field)); only the interior region is read / written.
 * Replaces set!(model, ...) (src/boundary_and_initial_conditions.jl:15) and
 * the reads done by the output writers (src/simulation_outputs.jl:14-37).   */
int ob200_set_field(ob200_handle* h, int32_t field, const ob200_real* buf, int32_t buf_on_device,
                    int32_t hx, int32_t hy, int32_t hz);
int ob200_get_field(ob200_handle* h, int32_t field, ob200_real* buf, int32_t buf_on_device,
                    int32_t hx, int32_t hy, int32_t hz);
/* One horizontal level k (0-based) of a 3-D field as a (ny + 2 hy) x (Nx + 2 hx) window: what the surface writer's
 * `Field(model.velocities.u; indices = (:, :, Nz))` views read (src/simulation_outputs.jl:17-25).                  */
int ob200_get_level(ob200_handle* h, int32_t field, int32_t k, ob200_real* buf, int32_t buf_on_device,
                    int32_t hx, int32_t hy);
/* --- borrowed arrays ("Julia owns, library borrows": SURVEY 8b) -------------------------------------------------
 * The reference's callbacks and writers read `model.velocities`, `model.tracers`, `model.free_surface.η` and
 * `model.timestepper.Gⁿ / G⁻` directly (src/near_global_simulation.jl:138-162, src/simulation_outputs.jl:14-37,47-64).
 * ob200_bind registers the caller-owned DEVICE parent array of a field (Oceananigans layout, halos (hx, hy, hz); the
 * library keeps its own 128-byte-aligned, pitched working copy, which TMA needs) for the lifetime of the handle or until
 * ob200_unbind; nothing is copied by the call itself.  ob200_sync_outputs refreshes every bound array from the library
 * (one device-to-device pass per bound field: u, v, w, eta, T, S at 1/6 degree = 3.5 ms, so the Julia glue calls it on
 * the schedule of the callbacks, not every step); ob200_sync_inputs goes the other way (after `set!(model, ...)`) and
 * ends with update_state!.  Both are ordered on the handle's stream; ob200_sync_outputs returns after the copies are done. */
int ob200_bind(ob200_handle* h, int32_t field, ob200_real* dev_ptr, int32_t hx, int32_t hy, int32_t hz);
int ob200_unbind(ob200_handle* h, int32_t field);
int ob200_sync_outputs(ob200_handle* h);
int ob200_sync_inputs(ob200_handle* h);

/* RealisticOcean forcing arrays (src/boundary_and_initial_conditions.jl:90-105):
 * which = 0 taux, 1 tauy, 2 Qs, 3 Fs, 4 Tr, 5 Sr; Nx x Ny(+1 for tauy) x 6.   */
int ob200_set_forcing(ob200_handle* h, int32_t which, const ob200_real* buf, int32_t buf_on_device);

/* --- the hot path --------------------------------------------------------- */
/* update_state!(model) -- mask, halos, w, kappa, pHY', tendencies.           */
int ob200_update_state(ob200_handle* h);
/* time_step!(model, dt): a step whose dt differs from the previous one (or the
 * first) is a forward-Euler step (chi = -0.5, G^- zeroed); at iteration 0
 * update_state! runs first.  Asynchronous on the handle's stream.            */
int ob200_time_step(ob200_handle* h, double dt);
/* n identical steps back to back (the profile loop, near_global_simulation.jl:192-200) */
int ob200_time_steps(ob200_handle* h, double dt, int32_t n);
int ob200_run_stage(ob200_handle* h, int32_t stage, double dt, int32_t arg);
int ob200_sync(ob200_handle* h);
int ob200_get_clock(ob200_handle* h, double* time, int64_t* iteration, double* last_dt);
int ob200_set_clock(ob200_handle* h, double time, int64_t iteration, double last_dt);
/* max |u|,|v|,|w|,|eta|,|T|,|S| (progress callback, near_global_simulation.jl:138-160)
 * and the advective CFL time scale used by TimeStepWizard (:161).            */
int ob200_diagnostics(ob200_handle* h, double maxabs[6], double* cfl_dt);
int ob200_get_timing(ob200_handle* h, ob200_timing* t);
int ob200_set_stream(ob200_handle* h, void* cuda_stream);
/* Options: "timing" (stage timing, above), "barotropic_graph" (default 1), "persistent_barotropic", "overlap",
 * "overlap_exchange" (tuning variants, default 0; bit-identical, measured not faster -- DESIGN.md), "transfer_x_halo" = n: ob200_set_field / ob200_get_field also copy n (<= 7, <= hx)
 * x-halo columns each side -- the checkpointer uses it for state whose halo columns are not recomputed by
 * update_state! (time-filtered diffusivities, barotropic averages; src/simulation_outputs.jl:36-40,45-70).      */
int ob200_set_option(ob200_handle* h, const char* name, int32_t value);

/* --- multi-GPU: x-slab ring over NCCL (replaces Oceananigans' MPI halo
 * exchange; experiments/run.jl:9-27, src/near_global_simulation.jl:52) ------ */
int ob200_comm_unique_id(void* id128);                       /* rank 0 */
int ob200_comm_init(ob200_handle* h, const void* id128);      /* all ranks */

#ifdef __cplusplus
}
#endif
#endif /* OCEAN_B200_H */
