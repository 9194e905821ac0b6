/*
 * breeze_b200.h — C ABI of libbreeze_b200.so, the B200-native replacement for ONE hot path of
 * NumericalEarth/Breeze.jl: the per-RK-stage tendency + pressure-correction step of
 * AtmosphereModel{<:AnelasticDynamics} with SSPRungeKutta3 and WENO(order=5).
 *
 * Breeze has no FFI: its extension mechanism is Julia multiple dispatch. Each entry point below names
 * the reference method it stands behind (paths relative to the reference repository); INTEGRATION.md
 * shows the `ccall` stubs a Breeze maintainer would add to dispatch to them.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types; no exception crosses the boundary;
 *  - every call returns BZ_OK (0) or a negative error code; bz_last_error() gives the text;
 *  - HOST arrays are interior-only, x fastest (Julia column-major `interior(field)`), Float64;
 *      centre fields  : Nx*Ny*Nz          index i + Nx*(j + Ny*k)
 *      rho_w (z-face) : Nx*Ny*(Nz+1)      faces k = 0..Nz, face 0 / Nz are the impenetrable walls
 *    With more than one rank, Nx in every HOST array is the rank-local slab width Nx/n_ranks;
 *  - the library never keeps a caller pointer after the call returns;
 *  - one host thread drives one context; work is queued on the context's CUDA stream(s) and the
 *    calls that copy to the host synchronise.
 *
 * The same ABI (prefix orc_ instead of bz_) is exported by the CPU oracle in oracle/ so that the parity
 * tests drive both through one harness. The oracle is test infrastructure, never a fallback.
 */
#ifndef BREEZE_B200_H
#define BREEZE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BZ_ABI_VERSION 1

enum { BZ_OK = 0, BZ_ERR_INVALID = -1, BZ_ERR_CUDA = -2, BZ_ERR_UNSUPPORTED = -3, BZ_ERR_NCCL = -4,
       BZ_ERR_NOMEM = -5, BZ_ERR_STATE = -6 };

/* Oceananigans topologies used by the path. z is always Bounded. */
enum { BZ_PERIODIC = 0, BZ_FLAT = 1 };

/* microphysics selector: `nothing` or SaturationAdjustment(equilibrium=WarmPhaseEquilibrium())
 * (src/Microphysics/saturation_adjustment.jl:23,55) */
enum { BZ_MICROPHYSICS_NONE = 0, BZ_MICROPHYSICS_WARM_SATURATION_ADJUSTMENT = 1 };

/* thermodynamic formulation: LiquidIcePotentialTemperatureFormulation (prognostic ρθ) or StaticEnergyFormulation (prognostic ρe,
 * e = cᵖᵐ T + g z - ℒˡqˡ - ℒⁱqⁱ; src/StaticEnergyFormulations/, src/Thermodynamics/dynamic_states.jl:283-312). With the static-energy
 * formulation the BZ_RHO_THETA slot carries ρe, BZ_THETA returns e, and microphysics must be NONE. */
enum { BZ_FORMULATION_POTENTIAL_TEMPERATURE = 0, BZ_FORMULATION_STATIC_ENERGY = 1 };

/* Field selectors for bz_get_field / bz_get_tendency. */
enum { BZ_RHO_U = 0, BZ_RHO_V = 1, BZ_RHO_W = 2, BZ_RHO_THETA = 3, BZ_RHO_Q = 4,   /* prognostic      */
       BZ_U = 5, BZ_V = 6, BZ_W = 7, BZ_THETA = 8, BZ_QV = 9, BZ_T = 10,            /* diagnostic      */
       BZ_PHI = 11,                                                                 /* p'/rho_r        */
       BZ_QL = 12 };                                                                /* cloud liquid    */

/*
 * bz_config — everything `RectilinearGrid(...)`, `ThermodynamicConstants()`, `ReferenceState(...)`,
 * `AnelasticDynamics(reference)` and `AtmosphereModel(grid; dynamics, advection=WENO(order=5))` fix at
 * construction time (src/AtmosphereModels/atmosphere_model.jl:114-314,
 * src/Thermodynamics/thermodynamics_constants.jl:182-212, src/Thermodynamics/reference_states.jl:402-445).
 * Fill with bz_default_config() first, then override.
 */
typedef struct bz_config {
    int32_t abi_version;          /* BZ_ABI_VERSION */
    int32_t Nx, Ny, Nz;           /* GLOBAL grid size; a Flat dimension has size 1 */
    int32_t topology_x, topology_y;
    double  x0, x1, y0, y1, z0, z1;   /* domain extents (uniform spacing) */

    /* ReferenceState(grid, constants; surface_pressure, potential_temperature, standard_pressure) */
    double surface_pressure;      /* 101325 */
    double potential_temperature; /* 288    */
    double standard_pressure;     /* 1e5    */

    /* ThermodynamicConstants() */
    double molar_gas_constant;          /* 8.314462618 */
    double gravitational_acceleration;  /* 9.81        */
    double energy_reference_temperature;/* 273.15      */
    double triple_point_temperature;    /* 273.16      */
    double triple_point_pressure;       /* 611.657     */
    double dry_air_molar_mass;          /* 0.02897     */
    double dry_air_heat_capacity;       /* 1005        */
    double vapor_molar_mass;            /* 0.018015    */
    double vapor_heat_capacity;         /* 1850        */
    double liquid_reference_latent_heat;/* 2500800     */
    double liquid_heat_capacity;        /* 4181        */
    double ice_reference_latent_heat;   /* 2834000     */
    double ice_heat_capacity;           /* 2108        */

    int32_t advection_order;      /* 5 (WENO(order=5)); the only supported value */
    int32_t microphysics;         /* BZ_MICROPHYSICS_* */

    /* x-slab decomposition: one context per rank / GPU. n_ranks == 1 needs nothing else. */
    int32_t n_ranks, rank;
    int32_t device;               /* CUDA device ordinal for this context (ignored by the oracle) */
    int32_t reserved0;
    uint8_t nccl_unique_id[128];  /* ncclUniqueId from rank 0, identical on all ranks (n_ranks > 1) */

    /* tuning knobs (0 = library default); never change results beyond FP64 round-off */
    int32_t use_tma;              /* stage kernel operand staging: 0 default, 1 TMA, 2 plain loads */
    int32_t z_chunks;             /* split the z march of the stage kernel into this many chunks   */
    int32_t formulation;          /* BZ_FORMULATION_* (AtmosphereModel(...; formulation = :StaticEnergy)); default 0 */
    int32_t reserved[5];
} bz_config;

/*
 * bz_forcing — the forcing / Coriolis / bottom flux-BC terms of an LES case such as BOMEX (examples/bomex.jl:77-196), all
 * zero in the bubble configurations. Pointers are HOST arrays read during the call (NULL = term absent):
 *   SubsidenceForcing(wˢ) on u, v, θ, qᵉ   F_ϕ = -ℑzb(wˢ ∂z ϕ̄), ϕ̄ = horizontal mean, recomputed every update_state!
 *                                          (src/Forcings/subsidence_forcing.jl:84-141); enters as ρ F_ϕ (specific_forcing.jl:70-74)
 *   geostrophic_forcings(uᵍ, vᵍ)            F_u = -f vᵍ, F_v = +f uᵍ (src/Forcings/geostrophic_forcings.jl:74-84), × ρ
 *   FPlane(f)                               -x_f_cross_U = +f ℑxy(ρv), -y_f_cross_U = -f ℑxy(ρu) (dynamics_kernel_functions.jl:79,99)
 *   Forcing(field) under `qᵉ` and `e`       ρ·q_tendency;  ρ·e_tendency / (cᵖᵐ Π) into ρθ (potential_temperature_tendency.jl:97-104)
 *   bottom FluxBoundaryConditions           G[i,j,1] += J/Δz (update_atmosphere_model_state.jl:418-434): constant J for ρθ, ρq;
 *                                           J = -ρ₀u★² ρu/|ρ𝐮ₕ| (and ρv) with the other component interpolated to the face.
 */
typedef struct bz_forcing {
    double coriolis_f;
    const double* subsidence_w;        /* Nz+1 values at z-faces */
    int32_t subsidence_mask;           /* bit 0: u, 1: v, 2: θ, 3: q */
    int32_t reserved0;
    const double* geostrophic_u;       /* Nz */
    const double* geostrophic_v;       /* Nz */
    const double* q_tendency;          /* Nz, specific (kg/kg/s) */
    const double* e_tendency;          /* Nz, specific energy tendency (J/kg/s) */
    double theta_flux;                 /* ρ₀ w'θ' */
    double q_flux;                     /* ρ₀ w'q' */
    double drag_rho_ustar2;            /* ρ₀ u★² */
} bz_forcing;

typedef struct bz_ctx bz_ctx;

void        bz_default_config(bz_config* cfg);
int         bz_abi_version(void);

/* AtmosphereModel(grid; dynamics=AnelasticDynamics(ReferenceState(grid; ...)), advection=WENO(order=5))
 * incl. TimeStepper(:SSPRungeKutta3, ...) (src/TimeSteppers/ssp_runge_kutta_3.jl:83-98) and
 * dynamics_pressure_solver (src/AnelasticEquations/anelastic_pressure_solver.jl:11-24).
 * Leaves θ = θ₀ (initialize_model_thermodynamics!, anelastic_time_stepping.jl:15-19). */
int         bz_create(const bz_config* cfg, bz_ctx** out);
void        bz_destroy(bz_ctx* ctx);
const char* bz_last_error(const bz_ctx* ctx);   /* ctx may be NULL: error of the last failed bz_create */

/* Reference-state profiles at cell centres, Nz values each (reference_states.jl:102-123,326-330).
 * Any pointer may be NULL. bz_set_reference_state overrides them the way the reference tests do
 * (`set!(reference_state.density, z -> z)`, test/anelastic_pressure_solver_analytic.jl:37) and rebuilds
 * the Poisson solver's diagonals. */
int bz_get_reference_state(bz_ctx* ctx, double* density, double* pressure, double* temperature);
int bz_set_reference_state(bz_ctx* ctx, const double* density, const double* pressure, const double* temperature);

/* set!(model; ρu, ρv, ρw, ρθ, ρqᵛ) (src/AtmosphereModels/set_atmosphere_model.jl:198-360): copies the HOST
 * arrays in (NULL keeps the current field), then update_state! → [enforce_mass_conservation: projection with
 * Δt = 1 → update_state!] so momentum is discretely divergence-free before the first step (:121-128,338,351). */
int bz_set_state(bz_ctx* ctx, const double* rho_u, const double* rho_v, const double* rho_w,
                 const double* rho_theta, const double* rho_q, int enforce_mass_conservation);

/* AtmosphereModel(grid; coriolis, forcing, boundary_conditions, ...): installs / replaces the terms above (NULL clears them). */
int bz_set_forcing(bz_ctx* ctx, const bz_forcing* forcing);

/* time_step!(model::AtmosphereModel{…,<:SSPRungeKutta3}, Δt) (src/TimeSteppers/ssp_runge_kutta_3.jl:209-278).
 * Asynchronous with respect to the device. bz_time_steps = many_time_steps! (benchmarking/src/timestepping.jl:11-16). */
int bz_time_step(bz_ctx* ctx, double dt);
int bz_time_steps(bz_ctx* ctx, double dt, int n);

/* Finer hooks, for per-hook parity tests.
 *  bz_compute_tendencies    = update_state!(model; compute_tendencies=true) → Gⁿ
 *                             (update_atmosphere_model_state.jl:41-68,294-387); fetch with bz_get_tendency.
 *  bz_pressure_correct      = compute_pressure_correction! + make_pressure_correction!
 *                             (src/AnelasticEquations/anelastic_time_stepping.jl:26-39,65-78) on the current momentum. */
int bz_compute_tendencies(bz_ctx* ctx);
int bz_get_tendency(bz_ctx* ctx, int field /* BZ_RHO_U..BZ_RHO_Q */, double* host_out);
int bz_pressure_correct(bz_ctx* ctx, double dt);

/* interior(field) → HOST; synchronises. rho_w / w: Nx*Ny*(Nz+1); everything else Nx*Ny*Nz. */
int bz_get_field(bz_ctx* ctx, int field, double* host_out);
int bz_get_state(bz_ctx* ctx, double* rho_u, double* rho_v, double* rho_w, double* rho_theta, double* rho_q);

/* Asynchronous forms of bz_set_state / bz_get_state for a caller that hands host buffers in and out every step: the copies are
 * strided 3-D transfers between the dense host arrays and the padded device fields, cut in z chunks and queued on two copy streams
 * (one per PCIe direction), so the download of step n and the upload of step n + 1 run full duplex; an upload from a buffer that a
 * pending download is still filling is ordered behind it chunk by chunk. The host buffers must stay valid (and should be page-locked)
 * until bz_synchronize; the synchronous forms above are these plus a wait. */
int bz_set_state_async(bz_ctx* ctx, const double* rho_u, const double* rho_v, const double* rho_w,
                       const double* rho_theta, const double* rho_q, int enforce_mass_conservation);
int bz_get_state_async(bz_ctx* ctx, double* rho_u, double* rho_v, double* rho_w, double* rho_theta, double* rho_q);

/* model.clock: time and iteration. */
int bz_get_clock(bz_ctx* ctx, double* time, int64_t* iteration);

/* cell_advection_timescale (src/AtmosphereModels/cell_advection_timescale.jl:46-65): min over cells of
 * 1/(|u|/Δx+|v|/Δy+|w|/Δz); what TimeStepWizard(cfl) multiplies. Global over ranks. */
int bz_cell_advection_timescale(bz_ctx* ctx, double* tau);

/* NaNChecker of `run!` (OceananigansDiagnostics.default_nan_checker(::AtmosphereModel), src/AtmosphereModels/atmosphere_model.jl:561-572):
 * *finite = 0 if any prognostic value is NaN or Inf (the reference checks the first prognostic field, ρu; a NaN spreads to the
 * others within one step, and the device reduction over all five costs the same launch). Global over ranks; synchronises. */
int bz_state_is_finite(bz_ctx* ctx, int* finite);

/* One 2-D slice of interior(field) → HOST without moving the whole field (what the examples' slice output writers save):
 * axis 0: x = index → Ny*Nz[+1] values (y fastest); axis 1: y = index → Nx*Nz[+1] (x fastest); axis 2: z = index → Nx*Ny (x fastest).
 * `index` is rank-local along x. Same field selectors as bz_get_field. */
int bz_get_slice(bz_ctx* ctx, int field, int axis, int index, double* host_out);

/* Discrete max |div(ρu)| over cells (the quantity test/anelastic_pressure_solver_nonhydrostatic.jl:40-46 bounds). */
int bz_max_abs_divergence(bz_ctx* ctx, double* out);

int bz_synchronize(bz_ctx* ctx);

/* Multi-GPU bootstrap: rank 0 obtains an ncclUniqueId (128 bytes) and hands it to the other ranks by any means
 * (torch.distributed broadcast in bench.py); every rank then puts it in bz_config.nccl_unique_id. */
int bz_nccl_unique_id(uint8_t* out128);

/* Peer memory over NVLink (optional, n_ranks > 1, one process per GPU): every rank exports ONE CUDA IPC handle of its
 * device arena (64 bytes); the rank-ordered concatenation of all handles (all_gather) is attached on every rank. From then
 * on ghost cells and the two transposes of the distributed FFT are peer LOADS issued by the consuming kernels, ordered by
 * a one-element NCCL all-reduce used as a stream barrier; without it the same exchanges run as NCCL send/recv. */
int bz_ipc_export(bz_ctx* ctx, uint8_t* out64);
int bz_ipc_attach(bz_ctx* ctx, const uint8_t* handles /* n_ranks * 64 bytes */);

/* Instrumentation for bench.py: CUDA-event time (ms) accumulated per kernel family since the last reset, and the
 * number of kernels this library launched. Families: 0 stage(tendency+RK), 1 Poisson forward (div+FFT), 2 Thomas,
 * 3 Poisson inverse, 4 projection(+halo), 5 halo exchange / transposes. Profiling adds event records only when enabled. */
int     bz_profile_enable(bz_ctx* ctx, int on);
int     bz_profile_read(bz_ctx* ctx, double* ms_per_family /* 8 */, int64_t* launches_per_family /* 8 */);
int64_t bz_kernel_launch_count(const bz_ctx* ctx);
/* raw CUDA stream handle (cudaStream_t) the context launches on, for event timing by the caller */
void*   bz_stream(bz_ctx* ctx);
/* bytes of device memory the context holds */
int64_t bz_device_bytes(const bz_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* BREEZE_B200_H */
