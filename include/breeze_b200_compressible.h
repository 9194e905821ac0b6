/*
 * breeze_b200_compressible.h — C ABI of the second hot-path family of libbreeze_b200.so (SURVEY.md §8a last row, §8f rank 2):
 * one Wicker–Skamarock RK3 step with linearized acoustic substepping of
 * AtmosphereModel{<:CompressibleDynamics{<:SplitExplicitTimeDiscretization}} — the reference's
 *   time_step!(model::CompressibleAcousticModel, Δt)           src/TimeSteppers/acoustic_runge_kutta_3.jl:264-319
 *   acoustic_rk3_substep!                                       :172-208
 *   acoustic_rk3_substep_loop!                                  src/CompressibleEquations/acoustic_substepping.jl:1404-1590
 * for dry or vapour-laden air (microphysics = nothing: the moisture density ρqᵛ is transported by the acoustic-mean velocities, enters
 * the mixture EOS, the linearized PGF coefficient γᵐRᵐ and — through the total density — the buoyancy), WENO(order=5),
 * LiquidIcePotentialTemperature formulation, no closure, optional UpperSponge, (Periodic | Flat, Periodic | Flat, Bounded) uniform grids.
 *
 * Same conventions as breeze_b200.h (plain pointers, HOST arrays interior-only with x fastest, 0 / negative return
 * codes, bzc_last_error for text). The CPU oracle exports the same ABI with prefix orcc_.
 */
#ifndef BREEZE_B200_COMPRESSIBLE_H
#define BREEZE_B200_COMPRESSIBLE_H

#include "breeze_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* dynamics.reference_state: `nothing` or an ExnerReferenceState (src/CompressibleEquations/compressible_dynamics.jl:114-177,223-301) */
enum { BZC_REFERENCE_NONE = 0, BZC_REFERENCE_EXNER = 1 };
/* damping strategy (src/CompressibleEquations/time_discretizations.jl:134-247) */
enum { BZC_NO_DIVERGENCE_DAMPING = 0, BZC_THERMAL_DIVERGENCE_DAMPING = 1 };
/* UpperSponge ramp shapes (time_discretizations.jl:397-437): ramp(z, Lz, depth) with s = clamp((z - (Lz - depth))/depth, 0, 1) */
enum { BZC_SPONGE_NONE = 0, BZC_SPONGE_LINEAR_RAMP = 1, BZC_SPONGE_CUBIC_RAMP = 2, BZC_SPONGE_SIN2_RAMP = 3 };
/* substep distribution (acoustic_substepping.jl:476-508) */
enum { BZC_PROPORTIONAL_SUBSTEPS = 0, BZC_CONSTANT_SUBSTEP_SIZE = 1, BZC_MONOLITHIC_FIRST_STAGE = 2 };

/* Field selectors for bzc_get_field (interior → HOST). z-face fields (RHO_W, W, RHO_W_PERT, SLOW_RHO_W, AVG_W) have Nz+1 levels. */
enum { BZC_RHO = 0, BZC_RHO_U = 1, BZC_RHO_V = 2, BZC_RHO_W = 3, BZC_RHO_THETA = 4,        /* prognostic (ρᵈ, momentum, ρθ) */
       BZC_U = 5, BZC_V = 6, BZC_W = 7, BZC_THETA = 8, BZC_T = 9, BZC_P = 10,                /* update_state! diagnostics      */
       BZC_G_RHO = 11, BZC_G_RHO_U = 12, BZC_G_RHO_V = 13, BZC_G_RHO_W = 14, BZC_G_RHO_THETA = 15,   /* slow tendencies Gⁿ     */
       BZC_SLOW_RHO_W = 16,                                                                  /* Gˢρw (z-faces)                  */
       BZC_EXNER_L = 17, BZC_THETA_L = 18, BZC_GAMMA_R_L = 19,                               /* Πᴸ, θᴸ, γᵐRᵐᴸ                  */
       BZC_RHO_PERT = 20, BZC_RHO_THETA_PERT = 21, BZC_RHO_U_PERT = 22, BZC_RHO_V_PERT = 23, BZC_RHO_W_PERT = 24,
       BZC_AVG_U = 25, BZC_AVG_V = 26, BZC_AVG_W = 27,                                       /* time-averaged velocities        */
       BZC_RHO_QV = 28, BZC_QV = 29, BZC_TOTAL_RHO = 30, BZC_G_RHO_QV = 31,                  /* moisture: ρqᵛ, qᵛ, ρ = ρᵈ + ρqᵛ, Gⁿ.ρqᵛ */
       BZC_N_FIELDS = 32 };

/*
 * bzc_config — RectilinearGrid + ThermodynamicConstants (the `base` part; base.microphysics must be NONE, base.n_ranks 1)
 * + CompressibleDynamics(SplitExplicitTimeDiscretization(...); standard_pressure, surface_pressure,
 *   reference_potential_temperature) (compressible_dynamics.jl:114-177; time_discretizations.jl:540-588).
 * base.potential_temperature is the constant θᵣ of the ExnerReferenceState (288 K for reference_state = :auto);
 * bzc_set_reference_potential_temperature installs a θᵣ(z) profile instead.
 */
typedef struct bzc_config {
    bz_config base;
    int32_t reference_state;                        /* BZC_REFERENCE_*; default EXNER (`:auto` on a Bounded z) */
    int32_t substeps;                               /* N per Δt; 0 = adaptive from acoustic_cfl (`substeps = nothing`) */
    int32_t damping;                                /* BZC_*_DIVERGENCE_DAMPING; default THERMAL */
    int32_t substep_distribution;                   /* BZC_PROPORTIONAL_SUBSTEPS */
    int32_t apply_first_substep_pressure_gradient;  /* false */
    int32_t damp_vertical;                          /* false */
    double  acoustic_cfl;                           /* 0.5  */
    double  forward_weight;                         /* 0.65 */
    double  damping_coefficient;                    /* 0.1  */
    double  damping_length_scale;                   /* 0 = `nothing`: mesh-local min(Δx, Δy) */
    double  thermodynamic_tendency_factor;          /* 1 */
    double  vertical_momentum_tendency_factor;      /* 1 */
    /* UpperSponge(damping_rate, depth, ramp) (time_discretizations.jl:440-520): implicit Rayleigh damping of (ρw)′ below the lid, folded
     * into the column tridiagonal (acoustic_substepping.jl:584-603): |δτᵐ⁺| rate ramp(z) on the diagonal, |δτˢ⁻| rate ramp(z) (ρw)′ on the rhs */
    int32_t sponge;                                 /* BZC_SPONGE_*; default NONE (`sponge = nothing`) */
    int32_t reserved1;
    double  sponge_damping_rate;                    /* 0.2 (1/s) */
    double  sponge_depth;                           /* 5e3 (m)   */
    int32_t reserved[2];
} bzc_config;

typedef struct bzc_ctx bzc_ctx;

void        bzc_default_config(bzc_config* cfg);

/* AtmosphereModel(grid; dynamics = CompressibleDynamics(SplitExplicitTimeDiscretization(...))), incl. the
 * AcousticRungeKutta3 stepper and its AcousticSubstepper (acoustic_runge_kutta_3.jl:86-113, acoustic_substepping.jl:180-270)
 * and the ExnerReferenceState in discrete hydrostatic balance (src/Thermodynamics/reference_states.jl:611-672,718-812). */
int         bzc_create(const bzc_config* cfg, bzc_ctx** out);
void        bzc_destroy(bzc_ctx* ctx);
const char* bzc_last_error(const bzc_ctx* ctx);

/* CompressibleDynamics(...; reference_potential_temperature = θᵣ(z)): Nz cell-centre values; rebuilds pᵣ, ρᵣ, πᵣ. */
int bzc_set_reference_potential_temperature(bzc_ctx* ctx, const double* theta_r);
/* ExnerReferenceState fields at cell centres (Nz each, any pointer may be NULL); BZ_ERR_STATE when reference_state = nothing. */
int bzc_get_reference_state(bzc_ctx* ctx, double* pressure, double* density, double* exner);

/* set!(model; ρᵈ, ρu, ρv, ρw, ρθ, ρqᵛ) for the prognostic fields themselves (set_atmosphere_model.jl:198-360; NULL keeps a field;
 * `rho` is the DRY density, the host side splits a given total density as establish_densities! does, compressible_time_stepping.jl:89-137),
 * followed by update_state! (update_atmosphere_model_state.jl:41-68): total density, halos, velocities, θ, qᵛ and the joint T, p
 * diagnosis of compressible_time_stepping.jl:167-235; with moisture also the first scalar tendency from the seeded transport velocity
 * (maybe_prepare_first_time_step!, acoustic_runge_kutta_3.jl:240-256). rho_w has Nz+1 levels. A context that never receives rho_qv is dry. */
int bzc_set_state(bzc_ctx* ctx, const double* rho, const double* rho_u, const double* rho_v, const double* rho_w,
                  const double* rho_theta, const double* rho_qv);

/* time_step!(model::CompressibleAcousticModel, Δt) (acoustic_runge_kutta_3.jl:264-319); asynchronous w.r.t. the device. */
int bzc_time_step(bzc_ctx* ctx, double dt);
int bzc_time_steps(bzc_ctx* ctx, double dt, int n);

/* Finer hooks for per-hook parity tests.
 *  bzc_compute_slow_tendencies = prepare_acoustic_cache! + compute_slow_momentum_tendencies! + compute_slow_scalar_tendencies!
 *      + assemble_slow_vertical_momentum_tendency! on the current state (acoustic_runge_kutta_3.jl:181-188,
 *      acoustic_substep_helpers.jl:55-149, acoustic_substepping.jl:322-370,689-748); results via bzc_get_field.
 *  bzc_stage_substep_count_and_size = stage_substep_count_and_size (acoustic_substepping.jl:476-508).
 *  bzc_acoustic_substep_loop = acoustic_rk3_substep_loop!(model, substepper, Δt, β, U⁰) with U⁰ = the state at the last
 *      bzc_set_state / the start of the last bzc_time_step (acoustic_substepping.jl:1404-1590), after
 *      bzc_compute_slow_tendencies; followed by update_state! as in time_step! (acoustic_runge_kutta_3.jl:284-285). */
int bzc_compute_slow_tendencies(bzc_ctx* ctx);
int bzc_stage_substep_count_and_size(bzc_ctx* ctx, double dt, double beta, int32_t* n_tau, double* d_tau);
int bzc_acoustic_substep_loop(bzc_ctx* ctx, double dt, double beta);

int bzc_get_field(bzc_ctx* ctx, int field, double* host_out);
/* interior of the five prognostics → HOST in one call (any pointer may be NULL); synchronises once */
int bzc_get_state(bzc_ctx* ctx, double* rho, double* rho_u, double* rho_v, double* rho_w, double* rho_theta, double* rho_qv);
int bzc_get_clock(bzc_ctx* ctx, double* time, int64_t* iteration);
int bzc_synchronize(bzc_ctx* ctx);

/* Instrumentation (CUDA library only): event time per kernel family since the last read. Families: 0 slow tendencies
 * (WENO5), 1 stage setup (linearization, Gˢρw, perturbation init), 2 horizontal step + damping, 3 column kernel
 * (predictors + tridiagonal solve + recovery), 4 stage end (averages, recovery, update_state). */
int     bzc_profile_enable(bzc_ctx* ctx, int on);
int     bzc_profile_read(bzc_ctx* ctx, double* ms_per_family /* 8 */, int64_t* launches_per_family /* 8 */);
int64_t bzc_kernel_launch_count(const bzc_ctx* ctx);
void*   bzc_stream(bzc_ctx* ctx);
int64_t bzc_device_bytes(const bzc_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* BREEZE_B200_COMPRESSIBLE_H */
