/*
 * breeze_b200_f32.h — C ABI of libbreeze_b200_f32.so: the Float32 build of the anelastic path of breeze_b200.h, the precision the
 * reference benchmarks in by default (`RectilinearGrid(GPU(), Float32; …)`, benchmarking/README.md:74).
 *
 * GENERATED from include/breeze_b200.h (the declarations below are that header's, retyped): the same entry points under the prefix
 * bzf_, every `double` array or scalar argument a `float` — host arrays are Julia `Array{Float32,3}` interiors, x fastest — except the
 * clock of bzf_get_clock, which stays double. The configuration and forcing structs (bz_config, bz_forcing) are shared with
 * breeze_b200.h and keep their double fields; bz_ctx is opaque in both. The library is compiled from a mechanically retyped copy of the
 * FP64 sources (breeze.jl_b200/make_f32.py); the compressible path's Float32 entry points (bzcf_*) are in breeze_b200_compressible_f32.h.
 * Each entry point stands behind the same reference method as its bz_ namesake (see breeze_b200.h for the file:line citations).
 */
#ifndef BREEZE_B200_F32_H
#define BREEZE_B200_F32_H

#include "breeze_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

void        bzf_default_config(bz_config* cfg);
int         bzf_abi_version(void);

/* AtmosphereModel(grid; dynamics=AnelasticDynamics(ReferenceState(grid; ...)), advection=WENO(order=5))
 * incl. TimeStepper(:SSPRungeKutta3, ...) (src/TimeSteppers/ssp_runge_kutta_3.jl:83-98) and
 * dynamics_pressure_solver (src/AnelasticEquations/anelastic_pressure_solver.jl:11-24).
 * Leaves θ = θ₀ (initialize_model_thermodynamics!, anelastic_time_stepping.jl:15-19). */
int         bzf_create(const bz_config* cfg, bz_ctx** out);
void        bzf_destroy(bz_ctx* ctx);
const char* bzf_last_error(const bz_ctx* ctx);   /* ctx may be NULL: error of the last failed bz_create */

/* Reference-state profiles at cell centres, Nz values each (reference_states.jl:102-123,326-330).
 * Any pointer may be NULL. bz_set_reference_state overrides them the way the reference tests do
 * (`set!(reference_state.density, z -> z)`, test/anelastic_pressure_solver_analytic.jl:37) and rebuilds
 * the Poisson solver's diagonals. */
int bzf_get_reference_state(bz_ctx* ctx, float* density, float* pressure, float* temperature);
int bzf_set_reference_state(bz_ctx* ctx, const float* density, const float* pressure, const float* temperature);

/* set!(model; ρu, ρv, ρw, ρθ, ρqᵛ) (src/AtmosphereModels/set_atmosphere_model.jl:198-360): copies the HOST
 * arrays in (NULL keeps the current field), then update_state! → [enforce_mass_conservation: projection with
 * Δt = 1 → update_state!] so momentum is discretely divergence-free before the first step (:121-128,338,351). */
int bzf_set_state(bz_ctx* ctx, const float* rho_u, const float* rho_v, const float* rho_w,
                 const float* rho_theta, const float* rho_q, int enforce_mass_conservation);

/* AtmosphereModel(grid; coriolis, forcing, boundary_conditions, ...): installs / replaces the terms above (NULL clears them). */
int bzf_set_forcing(bz_ctx* ctx, const bz_forcing* forcing);

/* time_step!(model::AtmosphereModel{…,<:SSPRungeKutta3}, Δt) (src/TimeSteppers/ssp_runge_kutta_3.jl:209-278).
 * Asynchronous with respect to the device. bz_time_steps = many_time_steps! (benchmarking/src/timestepping.jl:11-16). */
int bzf_time_step(bz_ctx* ctx, float dt);
int bzf_time_steps(bz_ctx* ctx, float dt, int n);

/* Finer hooks, for per-hook parity tests.
 *  bz_compute_tendencies    = update_state!(model; compute_tendencies=true) → Gⁿ
 *                             (update_atmosphere_model_state.jl:41-68,294-387); fetch with bz_get_tendency.
 *  bz_pressure_correct      = compute_pressure_correction! + make_pressure_correction!
 *                             (src/AnelasticEquations/anelastic_time_stepping.jl:26-39,65-78) on the current momentum. */
int bzf_compute_tendencies(bz_ctx* ctx);
int bzf_get_tendency(bz_ctx* ctx, int field /* BZ_RHO_U..BZ_RHO_Q */, float* host_out);
int bzf_pressure_correct(bz_ctx* ctx, float dt);

/* interior(field) → HOST; synchronises. rho_w / w: Nx*Ny*(Nz+1); everything else Nx*Ny*Nz. */
int bzf_get_field(bz_ctx* ctx, int field, float* host_out);
int bzf_get_state(bz_ctx* ctx, float* rho_u, float* rho_v, float* rho_w, float* rho_theta, float* rho_q);

/* Asynchronous forms of bz_set_state / bz_get_state for a caller that hands host buffers in and out every step: the copies are
 * strided 3-D transfers between the dense host arrays and the padded device fields, cut in z chunks and queued on two copy streams
 * (one per PCIe direction), so the download of step n and the upload of step n + 1 run full duplex; an upload from a buffer that a
 * pending download is still filling is ordered behind it chunk by chunk. The host buffers must stay valid (and should be page-locked)
 * until bz_synchronize; the synchronous forms above are these plus a wait. */
int bzf_set_state_async(bz_ctx* ctx, const float* rho_u, const float* rho_v, const float* rho_w,
                       const float* rho_theta, const float* rho_q, int enforce_mass_conservation);
int bzf_get_state_async(bz_ctx* ctx, float* rho_u, float* rho_v, float* rho_w, float* rho_theta, float* rho_q);

/* model.clock: time and iteration. */
int bzf_get_clock(bz_ctx* ctx, double* time, int64_t* iteration);

/* cell_advection_timescale (src/AtmosphereModels/cell_advection_timescale.jl:46-65): min over cells of
 * 1/(|u|/Δx+|v|/Δy+|w|/Δz); what TimeStepWizard(cfl) multiplies. Global over ranks. */
int bzf_cell_advection_timescale(bz_ctx* ctx, float* tau);

/* NaNChecker of `run!` (OceananigansDiagnostics.default_nan_checker(::AtmosphereModel), src/AtmosphereModels/atmosphere_model.jl:561-572):
 * *finite = 0 if any prognostic value is NaN or Inf (the reference checks the first prognostic field, ρu; a NaN spreads to the
 * others within one step, and the device reduction over all five costs the same launch). Global over ranks; synchronises. */
int bzf_state_is_finite(bz_ctx* ctx, int* finite);

/* One 2-D slice of interior(field) → HOST without moving the whole field (what the examples' slice output writers save):
 * axis 0: x = index → Ny*Nz[+1] values (y fastest); axis 1: y = index → Nx*Nz[+1] (x fastest); axis 2: z = index → Nx*Ny (x fastest).
 * `index` is rank-local along x. Same field selectors as bz_get_field. */
int bzf_get_slice(bz_ctx* ctx, int field, int axis, int index, float* host_out);

/* Discrete max |div(ρu)| over cells (the quantity test/anelastic_pressure_solver_nonhydrostatic.jl:40-46 bounds). */
int bzf_max_abs_divergence(bz_ctx* ctx, float* out);

int bzf_synchronize(bz_ctx* ctx);

/* Multi-GPU bootstrap: rank 0 obtains an ncclUniqueId (128 bytes) and hands it to the other ranks by any means
 * (torch.distributed broadcast in bench.py); every rank then puts it in bz_config.nccl_unique_id. */
int bzf_nccl_unique_id(uint8_t* out128);

/* Peer memory over NVLink (optional, n_ranks > 1, one process per GPU): every rank exports ONE CUDA IPC handle of its
 * device arena (64 bytes); the rank-ordered concatenation of all handles (all_gather) is attached on every rank. From then
 * on ghost cells and the two transposes of the distributed FFT are peer LOADS issued by the consuming kernels, ordered by
 * a one-element NCCL all-reduce used as a stream barrier; without it the same exchanges run as NCCL send/recv. */
int bzf_ipc_export(bz_ctx* ctx, uint8_t* out64);
int bzf_ipc_attach(bz_ctx* ctx, const uint8_t* handles /* n_ranks * 64 bytes */);

/* Instrumentation for bench.py: CUDA-event time (ms) accumulated per kernel family since the last reset, and the
 * number of kernels this library launched. Families: 0 stage(tendency+RK), 1 Poisson forward (div+FFT), 2 Thomas,
 * 3 Poisson inverse, 4 projection(+halo), 5 halo exchange / transposes. Profiling adds event records only when enabled. */
int     bzf_profile_enable(bz_ctx* ctx, int on);
int     bzf_profile_read(bz_ctx* ctx, float* ms_per_family /* 8 */, int64_t* launches_per_family /* 8 */);
int64_t bzf_kernel_launch_count(const bz_ctx* ctx);
/* raw CUDA stream handle (cudaStream_t) the context launches on, for event timing by the caller */
void*   bzf_stream(bz_ctx* ctx);
/* bytes of device memory the context holds */
int64_t bzf_device_bytes(const bz_ctx* ctx);


#ifdef __cplusplus
}
#endif
#endif /* BREEZE_B200_F32_H */
