/*
 * breeze_oracle_compressible.c — CPU ORACLE (test infrastructure, NOT a product path, NOT a fallback) for the second
 * hot-path family: one Wicker–Skamarock RK3 step with linearized acoustic substepping of
 * AtmosphereModel{<:CompressibleDynamics{<:SplitExplicitTimeDiscretization}} (dry air, WENO(order=5), no closure).
 *
 * Plain C11 + OpenMP, FP64, written the way the reference runs it: halo-padded fields, one loop nest per reference
 * kernel, the same order of launches and halo fills per substep. Every function cites the reference lines it follows
 * (paths relative to the reference repository). Exports the ABI of include/breeze_b200_compressible.h with prefix orcc_.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it.
 *
 * PARITY STATUS
 *   - The acoustic substepper (linearization, slow vertical tendency, kernels A–E, tridiagonal coefficients, stage
 *     rewind/recovery, substep counts), the WS-RK3 driver, the ExnerReferenceState and the compressible θ→(T, p)
 *     inversion are restated from files under /root/reference and pinned by the reference's own known-answer tests
 *     (test/acoustic_substepping_components.jl, test/substepper_rest_state.jl, test/substepper_structural.jl →
 *     tests/test_oracle_compressible.py).
 *   - The WENO5 slow advection tendencies, difference/interpolation operators, halo fills and the BatchedTridiagonalSolver
 *     sweep are Oceananigans.jl 0.110.14 code (not vendored, cannot run here): restated from the published algorithm
 *      — for those pieces **parity unpinned** (shared with breeze_oracle.c).
 *
 * Index conventions: 0-based; reference 1-based index = C index + 1. z-face k is the bottom face of cell k (0 and Nz
 * are the walls), so the reference's `(k > 1)` masks read `k > 0` here and tridiagonal row k is face k, k = 0..Nz-1.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stddef.h>
#include <stdarg.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/breeze_b200_compressible.h"
#include "oracle_weno.h"

#define HALO(order) (((order) + 1) / 2 + 1)   /* buffer + 1: 4 for WENO5, 6 for WENO9; results do not depend on a larger halo */
enum { C_RHO = 0, C_RU = 1, C_RV = 2, C_RW = 3, C_RTH = 4, NPROGC = 5 };
enum { LOC_CENTER = 0, LOC_ZFACE = 1 };

typedef struct orcc_ctx {
    bzc_config cfg;
    int Nx, Ny, Nz, Hx, Hy, Hz, Px, Py, Pz;
    int B, Bs;                     /* buffers of the biased (WENO) and symmetric (Centered) reconstructions: (order + 1) / 2, B - 1 */
    size_t n_padded;
    int flat_x, flat_y;
    double dx, dy, dz;
    double Rd, Rv, cpd, cpv, g, pst, p0;
    int has_ref;
    double *p_r, *rho_r, *pi_r, *theta_r;        /* z-only, index k + Hz */
    double* U[NPROGC];                            /* ρᵈ, ρu, ρv, ρw, ρθ */
    double* U0[NPROGC];
    double* G[NPROGC];                            /* Gⁿ.ρᵈ, ρu, ρv, ρw, ρθ */
    double *u, *v, *w, *theta, *T, *p;
    /* moisture (vapour only, microphysics = nothing): prognostic ρqᵛ, its U⁰ and Gⁿ, qᵛ = ρqᵛ/ρ, total density ρ = ρᵈ + ρqᵛ */
    int moist;
    double *rqv, *rqv0, *Grqv, *qv, *rho_tot;
    /* AcousticSubstepper fields (acoustic_substepping.jl:91-134) */
    double *PiL, *thL, *gRL;
    double *rho_p, *rth_p, *ru_p, *rv_p, *rw_p;
    double *rho_s, *rth_s, *rth_old;
    double *avg_u, *avg_v, *avg_w;
    double *Gs_rw, *rhs, *scratch;
    double time; int64_t iteration;
    char err[256];
} orcc_ctx;

static char g_err[256];
static void set_err(orcc_ctx* c, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(c ? c->err : g_err, 256, fmt, ap);
    va_end(ap);
}

#define IDX(c, i, j, k) ((size_t)((i) + (c)->Hx) + (size_t)(c)->Px * ((size_t)((j) + (c)->Hy) + (size_t)(c)->Py * (size_t)((k) + (c)->Hz)))
#define SX ((ptrdiff_t)1)
#define SY ((ptrdiff_t)c->Px)
#define SZ ((ptrdiff_t)c->Px * c->Py)
#define FOR_CELLS(kmax) _Pragma("omp parallel for collapse(2) schedule(static)") \
    for (int k = 0; k < (kmax); ++k) for (int j = 0; j < c->Ny; ++j) for (int i = 0; i < c->Nx; ++i)

void orc_default_config(bz_config* c);

/* SplitExplicitTimeDiscretization() defaults: time_discretizations.jl:540-588; CompressibleDynamics(): compressible_dynamics.jl:114-124 */
void orcc_default_config(bzc_config* c) {
    memset(c, 0, sizeof(*c));
    orc_default_config(&c->base);
    c->reference_state = BZC_REFERENCE_EXNER;
    c->substeps = 0;
    c->damping = BZC_THERMAL_DIVERGENCE_DAMPING;
    c->substep_distribution = BZC_PROPORTIONAL_SUBSTEPS;
    c->apply_first_substep_pressure_gradient = 0;
    c->damp_vertical = 0;
    c->acoustic_cfl = 0.5;
    c->forward_weight = 0.65;
    c->damping_coefficient = 0.1;
    c->damping_length_scale = 0.0;
    c->thermodynamic_tendency_factor = 1.0;
    c->vertical_momentum_tendency_factor = 1.0;
    c->sponge = BZC_SPONGE_NONE; c->sponge_damping_rate = 0.2; c->sponge_depth = 5e3;
}

static double* new_field(orcc_ctx* c) { return (double*)calloc(c->n_padded, sizeof(double)); }

/* ------------------------------------------------------------------------------------------------ */
/* halo fills (Oceananigans fill_halo_regions!): periodic x/y, zero-flux mirror in  */
/* Bounded z for centre fields, impenetrable wall faces (= 0) for the z-face fields                     */
/* ------------------------------------------------------------------------------------------------ */
static void fill_halos(const orcc_ctx* c, double* f, int loc) {
    const int Nx = c->Nx, Ny = c->Ny, Nz = c->Nz, Hx = c->Hx, Hy = c->Hy, Hz = c->Hz;
    if (loc == LOC_ZFACE)
        for (int j = 0; j < Ny; ++j) for (int i = 0; i < Nx; ++i) { f[IDX(c, i, j, 0)] = 0; f[IDX(c, i, j, Nz)] = 0; }
    for (int h = 1; h <= Hz; ++h)
        for (int j = 0; j < Ny; ++j) for (int i = 0; i < Nx; ++i) {
            if (loc == LOC_CENTER) {
                f[IDX(c, i, j, -h)] = f[IDX(c, i, j, h - 1)];
                f[IDX(c, i, j, Nz - 1 + h)] = f[IDX(c, i, j, Nz - h)];
            } else {
                f[IDX(c, i, j, -h)] = f[IDX(c, i, j, h)];
                f[IDX(c, i, j, Nz + h)] = f[IDX(c, i, j, Nz - h)];
            }
        }
    for (int k = -Hz; k <= Nz + Hz; ++k) {
        if (Hx) for (int j = 0; j < Ny; ++j) for (int h = 1; h <= Hx; ++h) {
            f[IDX(c, -h, j, k)] = f[IDX(c, Nx - h, j, k)];
            f[IDX(c, Nx - 1 + h, j, k)] = f[IDX(c, h - 1, j, k)];
        }
        if (Hy) for (int h = 1; h <= Hy; ++h) for (int i = -Hx; i < Nx + Hx; ++i) {
            f[IDX(c, i, -h, k)] = f[IDX(c, i, Ny - h, k)];
            f[IDX(c, i, Ny - 1 + h, k)] = f[IDX(c, i, h - 1, k)];
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* ExnerReferenceState, isentropic path: reference_states.jl:572-672 (dry: qᵛ ≡ 0 ⇒ Rᵐ = Rᵈ, cᵖᵐ = cᵖᵈ) */
/* ------------------------------------------------------------------------------------------------ */
static void build_exner_reference(orcc_ctx* c) {
    const int Nz = c->Nz, Hz = c->Hz;
    const double Rm = (1 - 0.0) * c->Rd + 0.0 * c->Rv, cpm = (1 - 0.0) * c->cpd + 0.0 * c->cpv, kap = Rm / cpm;
    const double g = c->g, pst = c->pst, p0 = c->p0;
    double* th = c->theta_r + Hz; double* pi = c->pi_r + Hz; double* p = c->p_r + Hz; double* rho = c->rho_r + Hz;
    /* _compute_exner_column!: anchor half a cell above the surface with the continuous Π recurrence (:611-631) */
    double pi_surface = pow(p0 / pst, kap);
    double Pi1 = pi_surface - g * c->dz / (2 * cpm * th[0]);
    double p1 = pst * pow(Pi1, 1 / kap);
    pi[0] = Pi1; p[0] = p1; rho[0] = p1 / (Rm * th[0] * Pi1);
    /* integrate_exner_column!: discrete balance (p_k - p_{k-1})/Δz + g (ρ_k + ρ_{k-1})/2 = 0 by Newton, FixedIterations(5) (:640-672,588-598) */
    double pm = p[0], rm = rho[0];
    for (int k = 1; k < Nz; ++k) {
        double dzf = c->dz;
        double th_face = (th[k] + th[k - 1]) / 2;
        double Pi_init = pi[k - 1] - g * dzf / (cpm * th_face);
        double pk = pst * pow(Pi_init, 1 / kap);
        double A = g * pow(pst, kap) / (2 * Rm * th[k]);
        double Cc = pm / dzf - g * rm / 2;
        for (int it = 0; it < 5; ++it) {
            double rp = pow(pk, -kap);
            double f = pk / dzf + A * pk * rp - Cc;
            double fp = 1 / dzf + A * (1 - kap) * rp;
            pk -= f / fp;
        }
        double Pik = pow(pk / pst, kap);
        double rk = pk / (Rm * th[k] * Pik);
        pi[k] = Pik; p[k] = pk; rho[k] = rk;
        pm = pk; rm = rk;
    }
    /* z-only halos: mirror (only ever read under a `k > 1` mask) */
    double* cols[4] = {c->theta_r, c->pi_r, c->p_r, c->rho_r};
    for (int f = 0; f < 4; ++f)
        for (int h = 1; h <= Hz; ++h) { cols[f][Hz - h] = cols[f][Hz + h - 1]; cols[f][Hz + Nz - 1 + h] = cols[f][Hz + Nz - h]; }
}

/* ------------------------------------------------------------------------------------------------ */
/* update_state!(model; compute_tendencies=false): update_atmosphere_model_state.jl:41-68, compressible   */
/* ------------------------------------------------------------------------------------------------ */

/* temperature(::LiquidIceDensityState) + p = ρ Rᵐ T: dynamic_states.jl:201-232 (NewtonSolver(reltol=0, abstol=1e-4, maxiter=8),
 * Solvers.jl:82), compressible_time_stepping.jl:215-235. Dry air: L = 0. */
static inline void temperature_and_pressure(const orcc_ctx* c, double rho, double theta, double qv, double* T_out, double* p_out) {
    /* mixture_gas_constant / mixture_heat_capacity of MoistureMassFractions(qᵛ, 0, 0) (thermodynamics_constants.jl:341-377) */
    const double qd = 1 - qv - 0.0 - 0.0;
    const double Rm = qd * c->Rd + qv * c->Rv, cpm = qd * c->cpd + qv * c->cpv + 0.0 + 0.0;
    const double kap = Rm / cpm, gam = cpm / (cpm - Rm), L = 0.0, pst = c->pst;
    double T = pow(theta, gam) * pow(rho * Rm / pst, gam - 1) + L;
    double dT = T; int iter = 0;
    while (fabs(dT) > fmax(1e-4, 0.0 * T) && iter < 8) {
        double Phi = pow(rho * Rm * T / pst, kap) * theta;
        dT = -(T - Phi - L) / (1 - kap * Phi / T);
        T += dT;
        ++iter;
    }
    *T_out = T; *p_out = rho * Rm * T;
}

/* compute_velocities!: update_atmosphere_model_state.jl:122-155,248-254 with dynamics_density = ρᵈ (3-D) */
static void compute_velocities(orcc_ctx* c) {
    fill_halos(c, c->U[C_RHO], LOC_CENTER);
    fill_halos(c, c->U[C_RU], LOC_CENTER);
    fill_halos(c, c->U[C_RV], LOC_CENTER);
    fill_halos(c, c->U[C_RW], LOC_ZFACE);
    const double* rho = c->U[C_RHO];
    FOR_CELLS(c->Nz + 1) {
        size_t n = IDX(c, i, j, k);
        if (k < c->Nz) {
            double rx = c->flat_x ? rho[n] : (rho[n] + rho[n - SX]) / 2;
            double ry = c->flat_y ? rho[n] : (rho[n] + rho[n - SY]) / 2;
            c->u[n] = c->U[C_RU][n] / rx;
            c->v[n] = c->U[C_RV][n] / ry;
        }
        c->w[n] = c->U[C_RW][n] / ((rho[n] + rho[n - SZ]) / 2);
    }
    fill_halos(c, c->u, LOC_CENTER);
    fill_halos(c, c->v, LOC_CENTER);
    fill_halos(c, c->w, LOC_ZFACE);
}

static void update_state(orcc_ctx* c) {
    /* compute_total_density!: ρ = ρᵈ + ρqᵛ (compressible_time_stepping.jl:49-67, microphysics_interface.jl:635-659) */
    FOR_CELLS(c->Nz) { size_t n = IDX(c, i, j, k); c->rho_tot[n] = c->U[C_RHO][n] + c->rqv[n]; }
    fill_halos(c, c->rho_tot, LOC_CENTER);
    fill_halos(c, c->U[C_RTH], LOC_CENTER);
    fill_halos(c, c->rqv, LOC_CENTER);
    compute_velocities(c);
    /* _compute_auxiliary_thermodynamic_variables! (θ = ρθ/ρᵈ, potential_temperature_formulation.jl) and
     * _compute_temperature_and_pressure! (compressible_time_stepping.jl:191-213) */
    FOR_CELLS(c->Nz) {
        size_t n = IDX(c, i, j, k);
        double rho = c->U[C_RHO][n];                 /* coupling density ρᵈ: θ = ρθ/ρᵈ */
        double th = c->U[C_RTH][n] / rho;
        c->theta[n] = th;
        double qv = c->rqv[n] / c->rho_tot[n];       /* mass fraction of the TOTAL density */
        c->qv[n] = qv;
        temperature_and_pressure(c, c->rho_tot[n], th, qv, &c->T[n], &c->p[n]);
    }
    fill_halos(c, c->qv, LOC_CENTER);
    fill_halos(c, c->theta, LOC_CENTER);
    fill_halos(c, c->T, LOC_CENTER);
    fill_halos(c, c->p, LOC_CENTER);
}

/* ------------------------------------------------------------------------------------------------ */
/* stage-entry linearization: acoustic_substepping.jl:322-415                                          */
/* ------------------------------------------------------------------------------------------------ */
static void refresh_linearization_basic_state(orcc_ctx* c) {
    const double kap = c->Rd / c->cpd;
    FOR_CELLS(c->Nz) {
        size_t n = IDX(c, i, j, k);
        c->PiL[n] = pow(c->p[n] / c->pst, kap);
        double rho = c->U[C_RHO][n];
        double rh = (rho == 0) ? 1.0 : rho;
        c->thL[n] = c->U[C_RTH][n] / rh;
        /* _compute_linearization_mixture_eos! with q = (qᵛ, 0, 0) */
        double qvl = c->qv[n];
        double qd = 1 - qvl - 0.0 - 0.0;
        double Rm = qd * c->Rd + qvl * c->Rv;
        double cpm = qd * c->cpd + qvl * c->cpv + 0.0 + 0.0;
        double cvm = cpm - Rm;
        c->gRL[n] = cpm * Rm / cvm;
    }
    fill_halos(c, c->PiL, LOC_CENTER);
    fill_halos(c, c->thL, LOC_CENTER);
    fill_halos(c, c->gRL, LOC_CENTER);
}

/* ------------------------------------------------------------------------------------------------ */
/* slow tendencies: acoustic_substep_helpers.jl:55-149 (SlowTendencyMode: PGF and buoyancy zeroed,     */
/* dynamics_interface.jl:397-411) → dynamics_kernel_functions.jl:64-130, compressible_density_tendency.jl:39-57, */
/* potential_temperature_tendency.jl:66-106 with ρ_field = ρᵈ, src/Advection.jl:20-35                  */
/* ------------------------------------------------------------------------------------------------ */
#define SYM_X(a, R) (c->flat_x ? (a)[0] : symmetric_interp((a), SX, (R)))
#define SYM_Y(a, R) (c->flat_y ? (a)[0] : symmetric_interp((a), SY, (R)))

static inline double flux_Uu(const orcc_ctx* c, int i, int j, int k) {
    size_t n1 = IDX(c, i + 1, j, k);
    double ut = c->dy * c->dz * symmetric_interp(c->U[C_RU] + n1, SX, c->Bs);
    return ut * biased_interp(c->u + n1, SX, c->B, ut > 0);
}
static inline double flux_Vu(const orcc_ctx* c, int i, int j, int k) {
    size_t n = IDX(c, i, j, k);
    double vt = c->dx * c->dz * SYM_X(c->U[C_RV] + n, c->Bs);
    return vt * biased_interp(c->u + n, SY, c->B, vt > 0);
}
static inline double flux_Wu(const orcc_ctx* c, int i, int j, int k) {
    if (k == 0 || k == c->Nz) return 0.0;
    size_t n = IDX(c, i, j, k);
    double wt = c->dx * c->dy * SYM_X(c->U[C_RW] + n, c->Bs);
    return wt * biased_interp(c->u + n, SZ, red_face(k, c->Nz, c->B), wt > 0);
}
static inline double flux_Uv(const orcc_ctx* c, int i, int j, int k) {
    size_t n = IDX(c, i, j, k);
    double ut = c->dy * c->dz * SYM_Y(c->U[C_RU] + n, c->Bs);
    return ut * biased_interp(c->v + n, SX, c->B, ut > 0);
}
static inline double flux_Vv(const orcc_ctx* c, int i, int j, int k) {
    size_t n1 = IDX(c, i, j + 1, k);
    double vt = c->dx * c->dz * symmetric_interp(c->U[C_RV] + n1, SY, c->Bs);
    return vt * biased_interp(c->v + n1, SY, c->B, vt > 0);
}
static inline double flux_Wv(const orcc_ctx* c, int i, int j, int k) {
    if (k == 0 || k == c->Nz) return 0.0;
    size_t n = IDX(c, i, j, k);
    double wt = c->dx * c->dy * SYM_Y(c->U[C_RW] + n, c->Bs);
    return wt * biased_interp(c->v + n, SZ, red_face(k, c->Nz, c->B), wt > 0);
}
static inline double flux_Uw(const orcc_ctx* c, int i, int j, int k) {
    size_t n = IDX(c, i, j, k);
    double ut = c->dy * c->dz * symmetric_interp(c->U[C_RU] + n, SZ, red_face(k, c->Nz, c->Bs));
    return ut * biased_interp(c->w + n, SX, c->B, ut > 0);
}
static inline double flux_Vw(const orcc_ctx* c, int i, int j, int k) {
    size_t n = IDX(c, i, j, k);
    double vt = c->dx * c->dz * symmetric_interp(c->U[C_RV] + n, SZ, red_face(k, c->Nz, c->Bs));
    return vt * biased_interp(c->w + n, SY, c->B, vt > 0);
}
static inline double flux_Ww(const orcc_ctx* c, int i, int j, int k) {
    size_t n1 = IDX(c, i, j, k + 1);
    double wt = c->dx * c->dy * symmetric_interp(c->U[C_RW] + n1, SZ, red_center(k, c->Nz, c->Bs));
    return wt * biased_interp(c->w + n1, SZ, red_center(k, c->Nz, c->B), wt > 0);
}
/* tracer_mass_flux_{x,y,z} with the 3-D coupling density ρᵈ */
static inline double tracer_flux_x(const orcc_ctx* c, const double* f, int i, int j, int k) {
    size_t n = IDX(c, i, j, k);
    const double* rho = c->U[C_RHO];
    double ut = c->u[n];
    return ((rho[n] + rho[n - SX]) / 2) * (c->dy * c->dz * ut * biased_interp(f + n, SX, c->B, ut > 0));
}
static inline double tracer_flux_y(const orcc_ctx* c, const double* f, int i, int j, int k) {
    size_t n = IDX(c, i, j, k);
    const double* rho = c->U[C_RHO];
    double vt = c->v[n];
    return ((rho[n] + rho[n - SY]) / 2) * (c->dx * c->dz * vt * biased_interp(f + n, SY, c->B, vt > 0));
}
static inline double tracer_flux_z(const orcc_ctx* c, const double* f, int i, int j, int k) {
    if (k == 0 || k == c->Nz) return 0.0;
    size_t n = IDX(c, i, j, k);
    const double* rho = c->U[C_RHO];
    double wt = c->w[n];
    return ((rho[n] + rho[n - SZ]) / 2) * (c->dx * c->dy * wt * biased_interp(f + n, SZ, red_face(k, c->Nz, c->B), wt > 0));
}

static void compute_slow_tendencies(orcc_ctx* c) {
    const int fx_ = c->flat_x, fy_ = c->flat_y;
    const double V = c->dx * c->dy * c->dz;
    const double Ax = c->dy * c->dz, Ay = c->dx * c->dz, Az = c->dx * c->dy;
    FOR_CELLS(c->Nz) {
        size_t n = IDX(c, i, j, k);
        {
            double fx = fx_ ? 0.0 : flux_Uu(c, i, j, k) - flux_Uu(c, i - 1, j, k);
            double fy = fy_ ? 0.0 : flux_Vu(c, i, j + 1, k) - flux_Vu(c, i, j, k);
            double fz = flux_Wu(c, i, j, k + 1) - flux_Wu(c, i, j, k);
            c->G[C_RU][n] = -((1 / V) * (fx + fy + fz));
        }
        {
            double fx = fx_ ? 0.0 : flux_Uv(c, i + 1, j, k) - flux_Uv(c, i, j, k);
            double fy = fy_ ? 0.0 : flux_Vv(c, i, j, k) - flux_Vv(c, i, j - 1, k);
            double fz = flux_Wv(c, i, j, k + 1) - flux_Wv(c, i, j, k);
            c->G[C_RV][n] = -((1 / V) * (fx + fy + fz));
        }
        if (k >= 1) {
            double fx = fx_ ? 0.0 : flux_Uw(c, i + 1, j, k) - flux_Uw(c, i, j, k);
            double fy = fy_ ? 0.0 : flux_Vw(c, i, j + 1, k) - flux_Vw(c, i, j, k);
            double fz = flux_Ww(c, i, j, k) - flux_Ww(c, i, j, k - 1);
            c->G[C_RW][n] = -((1 / V) * (fx + fy + fz));
        } else {
            c->G[C_RW][n] = 0.0;          /* wall face: masked by (k > 1) in Gˢρw and in the vertical rhs */
        }
        {   /* Gⁿ.ρᵈ = -divᶜᶜᶜ(ρu, ρv, ρw) */
            double dxu = fx_ ? 0.0 : Ax * c->U[C_RU][n + SX] - Ax * c->U[C_RU][n];
            double dyv = fy_ ? 0.0 : Ay * c->U[C_RV][n + SY] - Ay * c->U[C_RV][n];
            double dzw = Az * c->U[C_RW][n + SZ] - Az * c->U[C_RW][n];
            c->G[C_RHO][n] = -((1 / V) * (dxu + dyv + dzw));
        }
        {   /* Gⁿ.ρθ = -div_ρUc(ρᵈ, (u, v, w), θ) */
            double fx = fx_ ? 0.0 : tracer_flux_x(c, c->theta, i + 1, j, k) - tracer_flux_x(c, c->theta, i, j, k);
            double fy = fy_ ? 0.0 : tracer_flux_y(c, c->theta, i, j + 1, k) - tracer_flux_y(c, c->theta, i, j, k);
            double fz = tracer_flux_z(c, c->theta, i, j, k + 1) - tracer_flux_z(c, c->theta, i, j, k);
            c->G[C_RTH][n] = -((1 / V) * (fx + fy + fz));
        }
    }
}

/* compute_scalar_tendency! for the moisture density inside update_state!(compute_tendencies=true)
 * (update_atmosphere_model_state.jl:343, dynamics_kernel_functions.jl:132-159): -div_ρUc(ρ_total, ⟨𝐮⟩, qᵛ) with the acoustic-mean
 * transport velocities (acoustic_runge_kutta_3.jl:352-358) */
static void compute_moisture_tendency(orcc_ctx* c) {
    if (!c->moist) return;
    const double V = c->dx * c->dy * c->dz, Ax = c->dy * c->dz, Ay = c->dx * c->dz, Az = c->dx * c->dy;
    const double* rho = c->rho_tot; const double* q = c->qv;
    FOR_CELLS(c->Nz) {
        size_t n = IDX(c, i, j, k);
        double fx = 0.0, fy = 0.0, fz;
        if (!c->flat_x) {
            double ue = c->avg_u[n + SX], uw = c->avg_u[n];
            double Fe = ((rho[n + SX] + rho[n]) / 2) * (Ax * ue * biased_interp(q + n + SX, SX, c->B, ue > 0));
            double Fw = ((rho[n] + rho[n - SX]) / 2) * (Ax * uw * biased_interp(q + n, SX, c->B, uw > 0));
            fx = Fe - Fw;
        }
        if (!c->flat_y) {
            double vn = c->avg_v[n + SY], vs = c->avg_v[n];
            double Fn = ((rho[n + SY] + rho[n]) / 2) * (Ay * vn * biased_interp(q + n + SY, SY, c->B, vn > 0));
            double Fs = ((rho[n] + rho[n - SY]) / 2) * (Ay * vs * biased_interp(q + n, SY, c->B, vs > 0));
            fy = Fn - Fs;
        }
        {
            double Ft = 0.0, Fb = 0.0;
            if (k + 1 < c->Nz) { double wt = c->avg_w[n + SZ]; Ft = ((rho[n + SZ] + rho[n]) / 2) * (Az * wt * biased_interp(q + n + SZ, SZ, red_face(k + 1, c->Nz, c->B), wt > 0)); }
            if (k > 0) { double wb = c->avg_w[n]; Fb = ((rho[n] + rho[n - SZ]) / 2) * (Az * wb * biased_interp(q + n, SZ, red_face(k, c->Nz, c->B), wb > 0)); }
            fz = Ft - Fb;
        }
        c->Grqv[n] = -((1 / V) * (fx + fy + fz));
    }
}

/* assemble_slow_vertical_momentum_tendency!: acoustic_substepping.jl:689-748 */
static void assemble_slow_vertical_momentum_tendency(orcc_ctx* c) {
    const double g = c->g, rdz = 1 / c->dz;
    const int Hz = c->Hz;
    FOR_CELLS(c->Nz) {
        size_t n = IDX(c, i, j, k);
        double v;
        if (c->has_ref) {
            double dp_k = c->p[n] - c->p_r[k + Hz], dp_m = c->p[n - SZ] - c->p_r[k - 1 + Hz];
            double dr_k = c->rho_tot[n] - c->rho_r[k + Hz], dr_m = c->rho_tot[n - SZ] - c->rho_r[k - 1 + Hz];
            v = c->G[C_RW][n] - (dp_k - dp_m) * rdz - g * ((dr_k + dr_m) / 2);
        } else {
            v = c->G[C_RW][n] - (c->p[n] - c->p[n - SZ]) * rdz - g * ((c->rho_tot[n] + c->rho_tot[n - SZ]) / 2);
        }
        c->Gs_rw[n] = v * (k > 0);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* substep counts: acoustic_substepping.jl:451-508                                                     */
/* ------------------------------------------------------------------------------------------------ */
static int compute_acoustic_substeps(const orcc_ctx* c, double dt) {
    double gamd = c->cpd / (c->cpd - c->Rd);
    double cs = sqrt(gamd * c->Rd * 300.0);
    double dxm = c->flat_x ? INFINITY : c->dx, dym = c->flat_y ? INFINITY : c->dy;
    double dmin = fmin(dxm, dym);
    double N = ceil(fabs(dt) * cs / (c->cfg.acoustic_cfl * dmin));
    return N < 1 ? 1 : (int)N;
}
static int imax_(int a, int b) { return a > b ? a : b; }
static void stage_substep_count_and_size(const orcc_ctx* c, double beta, double dt, int* n_tau, double* d_tau) {
    const int S = c->cfg.substeps;
    if (c->cfg.substep_distribution == BZC_PROPORTIONAL_SUBSTEPS) {
        double dt_stage = beta * dt;
        int N = S > 0 ? imax_(1, (int)ceil(beta * S)) : compute_acoustic_substeps(c, dt_stage);
        *n_tau = N; *d_tau = dt_stage / N;
        return;
    }
    if (c->cfg.substep_distribution == BZC_MONOLITHIC_FIRST_STAGE && beta < (1.0 / 3 + 1.0 / 2) / 2) { *n_tau = 1; *d_tau = dt / 3; return; }
    int Nraw = S > 0 ? S : compute_acoustic_substeps(c, dt);
    int N = imax_(6, 6 * ((Nraw + 5) / 6));
    *n_tau = imax_(1, (int)nearbyint(beta * N)); *d_tau = dt / N;
}

/* ------------------------------------------------------------------------------------------------ */
/* acoustic_rk3_substep_loop!: acoustic_substepping.jl:1404-1590                                        */
/* ------------------------------------------------------------------------------------------------ */

/* ℑbzᵃᵃᶠ: boundary-aware centre → z-face interpolation (:539-550); face k lies between cells k-1 and k */
static inline double interp_bz(const orcc_ctx* c, const double* f, size_t n, int k) {
    double fp = f[n], fm = f[n - SZ];
    int pp = (k > c->Nz - 1) || (k < 0), pm = (k - 1 < 0) || (k - 1 > c->Nz - 1);
    fp = pp ? fm : fp;
    fm = pm ? fp : fm;
    return (fp + fm) / 2;
}

/* UpperSponge: rate · ramp(z_face, Lz, depth) at z-face k (acoustic_substepping.jl:591-603, time_discretizations.jl:397-437);
 * the ramp receives grid.Lz as the sponge top */
static inline double sponge_rate_at_face(const orcc_ctx* c, int k) {
    if (c->cfg.sponge == BZC_SPONGE_NONE) return 0.0;
    double z = c->cfg.base.z0 + k * c->dz, Lz = c->cfg.base.z1 - c->cfg.base.z0, depth = c->cfg.sponge_depth;
    double sfrac = (z - (Lz - depth)) / depth;
    sfrac = sfrac < 0 ? 0 : (sfrac > 1 ? 1 : sfrac);
    double ramp = sfrac;
    if (c->cfg.sponge == BZC_SPONGE_CUBIC_RAMP) ramp = sfrac * sfrac * (3 - 2 * sfrac);
    else if (c->cfg.sponge == BZC_SPONGE_SIN2_RAMP) { double sn = sin(M_PI / 2 * sfrac); ramp = sn * sn; }
    return c->cfg.sponge_damping_rate * ramp;
}
double orcc_test_sponge_term_diag(const bzc_config* cfg, int k_ref, double dtm) {
    orcc_ctx c; memset(&c, 0, sizeof(c)); c.cfg = *cfg; c.dz = (cfg->base.z1 - cfg->base.z0) / cfg->base.Nz;
    return fabs(dtm) * sponge_rate_at_face(&c, k_ref - 1);
}
double orcc_test_sponge_rhs(const bzc_config* cfg, int k_ref, double dts, double rho_w_old) {
    orcc_ctx c; memset(&c, 0, sizeof(c)); c.cfg = *cfg; c.dz = (cfg->base.z1 - cfg->base.z0) / cfg->base.Nz;
    return fabs(dts) * sponge_rate_at_face(&c, k_ref - 1) * rho_w_old;
}

/* get_coefficient for the three diagonal tags (:605-659); row = z-face k (0-based). */
static inline double tri_lower_for_row(const orcc_ctx* c, size_t n, int k, double dtm, double g, double dm) {
    /* reference: get_coefficient(k_ref - 1, ::AcousticTridiagLower) with kᶠ = k_ref */
    double rdzf = 1 / c->dz, rdzc = 1 / c->dz;
    double Ck = c->gRL[n - SZ] * c->PiL[n - SZ];
    double th = interp_bz(c, c->thL, n - SZ, k - 1);
    double pgf = -(dtm * dtm) * Ck * th * rdzc * rdzf;
    double buoy = (dtm * dtm) * g * rdzc / 2;
    double damp = -dm * rdzc * rdzf;
    return pgf + buoy + damp;
}
static inline double tri_diag(const orcc_ctx* c, size_t n, int k, double dtm, double g, double dm) {
    double rdzf = 1 / c->dz, rdzp = 1 / c->dz, rdzm = 1 / c->dz;
    double Cp = c->gRL[n] * c->PiL[n], Cm = c->gRL[n - SZ] * c->PiL[n - SZ];
    double th = interp_bz(c, c->thL, n, k);
    double pgf = (dtm * dtm) * th * (Cp * rdzp + Cm * rdzm) * rdzf;
    double buoy = (dtm * dtm) * g * (rdzp - rdzm) / 2;
    double damp = dm * (rdzp + rdzm) * rdzf;
    double sponge = fabs(dtm) * sponge_rate_at_face(c, k);
    return 1 + (pgf + buoy + damp + sponge) * (k > 0);
}
static inline double tri_upper(const orcc_ctx* c, size_t n, int k, double dtm, double g, double dm) {
    double rdzf = 1 / c->dz, rdzp = 1 / c->dz;
    double Cp = c->gRL[n] * c->PiL[n];
    double th = interp_bz(c, c->thL, n + SZ, k + 1);
    double pgf = -(dtm * dtm) * Cp * th * rdzp * rdzf;
    double buoy = -(dtm * dtm) * g * rdzp / 2;
    double damp = -dm * rdzp * rdzf;
    return (pgf + buoy + damp) * (k > 0);
}

/* apply_horizontal_pressure_gradient_substep (:887-891); substep is 1-based */
static inline int apply_pgf_substep(int substep, int n_tau, int apply_first) { return apply_first | (substep != 1) | (n_tau == 1); }
int orcc_apply_horizontal_pressure_gradient_substep(int substep, int n_tau, int apply_first) { return apply_pgf_substep(substep, n_tau, apply_first); }

/* Step A — _explicit_horizontal_step! (:859-876) */
static void explicit_horizontal_step(orcc_ctx* c, double dtau, int apply) {
    const double rdx = 1 / c->dx, rdy = 1 / c->dy;
    const double factor = apply ? 1.0 : 0.0;
    FOR_CELLS(c->Nz) {
        size_t n = IDX(c, i, j, k);
        double dxpL = c->flat_x ? 0.0 : (c->p[n] - c->p[n - SX]) * rdx;
        double dxpp = c->flat_x ? 0.0 : (c->gRL[n] * c->PiL[n] * c->rth_p[n] - c->gRL[n - SX] * c->PiL[n - SX] * c->rth_p[n - SX]) * rdx;
        double dypL = c->flat_y ? 0.0 : (c->p[n] - c->p[n - SY]) * rdy;
        double dypp = c->flat_y ? 0.0 : (c->gRL[n] * c->PiL[n] * c->rth_p[n] - c->gRL[n - SY] * c->PiL[n - SY] * c->rth_p[n - SY]) * rdy;
        double dxp = dxpL + factor * dxpp;
        double dyp = dypL + factor * dypp;
        c->ru_p[n] += dtau * (c->G[C_RU][n] - dxp);
        c->rv_p[n] += dtau * (c->G[C_RV][n] - dyp);
    }
}

/* Step B — _build_predictors! (:902-920) */
static void build_predictors(orcc_ctx* c, double dtau, double dts) {
    const double V = c->dx * c->dy * c->dz, Vinv = 1 / V, rdz = 1 / c->dz;
    const double Ax = c->dy * c->dz, Ay = c->dx * c->dz;
    const double fth = c->cfg.thermodynamic_tendency_factor;
    FOR_CELLS(c->Nz) {
        size_t n = IDX(c, i, j, k);
        c->rth_old[n] = c->rth_p[n];
        double dxM = c->flat_x ? 0.0 : Ax * c->ru_p[n + SX] - Ax * c->ru_p[n];
        double dyM = c->flat_y ? 0.0 : Ay * c->rv_p[n + SY] - Ay * c->rv_p[n];
        double divM = Vinv * (dxM + dyM);
        double dxT = c->flat_x ? 0.0 : Ax * ((c->thL[n + SX] + c->thL[n]) / 2) * c->ru_p[n + SX] - Ax * ((c->thL[n] + c->thL[n - SX]) / 2) * c->ru_p[n];
        double dyT = c->flat_y ? 0.0 : Ay * ((c->thL[n + SY] + c->thL[n]) / 2) * c->rv_p[n + SY] - Ay * ((c->thL[n] + c->thL[n - SY]) / 2) * c->rv_p[n];
        double divT = Vinv * (dxT + dyT);
        double dzw = (c->rw_p[n + SZ] - c->rw_p[n]) * rdz;
        double dzT = (interp_bz(c, c->thL, n + SZ, k + 1) * c->rw_p[n + SZ] - interp_bz(c, c->thL, n, k) * c->rw_p[n]) * rdz;
        c->rho_s[n] = c->rho_p[n] + dtau * (c->G[C_RHO][n] - divM) - dts * dzw;
        c->rth_s[n] = c->rth_p[n] + dtau * (fth * c->G[C_RTH][n] - divT) - dts * dzT;
    }
}

/* _build_vertical_rhs! over the Nz+1 faces (:926-958) */
static void build_vertical_rhs(orcc_ctx* c, double dtau, double dtm, double dts, double ds) {
    const double g = c->g, rdz = 1 / c->dz, fw = c->cfg.vertical_momentum_tendency_factor;
    const int Nz = c->Nz;
    FOR_CELLS(Nz + 1) {
        size_t n = IDX(c, i, j, k);
        double Ck = c->gRL[n] * c->PiL[n], Cm = c->gRL[n - SZ] * c->PiL[n - SZ];
        double dp_s = (Ck * c->rth_s[n] - Cm * c->rth_s[n - SZ]) * rdz;
        double dp_o = (Ck * c->rth_p[n] - Cm * c->rth_p[n - SZ]) * rdz;
        double Gp = dts * dp_o + dtm * dp_s;
        double r_s = (c->rho_s[n] + c->rho_s[n - SZ]) / 2;
        double r_o = (c->rho_p[n] + c->rho_p[n - SZ]) / 2;
        double Gb = g * (dts * r_o + dtm * r_s);
        double d2 = ((c->rw_p[n + SZ] - c->rw_p[n]) * rdz - (c->rw_p[n] - c->rw_p[n - SZ]) * rdz) * rdz;
        double Gd = -ds * d2;
        double Gsp = fabs(dts) * sponge_rate_at_face(c, k) * c->rw_p[n];
        double rhs = c->rw_p[n] + dtau * fw * c->Gs_rw[n] - Gp - Gb - Gd - Gsp;
        c->rhs[n] = ((k != 0) & (k != Nz)) ? rhs : 0.0;
    }
}

/* Step C — solve!(::BatchedTridiagonalSolver) with on-the-fly coefficients, rows = faces 0..Nz-1 */
static void solve_vertical(orcc_ctx* c, double dtm, double dm) {
    const double g = c->g;
    const int Nz = c->Nz;
    const double tiny = 10 * 2.220446049250313e-16;
#pragma omp parallel for collapse(2) schedule(static)
    for (int j = 0; j < c->Ny; ++j)
        for (int i = 0; i < c->Nx; ++i) {
            size_t n0 = IDX(c, i, j, 0);
            double beta = tri_diag(c, n0, 0, dtm, g, dm);
            c->rw_p[n0] = c->rhs[n0] / beta;
            for (int k = 1; k < Nz; ++k) {
                size_t n = n0 + (size_t)k * SZ;
                double cu = tri_upper(c, n - SZ, k - 1, dtm, g, dm);
                double al = tri_lower_for_row(c, n, k, dtm, g, dm);
                double b = tri_diag(c, n, k, dtm, g, dm);
                double t = cu / beta;
                c->scratch[n] = t;
                beta = b - al * t;
                if (fabs(beta) > tiny) c->rw_p[n] = (c->rhs[n] - al * c->rw_p[n - SZ]) / beta;
            }
            for (int k = Nz - 2; k >= 0; --k) {
                size_t n = n0 + (size_t)k * SZ;
                c->rw_p[n] -= c->scratch[n + SZ] * c->rw_p[n + SZ];
            }
        }
}

/* Step D — _post_solve_recovery! (:978-987) */
static void post_solve_recovery(orcc_ctx* c, double dtm) {
    const double rdz = 1 / c->dz;
    FOR_CELLS(c->Nz) {
        size_t n = IDX(c, i, j, k);
        double dzw = (c->rw_p[n + SZ] - c->rw_p[n]) * rdz;
        double dzT = (interp_bz(c, c->thL, n + SZ, k + 1) * c->rw_p[n + SZ] - interp_bz(c, c->thL, n, k) * c->rw_p[n]) * rdz;
        c->rho_p[n] = c->rho_s[n] - dtm * dzw;
        c->rth_p[n] = c->rth_s[n] - dtm * dzT;
        c->avg_u[n] += c->ru_p[n];
        c->avg_v[n] += c->rv_p[n];
        c->avg_w[n] += c->rw_p[n];
    }
}

/* Step E — _thermal_divergence_damping! (:1045-1063,1105-1144) */
static void thermal_divergence_damping(orcc_ctx* c, double dtau) {
    const double alpha = c->cfg.damping_coefficient;
    double kx, ky;
    if (c->cfg.damping_length_scale > 0) {
        double coef = alpha * (c->cfg.damping_length_scale * c->cfg.damping_length_scale);
        kx = ky = coef / dtau;
    } else {
        double dxm = c->flat_x ? INFINITY : c->dx, dym = c->flat_y ? INFINITY : c->dy;
        double l = fmin(dxm, dym);
        kx = ky = alpha * (l * l) / dtau;
    }
    if (c->flat_x) kx = 0;
    if (c->flat_y) ky = 0;
    const double rdx = 1 / c->dx, rdy = 1 / c->dy;
    FOR_CELLS(c->Nz) {
        size_t n = IDX(c, i, j, k);
        double d0 = c->rth_p[n] - c->rth_old[n];
        if (!c->flat_x) {
            double dxd = (d0 - (c->rth_p[n - SX] - c->rth_old[n - SX])) * rdx;
            double thf = (c->thL[n] + c->thL[n - SX]) / 2;
            c->ru_p[n] -= kx * dxd / thf;
        }
        if (!c->flat_y) {
            double dyd = (d0 - (c->rth_p[n - SY] - c->rth_old[n - SY])) * rdy;
            double thf = (c->thL[n] + c->thL[n - SY]) / 2;
            c->rv_p[n] -= ky * dyd / thf;
        }
    }
}

static void acoustic_substep_loop(orcc_ctx* c, double dt, double beta) {
    int n_tau; double dtau;
    stage_substep_count_and_size(c, beta, dt, &n_tau, &dtau);
    const double om = c->cfg.forward_weight;
    const double dtm = om * dtau, dts = (1 - om) * dtau;

    assemble_slow_vertical_momentum_tendency(c);

    /* initialize_stage_perturbations! (:765-842): zero workspaces, rewind-initialise the perturbations */
    FOR_CELLS(c->Nz) {
        size_t n = IDX(c, i, j, k);
        c->rth_old[n] = 0; c->rho_s[n] = 0; c->rth_s[n] = 0; c->avg_u[n] = 0; c->avg_v[n] = 0; c->avg_w[n] = 0;
        c->rho_p[n] = c->U0[C_RHO][n] - c->U[C_RHO][n];
        c->rth_p[n] = c->U0[C_RTH][n] - c->U[C_RTH][n];
        c->ru_p[n] = c->U0[C_RU][n] - c->U[C_RU][n];
        c->rv_p[n] = c->U0[C_RV][n] - c->U[C_RV][n];
        c->rw_p[n] = c->U0[C_RW][n] - c->U[C_RW][n];
    }
    fill_halos(c, c->rho_p, LOC_CENTER);
    fill_halos(c, c->rth_p, LOC_CENTER);
    fill_halos(c, c->ru_p, LOC_CENTER);
    fill_halos(c, c->rv_p, LOC_CENTER);
    fill_halos(c, c->rw_p, LOC_ZFACE);

    /* implicit_damping_factors (:1003-1011) */
    double dm = 0, ds = 0;
    if (c->cfg.damping == BZC_THERMAL_DIVERGENCE_DAMPING && c->cfg.damp_vertical) {
        double base = c->cfg.damping_coefficient * (c->dz * c->dz);
        dm = om * base; ds = (1 - om) * base;
    }

    for (int substep = 1; substep <= n_tau; ++substep) {
        int apply = apply_pgf_substep(substep, n_tau, c->cfg.apply_first_substep_pressure_gradient);
        explicit_horizontal_step(c, dtau, apply);
        fill_halos(c, c->ru_p, LOC_CENTER);
        fill_halos(c, c->rv_p, LOC_CENTER);
        build_predictors(c, dtau, dts);
        fill_halos(c, c->rth_old, LOC_CENTER);
        build_vertical_rhs(c, dtau, dtm, dts, ds);
        solve_vertical(c, dtm, dm);
        post_solve_recovery(c, dtm);
        fill_halos(c, c->rho_p, LOC_CENTER);
        fill_halos(c, c->rth_p, LOC_CENTER);
        if (c->cfg.damping == BZC_THERMAL_DIVERGENCE_DAMPING) thermal_divergence_damping(c, dtau);
        fill_halos(c, c->ru_p, LOC_CENTER);
        fill_halos(c, c->rv_p, LOC_CENTER);
    }

    /* finalize_time_averaged_velocity! (:1208-1253) */
    const double invN = 1.0 / (double)n_tau;
    const double* rho = c->U[C_RHO];
    FOR_CELLS(c->Nz) {
        size_t n = IDX(c, i, j, k);
        double ru_t = c->U[C_RU][n] + c->avg_u[n] * invN;
        double rv_t = c->U[C_RV][n] + c->avg_v[n] * invN;
        double rw_t = c->U[C_RW][n] + c->avg_w[n] * invN;
        double rx = c->flat_x ? rho[n] : (rho[n] + rho[n - SX]) / 2;
        double ry = c->flat_y ? rho[n] : (rho[n] + rho[n - SY]) / 2;
        double rz = (rho[n] + rho[n - SZ]) / 2;
        rx = (rx == 0) ? 1.0 : rx; ry = (ry == 0) ? 1.0 : ry; rz = (rz == 0) ? 1.0 : rz;
        c->avg_u[n] = ru_t / rx;
        c->avg_v[n] = rv_t / ry;
        c->avg_w[n] = rw_t / rz * (k > 0);
    }
    fill_halos(c, c->avg_u, LOC_CENTER);
    fill_halos(c, c->avg_v, LOC_CENTER);
    fill_halos(c, c->avg_w, LOC_ZFACE);

    /* _recover_full_state! (:1274-1293), halos, compute_velocities! (:1584-1587) */
    FOR_CELLS(c->Nz) {
        size_t n = IDX(c, i, j, k);
        c->U[C_RHO][n] = c->U[C_RHO][n] + c->rho_p[n];
        c->U[C_RTH][n] = c->U[C_RTH][n] + c->rth_p[n];
        c->U[C_RU][n] = c->U[C_RU][n] + c->ru_p[n];
        c->U[C_RV][n] = c->U[C_RV][n] + c->rv_p[n];
        c->U[C_RW][n] = c->U[C_RW][n] + c->rw_p[n];
    }
    fill_halos(c, c->U[C_RTH], LOC_CENTER);
    compute_velocities(c);
}

/* prepare_acoustic_cache! + slow tendencies (acoustic_runge_kutta_3.jl:181-193) */
static void stage_tendencies(orcc_ctx* c) {
    refresh_linearization_basic_state(c);
    compute_slow_tendencies(c);
}

static void store_initial_state(orcc_ctx* c) {
    for (int f = 0; f < NPROGC; ++f) memcpy(c->U0[f], c->U[f], c->n_padded * sizeof(double));
    memcpy(c->rqv0, c->rqv, c->n_padded * sizeof(double));
}

/* scalar_rk3_substep! (acoustic_runge_kutta_3.jl:214-223): ρqᵛ = ρqᵛ⁰ + βΔt Gⁿ.ρqᵛ, Gⁿ from the preceding update_state! */
static void scalar_rk3_substep(orcc_ctx* c, double dt_stage) {
    if (!c->moist) return;
    FOR_CELLS(c->Nz) { size_t n = IDX(c, i, j, k); c->rqv[n] = c->rqv0[n] + dt_stage * c->Grqv[n]; }
}

/* time_step!(model::CompressibleAcousticModel, Δt): acoustic_runge_kutta_3.jl:264-319. The full (non-slow) tendencies that
 * update_state!(compute_tendencies=true) evaluates between stages are overwritten by the next stage's slow tendencies
 * before anything reads them in a dry model without tracers, so they are not evaluated here. */
static void time_step(orcc_ctx* c, double dt) {
    const double betas[3] = {1.0 / 3.0, 1.0 / 2.0, 1.0};
    store_initial_state(c);
    refresh_linearization_basic_state(c);               /* freeze_linearization_state! */
    memcpy(c->avg_u, c->u, c->n_padded * sizeof(double));  /* seed_time_averaged_velocities! */
    memcpy(c->avg_v, c->v, c->n_padded * sizeof(double));
    memcpy(c->avg_w, c->w, c->n_padded * sizeof(double));
    for (int s = 0; s < 3; ++s) {
        stage_tendencies(c);
        acoustic_substep_loop(c, dt, betas[s]);
        scalar_rk3_substep(c, betas[s] * dt);
        update_state(c);
        compute_moisture_tendency(c);                    /* update_state!(compute_tendencies = true): consumed by the NEXT stage */
    }
    c->time += dt; c->iteration += 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* ABI                                                                                               */
/* ------------------------------------------------------------------------------------------------ */
const char* orcc_last_error(const orcc_ctx* c) { return c ? c->err : g_err; }

void orcc_destroy(orcc_ctx* c) {
    if (!c) return;
    double** all[] = {&c->p_r, &c->rho_r, &c->pi_r, &c->theta_r, &c->u, &c->v, &c->w, &c->theta, &c->T, &c->p, &c->PiL, &c->thL, &c->gRL,
                      &c->rho_p, &c->rth_p, &c->ru_p, &c->rv_p, &c->rw_p, &c->rho_s, &c->rth_s, &c->rth_old, &c->avg_u, &c->avg_v, &c->avg_w,
                      &c->Gs_rw, &c->rhs, &c->scratch};
    for (size_t a = 0; a < sizeof(all) / sizeof(all[0]); ++a) free(*all[a]);
    for (int f = 0; f < NPROGC; ++f) { free(c->U[f]); free(c->U0[f]); free(c->G[f]); }
    free(c->rqv); free(c->rqv0); free(c->Grqv); free(c->qv); free(c->rho_tot);
    free(c);
}

int orcc_create(const bzc_config* cfg, orcc_ctx** out) {
    if (!cfg || !out) { set_err(NULL, "null argument"); return BZ_ERR_INVALID; }
    const bz_config* b = &cfg->base;
    if (b->abi_version != BZ_ABI_VERSION) { set_err(NULL, "ABI version mismatch"); return BZ_ERR_INVALID; }
    if (b->microphysics != BZ_MICROPHYSICS_NONE || b->n_ranks > 1) { set_err(NULL, "compressible path: dry air, one rank"); return BZ_ERR_UNSUPPORTED; }
    /* the oracle also restates WENO(order = 7 / 9) (the CUDA path rejects them) */
    if (b->advection_order != 5 && b->advection_order != 7 && b->advection_order != 9) { set_err(NULL, "WENO(order = 5, 7 or 9)"); return BZ_ERR_UNSUPPORTED; }
    if (b->Nz < 4) { set_err(NULL, "Nz >= 4 required"); return BZ_ERR_INVALID; }
    orcc_ctx* c = (orcc_ctx*)calloc(1, sizeof(orcc_ctx));
    c->cfg = *cfg;
    c->flat_x = b->topology_x == BZ_FLAT; c->flat_y = b->topology_y == BZ_FLAT;
    c->Nx = c->flat_x ? 1 : b->Nx; c->Ny = c->flat_y ? 1 : b->Ny; c->Nz = b->Nz;
    c->Hx = c->flat_x ? 0 : HALO(b->advection_order); c->Hy = c->flat_y ? 0 : HALO(b->advection_order); c->Hz = HALO(b->advection_order);
    c->B = (b->advection_order + 1) / 2; c->Bs = c->B - 1;
    c->Px = c->Nx + 2 * c->Hx; c->Py = c->Ny + 2 * c->Hy; c->Pz = c->Nz + 1 + 2 * c->Hz;
    c->n_padded = (size_t)c->Px * c->Py * c->Pz;
    c->dx = c->flat_x ? 1.0 : (b->x1 - b->x0) / c->Nx;
    c->dy = c->flat_y ? 1.0 : (b->y1 - b->y0) / c->Ny;
    c->dz = (b->z1 - b->z0) / c->Nz;
    c->Rd = b->molar_gas_constant / b->dry_air_molar_mass;
    c->Rv = b->molar_gas_constant / b->vapor_molar_mass;
    c->cpd = b->dry_air_heat_capacity; c->cpv = b->vapor_heat_capacity;
    c->g = b->gravitational_acceleration; c->pst = b->standard_pressure; c->p0 = b->surface_pressure;
    c->has_ref = cfg->reference_state == BZC_REFERENCE_EXNER;
    size_t nz = (size_t)c->Nz + 2 * c->Hz + 1;
    c->p_r = calloc(nz, 8); c->rho_r = calloc(nz, 8); c->pi_r = calloc(nz, 8); c->theta_r = calloc(nz, 8);
    for (int f = 0; f < NPROGC; ++f) { c->U[f] = new_field(c); c->U0[f] = new_field(c); c->G[f] = new_field(c); }
    c->rqv = new_field(c); c->rqv0 = new_field(c); c->Grqv = new_field(c); c->qv = new_field(c); c->rho_tot = new_field(c);
    double** fs[] = {&c->u, &c->v, &c->w, &c->theta, &c->T, &c->p, &c->PiL, &c->thL, &c->gRL, &c->rho_p, &c->rth_p, &c->ru_p, &c->rv_p, &c->rw_p,
                     &c->rho_s, &c->rth_s, &c->rth_old, &c->avg_u, &c->avg_v, &c->avg_w, &c->Gs_rw, &c->rhs, &c->scratch};
    for (size_t a = 0; a < sizeof(fs) / sizeof(fs[0]); ++a) *fs[a] = new_field(c);
    if (c->has_ref) {
        for (int k = 0; k < c->Nz; ++k) c->theta_r[k + c->Hz] = b->potential_temperature;
        build_exner_reference(c);
        /* seed_pressure! from the reference (compressible_dynamics.jl:290-294) */
        FOR_CELLS(c->Nz) c->p[IDX(c, i, j, k)] = c->p_r[k + c->Hz];
    } else {
        FOR_CELLS(c->Nz) c->p[IDX(c, i, j, k)] = c->p0;
    }
    *out = c;
    return BZ_OK;
}

int orcc_set_reference_potential_temperature(orcc_ctx* c, const double* theta_r) {
    if (!c->has_ref) { set_err(c, "reference_state = nothing"); return BZ_ERR_STATE; }
    for (int k = 0; k < c->Nz; ++k) c->theta_r[k + c->Hz] = theta_r[k];
    build_exner_reference(c);
    return BZ_OK;
}

int orcc_get_reference_state(orcc_ctx* c, double* p, double* rho, double* pi) {
    if (!c->has_ref) { set_err(c, "reference_state = nothing"); return BZ_ERR_STATE; }
    for (int k = 0; k < c->Nz; ++k) {
        if (p) p[k] = c->p_r[k + c->Hz];
        if (rho) rho[k] = c->rho_r[k + c->Hz];
        if (pi) pi[k] = c->pi_r[k + c->Hz];
    }
    return BZ_OK;
}

static void copy_in(orcc_ctx* c, double* dst, const double* src, int nzl) {
    for (int k = 0; k < nzl; ++k) for (int j = 0; j < c->Ny; ++j) for (int i = 0; i < c->Nx; ++i)
        dst[IDX(c, i, j, k)] = src[(size_t)i + (size_t)c->Nx * ((size_t)j + (size_t)c->Ny * k)];
}
static void copy_out(const orcc_ctx* c, double* dst, const double* src, int nzl) {
    for (int k = 0; k < nzl; ++k) for (int j = 0; j < c->Ny; ++j) for (int i = 0; i < c->Nx; ++i)
        dst[(size_t)i + (size_t)c->Nx * ((size_t)j + (size_t)c->Ny * k)] = src[IDX(c, i, j, k)];
}

int orcc_set_state(orcc_ctx* c, const double* rho, const double* ru, const double* rv, const double* rw, const double* rth, const double* rqv) {
    const double* src[NPROGC] = {rho, ru, rv, rw, rth};
    for (int f = 0; f < NPROGC; ++f) if (src[f]) copy_in(c, c->U[f], src[f], f == C_RW ? c->Nz + 1 : c->Nz);
    if (rqv) { copy_in(c, c->rqv, rqv, c->Nz); c->moist = 1; }
    update_state(c);
    store_initial_state(c);
    /* maybe_prepare_first_time_step! (acoustic_runge_kutta_3.jl:240-256): seed ⟨𝐮⟩ with the velocities, then the first tendencies */
    memcpy(c->avg_u, c->u, c->n_padded * sizeof(double));
    memcpy(c->avg_v, c->v, c->n_padded * sizeof(double));
    memcpy(c->avg_w, c->w, c->n_padded * sizeof(double));
    compute_moisture_tendency(c);
    return BZ_OK;
}

int orcc_time_step(orcc_ctx* c, double dt) { time_step(c, dt); return BZ_OK; }
int orcc_time_steps(orcc_ctx* c, double dt, int n) { for (int s = 0; s < n; ++s) time_step(c, dt); return BZ_OK; }

int orcc_compute_slow_tendencies(orcc_ctx* c) {
    stage_tendencies(c);
    assemble_slow_vertical_momentum_tendency(c);
    return BZ_OK;
}
int orcc_stage_substep_count_and_size(orcc_ctx* c, double dt, double beta, int32_t* n_tau, double* d_tau) {
    int n; double d;
    stage_substep_count_and_size(c, beta, dt, &n, &d);
    if (n_tau) *n_tau = n;
    if (d_tau) *d_tau = d;
    return BZ_OK;
}
int orcc_acoustic_substep_loop(orcc_ctx* c, double dt, double beta) {
    acoustic_substep_loop(c, dt, beta);
    scalar_rk3_substep(c, beta * dt);
    update_state(c);
    compute_moisture_tendency(c);
    return BZ_OK;
}

int orcc_get_field(orcc_ctx* c, int f, double* out) {
    const double* src = NULL; int zf = 0;
    switch (f) {
        case BZC_RHO: src = c->U[C_RHO]; break;
        case BZC_RHO_U: src = c->U[C_RU]; break;
        case BZC_RHO_V: src = c->U[C_RV]; break;
        case BZC_RHO_W: src = c->U[C_RW]; zf = 1; break;
        case BZC_RHO_THETA: src = c->U[C_RTH]; break;
        case BZC_U: src = c->u; break;
        case BZC_V: src = c->v; break;
        case BZC_W: src = c->w; zf = 1; break;
        case BZC_THETA: src = c->theta; break;
        case BZC_T: src = c->T; break;
        case BZC_P: src = c->p; break;
        case BZC_G_RHO: src = c->G[C_RHO]; break;
        case BZC_G_RHO_U: src = c->G[C_RU]; break;
        case BZC_G_RHO_V: src = c->G[C_RV]; break;
        case BZC_G_RHO_W: src = c->G[C_RW]; break;
        case BZC_G_RHO_THETA: src = c->G[C_RTH]; break;
        case BZC_SLOW_RHO_W: src = c->Gs_rw; zf = 1; break;
        case BZC_EXNER_L: src = c->PiL; break;
        case BZC_THETA_L: src = c->thL; break;
        case BZC_GAMMA_R_L: src = c->gRL; break;
        case BZC_RHO_PERT: src = c->rho_p; break;
        case BZC_RHO_THETA_PERT: src = c->rth_p; break;
        case BZC_RHO_U_PERT: src = c->ru_p; break;
        case BZC_RHO_V_PERT: src = c->rv_p; break;
        case BZC_RHO_W_PERT: src = c->rw_p; zf = 1; break;
        case BZC_AVG_U: src = c->avg_u; break;
        case BZC_AVG_V: src = c->avg_v; break;
        case BZC_AVG_W: src = c->avg_w; zf = 1; break;
        case BZC_RHO_QV: src = c->rqv; break;
        case BZC_QV: src = c->qv; break;
        case BZC_TOTAL_RHO: src = c->rho_tot; break;
        case BZC_G_RHO_QV: src = c->Grqv; break;
        default: set_err(c, "unknown field %d", f); return BZ_ERR_INVALID;
    }
    copy_out(c, out, src, zf ? c->Nz + 1 : c->Nz);
    return BZ_OK;
}
int orcc_get_state(orcc_ctx* c, double* rho, double* ru, double* rv, double* rw, double* rth, double* rqv) {
    double* dst[NPROGC] = {rho, ru, rv, rw, rth};
    for (int f = 0; f < NPROGC; ++f) if (dst[f]) copy_out(c, dst[f], c->U[f], f == C_RW ? c->Nz + 1 : c->Nz);
    if (rqv) copy_out(c, rqv, c->rqv, c->Nz);
    return BZ_OK;
}
int orcc_get_clock(orcc_ctx* c, double* t, int64_t* it) { if (t) *t = c->time; if (it) *it = c->iteration; return BZ_OK; }
int orcc_synchronize(orcc_ctx* c) { (void)c; return BZ_OK; }

/* ---- test hooks (oracle only): the pieces test/acoustic_substepping_components.jl probes directly ---------------- */

/* `get_coefficient` of the three tags for tridiagonal row k_ref (1-based), given column profiles Πᴸ, θᴸ, γRᵐᴸ (Nz values)
 * on a uniform column of spacing dz (test/acoustic_substepping_components.jl:95-166). out = {lower_for_row, diag, upper}. */
int orcc_test_tridiagonal_coefficients(int Nz, double dz, const double* Pi, const double* th, const double* gR, double g,
                                       double dtm, double dm, int k_ref, double* out3) {
    bzc_config cfg; orcc_default_config(&cfg);
    cfg.base.Nx = 1; cfg.base.Ny = 1; cfg.base.Nz = Nz; cfg.base.topology_x = cfg.base.topology_y = BZ_FLAT;
    cfg.base.z0 = 0; cfg.base.z1 = dz * Nz; cfg.reference_state = BZC_REFERENCE_NONE;
    orcc_ctx* c; int rc = orcc_create(&cfg, &c);
    if (rc) return rc;
    for (int k = 0; k < Nz; ++k) { size_t n = IDX(c, 0, 0, k); c->PiL[n] = Pi[k]; c->thL[n] = th[k]; c->gRL[n] = gR[k]; }
    fill_halos(c, c->PiL, LOC_CENTER); fill_halos(c, c->thL, LOC_CENTER); fill_halos(c, c->gRL, LOC_CENTER);
    int k = k_ref - 1;
    size_t n = IDX(c, 0, 0, k);
    out3[0] = (k >= 1) ? tri_lower_for_row(c, n, k, dtm, g, dm) : 0.0;
    out3[1] = tri_diag(c, n, k, dtm, g, dm);
    out3[2] = tri_upper(c, n, k, dtm, g, dm);
    orcc_destroy(c);
    return BZ_OK;
}

/* One `_explicit_horizontal_step!` launch on user-supplied fields (interior arrays; NULL = zeros): the known-answer check of
 * test/acoustic_substepping_components.jl:58-93. Overwrites ρu′, ρv′ (fetch with orcc_get_field). */
int orcc_test_explicit_horizontal_step(orcc_ctx* c, const double* p, const double* rth_p, const double* PiL, const double* gRL,
                                       double dtau, int apply) {
    double* dst[4] = {c->p, c->rth_p, c->PiL, c->gRL};
    const double* src[4] = {p, rth_p, PiL, gRL};
    for (int f = 0; f < 4; ++f) {
        memset(dst[f], 0, c->n_padded * sizeof(double));
        if (src[f]) copy_in(c, dst[f], src[f], c->Nz);
        fill_halos(c, dst[f], LOC_CENTER);
    }
    memset(c->ru_p, 0, c->n_padded * 8); memset(c->rv_p, 0, c->n_padded * 8);
    memset(c->G[C_RU], 0, c->n_padded * 8); memset(c->G[C_RV], 0, c->n_padded * 8);
    explicit_horizontal_step(c, dtau, apply);
    return BZ_OK;
}
