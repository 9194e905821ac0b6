/*
 * breeze_oracle.c — CPU ORACLE (test infrastructure, NOT a product path, NOT a fallback).
 *
 * A plain-C (C11 + OpenMP), FP64 restatement of the Breeze.jl hot path that libbreeze_b200.so replaces:
 * one SSP-RK3 step of AtmosphereModel{<:AnelasticDynamics} with WENO(order=5), liquid-ice potential
 * temperature formulation, closure = nothing, default boundary conditions. It is written the way the
 * reference runs it — unfused, one loop nest per reference kernel, halo-padded fields, the same order of
 * operations per stage — so that (i) the CUDA path can be checked against it field by field and hook by
 * hook, and (ii) it can be timed as the "CPU restatement of the reference algorithm" baseline.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * PARITY STATUS
 *   - Thermodynamics, ReferenceState, SSP-RK3 driver, tendencies assembly, anelastic buoyancy, Poisson
 *     source / diagonals / projection: restated from files under /root/reference (cited per function,
 *     paths relative to the reference repo) and pinned by the reference's own known-answer tests
 *     (tests/test_oracle_reference_vectors.py).
 *   - WENO5 reconstruction, centred-4 advecting interpolation, momentum-flux form, boundary order
 *     reduction, FourierTridiagonalPoissonSolver / batched Thomas: these live in Oceananigans.jl 0.110.14
 *     (Project.toml:43), which is NOT vendored in /root/reference and cannot be executed here (no Julia).
 *     They restate the published upstream algorithm. No reference test checks a
 *     WENO number => for those pieces: **parity unpinned** (invariants only: conservation, order of
 *     accuracy, divergence-free projection, analytic Poisson solution).
 *
 * Index conventions: 0-based C indices; reference 1-based index = C index + 1. Centre (i,j,k); x-face i is the
 * face between cells i-1 and i; z-face k is the bottom face of cell k, k = 0..Nz (0 and Nz are walls).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <complex.h>
#include <stdint.h>
#include <stddef.h>
#include <stdarg.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/breeze_b200.h"

#define HALO(order) (((order) + 1) / 2 + 1)   /* buffer + 1: 4 for WENO5 (any larger halo gives identical results; the reference default is 3), 6 for WENO9 */
#define NPROG 5

typedef double complex cplx;

typedef struct fft_plan {
    int n;
    cplx* w;              /* w[k] = exp(-2*pi*i*k/n) */
} fft_plan;

typedef struct orc_ctx {
    bz_config cfg;
    int Nx, Ny, Nz;       /* local == global (the oracle is single-process) */
    int Hx, Hy, Hz;
    int B, Bs;               /* buffers of the biased (WENO) and the symmetric (Centered) reconstructions: (order + 1) / 2 and B - 1 */
    int Px, Py, Pz;       /* padded sizes; Pz covers Nz+1 faces */
    size_t n_padded;
    double dx, dy, dz;    /* Flat dimension: spacing 1, as in Oceananigans */
    double Rd, Rv, cpd, cpv, g;
    /* reference state, z-only, with halos: index k + Hz */
    double *rho_r, *p_r, *T_r;
    /* prognostic (momentum, ρθ, ρq), U0, tendencies */
    double* U[NPROG];
    double* U0[NPROG];
    double* G[NPROG];
    /* diagnostics */
    double *u, *v, *w, *theta, *qv, *ql, *T, *phi;
    /* Poisson solver */
    cplx *rhs, *sol;      /* Nx*Ny*Nz, x fastest */
    double *lam_x, *lam_y;
    double *lower;        /* Nz-1 */
    double *diag;         /* Nx*Ny*Nz */
    fft_plan px, py;
    /* forcing / Coriolis / bottom flux BCs (bz_forcing); all absent by default */
    int stale;            /* tendencies must be rebuilt before the next stage (after set! / a change of forcing) */
    int has_forcing;
    double coriolis_f, theta_flux, q_flux, drag_rho_ustar2;
    int subsidence_mask;
    double *ws, *ug, *vg, *q_tend, *e_tend;      /* ws: Nz+1 faces; others Nz; NULL = absent */
    double *mean[4];                              /* horizontal means of u, v, θ, q (Nz) */
    double time; int64_t iteration;
    char err[256];
} orc_ctx;

static char g_create_err[256];

static void set_err(orc_ctx* c, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(c ? c->err : g_create_err, 256, fmt, ap);
    va_end(ap);
}

/* ------------------------------------------------------------------------------------------------ */
/* config                                                                                            */
/* ------------------------------------------------------------------------------------------------ */

/* ThermodynamicConstants() defaults: src/Thermodynamics/thermodynamics_constants.jl:182-212 and
 * liquid_water / water_ice :92-93; ReferenceState defaults: src/Thermodynamics/reference_states.jl:402-405 */
void orc_default_config(bz_config* c) {
    memset(c, 0, sizeof(*c));
    c->abi_version = BZ_ABI_VERSION;
    c->Nx = c->Ny = c->Nz = 8;
    c->topology_x = c->topology_y = BZ_PERIODIC;
    c->x0 = 0; c->x1 = 1; c->y0 = 0; c->y1 = 1; c->z0 = 0; c->z1 = 1;
    c->surface_pressure = 101325.0;
    c->potential_temperature = 288.0;
    c->standard_pressure = 1e5;
    c->molar_gas_constant = 8.314462618;
    c->gravitational_acceleration = 9.81;
    c->energy_reference_temperature = 273.15;
    c->triple_point_temperature = 273.16;
    c->triple_point_pressure = 611.657;
    c->dry_air_molar_mass = 0.02897;
    c->dry_air_heat_capacity = 1005.0;
    c->vapor_molar_mass = 0.018015;
    c->vapor_heat_capacity = 1850.0;
    c->liquid_reference_latent_heat = 2500800.0;
    c->liquid_heat_capacity = 4181.0;
    c->ice_reference_latent_heat = 2834000.0;
    c->ice_heat_capacity = 2108.0;
    c->advection_order = 5;
    c->microphysics = BZ_MICROPHYSICS_NONE;
    c->n_ranks = 1; c->rank = 0; c->device = 0;
}

int orc_abi_version(void) { return BZ_ABI_VERSION; }

/* ------------------------------------------------------------------------------------------------ */
/* indexing                                                                                          */
/* ------------------------------------------------------------------------------------------------ */
#define IDX(c, i, j, k) ((size_t)((i) + (c)->Hx) + (size_t)(c)->Px * ((size_t)((j) + (c)->Hy) + (size_t)(c)->Py * (size_t)((k) + (c)->Hz)))

static double* new_field(orc_ctx* c) { return (double*)calloc(c->n_padded, sizeof(double)); }

/* ------------------------------------------------------------------------------------------------ */
/* thermodynamics                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

/* mixture_gas_constant, mixture_heat_capacity: thermodynamics_constants.jl:341-347, 367-377 */
static inline double mixture_gas_constant(const orc_ctx* c, double qv, double ql, double qi) {
    double qd = 1 - (qv + ql + qi);
    return qd * c->Rd + qv * c->Rv;
}
static inline double mixture_heat_capacity(const orc_ctx* c, double qv, double ql, double qi) {
    double qd = 1 - (qv + ql + qi);
    return qd * c->cpd + qv * c->cpv + ql * c->cfg.liquid_heat_capacity + qi * c->cfg.ice_heat_capacity;
}

/* saturation_vapor_pressure over a planar liquid (mix = 1) / ice (mix = 0) / mixed surface:
 * src/Thermodynamics/clausius_clapeyron.jl:59-68 with absolute_zero_latent_heat and specific_heat_difference
 * (thermodynamics_constants.jl:262-274); mixed-phase surface blends ℒ₀ and Δc linearly in the liquid fraction. */
double orc_saturation_vapor_pressure(const bz_config* cfg, double T, double liquid_fraction) {
    double Rv = cfg->molar_gas_constant / cfg->vapor_molar_mass;
    double dcl = cfg->vapor_heat_capacity - cfg->liquid_heat_capacity;
    double dci = cfg->vapor_heat_capacity - cfg->ice_heat_capacity;
    double Tr = cfg->energy_reference_temperature;
    double L0l = cfg->liquid_reference_latent_heat - dcl * Tr;
    double L0i = cfg->ice_reference_latent_heat - dci * Tr;
    double lam = liquid_fraction;
    double L0 = lam * L0l + (1 - lam) * L0i;
    double dc = lam * dcl + (1 - lam) * dci;
    double Ttr = cfg->triple_point_temperature, ptr = cfg->triple_point_pressure;
    return ptr * pow(T / Ttr, dc / Rv) * exp((1 / Ttr - 1 / T) * L0 / Rv);
}

/* saturation_specific_humidity(T, ρ, constants, surface): vapor_saturation.jl:93-97 */
double orc_saturation_specific_humidity(const bz_config* cfg, double T, double rho, double liquid_fraction) {
    double Rv = cfg->molar_gas_constant / cfg->vapor_molar_mass;
    double pvs = orc_saturation_vapor_pressure(cfg, T, liquid_fraction);
    return pvs / (rho * Rv * T);
}

/* density(T, p, q, constants) = p / (Rᵐ T): thermodynamics_constants.jl:383-386 */
double orc_density(const bz_config* cfg, double T, double p, double qv) {
    double Rd = cfg->molar_gas_constant / cfg->dry_air_molar_mass;
    double Rv = cfg->molar_gas_constant / cfg->vapor_molar_mass;
    double Rm = (1 - qv) * Rd + qv * Rv;
    return p / (Rm * T);
}

/* ReferenceState closed forms for constant θ₀: reference_states.jl:88-123 (surface_density,
 * adiabatic_hydrostatic_pressure, adiabatic_hydrostatic_density) and :326-330 (hydrostatic_temperature). */
static double surface_density(const orc_ctx* c) {
    double p0 = c->cfg.surface_pressure, th0 = c->cfg.potential_temperature, pst = c->cfg.standard_pressure;
    double Pi0 = pow(p0 / pst, c->Rd / c->cpd);
    double T0 = Pi0 * th0;
    return p0 / (c->Rd * T0);
}
static double adiabatic_hydrostatic_pressure(const orc_ctx* c, double z) {
    double p0 = c->cfg.surface_pressure, th0 = c->cfg.potential_temperature, pst = c->cfg.standard_pressure;
    double T0 = th0 * pow(p0 / pst, c->Rd / c->cpd);
    return p0 * pow(1 - c->g * z / (c->cpd * T0), c->cpd / c->Rd);
}
static double adiabatic_hydrostatic_density(const orc_ctx* c, double z) {
    double p0 = c->cfg.surface_pressure;
    double pr = adiabatic_hydrostatic_pressure(c, z);
    double rho0 = surface_density(c);
    return rho0 * pow(pr / p0, 1 - c->Rd / c->cpd);
}
static double hydrostatic_temperature(const orc_ctx* c, double z) {
    double kappa = c->Rd / c->cpd;
    double p = adiabatic_hydrostatic_pressure(c, z);
    return c->cfg.potential_temperature * pow(p / c->cfg.standard_pressure, kappa);
}

/* z-only field halos: bottom ValueBoundaryCondition(v0) => c[-1] = 2 v0 - c[0]; other sides mirror
 * (reference_states.jl:425-442). Only ever multiplied by w = 0 at the walls. */
static void fill_column_halo(const orc_ctx* c, double* col, int has_value_bc, double v0) {
    int Hz = c->Hz, Nz = c->Nz;
    for (int h = 1; h <= Hz; ++h) {
        col[Hz - h] = has_value_bc ? (2 * v0 - col[Hz + h - 1]) : col[Hz + h - 1];
        col[Hz + Nz - 1 + h] = col[Hz + Nz - h];
    }
}

static void build_reference_state(orc_ctx* c) {
    for (int k = 0; k < c->Nz; ++k) {
        double z = c->cfg.z0 + (k + 0.5) * c->dz;
        c->rho_r[k + c->Hz] = adiabatic_hydrostatic_density(c, z);
        c->p_r[k + c->Hz] = adiabatic_hydrostatic_pressure(c, z);
        c->T_r[k + c->Hz] = hydrostatic_temperature(c, z);
    }
    fill_column_halo(c, c->rho_r, 1, surface_density(c));
    fill_column_halo(c, c->p_r, 1, c->cfg.surface_pressure);
    fill_column_halo(c, c->T_r, 0, 0);
}

/* ------------------------------------------------------------------------------------------------ */
/* halo fills (Oceananigans fill_halo_regions!)                                  */
/* ------------------------------------------------------------------------------------------------ */
enum { LOC_CENTER = 0, LOC_ZFACE = 1 };

static void fill_halos(const orc_ctx* c, double* f, int loc) {
    const int Nx = c->Nx, Ny = c->Ny, Nz = c->Nz, Hx = c->Hx, Hy = c->Hy, Hz = c->Hz;
    const int nzl = (loc == LOC_ZFACE) ? Nz + 1 : Nz;
    /* impenetrable walls for the wall-normal face field */
    if (loc == LOC_ZFACE) {
        for (int j = 0; j < Ny; ++j) for (int i = 0; i < Nx; ++i) { f[IDX(c, i, j, 0)] = 0; f[IDX(c, i, j, Nz)] = 0; }
    }
    /* z: zero-flux mirror for centre fields; face fields: mirror about the wall */
    (void)nzl;
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
    /* periodic x then y, over the full padded z range (corners included) */
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

#include "oracle_weno.h"
int g_beta_form = 0;
void orc_set_beta_form(int form) { g_beta_form = form; }

/* ------------------------------------------------------------------------------------------------ */
/* diagnostics: update_state! pieces                                                                 */
/* ------------------------------------------------------------------------------------------------ */

/* _compute_velocities!: update_atmosphere_model_state.jl:248-254 (launch covers the wall faces in Bounded z, :138-145) */
static void compute_velocities(orc_ctx* c) {
    fill_halos(c, c->U[BZ_RHO_U], LOC_CENTER);
    fill_halos(c, c->U[BZ_RHO_V], LOC_CENTER);
    fill_halos(c, c->U[BZ_RHO_W], LOC_ZFACE);
    const int Hz = c->Hz;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k <= c->Nz; ++k)
        for (int j = 0; j < c->Ny; ++j)
            for (int i = 0; i < c->Nx; ++i) {
                size_t n = IDX(c, i, j, k);
                double rc = c->rho_r[k + Hz];
                double rf = 0.5 * (c->rho_r[k + Hz] + c->rho_r[k - 1 + Hz]);
                if (k < c->Nz) {
                    c->u[n] = c->U[BZ_RHO_U][n] / (0.5 * (rc + rc));
                    c->v[n] = c->U[BZ_RHO_V][n] / (0.5 * (rc + rc));
                }
                c->w[n] = c->U[BZ_RHO_W][n] / rf;
            }
    fill_halos(c, c->u, LOC_CENTER);
    fill_halos(c, c->v, LOC_CENTER);
    fill_halos(c, c->w, LOC_ZFACE);
}

/* exner_function + temperature of a LiquidIcePotentialTemperatureState: dynamic_states.jl:31-58 */
static inline double lipt_temperature(const orc_ctx* c, double theta, double pr, double qv, double ql, double qi) {
    double Rm = mixture_gas_constant(c, qv, ql, qi);
    double cpm = mixture_heat_capacity(c, qv, ql, qi);
    double Pi = pow(pr / c->cfg.standard_pressure, Rm / cpm);
    return Pi * theta + (c->cfg.liquid_reference_latent_heat * ql + c->cfg.ice_reference_latent_heat * qi) / cpm;
}

/* adjustment_saturation_specific_humidity for WarmPhaseEquilibrium (vapor_saturation.jl:216-256):
 * qᵛ⁺ = ϵ (1 - qᵗ) pᵛ⁺ / (p - pᵛ⁺), ϵ = Rᵈ/Rᵛ — saturation specific humidity of a saturated parcel. */
static inline double adjustment_saturation_specific_humidity(const orc_ctx* c, double T, double pr, double qt) {
    double pvs = orc_saturation_vapor_pressure(&c->cfg, T, 1.0);
    double eps = c->Rd / c->Rv;
    return eps * (1 - qt) * pvs / (pr - pvs);
}

/* adjust_thermodynamic_state (saturation_adjustment.jl:182-231) + secant_solve (src/Solvers.jl:243-262)
 * for the warm-phase equilibrium; returns T and the adjusted (qv, ql). */
static void saturation_adjust(const orc_ctx* c, double theta, double pr, double qt, double* T_out, double* qv_out, double* ql_out) {
    const double Ll = c->cfg.liquid_reference_latent_heat;
    if (theta == 0) { *T_out = 0; *qv_out = qt; *ql_out = 0; return; }
    double T1 = lipt_temperature(c, theta, pr, qt, 0, 0);
    double rho1 = pr / (mixture_gas_constant(c, qt, 0, 0) * T1);
    double qvs1 = orc_saturation_specific_humidity(&c->cfg, T1, rho1, 1.0);
    if (qt <= qvs1) { *T_out = T1; *qv_out = qt; *ql_out = 0; return; }
    /* saturated: first guess */
    double qvp = adjustment_saturation_specific_humidity(c, T1, pr, qt);
    double ql1 = fmax(0.0, qt - qvp), qv1 = qt - ql1;
    double cpm = mixture_heat_capacity(c, qv1, ql1, 0);
    double dT = (Ll * ql1) / cpm;
    double T2 = T1 + fmax(0.01, dT / 2);
    /* residual r(T) = T - temperature(adjust_state(T)) */
#define SA_RESIDUAL(Tx, rx) do { double qs_ = adjustment_saturation_specific_humidity(c, (Tx), pr, qt); \
        double ql_ = fmax(0.0, qt - qs_), qv_ = qt - ql_; (rx) = (Tx) - lipt_temperature(c, theta, pr, qv_, ql_, 0); } while (0)
    /* secant_solve(f, solver, x₀ = T1, x₁ = T2, fallback): abstol 1e-4 (SaturationAdjustment default tolerance), maxiter 20 */
    double x1 = T1, x2 = T2, r1, r2;
    SA_RESIDUAL(x1, r1); SA_RESIDUAL(x2, r2);
    const double abstol = 1e-4; const int maxiter = 20;
    int iter = 0;
    while (fabs(r2) > abstol && iter < maxiter) {
        double slope = (x2 - x1) / (r2 - r1);
        int valid = isfinite(slope);
        if (!valid) slope = 0;
        x1 = x2; r1 = r2;
        x2 -= r2 * slope;
        SA_RESIDUAL(x2, r2);
        if (!valid) r2 = 0;
        ++iter;
    }
    x1 = x2;
#undef SA_RESIDUAL
    double qs = adjustment_saturation_specific_humidity(c, x1, pr, qt);
    double ql = fmax(0.0, qt - qs), qv = qt - ql;
    *T_out = lipt_temperature(c, theta, pr, qv, ql, 0);
    *qv_out = qv; *ql_out = ql;
}

/* _compute_auxiliary_thermodynamic_variables!: update_atmosphere_model_state.jl:256-292 with
 * compute_auxiliary_thermodynamic_variables! (potential_temperature_formulation.jl:115-123) */
static void compute_auxiliary_thermodynamics(orc_ctx* c) {
    const int Hz = c->Hz;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < c->Nz; ++k)
        for (int j = 0; j < c->Ny; ++j)
            for (int i = 0; i < c->Nx; ++i) {
                size_t n = IDX(c, i, j, k);
                double rho = c->rho_r[k + Hz];
                double theta = c->U[BZ_RHO_THETA][n] / rho;
                c->theta[n] = theta;
                double qve = c->U[BZ_RHO_Q][n] / rho;
                double pr = c->p_r[k + Hz];
                if (c->cfg.formulation == BZ_FORMULATION_STATIC_ENERGY) {
                    /* StaticEnergyFormulation (static_energy_formulation.jl:60-88): e = ρe/ρ, held in c->theta;
                     * temperature(::StaticEnergyState) = (e - g z + ℒˡqˡ + ℒⁱqⁱ)/cᵖᵐ (dynamic_states.jl:283-298), vapour only */
                    double z = c->cfg.z0 + (k + 0.5) * c->dz;
                    c->qv[n] = qve; c->ql[n] = 0;
                    c->T[n] = (theta - c->g * z + 0.0 + 0.0) / mixture_heat_capacity(c, qve, 0, 0);
                } else if (c->cfg.microphysics == BZ_MICROPHYSICS_NONE) {
                    c->qv[n] = qve; c->ql[n] = 0;
                    c->T[n] = lipt_temperature(c, theta, pr, qve, 0, 0);
                } else {
                    double T, qv, ql;
                    saturation_adjust(c, theta, pr, qve, &T, &qv, &ql);
                    c->qv[n] = qv; c->ql[n] = ql; c->T[n] = T;
                }
            }
    fill_halos(c, c->T, LOC_CENTER);
    fill_halos(c, c->qv, LOC_CENTER);
    fill_halos(c, c->ql, LOC_CENTER);
    fill_halos(c, c->theta, LOC_CENTER);
}

/* ------------------------------------------------------------------------------------------------ */
/* tendencies                                                                                        */
/* ------------------------------------------------------------------------------------------------ */

/* buoyancy_forceᶜᶜᶜ for AnelasticDynamics: src/AnelasticEquations/anelastic_buoyancy.jl:36-72
 * (reference moisture fractions are ZeroFields: reference_states.jl:412-414) */
static inline double buoyancy_ccc(const orc_ctx* c, int i, int j, int k) {
    size_t n = IDX(c, i, j, k);
    double rho_r = c->rho_r[k + c->Hz], T_r = c->T_r[k + c->Hz];
    double T = c->T[n];
    double Rm_r = mixture_gas_constant(c, 0, 0, 0);
    double Rm = mixture_gas_constant(c, c->qv[n], c->ql[n], 0);
    double rho_p = rho_r * (Rm_r * T_r / (Rm * T) - 1);
    return -c->g * rho_p;
}

/* prognostic moisture seen by the scalar advection: specific_prognostic_moisture = qᵛ (no microphysics) or qᵉ = ρqᵉ/ρ
 * (saturation adjustment, saturation_adjustment.jl:108,118) — both equal ρq/ρ, stored in c->qe below. */

typedef struct { double *qe; } aux_fields;

/* Flux helpers. A* are face areas, V the cell volume (uniform grid; a Flat dimension has spacing 1). */

/* tracer_mass_flux_{x,y,z} = ℑ(ρ) · Oceananigans _advective_tracer_flux: src/Advection.jl:20-27 */
static inline double tracer_flux_x(const orc_ctx* c, const double* cfield, int i, int j, int k) {
    size_t n = IDX(c, i, j, k);
    double ut = c->u[n];
    double rho = c->rho_r[k + c->Hz];
    double cR = biased_interp(cfield + n, 1, c->B, ut > 0);
    return (0.5 * (rho + rho)) * (c->dy * c->dz * ut * cR);
}
static inline double tracer_flux_y(const orc_ctx* c, const double* cfield, int i, int j, int k) {
    size_t n = IDX(c, i, j, k);
    double vt = c->v[n];
    double rho = c->rho_r[k + c->Hz];
    double cR = biased_interp(cfield + n, c->Px, c->B, vt > 0);
    return (0.5 * (rho + rho)) * (c->dx * c->dz * vt * cR);
}
static inline double tracer_flux_z(const orc_ctx* c, const double* cfield, int i, int j, int k) {
    if (k == 0 || k == c->Nz) return 0.0;    /* w = 0 exactly on the walls */
    size_t n = IDX(c, i, j, k);
    double wt = c->w[n];
    double rho_f = 0.5 * (c->rho_r[k + c->Hz] + c->rho_r[k - 1 + c->Hz]);
    double cR = biased_interp(cfield + n, (ptrdiff_t)c->Px * c->Py, red_face(k, c->Nz, c->B), wt > 0);
    return rho_f * (c->dx * c->dy * wt * cR);
}

/* div_ρUc: src/Advection.jl:30-35 */
static inline double div_rhoUc(const orc_ctx* c, const double* cfield, int i, int j, int k) {
    double V = c->dx * c->dy * c->dz;
    double fx = (c->cfg.topology_x == BZ_FLAT) ? 0.0 : tracer_flux_x(c, cfield, i + 1, j, k) - tracer_flux_x(c, cfield, i, j, k);
    double fy = (c->cfg.topology_y == BZ_FLAT) ? 0.0 : tracer_flux_y(c, cfield, i, j + 1, k) - tracer_flux_y(c, cfield, i, j, k);
    double fz = tracer_flux_z(c, cfield, i, j, k + 1) - tracer_flux_z(c, cfield, i, j, k);
    return (1 / V) * (fx + fy + fz);
}

/* Oceananigans advective_momentum_flux_* for an UpwindScheme (Appendix A.3): advecting = centred-4 interpolation of
 * area-weighted momentum, advected = WENO5 of velocity biased by the sign of the advecting flux. */
#define SX 1
#define SY ((ptrdiff_t)c->Px)
#define SZ ((ptrdiff_t)c->Px * c->Py)
/* interpolation along a Flat dimension is the identity (Oceananigans Flat topology) */
#define SYM_X(a, R) ((c->cfg.topology_x == BZ_FLAT) ? (a)[0] : symmetric_interp((a), SX, (R)))
#define SYM_Y(a, R) ((c->cfg.topology_y == BZ_FLAT) ? (a)[0] : symmetric_interp((a), SY, (R)))

static inline double flux_Uu(const orc_ctx* c, int i, int j, int k) {        /* at centre i */
    size_t n1 = IDX(c, i + 1, j, k);
    double ut = c->dy * c->dz * symmetric_interp(c->U[BZ_RHO_U] + n1, SX, c->Bs);
    return ut * biased_interp(c->u + n1, SX, c->B, ut > 0);
}
static inline double flux_Vu(const orc_ctx* c, int i, int j, int k) {        /* at (face i, face j) */
    size_t n = IDX(c, i, j, k);
    double vt = c->dx * c->dz * SYM_X(c->U[BZ_RHO_V] + n, c->Bs);
    return vt * biased_interp(c->u + n, SY, c->B, vt > 0);
}
static inline double flux_Wu(const orc_ctx* c, int i, int j, int k) {        /* at (face i, z-face k) */
    if (k == 0 || k == c->Nz) return 0.0;
    size_t n = IDX(c, i, j, k);
    double wt = c->dx * c->dy * SYM_X(c->U[BZ_RHO_W] + n, c->Bs);
    return wt * biased_interp(c->u + n, SZ, red_face(k, c->Nz, c->B), wt > 0);
}
static inline double flux_Uv(const orc_ctx* c, int i, int j, int k) {        /* at (face i, face j) */
    size_t n = IDX(c, i, j, k);
    double ut = c->dy * c->dz * SYM_Y(c->U[BZ_RHO_U] + n, c->Bs);
    return ut * biased_interp(c->v + n, SX, c->B, ut > 0);
}
static inline double flux_Vv(const orc_ctx* c, int i, int j, int k) {        /* at centre j */
    size_t n1 = IDX(c, i, j + 1, k);
    double vt = c->dx * c->dz * symmetric_interp(c->U[BZ_RHO_V] + n1, SY, c->Bs);
    return vt * biased_interp(c->v + n1, SY, c->B, vt > 0);
}
static inline double flux_Wv(const orc_ctx* c, int i, int j, int k) {        /* at (face j, z-face k) */
    if (k == 0 || k == c->Nz) return 0.0;
    size_t n = IDX(c, i, j, k);
    double wt = c->dx * c->dy * SYM_Y(c->U[BZ_RHO_W] + n, c->Bs);
    return wt * biased_interp(c->v + n, SZ, red_face(k, c->Nz, c->B), wt > 0);
}
static inline double flux_Uw(const orc_ctx* c, int i, int j, int k) {        /* at (face i, z-face k), 1 <= k <= Nz-1 */
    size_t n = IDX(c, i, j, k);
    double ut = c->dy * c->dz * symmetric_interp(c->U[BZ_RHO_U] + n, SZ, red_face(k, c->Nz, c->Bs));
    return ut * biased_interp(c->w + n, SX, c->B, ut > 0);
}
static inline double flux_Vw(const orc_ctx* c, int i, int j, int k) {
    size_t n = IDX(c, i, j, k);
    double vt = c->dx * c->dz * symmetric_interp(c->U[BZ_RHO_V] + n, SZ, red_face(k, c->Nz, c->Bs));
    return vt * biased_interp(c->w + n, SY, c->B, vt > 0);
}
static inline double flux_Ww(const orc_ctx* c, int i, int j, int k) {        /* at centre k, 0 <= k <= Nz-1 */
    size_t n1 = IDX(c, i, j, k + 1);
    double wt = c->dx * c->dy * symmetric_interp(c->U[BZ_RHO_W] + n1, SZ, red_center(k, c->Nz, c->Bs));
    return wt * biased_interp(c->w + n1, SZ, red_center(k, c->Nz, c->B), wt > 0);
}

/* compute_forcing!(::SubsidenceForcing): horizontal averages of the specific fields (subsidence_forcing.jl:137-141) */
static void compute_forcing_means(orc_ctx* c, const double* qe) {
    if (!c->has_forcing || !c->ws) return;
    const double* src[4] = {c->u, c->v, c->theta, qe};
    const double n = (double)c->Nx * c->Ny;
    for (int f = 0; f < 4; ++f)
        for (int k = 0; k < c->Nz; ++k) {
            double s = 0;
            for (int j = 0; j < c->Ny; ++j) for (int i = 0; i < c->Nx; ++i) s += src[f][IDX(c, i, j, k)];
            c->mean[f][k] = s / n;
        }
}

/* SubsidenceForcing kernel: -ℑzb(wˢ ∂z ϕ̄) with the one-sided top / bottom rule (subsidence_forcing.jl:84-100) */
static inline double subsidence_tendency(const orc_ctx* c, const double* phibar, int k) {
    const int Nz = c->Nz;
    double up = (k + 1 < Nz) ? c->ws[k + 1] * ((phibar[k + 1] - phibar[k]) / c->dz) : 0.0;   /* face k+1 */
    double lo = (k > 0) ? c->ws[k] * ((phibar[k] - phibar[k - 1]) / c->dz) : 0.0;            /* face k   */
    double v = (k == Nz - 1) ? lo : ((k == 0) ? up : (up + lo) / 2);
    return -v;
}

/* ρ × (sum of the specific forcings of field f = 0 u, 1 v, 2 θ, 3 q) at level k, horizontally uniform part */
static inline double column_forcing(const orc_ctx* c, int f, int k) {
    double F = 0;
    if (c->ws && (c->subsidence_mask >> f & 1)) F += subsidence_tendency(c, c->mean[f], k);
    if (f == 0 && c->vg) F += -c->coriolis_f * c->vg[k];
    if (f == 1 && c->ug) F += c->coriolis_f * c->ug[k];
    if (f == 3 && c->q_tend) F += c->q_tend[k];
    return c->rho_r[k + c->Hz] * F;
}

/* compute_tendencies!: update_atmosphere_model_state.jl:294-387 → x/y/z_momentum_tendency
 * (dynamics_kernel_functions.jl:64-130), potential_temperature_tendency (potential_temperature_tendency.jl:66-106),
 * scalar_tendency (dynamics_kernel_functions.jl:132-159). Zero terms of the configs on the path (Coriolis, closure,
 * forcing, PGF for anelastic, metric terms on a rectilinear grid) are omitted. */
static void compute_tendencies(orc_ctx* c, const double* qe) {
    const int flat_x = c->cfg.topology_x == BZ_FLAT, flat_y = c->cfg.topology_y == BZ_FLAT;
    const double V = c->dx * c->dy * c->dz;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < c->Nz; ++k)
        for (int j = 0; j < c->Ny; ++j)
            for (int i = 0; i < c->Nx; ++i) {
                size_t n = IDX(c, i, j, k);
                /* x momentum: -div_𝐯u */
                {
                    double fx = flat_x ? 0.0 : flux_Uu(c, i, j, k) - flux_Uu(c, i - 1, j, k);
                    double fy = flat_y ? 0.0 : flux_Vu(c, i, j + 1, k) - flux_Vu(c, i, j, k);
                    double fz = flux_Wu(c, i, j, k + 1) - flux_Wu(c, i, j, k);
                    c->G[BZ_RHO_U][n] = -((1 / V) * (fx + fy + fz));
                }
                /* y momentum */
                {
                    double fx = flat_x ? 0.0 : flux_Uv(c, i + 1, j, k) - flux_Uv(c, i, j, k);
                    double fy = flat_y ? 0.0 : flux_Vv(c, i, j, k) - flux_Vv(c, i, j - 1, k);
                    double fz = flux_Wv(c, i, j, k + 1) - flux_Wv(c, i, j, k);
                    c->G[BZ_RHO_V][n] = -((1 / V) * (fx + fy + fz));
                }
                /* z momentum: -div_𝐯w + ℑz(buoyancy); the wall face k = 0 is overwritten by the impenetrable halo fill */
                if (k >= 1) {
                    double fx = flat_x ? 0.0 : flux_Uw(c, i + 1, j, k) - flux_Uw(c, i, j, k);
                    double fy = flat_y ? 0.0 : flux_Vw(c, i, j + 1, k) - flux_Vw(c, i, j, k);
                    double fz = flux_Ww(c, i, j, k) - flux_Ww(c, i, j, k - 1);
                    double b = 0.5 * (buoyancy_ccc(c, i, j, k) + buoyancy_ccc(c, i, j, k - 1));
                    c->G[BZ_RHO_W][n] = -((1 / V) * (fx + fy + fz)) + b;
                } else {
                    c->G[BZ_RHO_W][n] = 0.0;
                }
                c->G[BZ_RHO_THETA][n] = -div_rhoUc(c, c->theta, i, j, k);
                if (c->cfg.formulation == BZ_FORMULATION_STATIC_ENERGY) {
                    /* static_energy_tendency (static_energy_tendency.jl:39-72): - ℑzᵃᵃᶜ(w · ℑzᵃᵃᶠ(buoyancy_forceᶜᶜᶜ)); w = 0 on the walls */
                    double bf_lo = (k > 0) ? c->w[n] * (0.5 * (buoyancy_ccc(c, i, j, k) + buoyancy_ccc(c, i, j, k - 1))) : 0.0;
                    double bf_hi = (k + 1 < c->Nz) ? c->w[IDX(c, i, j, k + 1)] * (0.5 * (buoyancy_ccc(c, i, j, k + 1) + buoyancy_ccc(c, i, j, k))) : 0.0;
                    c->G[BZ_RHO_THETA][n] -= 0.5 * (bf_lo + bf_hi);
                }
                c->G[BZ_RHO_Q][n] = -div_rhoUc(c, qe, i, j, k);
                if (c->has_forcing) {
                    const double f = c->coriolis_f;
                    const double* ru = c->U[BZ_RHO_U]; const double* rv = c->U[BZ_RHO_V];
                    /* FPlane: -x_f_cross_U = +f ℑxyᶠᶜᵃ(ρv), -y_f_cross_U = -f ℑxyᶜᶠᵃ(ρu) */
                    double rv_fc = 0.25 * (rv[IDX(c, i - 1, j, k)] + rv[n] + rv[IDX(c, i - 1, j + 1, k)] + rv[IDX(c, i, j + 1, k)]);
                    double ru_cf = 0.25 * (ru[IDX(c, i, j - 1, k)] + ru[IDX(c, i + 1, j - 1, k)] + ru[n] + ru[IDX(c, i + 1, j, k)]);
                    c->G[BZ_RHO_U][n] += f * rv_fc + column_forcing(c, 0, k);
                    c->G[BZ_RHO_V][n] += -f * ru_cf + column_forcing(c, 1, k);
                    c->G[BZ_RHO_THETA][n] += column_forcing(c, 2, k);
                    c->G[BZ_RHO_Q][n] += column_forcing(c, 3, k);
                    if (c->e_tend) {   /* (Fρe) / (cᵖᵐ Π), Fρe = ρ e_tendency */
                        double qv = c->qv[n], ql = c->ql[n];
                        double Rm = mixture_gas_constant(c, qv, ql, 0), cpm = mixture_heat_capacity(c, qv, ql, 0);
                        double Pi = pow(c->p_r[k + c->Hz] / c->cfg.standard_pressure, Rm / cpm);
                        c->G[BZ_RHO_THETA][n] += c->rho_r[k + c->Hz] * c->e_tend[k] / (cpm * Pi);
                    }
                    if (k == 0) {      /* compute_flux_bc_tendencies!: bottom fluxes, G += J Az / V */
                        c->G[BZ_RHO_THETA][n] += c->theta_flux / c->dz;
                        c->G[BZ_RHO_Q][n] += c->q_flux / c->dz;
                        if (c->drag_rho_ustar2 != 0) {
                            double a = c->drag_rho_ustar2;
                            c->G[BZ_RHO_U][n] += (-a * ru[n] / sqrt(ru[n] * ru[n] + rv_fc * rv_fc)) / c->dz;
                            c->G[BZ_RHO_V][n] += (-a * rv[n] / sqrt(ru_cf * ru_cf + rv[n] * rv[n])) / c->dz;
                        }
                    }
                }
            }
}

/* ------------------------------------------------------------------------------------------------ */
/* FFT (any length: mixed radix, O(n·Σp)); the reference uses FFTW through Oceananigans               */
/* ------------------------------------------------------------------------------------------------ */
static void fft_plan_init(fft_plan* p, int n) {
    p->n = n;
    p->w = (cplx*)malloc(sizeof(cplx) * (size_t)(n > 0 ? n : 1));
    for (int k = 0; k < n; ++k) {
        double a = -2.0 * M_PI * (double)k / (double)n;
        p->w[k] = cos(a) + I * sin(a);
    }
}
static void fft_plan_free(fft_plan* p) { free(p->w); p->w = NULL; }

/* out[0..n) = DFT of in[0], in[s], ...; sign = -1 forward / +1 backward (unnormalised). tw_stride = N/n. */
static void fft_rec(const fft_plan* p, int n, const cplx* in, ptrdiff_t s, cplx* out, int sign) {
    if (n == 1) { out[0] = in[0]; return; }
    int r = 2;
    while (r * r <= n && n % r) ++r;
    if (n % r) r = n;                  /* prime */
    if (n % 4 == 0) r = 4;
    int m = n / r;
    for (int q = 0; q < r; ++q) fft_rec(p, m, in + q * s, s * r, out + q * m, sign);
    int tws = p->n / n;
    cplx tmp[64];
    cplx* t = (r <= 64) ? tmp : (cplx*)malloc(sizeof(cplx) * (size_t)r);
    for (int k = 0; k < m; ++k) {
        for (int q = 0; q < r; ++q) {
            int e = (int)(((long long)q * k * tws) % p->n);
            cplx w = p->w[e];
            if (sign > 0) w = conj(w);
            t[q] = out[q * m + k] * w;
        }
        for (int a = 0; a < r; ++a) {
            cplx acc = 0;
            for (int q = 0; q < r; ++q) {
                int e = (int)(((long long)a * q * m * tws) % p->n);
                cplx w = p->w[e];
                if (sign > 0) w = conj(w);
                acc += t[q] * w;
            }
            out[a * m + k] = acc;
        }
    }
    if (t != tmp) free(t);
}

/* in-place strided transform of `count` lines */
static void fft_lines(const fft_plan* p, cplx* data, ptrdiff_t elem_stride, int sign, cplx* scratch_in, cplx* scratch_out) {
    int n = p->n;
    for (int a = 0; a < n; ++a) scratch_in[a] = data[a * elem_stride];
    fft_rec(p, n, scratch_in, 1, scratch_out, sign);
    for (int a = 0; a < n; ++a) data[a * elem_stride] = scratch_out[a];
}

static void fft_xy(orc_ctx* c, cplx* a, int sign) {
    const int Nx = c->Nx, Ny = c->Ny, Nz = c->Nz;
    if (c->cfg.topology_x == BZ_PERIODIC && Nx > 1) {
#pragma omp parallel
        {
            cplx* si = (cplx*)malloc(sizeof(cplx) * Nx); cplx* so = (cplx*)malloc(sizeof(cplx) * Nx);
#pragma omp for collapse(2) schedule(static)
            for (int k = 0; k < Nz; ++k) for (int j = 0; j < Ny; ++j)
                fft_lines(&c->px, a + (size_t)Nx * (j + (size_t)Ny * k), 1, sign, si, so);
            free(si); free(so);
        }
    }
    if (c->cfg.topology_y == BZ_PERIODIC && Ny > 1) {
#pragma omp parallel
        {
            cplx* si = (cplx*)malloc(sizeof(cplx) * Ny); cplx* so = (cplx*)malloc(sizeof(cplx) * Ny);
#pragma omp for collapse(2) schedule(static)
            for (int k = 0; k < Nz; ++k) for (int i = 0; i < Nx; ++i)
                fft_lines(&c->py, a + i + (size_t)Nx * Ny * k, Nx, sign, si, so);
            free(si); free(so);
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* anelastic pressure solver                                                                         */
/* ------------------------------------------------------------------------------------------------ */

/* Oceananigans poisson_eigenvalues (Appendix A.4): Periodic λ_i = (2 sin(π i / N) / Δ)², Flat 0. */
static void build_solver(orc_ctx* c) {
    const int Nx = c->Nx, Ny = c->Ny, Nz = c->Nz, Hz = c->Hz;
    for (int i = 0; i < Nx; ++i) {
        double s = 2 * sin(M_PI * i / Nx) / c->dx;
        c->lam_x[i] = (c->cfg.topology_x == BZ_FLAT) ? 0.0 : s * s;
    }
    for (int j = 0; j < Ny; ++j) {
        double s = 2 * sin(M_PI * j / Ny) / c->dy;
        c->lam_y[j] = (c->cfg.topology_y == BZ_FLAT) ? 0.0 : s * s;
    }
    /* _compute_anelastic_lower_diagonal!: anelastic_pressure_solver.jl:72-78 (lower == upper, symmetric) */
    for (int k = 0; k < Nz - 1; ++k) {
        double rf = 0.5 * (c->rho_r[k + 1 + Hz] + c->rho_r[k + Hz]);
        c->lower[k] = rf / c->dz;
    }
    /* _compute_anelastic_main_diagonal!: anelastic_pressure_solver.jl:39-62 */
    for (int j = 0; j < Ny; ++j)
        for (int i = 0; i < Nx; ++i) {
            double lam = c->lam_x[i] + c->lam_y[j];
            for (int k = 0; k < Nz; ++k) {
                double rk = c->rho_r[k + Hz];
                double d;
                if (Nz == 1) {
                    d = -rk * c->dz * lam;
                } else if (k == 0) {
                    double r2 = 0.5 * (c->rho_r[1 + Hz] + c->rho_r[0 + Hz]);
                    d = -r2 / c->dz - rk * c->dz * lam;
                } else if (k == Nz - 1) {
                    double rN = 0.5 * (c->rho_r[Nz - 1 + Hz] + c->rho_r[Nz - 2 + Hz]);
                    d = -rN / c->dz - rk * c->dz * lam;
                } else {
                    double rp = 0.5 * (c->rho_r[k + 1 + Hz] + c->rho_r[k + Hz]);
                    double rm = 0.5 * (c->rho_r[k + Hz] + c->rho_r[k - 1 + Hz]);
                    d = -(rp / c->dz + rm / c->dz) - rk * c->dz * lam;
                }
                c->diag[i + (size_t)Nx * (j + (size_t)Ny * k)] = d;
            }
        }
}

/* compute_pressure_correction!: anelastic_time_stepping.jl:26-39 with solve_for_anelastic_pressure!
 * (anelastic_pressure_solver.jl:84-105) and Oceananigans solve!(::FourierTridiagonalPoissonSolver) (Appendix A.4). */
static void compute_pressure_correction(orc_ctx* c, double dt) {
    const int Nx = c->Nx, Ny = c->Ny, Nz = c->Nz;
    const int flat_x = c->cfg.topology_x == BZ_FLAT, flat_y = c->cfg.topology_y == BZ_FLAT;
    fill_halos(c, c->U[BZ_RHO_U], LOC_CENTER);
    fill_halos(c, c->U[BZ_RHO_V], LOC_CENTER);
    fill_halos(c, c->U[BZ_RHO_W], LOC_ZFACE);
    const double V = c->dx * c->dy * c->dz;
    /* _compute_anelastic_source_term!: rhs = Δzᶜ · divᶜᶜᶜ(ρu, ρv, ρw) / Δt */
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) {
                double dxu = flat_x ? 0.0 : c->dy * c->dz * c->U[BZ_RHO_U][IDX(c, i + 1, j, k)] - c->dy * c->dz * c->U[BZ_RHO_U][IDX(c, i, j, k)];
                double dyv = flat_y ? 0.0 : c->dx * c->dz * c->U[BZ_RHO_V][IDX(c, i, j + 1, k)] - c->dx * c->dz * c->U[BZ_RHO_V][IDX(c, i, j, k)];
                double dzw = c->dx * c->dy * c->U[BZ_RHO_W][IDX(c, i, j, k + 1)] - c->dx * c->dy * c->U[BZ_RHO_W][IDX(c, i, j, k)];
                double div = (1 / V) * (dxu + dyv + dzw);
                c->rhs[i + (size_t)Nx * (j + (size_t)Ny * k)] = c->dz * div / dt;
            }
    fft_xy(c, c->rhs, -1);
    /* batched Thomas in z, symmetric off-diagonals a = c = lower; stale-value elision when |β| < 10 eps */
#pragma omp parallel
    {
        double* t = (double*)malloc(sizeof(double) * (size_t)Nz);
#pragma omp for collapse(2) schedule(static)
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) {
                size_t col = i + (size_t)Nx * j, st = (size_t)Nx * Ny;
                double beta = c->diag[col];
                c->sol[col] = c->rhs[col] / beta;
                for (int k = 1; k < Nz; ++k) {
                    double ck = c->lower[k - 1], ak = c->lower[k - 1];
                    t[k] = ck / beta;
                    beta = c->diag[col + st * k] - ak * t[k];
                    cplx star = (c->rhs[col + st * k] - ak * c->sol[col + st * (k - 1)]) / beta;
                    if (fabs(beta) > 10 * 2.220446049250313e-16) c->sol[col + st * k] = star;
                    /* else: keep the stale value (the singular (0,0) mode; removed by the mean subtraction) */
                }
                for (int k = Nz - 2; k >= 0; --k) c->sol[col + st * k] -= t[k + 1] * c->sol[col + st * (k + 1)];
            }
        free(t);
    }
    fft_xy(c, c->sol, +1);
    /* backward transform normalisation, then φ .= φ .- mean(φ) */
    double norm = 1.0;
    if (!flat_x) norm *= Nx;
    if (!flat_y) norm *= Ny;
    size_t N = (size_t)Nx * Ny * Nz;
    cplx mean = 0;
    for (size_t a = 0; a < N; ++a) { c->sol[a] /= norm; mean += c->sol[a]; }
    mean /= (double)N;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) {
                size_t a = i + (size_t)Nx * (j + (size_t)Ny * k);
                c->sol[a] -= mean;
                c->phi[IDX(c, i, j, k)] = creal(c->sol[a]);
            }
    fill_halos(c, c->phi, LOC_CENTER);
}

/* _pressure_correct_momentum!: anelastic_time_stepping.jl:45-54 */
static void make_pressure_correction(orc_ctx* c, double dt) {
    const int Hz = c->Hz;
    const int flat_x = c->cfg.topology_x == BZ_FLAT, flat_y = c->cfg.topology_y == BZ_FLAT;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k < c->Nz; ++k)
        for (int j = 0; j < c->Ny; ++j)
            for (int i = 0; i < c->Nx; ++i) {
                size_t n = IDX(c, i, j, k);
                double rf = 0.5 * (c->rho_r[k + Hz] + c->rho_r[k - 1 + Hz]);
                double rc = c->rho_r[k + Hz];
                if (!flat_x) c->U[BZ_RHO_U][n] -= rc * dt * ((c->phi[n] - c->phi[IDX(c, i - 1, j, k)]) / c->dx);
                if (!flat_y) c->U[BZ_RHO_V][n] -= rc * dt * ((c->phi[n] - c->phi[IDX(c, i, j - 1, k)]) / c->dy);
                c->U[BZ_RHO_W][n] -= rf * dt * ((c->phi[n] - c->phi[IDX(c, i, j, k - 1)]) / c->dz);
            }
}

/* ------------------------------------------------------------------------------------------------ */
/* update_state! and the SSP-RK3 driver                                                              */
/* ------------------------------------------------------------------------------------------------ */

/* update_state!: update_atmosphere_model_state.jl:41-68 */
static void update_state(orc_ctx* c, int with_tendencies) {
    for (int f = 0; f < NPROG; ++f) fill_halos(c, c->U[f], f == BZ_RHO_W ? LOC_ZFACE : LOC_CENTER);
    compute_velocities(c);
    compute_auxiliary_thermodynamics(c);
    if (with_tendencies) {
        /* specific prognostic moisture ρq/ρ with halos */
        double* qe = new_field(c);
        for (int k = 0; k < c->Nz; ++k) for (int j = 0; j < c->Ny; ++j) for (int i = 0; i < c->Nx; ++i) {
            size_t n = IDX(c, i, j, k);
            qe[n] = c->U[BZ_RHO_Q][n] / c->rho_r[k + c->Hz];
        }
        fill_halos(c, qe, LOC_CENTER);
        compute_forcing_means(c, qe);
        compute_tendencies(c, qe);
        free(qe);
    }
}

/* _ssp_rk3_substep!: ssp_runge_kutta_3.jl:167-173 */
static void ssp_rk3_substep(orc_ctx* c, double dt, double alpha) {
    for (int f = 0; f < NPROG; ++f) {
        double *u = c->U[f], *u0 = c->U0[f], *G = c->G[f];
#pragma omp parallel for collapse(2) schedule(static)
        for (int k = 0; k < c->Nz; ++k)
            for (int j = 0; j < c->Ny; ++j)
                for (int i = 0; i < c->Nx; ++i) {
                    size_t n = IDX(c, i, j, k);
                    u[n] = (1 - alpha) * u0[n] + alpha * (u[n] + dt * G[n]);
                }
    }
}

/* time_step!: ssp_runge_kutta_3.jl:209-278 */
static void time_step(orc_ctx* c, double dt) {
    if (c->iteration == 0 || c->stale) { update_state(c, 1); c->stale = 0; }   /* maybe_prepare_first_time_step! */
    for (int f = 0; f < NPROG; ++f) memcpy(c->U0[f], c->U[f], c->n_padded * sizeof(double));   /* store_initial_state! */
    const double alphas[3] = {1.0, 1.0 / 4.0, 2.0 / 3.0};
    for (int s = 0; s < 3; ++s) {
        double a = alphas[s];
        ssp_rk3_substep(c, dt, a);
        compute_pressure_correction(c, a * dt);
        make_pressure_correction(c, a * dt);
        update_state(c, 1);
    }
    c->time += dt;
    c->iteration += 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* C ABI (mirror of include/breeze_b200.h with the orc_ prefix)                                      */
/* ------------------------------------------------------------------------------------------------ */
const char* orc_last_error(const orc_ctx* c) { return c ? c->err : g_create_err; }

int orc_create(const bz_config* cfg, orc_ctx** out) {
    if (!cfg || !out) return BZ_ERR_INVALID;
    if (cfg->abi_version != BZ_ABI_VERSION) { set_err(NULL, "abi_version mismatch"); return BZ_ERR_INVALID; }
    if (cfg->Nx < 1 || cfg->Ny < 1 || cfg->Nz < 1) { set_err(NULL, "grid size must be positive"); return BZ_ERR_INVALID; }
    /* the oracle also restates WENO(order = 7 / 9) (not yet on the CUDA path, which rejects it) */
    if (cfg->advection_order != 5 && cfg->advection_order != 7 && cfg->advection_order != 9) { set_err(NULL, "WENO(order = 5, 7 or 9)"); return BZ_ERR_UNSUPPORTED; }
    if (cfg->formulation == BZ_FORMULATION_STATIC_ENERGY && cfg->microphysics != BZ_MICROPHYSICS_NONE) {
        set_err(NULL, "StaticEnergyFormulation is on the path without microphysics only"); return BZ_ERR_UNSUPPORTED; }
    if ((cfg->topology_x == BZ_FLAT && cfg->Nx != 1) || (cfg->topology_y == BZ_FLAT && cfg->Ny != 1)) {
        set_err(NULL, "a Flat dimension must have size 1"); return BZ_ERR_INVALID; }
    orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
    if (!c) return BZ_ERR_NOMEM;
    c->cfg = *cfg;
    c->Nx = cfg->Nx; c->Ny = cfg->Ny; c->Nz = cfg->Nz;
    c->Hx = cfg->topology_x == BZ_FLAT ? 0 : HALO(cfg->advection_order);
    c->Hy = cfg->topology_y == BZ_FLAT ? 0 : HALO(cfg->advection_order);
    c->Hz = HALO(cfg->advection_order);
    c->B = (cfg->advection_order + 1) / 2; c->Bs = c->B - 1;
    c->Px = c->Nx + 2 * c->Hx; c->Py = c->Ny + 2 * c->Hy; c->Pz = c->Nz + 1 + 2 * c->Hz;
    c->n_padded = (size_t)c->Px * c->Py * c->Pz;
    c->dx = cfg->topology_x == BZ_FLAT ? 1.0 : (cfg->x1 - cfg->x0) / cfg->Nx;
    c->dy = cfg->topology_y == BZ_FLAT ? 1.0 : (cfg->y1 - cfg->y0) / cfg->Ny;
    c->dz = (cfg->z1 - cfg->z0) / cfg->Nz;
    c->Rd = cfg->molar_gas_constant / cfg->dry_air_molar_mass;
    c->Rv = cfg->molar_gas_constant / cfg->vapor_molar_mass;
    c->cpd = cfg->dry_air_heat_capacity; c->cpv = cfg->vapor_heat_capacity;
    c->g = cfg->gravitational_acceleration;
    size_t ncol = (size_t)c->Nz + 2 * c->Hz + 1;
    c->rho_r = (double*)calloc(ncol, sizeof(double));
    c->p_r = (double*)calloc(ncol, sizeof(double));
    c->T_r = (double*)calloc(ncol, sizeof(double));
    for (int f = 0; f < NPROG; ++f) { c->U[f] = new_field(c); c->U0[f] = new_field(c); c->G[f] = new_field(c); }
    c->u = new_field(c); c->v = new_field(c); c->w = new_field(c); c->theta = new_field(c);
    c->qv = new_field(c); c->ql = new_field(c); c->T = new_field(c); c->phi = new_field(c);
    size_t N = (size_t)c->Nx * c->Ny * c->Nz;
    c->rhs = (cplx*)calloc(N, sizeof(cplx)); c->sol = (cplx*)calloc(N, sizeof(cplx));
    c->lam_x = (double*)calloc(c->Nx, sizeof(double)); c->lam_y = (double*)calloc(c->Ny, sizeof(double));
    c->lower = (double*)calloc(c->Nz > 1 ? c->Nz - 1 : 1, sizeof(double));
    c->diag = (double*)calloc(N, sizeof(double));
    fft_plan_init(&c->px, c->Nx); fft_plan_init(&c->py, c->Ny);
    build_reference_state(c);
    build_solver(c);
    /* initialize_model_thermodynamics!: θ = θ₀ (anelastic_time_stepping.jl:15-19) */
    for (int k = 0; k < c->Nz; ++k) for (int j = 0; j < c->Ny; ++j) for (int i = 0; i < c->Nx; ++i) {
        double th = cfg->potential_temperature;
        if (cfg->formulation == BZ_FORMULATION_STATIC_ENERGY) {
            /* set!(model, θ = θ₀) → _energy_density_from_potential_temperature! (static_energy_tendency.jl:113-146): e = cᵖᵐ Π θ + g z */
            double T = lipt_temperature(c, th, c->p_r[k + c->Hz], 0, 0, 0);
            th = mixture_heat_capacity(c, 0, 0, 0) * T + c->g * (cfg->z0 + (k + 0.5) * c->dz) - 0.0 - 0.0;
        }
        c->U[BZ_RHO_THETA][IDX(c, i, j, k)] = c->rho_r[k + c->Hz] * th;
    }
    update_state(c, 0);
    *out = c;
    return BZ_OK;
}

void orc_destroy(orc_ctx* c) {
    if (!c) return;
    free(c->rho_r); free(c->p_r); free(c->T_r);
    for (int f = 0; f < NPROG; ++f) { free(c->U[f]); free(c->U0[f]); free(c->G[f]); }
    free(c->u); free(c->v); free(c->w); free(c->theta); free(c->qv); free(c->ql); free(c->T); free(c->phi);
    free(c->rhs); free(c->sol); free(c->lam_x); free(c->lam_y); free(c->lower); free(c->diag);
    fft_plan_free(&c->px); fft_plan_free(&c->py);
    free(c->ws); free(c->ug); free(c->vg); free(c->q_tend); free(c->e_tend);
    for (int f = 0; f < 4; ++f) free(c->mean[f]);
    free(c);
}

int orc_get_reference_state(orc_ctx* c, double* rho, double* p, double* T) {
    for (int k = 0; k < c->Nz; ++k) {
        if (rho) rho[k] = c->rho_r[k + c->Hz];
        if (p) p[k] = c->p_r[k + c->Hz];
        if (T) T[k] = c->T_r[k + c->Hz];
    }
    return BZ_OK;
}

int orc_set_reference_state(orc_ctx* c, const double* rho, const double* p, const double* T) {
    for (int k = 0; k < c->Nz; ++k) {
        if (rho) c->rho_r[k + c->Hz] = rho[k];
        if (p) c->p_r[k + c->Hz] = p[k];
        if (T) c->T_r[k + c->Hz] = T[k];
    }
    /* halos: keep the construction-time boundary values (they only ever multiply w = 0) */
    if (rho) fill_column_halo(c, c->rho_r, 0, 0);
    if (p) fill_column_halo(c, c->p_r, 0, 0);
    if (T) fill_column_halo(c, c->T_r, 0, 0);
    build_solver(c);
    return BZ_OK;
}

static void copy_in(orc_ctx* c, double* dst, const double* src, int nzl) {
    for (int k = 0; k < nzl; ++k) for (int j = 0; j < c->Ny; ++j) for (int i = 0; i < c->Nx; ++i)
        dst[IDX(c, i, j, k)] = src[i + (size_t)c->Nx * (j + (size_t)c->Ny * k)];
}
static void copy_out(const orc_ctx* c, double* dst, const double* src, int nzl) {
    for (int k = 0; k < nzl; ++k) for (int j = 0; j < c->Ny; ++j) for (int i = 0; i < c->Nx; ++i)
        dst[i + (size_t)c->Nx * (j + (size_t)c->Ny * k)] = src[IDX(c, i, j, k)];
}

/* set!(model; ...): set_atmosphere_model.jl:198-360, enforce_mass_conservation! :121-128 */
int orc_set_state(orc_ctx* c, const double* ru, const double* rv, const double* rw, const double* rth, const double* rq, int enforce) {
    if (ru) copy_in(c, c->U[BZ_RHO_U], ru, c->Nz);
    if (rv) copy_in(c, c->U[BZ_RHO_V], rv, c->Nz);
    if (rw) copy_in(c, c->U[BZ_RHO_W], rw, c->Nz + 1);
    if (rth) copy_in(c, c->U[BZ_RHO_THETA], rth, c->Nz);
    if (rq) copy_in(c, c->U[BZ_RHO_Q], rq, c->Nz);
    update_state(c, 0);
    if (enforce) {
        compute_pressure_correction(c, 1.0);
        make_pressure_correction(c, 1.0);
        update_state(c, 0);
    }
    c->stale = 1;        /* tendencies are recomputed by the next time_step! */
    return BZ_OK;
}

static void replace_profile(double** dst, const double* src, int n) {
    free(*dst); *dst = NULL;
    if (src) { *dst = (double*)malloc(sizeof(double) * (size_t)n); memcpy(*dst, src, sizeof(double) * (size_t)n); }
}

int orc_set_forcing(orc_ctx* c, const bz_forcing* F) {
    if (!F) { c->has_forcing = 0; c->stale = 1; return BZ_OK; }
    if (c->cfg.formulation == BZ_FORMULATION_STATIC_ENERGY) { set_err(c, "forcings are on the path for the potential-temperature formulation only"); return BZ_ERR_UNSUPPORTED; }
    c->has_forcing = 1;
    c->coriolis_f = F->coriolis_f; c->theta_flux = F->theta_flux; c->q_flux = F->q_flux; c->drag_rho_ustar2 = F->drag_rho_ustar2;
    c->subsidence_mask = F->subsidence_mask;
    replace_profile(&c->ws, F->subsidence_w, c->Nz + 1);
    replace_profile(&c->ug, F->geostrophic_u, c->Nz);
    replace_profile(&c->vg, F->geostrophic_v, c->Nz);
    replace_profile(&c->q_tend, F->q_tendency, c->Nz);
    replace_profile(&c->e_tend, F->e_tendency, c->Nz);
    for (int f = 0; f < 4; ++f) if (!c->mean[f]) c->mean[f] = (double*)calloc((size_t)c->Nz, sizeof(double));
    c->stale = 1;          /* tendencies are recomputed by the next time_step! */
    return BZ_OK;
}

int orc_time_step(orc_ctx* c, double dt) { time_step(c, dt); return BZ_OK; }
int orc_time_steps(orc_ctx* c, double dt, int n) { for (int s = 0; s < n; ++s) time_step(c, dt); return BZ_OK; }

int orc_compute_tendencies(orc_ctx* c) { update_state(c, 1); return BZ_OK; }

int orc_get_tendency(orc_ctx* c, int f, double* out) {
    if (f < 0 || f >= NPROG) return BZ_ERR_INVALID;
    if (f == BZ_RHO_W) {
        copy_out(c, out, c->G[f], c->Nz);
        memset(out + (size_t)c->Nx * c->Ny * c->Nz, 0, sizeof(double) * (size_t)c->Nx * c->Ny);
    } else copy_out(c, out, c->G[f], c->Nz);
    return BZ_OK;
}

int orc_pressure_correct(orc_ctx* c, double dt) {
    compute_pressure_correction(c, dt);
    make_pressure_correction(c, dt);
    update_state(c, 0);
    return BZ_OK;
}

int orc_get_field(orc_ctx* c, int f, double* out) {
    const double* src = NULL; int nzl = c->Nz;
    switch (f) {
        case BZ_RHO_U: case BZ_RHO_V: case BZ_RHO_THETA: case BZ_RHO_Q: src = c->U[f]; break;
        case BZ_RHO_W: src = c->U[f]; nzl = c->Nz + 1; break;
        case BZ_U: src = c->u; break;
        case BZ_V: src = c->v; break;
        case BZ_W: src = c->w; nzl = c->Nz + 1; break;
        case BZ_THETA: src = c->theta; break;
        case BZ_QV: src = c->qv; break;
        case BZ_QL: src = c->ql; break;
        case BZ_T: src = c->T; break;
        case BZ_PHI: src = c->phi; break;
        default: return BZ_ERR_INVALID;
    }
    copy_out(c, out, src, nzl);
    return BZ_OK;
}

int orc_get_state(orc_ctx* c, double* ru, double* rv, double* rw, double* rth, double* rq) {
    if (ru) orc_get_field(c, BZ_RHO_U, ru);
    if (rv) orc_get_field(c, BZ_RHO_V, rv);
    if (rw) orc_get_field(c, BZ_RHO_W, rw);
    if (rth) orc_get_field(c, BZ_RHO_THETA, rth);
    if (rq) orc_get_field(c, BZ_RHO_Q, rq);
    return BZ_OK;
}

int orc_get_clock(orc_ctx* c, double* time, int64_t* iteration) {
    if (time) *time = c->time;
    if (iteration) *iteration = c->iteration;
    return BZ_OK;
}

/* cell_advection_timescale: src/AtmosphereModels/cell_advection_timescale.jl:46-65 */
int orc_cell_advection_timescale(orc_ctx* c, double* tau) {
    double mx = 0;
    for (int k = 0; k < c->Nz; ++k) for (int j = 0; j < c->Ny; ++j) for (int i = 0; i < c->Nx; ++i) {
        size_t n = IDX(c, i, j, k);
        double s = 0;
        if (c->cfg.topology_x != BZ_FLAT) s += fabs(c->u[n]) / c->dx;
        if (c->cfg.topology_y != BZ_FLAT) s += fabs(c->v[n]) / c->dy;
        s += fabs(c->w[n]) / c->dz;
        if (s > mx) mx = s;
    }
    *tau = 1 / mx;
    return BZ_OK;
}

int orc_max_abs_divergence(orc_ctx* c, double* out) {
    double mx = 0;
    const int flat_x = c->cfg.topology_x == BZ_FLAT, flat_y = c->cfg.topology_y == BZ_FLAT;
    fill_halos(c, c->U[BZ_RHO_U], LOC_CENTER);
    fill_halos(c, c->U[BZ_RHO_V], LOC_CENTER);
    fill_halos(c, c->U[BZ_RHO_W], LOC_ZFACE);
    for (int k = 0; k < c->Nz; ++k) for (int j = 0; j < c->Ny; ++j) for (int i = 0; i < c->Nx; ++i) {
        double d = 0;
        if (!flat_x) d += (c->U[BZ_RHO_U][IDX(c, i + 1, j, k)] - c->U[BZ_RHO_U][IDX(c, i, j, k)]) / c->dx;
        if (!flat_y) d += (c->U[BZ_RHO_V][IDX(c, i, j + 1, k)] - c->U[BZ_RHO_V][IDX(c, i, j, k)]) / c->dy;
        d += (c->U[BZ_RHO_W][IDX(c, i, j, k + 1)] - c->U[BZ_RHO_W][IDX(c, i, j, k)]) / c->dz;
        if (fabs(d) > mx) mx = fabs(d);
    }
    *out = mx;
    return BZ_OK;
}

/* NaNChecker on the first prognostic field and the others (atmosphere_model.jl:561-572) */
int orc_state_is_finite(orc_ctx* c, int* finite) {
    int ok = 1;
    for (int f = 0; f < NPROG; ++f)
        for (int k = 0; k < c->Nz; ++k) for (int j = 0; j < c->Ny; ++j) for (int i = 0; i < c->Nx; ++i)
            if (!isfinite(c->U[f][IDX(c, i, j, k)])) ok = 0;
    *finite = ok;
    return BZ_OK;
}

int orc_get_slice(orc_ctx* c, int field, int axis, int index, double* out) {
    if (axis < 0 || axis > 2) return BZ_ERR_INVALID;
    const int nzl = (field == BZ_RHO_W || field == BZ_W) ? c->Nz + 1 : c->Nz;
    double* full = (double*)malloc(sizeof(double) * (size_t)c->Nx * c->Ny * nzl);
    int rc = orc_get_field(c, field, full);
    if (rc == BZ_OK) {
        const int n0 = axis == 0 ? c->Ny : c->Nx, n1 = axis == 2 ? c->Ny : nzl;
        const int lim = axis == 0 ? c->Nx : (axis == 1 ? c->Ny : nzl);
        if (index < 0 || index >= lim) rc = BZ_ERR_INVALID;
        else for (int b = 0; b < n1; ++b) for (int a = 0; a < n0; ++a) {
            int i = axis == 0 ? index : a, j = axis == 0 ? a : (axis == 1 ? index : b), k = axis == 2 ? index : b;
            out[(size_t)b * n0 + a] = full[((size_t)k * c->Ny + j) * c->Nx + i];
        }
    }
    free(full);
    return rc;
}

int orc_synchronize(orc_ctx* c) { (void)c; return BZ_OK; }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* Stand-alone reconstructions exported for unit tests of the GPU device functions */
double orc_weno5_biased(const double* s) { return weno5_biased(s[0], s[1], s[2], s[3], s[4]); }
double orc_weno3_biased(const double* s) { return weno3_biased(s[0], s[1], s[2]); }
/* order = 5, 7, 9: left-biased value at the face between w[R-1] and w[R] from the window w[0 .. 2R-2], R = (order + 1) / 2 */
double orc_weno_biased_window(const double* w, int order) {
    int R = (order + 1) / 2;
    if (R == 3) return weno5_biased(w[0], w[1], w[2], w[3], w[4]);
    return weno_hi_window(w, R);
}
/* Centered(order) value at the face between a[order/2 - 1] and a[order/2] from the window a[0 .. order-1] */
double orc_centered_window(const double* a, int order) { return symmetric_interp(a + order / 2, 1, order / 2); }
