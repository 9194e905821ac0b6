// compressible_api.cuh — the C ABI of include/breeze_b200_compressible.h: context, ExnerReferenceState, orchestration of one
// WS-RK3 step with acoustic substepping (reference: src/TimeSteppers/acoustic_runge_kutta_3.jl:264-319,172-208;
// src/CompressibleEquations/acoustic_substepping.jl:1404-1590). Included at the end of api.cu (one translation unit).
#pragma once
#include "../../include/breeze_b200_compressible.h"
#include "compressible.cuh"

enum { CF_RHO = 0, CF_RU = 1, CF_RV = 2, CF_RW = 3, CF_RTH = 4 };
#define C_NFAM 8

struct bzc_ctx {
    bzc_config cfg;
    Layout L;
    CEos eos;
    int has_ref = 0;
    int buf = 3;                                       // buffer of the WENO scheme: (order + 1) / 2
    std::vector<double> h_p, h_rho, h_pi, h_theta;     // ExnerReferenceState columns (host, Nz)
    double* d_cols = nullptr;                          // p_r | rho_r | sponge rate·ramp at the Nz + 1 faces, on the device
    size_t fsize = 0;                                  // doubles per field: plane * (Nz + 1)
    double* arena = nullptr;
    double *U[5] = {}, *U0[5] = {}, *G[5] = {}, *P[5] = {};          // prognostics, step-start copy, slow tendencies, perturbations
    double *u = nullptr, *v = nullptr, *w = nullptr, *theta = nullptr, *T = nullptr, *p = nullptr;
    double *PiL = nullptr, *thL = nullptr, *CL = nullptr;
    double *rho_s = nullptr, *rth_s = nullptr, *rth_old = nullptr, *tfac = nullptr;
    double *avg[3] = {};
    double* Gs_rw = nullptr;
    int moist = 0;                                     // set once bzc_set_state receives a moisture density
    double *rqv = nullptr, *rqv0 = nullptr, *Grqv = nullptr, *qv = nullptr, *rho_tot = nullptr;
    double* dense = nullptr;
    cudaStream_t stream = nullptr;
    double time = 0.0;                   // BZ_KEEP_F64 (the clock stays double in the Float32 build)
    int64_t iteration = 0, launches = 0, bytes = 0;
    int prof_on = 0;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<cudaEvent_t> prof_pool;  // recycled events (see ProfScope in api.cu)
    StepGraphCache graphs;               // captured time steps (common.cuh)
    std::vector<int> prof_fam;
    double prof_ms[C_NFAM] = {};
    int64_t prof_n[C_NFAM] = {};
    char err[512] = {};
};

static char gc_err[512];
static void bzc_set_error(bzc_ctx* c, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c ? c->err : gc_err, 512, fmt, ap);
    va_end(ap);
}
#define CC_TRY(ctx, call)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            bzc_set_error((ctx), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return BZ_ERR_CUDA;                                                                       \
        }                                                                                             \
    } while (0)

struct CProfScope {
    bzc_ctx* c; int fam; cudaEvent_t a = nullptr, b = nullptr; bool on;
    static cudaEvent_t take(bzc_ctx* c) {
        cudaEvent_t e = nullptr;
        if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); } else cudaEventCreate(&e);
        return e;
    }
    CProfScope(bzc_ctx* c_, int fam_) : c(c_), fam(fam_), on(c_->prof_on && c_->prof_fam.size() < PROF_MAX_SCOPES) {
        if (!on) return;
        a = take(c); b = take(c);
        cudaEventRecord(a, c->stream);
    }
    ~CProfScope() {
        if (!on) return;
        cudaEventRecord(b, c->stream);
        c->prof_ev.push_back(a); c->prof_ev.push_back(b); c->prof_fam.push_back(fam);
    }
};

// ExnerReferenceState, isentropic path (src/Thermodynamics/reference_states.jl:572-672), dry: Rᵐ = Rᵈ, cᵖᵐ = cᵖᵈ. Host code:
// Nz values, built once; the same glibc pow the CPU restatement uses.
static void c_build_exner_reference(bzc_ctx* c) {
    const int Nz = c->L.Nz;
    const double Rm = c->eos.Rd, cpm = c->eos.cpd, kap = Rm / cpm, g = c->eos.g, pst = c->eos.pst, p0 = c->cfg.base.surface_pressure;
    const double dz = c->L.dz;
    std::vector<double>& th = c->h_theta; std::vector<double>& pi = c->h_pi; std::vector<double>& p = c->h_p; std::vector<double>& rho = c->h_rho;
    double pi_surface = pow(p0 / pst, kap);
    double Pi1 = pi_surface - g * dz / (2 * cpm * th[0]);
    double p1 = pst * pow(Pi1, 1 / kap);
    pi[0] = Pi1; p[0] = p1; rho[0] = p1 / (Rm * th[0] * Pi1);
    double pm = p[0], rm = rho[0];
    for (int k = 1; k < Nz; ++k) {
        double th_face = (th[k] + th[k - 1]) / 2;
        double Pi_init = pi[k - 1] - g * dz / (cpm * th_face);
        double pk = pst * pow(Pi_init, 1 / kap);
        double A = g * pow(pst, kap) / (2 * Rm * th[k]);
        double Cc = pm / dz - g * rm / 2;
        for (int it = 0; it < 5; ++it) {                       // newton_hydrostatic_pressure, FixedIterations(5)
            double rp = pow(pk, -kap);
            double f = pk / dz + A * pk * rp - Cc;
            double fp = 1 / dz + A * (1 - kap) * rp;
            pk -= f / fp;
        }
        double Pik = pow(pk / pst, kap);
        double rk = pk / (Rm * th[k] * Pik);
        pi[k] = Pik; p[k] = pk; rho[k] = rk;
        pm = pk; rm = rk;
    }
}

static int c_upload_reference(bzc_ctx* c) {
    const int Nz = c->L.Nz;
    std::vector<double> h((size_t)2 * Nz);
    for (int k = 0; k < Nz; ++k) { h[k] = c->h_p[k]; h[Nz + k] = c->h_rho[k]; }
    CC_TRY(c, cudaMemcpyAsync(c->d_cols, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CC_TRY(c, cudaStreamSynchronize(c->stream));
    return BZ_OK;
}

static dim3 c_grid(const Layout& L, int nz) { return dim3((L.nx + 127) / 128, L.Ny, nz); }

static int c_fill_ghosts(bzc_ctx* c, double* const* fields, int nf) {
    const Layout& L = c->L;
    if (L.flat_x && L.flat_y) return BZ_OK;
    const int mode = (L.flat_x ? 0 : 1) | (L.flat_y ? 0 : 2);
    for (int f0 = 0; f0 < nf; f0 += NPROG + 1) {
        FieldSet F; F.n = 0;
        for (int f = f0; f < nf && F.n < NPROG + 1; ++f) F.f[F.n++] = fields[f];
        long long per_level = (long long)((mode & 2) ? 2 * L.HY : 0) * L.PX + (long long)((mode & 1) ? 2 * L.HX : 0) * L.Ny;
        long long total = per_level * L.Nz;
        int blocks = (int)((total + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16; if (blocks < 1) blocks = 1;
        halo_fill_periodic<<<blocks, 256, 0, c->stream>>>(L, F, mode);
        c->launches++;
    }
    CC_TRY(c, cudaGetLastError());
    return BZ_OK;
}

// update_state!(model; compute_tendencies=false) for the compressible model (update_atmosphere_model_state.jl:41-68)
static int c_update_state_host(bzc_ctx* c) {
    const Layout& L = c->L;
    int rc;
    if ((rc = c_fill_ghosts(c, c->U, 5))) return rc;
    c_update_state<<<c_grid(L, L.Nz + 1), 128, 0, c->stream>>>(L, c->eos, c->U[CF_RHO], c->U[CF_RU], c->U[CF_RV], c->U[CF_RW], c->U[CF_RTH],
                                                                c->moist ? c->rqv : nullptr, c->u, c->v, c->w, c->theta, c->T, c->p,
                                                                c->rho_tot, c->qv, c->PiL, c->CL);
    c->launches++;
    CC_TRY(c, cudaGetLastError());
    double* diag[7] = {c->u, c->v, c->w, c->theta, c->p, c->rho_tot, c->qv};
    return c_fill_ghosts(c, diag, c->moist ? 7 : 5);
}

// update_state!(compute_tendencies = true) for the moisture density: consumed by the NEXT stage's scalar_rk3_substep!
static int c_moisture_tendency_host(bzc_ctx* c) {
    if (!c->moist) return BZ_OK;
    const Layout& L = c->L;
    int rc;
    if ((rc = c_fill_ghosts(c, c->avg, 3))) return rc;
    if (c->buf == 3) c_moisture_tendency<<<c_grid(L, L.Nz), 128, 0, c->stream>>>(L, c->rho_tot, c->qv, c->avg[0], c->avg[1], c->avg[2], c->Grqv);
    else if (c->buf == 4) c_moisture_tendency<4><<<c_grid(L, L.Nz), 128, 0, c->stream>>>(L, c->rho_tot, c->qv, c->avg[0], c->avg[1], c->avg[2], c->Grqv);
    else c_moisture_tendency<5><<<c_grid(L, L.Nz), 128, 0, c->stream>>>(L, c->rho_tot, c->qv, c->avg[0], c->avg[1], c->avg[2], c->Grqv);
    c->launches++;
    CC_TRY(c, cudaGetLastError());
    return BZ_OK;
}

// prepare_acoustic_cache! + compute_slow_momentum_tendencies! + compute_slow_scalar_tendencies! +
// assemble_slow_vertical_momentum_tendency! (acoustic_runge_kutta_3.jl:181-188, acoustic_substepping.jl:1431)
static int c_stage_tendencies(bzc_ctx* c) {
    const Layout& L = c->L;
    // prepare_acoustic_cache!: Πᴸ, θᴸ (= θ), Cᴸ were written by the update_state! that produced this stage-entry state
    CProfScope ps(c, 0);
    CSlowArgs A;
    A.rho = c->U[CF_RHO]; A.ru = c->U[CF_RU]; A.rv = c->U[CF_RV]; A.rw = c->U[CF_RW];
    A.u = c->u; A.v = c->v; A.w = c->w; A.theta = c->theta; A.p = c->p;
    A.rho_tot = c->moist ? c->rho_tot : c->U[CF_RHO];
    A.p_r = c->has_ref ? c->d_cols : nullptr; A.rho_r = c->has_ref ? c->d_cols + L.Nz : nullptr;
    A.Grho = c->G[CF_RHO]; A.Gru = c->G[CF_RU]; A.Grv = c->G[CF_RV]; A.Grw = c->G[CF_RW]; A.Grth = c->G[CF_RTH]; A.Gs_rw = c->Gs_rw;
    {
        const int gx = (L.nx + CS_TX - 1) / CS_TX, gy = (L.Ny + CS_TY - 1) / CS_TY;
        int chunks = (2 * 148 * 8 + gx * gy - 1) / (gx * gy);                    // enough CTAs for ~2 waves of 8 resident CTAs per SM
        const int max_chunks = (L.Nz + 7) / 8;                                    // at least 8 levels per chunk: the carried-flux replay stays < 7 %
        if (chunks > max_chunks) chunks = max_chunks;
        if (chunks < 1) chunks = 1;
        const int k_chunk = (L.Nz + chunks - 1) / chunks;
        const dim3 grid(gx, gy, (L.Nz + k_chunk - 1) / k_chunk), block(32, CS_TY);
        if (c->buf == 3) c_slow_tendencies<<<grid, block, 0, c->stream>>>(L, A, c->eos.g, k_chunk);
        else if (c->buf == 4) c_slow_tendencies<4><<<grid, block, 0, c->stream>>>(L, A, c->eos.g, k_chunk);
        else c_slow_tendencies<5><<<grid, block, 0, c->stream>>>(L, A, c->eos.g, k_chunk);
    }
    c->launches++;
    CC_TRY(c, cudaGetLastError());
    return BZ_OK;
}

// compute_acoustic_substeps / stage_substep_count_and_size (acoustic_substepping.jl:451-508)
static int c_acoustic_substeps(const bzc_ctx* c, double dt) {
    const double gamd = c->eos.cpd / (c->eos.cpd - c->eos.Rd);
    const double cs = sqrt(gamd * c->eos.Rd * 300.0);
    const double dxm = c->L.flat_x ? INFINITY : c->L.dx, dym = c->L.flat_y ? INFINITY : c->L.dy;
    const double N = ceil(fabs(dt) * cs / (c->cfg.acoustic_cfl * fmin(dxm, dym)));
    return N < 1 ? 1 : (int)N;
}
static void c_stage_substeps(const bzc_ctx* c, double beta, double dt, int* n_tau, double* d_tau) {
    const int S = c->cfg.substeps;
    auto imax = [](int a, int b) { return a > b ? a : b; };
    if (c->cfg.substep_distribution == BZC_PROPORTIONAL_SUBSTEPS) {
        double dt_stage = beta * dt;
        int N = S > 0 ? imax(1, (int)ceil(beta * S)) : c_acoustic_substeps(c, dt_stage);
        *n_tau = N; *d_tau = dt_stage / N;
        return;
    }
    if (c->cfg.substep_distribution == BZC_MONOLITHIC_FIRST_STAGE && beta < (1.0 / 3 + 1.0 / 2) / 2) { *n_tau = 1; *d_tau = dt / 3; return; }
    int Nraw = S > 0 ? S : c_acoustic_substeps(c, dt);
    int N = imax(6, 6 * ((Nraw + 5) / 6));
    *n_tau = imax(1, (int)nearbyint(beta * N)); *d_tau = dt / N;
}

// acoustic_rk3_substep_loop! (acoustic_substepping.jl:1404-1590): 2 launches per substep (compressible.cuh)
static int c_substep_loop(bzc_ctx* c, double dt, double beta) {
    const Layout& L = c->L;
    int n_tau; double dtau;
    c_stage_substeps(c, beta, dt, &n_tau, &dtau);
    const double om = c->cfg.forward_weight, dtm = om * dtau, dts = (1 - om) * dtau;
    CConst5 Pc; CFields5 Um;
    for (int f = 0; f < 5; ++f) { Pc.f[f] = c->P[f]; Um.f[f] = c->U[f]; }
    // initialize_stage_perturbations!: no separate pass — the first substep's kernels form U⁰ - U_stage on the fly (compressible.cuh);
    // the top wall face of (ρw)′ (level Nz, never written by the kernels) is reset because the perturbation fields double as staging
    // buffers of bzc_set_state / bzc_get_state
    CC_TRY(c, cudaMemsetAsync(c->P[CF_RW] + (size_t)L.plane * L.Nz, 0, (size_t)L.plane * sizeof(double), c->stream));
    double dm = 0, ds = 0;                                 // implicit_damping_factors (:1003-1011)
    const bool thermal = c->cfg.damping == BZC_THERMAL_DIVERGENCE_DAMPING;
    if (thermal && c->cfg.damp_vertical) { double base = c->cfg.damping_coefficient * (L.dz * L.dz); dm = om * base; ds = (1 - om) * base; }
    double kx = 0, ky = 0;                                 // κˣ, κʸ (:1090-1116)
    if (thermal) {
        if (c->cfg.damping_length_scale > 0) kx = ky = c->cfg.damping_coefficient * (c->cfg.damping_length_scale * c->cfg.damping_length_scale) / dtau;
        else { double l = fmin(L.flat_x ? INFINITY : L.dx, L.flat_y ? INFINITY : L.dy); kx = ky = c->cfg.damping_coefficient * (l * l) / dtau; }
        if (L.flat_x) kx = 0;
        if (L.flat_y) ky = 0;
    }
    CHorizArgs H;
    H.ru_p = c->P[CF_RU]; H.rv_p = c->P[CF_RV]; H.rth_p = c->P[CF_RTH]; H.rth_old = c->rth_old; H.thL = c->thL; H.CL = c->CL; H.p = c->p;
    H.Gru = c->G[CF_RU]; H.Grv = c->G[CF_RV]; H.kx = kx; H.ky = ky; H.dtau = dtau;
    CColumnArgs K;
    K.rho_p = c->P[CF_RHO]; K.rth_p = c->P[CF_RTH]; K.rw_p = c->P[CF_RW]; K.ru_p = c->P[CF_RU]; K.rv_p = c->P[CF_RV];
    K.rho_s = c->rho_s; K.rth_s = c->rth_s; K.rth_old = c->rth_old; K.tfac = c->tfac;
    K.avg_u = c->avg[0]; K.avg_v = c->avg[1]; K.avg_w = c->avg[2];
    K.Grho = c->G[CF_RHO]; K.Grth = c->G[CF_RTH]; K.Gs_rw = c->Gs_rw; K.thL = c->thL; K.CL = c->CL;
    K.sponge = c->cfg.sponge != BZC_SPONGE_NONE ? c->d_cols + 2 * L.Nz : nullptr;
    K.dtau = dtau; K.dtm = dtm; K.dts = dts; K.dm = dm; K.ds = ds; K.g = c->eos.g;
    K.fth = c->cfg.thermodynamic_tendency_factor; K.fw = c->cfg.vertical_momentum_tendency_factor;
    const int cb = L.nx >= 128 ? 128 : (L.nx >= 64 ? 64 : 32);
    for (int substep = 1; substep <= n_tau + 1; ++substep) {
        // E of substep - 1 (if any) fused with A of this substep (if any)
        H.do_damp = (substep > 1) && thermal;
        H.do_step = substep <= n_tau;
        const int apply = c->cfg.apply_first_substep_pressure_gradient | (substep != 1) | (n_tau == 1);
        H.factor = apply ? 1.0 : 0.0;
        const bool first = (substep == 1);
        H.ru0 = first ? c->U0[CF_RU] : nullptr; H.ru = c->U[CF_RU]; H.rv0 = c->U0[CF_RV]; H.rv = c->U[CF_RV];
        H.rth0 = c->U0[CF_RTH]; H.rth = c->U[CF_RTH];
        K.rho0 = first ? c->U0[CF_RHO] : nullptr; K.rho = c->U[CF_RHO]; K.rth0 = c->U0[CF_RTH]; K.rth = c->U[CF_RTH];
        K.rw0 = c->U0[CF_RW]; K.rw = c->U[CF_RW];
        if (H.do_damp || H.do_step) {
            CProfScope ps(c, 2);
            if (first) c_acoustic_horizontal<true><<<c_grid(L, L.Nz), 128, 0, c->stream>>>(L, H);
            else c_acoustic_horizontal<false><<<c_grid(L, L.Nz), 128, 0, c->stream>>>(L, H);
            c->launches++;
        }
        if (substep <= n_tau) {
            CProfScope ps(c, 3);
            if (first) c_acoustic_column<true><<<dim3((L.nx + cb - 1) / cb, L.Ny), cb, 0, c->stream>>>(L, K);
            else c_acoustic_column<false><<<dim3((L.nx + cb - 1) / cb, L.Ny), cb, 0, c->stream>>>(L, K);
            c->launches++;
        }
    }
    CC_TRY(c, cudaGetLastError());
    // stage end: ⟨u⟩, full-state recovery, compute_velocities! (:1560-1587)
    CProfScope ps(c, 4);
    c_finalize_average<<<c_grid(L, L.Nz), 128, 0, c->stream>>>(L, c->U[CF_RHO], c->U[CF_RU], c->U[CF_RV], c->U[CF_RW], c->avg[0], c->avg[1],
                                                                c->avg[2], 1.0 / (double)n_tau);
    c_recover<<<c_grid(L, L.Nz), 128, 0, c->stream>>>(L, Um, Pc, c->moist ? c->rqv : nullptr, c->rqv0, c->Grqv, beta * dt);
    c->launches += 2;
    CC_TRY(c, cudaGetLastError());
    return BZ_OK;
}

static int c_store_initial_state(bzc_ctx* c) {
    for (int f = 0; f < 5; ++f)
        CC_TRY(c, cudaMemcpyAsync(c->U0[f], c->U[f], c->fsize * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    if (c->moist) CC_TRY(c, cudaMemcpyAsync(c->rqv0, c->rqv, c->fsize * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return BZ_OK;
}

static int c_time_step(bzc_ctx* c, double dt) {
    const double betas[3] = {1.0 / 3.0, 1.0 / 2.0, 1.0};
    int rc;
    if ((rc = c_store_initial_state(c))) return rc;
    // freeze_linearization_state!: the linearization is refreshed again at stage entry; seed ⟨u⟩ with the step-start velocities
    const double* vel[3] = {c->u, c->v, c->w};
    for (int a = 0; a < 3; ++a) CC_TRY(c, cudaMemcpyAsync(c->avg[a], vel[a], c->fsize * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    for (int s = 0; s < 3; ++s) {
        if ((rc = c_stage_tendencies(c))) return rc;
        if ((rc = c_substep_loop(c, dt, betas[s]))) return rc;
        CProfScope ps(c, 4);
        if ((rc = c_update_state_host(c))) return rc;
        if ((rc = c_moisture_tendency_host(c))) return rc;
    }
    c->time += dt; c->iteration += 1;
    return BZ_OK;
}

extern "C" {

void bzc_default_config(bzc_config* c) {
    memset(c, 0, sizeof(*c));
    bz_default_config(&c->base);
    c->reference_state = BZC_REFERENCE_EXNER;
    c->substeps = 0;
    c->damping = BZC_THERMAL_DIVERGENCE_DAMPING;
    c->substep_distribution = BZC_PROPORTIONAL_SUBSTEPS;
    c->acoustic_cfl = 0.5; c->forward_weight = 0.65; c->damping_coefficient = 0.1; c->damping_length_scale = 0.0;
    c->thermodynamic_tendency_factor = 1.0; c->vertical_momentum_tendency_factor = 1.0;
    c->sponge = BZC_SPONGE_NONE; c->sponge_damping_rate = 0.2; c->sponge_depth = 5e3;
}

const char* bzc_last_error(const bzc_ctx* c) { return c ? c->err : gc_err; }

void bzc_destroy(bzc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.base.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->graphs.clear();
    cudaFree(c->arena); cudaFree(c->dense); cudaFree(c->d_cols);
    for (auto e : c->prof_ev) cudaEventDestroy(e);
    for (auto e : c->prof_pool) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int bzc_create(const bzc_config* cfg, bzc_ctx** out) {
    if (!cfg || !out) return BZ_ERR_INVALID;
    *out = nullptr;
    const bz_config* b = &cfg->base;
#define FAIL(code, ...) do { bzc_set_error(nullptr, __VA_ARGS__); return (code); } while (0)
    if (b->abi_version != BZ_ABI_VERSION) FAIL(BZ_ERR_INVALID, "abi_version %d != %d", b->abi_version, BZ_ABI_VERSION);
    if (b->microphysics != BZ_MICROPHYSICS_NONE) FAIL(BZ_ERR_UNSUPPORTED, "the compressible path carries vapour only (microphysics = nothing)");
    if (b->n_ranks > 1) FAIL(BZ_ERR_UNSUPPORTED, "the compressible path runs on one GPU");
    // WENO(order = 7 / 9): c_slow_tendencies<4 / 5>, c_moisture_tendency<4 / 5>; verified against the oracle on a B200 (tests/test_gpu_weno_high_order.py)
    if (b->advection_order != 5 && b->advection_order != 7 && b->advection_order != 9)
        FAIL(BZ_ERR_UNSUPPORTED, "WENO(order = 5, 7 or 9) is on the path, got order %d", b->advection_order);
    const int buf = (b->advection_order + 1) / 2;                      // 3, 4 or 5
    const int fx = b->topology_x == BZ_FLAT, fy = b->topology_y == BZ_FLAT;
    if ((fx && b->Nx != 1) || (fy && b->Ny != 1)) FAIL(BZ_ERR_INVALID, "a Flat dimension must have size 1");
    if ((!fx && b->Nx < 4) || (!fy && b->Ny < 4) || b->Nz < 4) FAIL(BZ_ERR_INVALID, "at least 4 cells per non-Flat dimension");
    if (buf > 3 && ((!fx && b->Nx < buf + 1) || (!fy && b->Ny < buf + 1))) FAIL(BZ_ERR_INVALID, "at least %d cells per periodic dimension for this order", buf + 1);
    if (!(cfg->acoustic_cfl > 0)) FAIL(BZ_ERR_INVALID, "`acoustic_cfl` must be positive");
    if (cfg->sponge < BZC_SPONGE_NONE || cfg->sponge > BZC_SPONGE_SIN2_RAMP) FAIL(BZ_ERR_INVALID, "unknown sponge ramp %d", cfg->sponge);
    if (cfg->sponge != BZC_SPONGE_NONE && !(cfg->sponge_depth > 0)) FAIL(BZ_ERR_INVALID, "`sponge_depth` must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) FAIL(BZ_ERR_CUDA, "no CUDA device: libbreeze_b200 has no CPU fallback");
    if (b->device < 0 || b->device >= ndev) FAIL(BZ_ERR_INVALID, "device ordinal %d out of range (%d devices)", b->device, ndev);
    if (cudaSetDevice(b->device) != cudaSuccess) FAIL(BZ_ERR_CUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, b->device);
    if (prop.major < 10) FAIL(BZ_ERR_UNSUPPORTED, "compute capability %d.%d: this library is built for sm_100a (B200) only", prop.major, prop.minor);
#undef FAIL
    bzc_ctx* c = new bzc_ctx();
    c->cfg = *cfg;
    Layout& L = c->L;
    L.nx = fx ? 1 : b->Nx; L.Ny = fy ? 1 : b->Ny; L.Nz = b->Nz;
    L.flat_x = fx; L.flat_y = fy;
    c->buf = buf;
    L.HX = fx ? 0 : buf + 1; L.HY = fy ? 0 : buf + 1;                  // BZ_HALO (= 4) for the path of record
    L.PX = L.nx + 2 * L.HX; L.PY = L.Ny + 2 * L.HY;
    L.plane = (long long)L.PX * L.PY; L.n = L.plane * L.Nz;
    L.dx = fx ? 1.0 : (b->x1 - b->x0) / b->Nx;
    L.dy = fy ? 1.0 : (b->y1 - b->y0) / b->Ny;
    L.dz = (b->z1 - b->z0) / b->Nz;
    L.rdx = fx ? 0.0 : 1.0 / L.dx; L.rdy = fy ? 0.0 : 1.0 / L.dy; L.rdz = 1.0 / L.dz;
    c->eos.Rd = b->molar_gas_constant / b->dry_air_molar_mass; c->eos.cpd = b->dry_air_heat_capacity;
    c->eos.pst = b->standard_pressure; c->eos.g = b->gravitational_acceleration;
    c->eos.Rv = b->molar_gas_constant / b->vapor_molar_mass; c->eos.cpv = b->vapor_heat_capacity;
    c->has_ref = cfg->reference_state == BZC_REFERENCE_EXNER;
    c->fsize = (size_t)L.plane * (L.Nz + 1);
#define TRYCUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { bzc_set_error(nullptr, "%s: %s", #x, cudaGetErrorString(e_)); bzc_destroy(c); return BZ_ERR_CUDA; } } while (0)
    TRYCUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    const int nfields = 5 * 4 + 6 + 2 + 4 + 3 + 1 + 5;     // U, U0, G, P | u v w θ T p | Πᴸ Cᴸ | ρ′★ (ρθ)′★ (ρθ)′ˢ⁻ tfac | ⟨u v w⟩ | Gˢρw | moisture
    const size_t abytes = (size_t)nfields * c->fsize * sizeof(double);
    TRYCUDA(cudaMalloc((void**)&c->arena, abytes));
    TRYCUDA(cudaMemsetAsync(c->arena, 0, abytes, c->stream));
    c->bytes += (int64_t)abytes;
    {
        double* q = c->arena;
        auto take = [&]() { double* r = q; q += c->fsize; return r; };
        for (int f = 0; f < 5; ++f) c->U[f] = take();
        for (int f = 0; f < 5; ++f) c->U0[f] = take();
        for (int f = 0; f < 5; ++f) c->G[f] = take();
        for (int f = 0; f < 5; ++f) c->P[f] = take();
        c->u = take(); c->v = take(); c->w = take(); c->theta = take(); c->T = take(); c->p = take();
        c->PiL = take(); c->CL = take();
        c->thL = c->theta;                              // θᴸ = ρθᴸ/ρᴸ is the diagnosed θ of the stage-entry state
        c->rho_s = take(); c->rth_s = take(); c->rth_old = take(); c->tfac = take();
        for (int a = 0; a < 3; ++a) c->avg[a] = take();
        c->Gs_rw = take();
        c->rqv = take(); c->rqv0 = take(); c->Grqv = take(); c->qv = take(); c->rho_tot = take();
    }
    TRYCUDA(cudaMalloc((void**)&c->dense, (size_t)L.nx * L.Ny * (L.Nz + 1) * sizeof(double)));
    TRYCUDA(cudaMalloc((void**)&c->d_cols, (size_t)(3 * L.Nz + 1) * sizeof(double)));
    if (cfg->sponge != BZC_SPONGE_NONE) {               // sponge_term_diag / sponge_rhs profile (acoustic_substepping.jl:591-603)
        std::vector<double> sp(L.Nz + 1);
        const double Lz = b->z1 - b->z0;
        for (int k = 0; k <= L.Nz; ++k) {
            double sf = ((b->z0 + k * L.dz) - (Lz - cfg->sponge_depth)) / cfg->sponge_depth;
            sf = sf < 0 ? 0 : (sf > 1 ? 1 : sf);
            double ramp = sf;
            if (cfg->sponge == BZC_SPONGE_CUBIC_RAMP) ramp = sf * sf * (3 - 2 * sf);
            else if (cfg->sponge == BZC_SPONGE_SIN2_RAMP) { double sn = sin(M_PI / 2 * sf); ramp = sn * sn; }
            sp[k] = cfg->sponge_damping_rate * ramp;
        }
        TRYCUDA(cudaMemcpyAsync(c->d_cols + 2 * L.Nz, sp.data(), sp.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        TRYCUDA(cudaStreamSynchronize(c->stream));
    }
    c->bytes += (int64_t)((size_t)L.nx * L.Ny * (L.Nz + 1) + 2 * L.Nz) * 8;
    c->h_p.assign(L.Nz, b->surface_pressure); c->h_rho.assign(L.Nz, 0.0); c->h_pi.assign(L.Nz, 0.0);
    c->h_theta.assign(L.Nz, b->potential_temperature);
    if (c->has_ref) {
        c_build_exner_reference(c);
        int rc = c_upload_reference(c);
        if (rc) { strncpy(gc_err, c->err, 511); bzc_destroy(c); return rc; }
        // seed_pressure! from the reference (compressible_dynamics.jl:290-294)
        c_fill<<<c_grid(L, L.Nz), 128, 0, c->stream>>>(L, c->p, c->d_cols, 0.0, L.Nz);
    } else {
        c_fill<<<c_grid(L, L.Nz), 128, 0, c->stream>>>(L, c->p, nullptr, b->surface_pressure, L.Nz);
    }
    c->launches++;
    TRYCUDA(cudaGetLastError());
    TRYCUDA(cudaStreamSynchronize(c->stream));
#undef TRYCUDA
    *out = c;
    return BZ_OK;
}

int bzc_set_reference_potential_temperature(bzc_ctx* c, const double* theta_r) {
    if (c) c->graphs.clear();
    if (!c || !theta_r) return BZ_ERR_INVALID;
    if (!c->has_ref) { bzc_set_error(c, "reference_state = nothing"); return BZ_ERR_STATE; }
    cudaSetDevice(c->cfg.base.device);
    for (int k = 0; k < c->L.Nz; ++k) c->h_theta[k] = theta_r[k];
    c_build_exner_reference(c);
    return c_upload_reference(c);
}

int bzc_get_reference_state(bzc_ctx* c, double* p, double* rho, double* pi) {
    if (!c) return BZ_ERR_INVALID;
    if (!c->has_ref) { bzc_set_error(c, "reference_state = nothing"); return BZ_ERR_STATE; }
    for (int k = 0; k < c->L.Nz; ++k) {
        if (p) p[k] = c->h_p[k];
        if (rho) rho[k] = c->h_rho[k];
        if (pi) pi[k] = c->h_pi[k];
    }
    return BZ_OK;
}

int bzc_set_state(bzc_ctx* c, const double* rho, const double* ru, const double* rv, const double* rw, const double* rth, const double* rqv) {
    if (c) c->graphs.clear();                          // the moist path adds kernels to the step
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.base.device);
    const Layout& L = c->L;
    const double* src[5] = {rho, ru, rv, rw, rth};
    if (rqv) {                                          // staged through ρ′★ (scratch between steps)
        CC_TRY(c, cudaMemcpyAsync(c->rho_s, rqv, (size_t)L.nx * L.Ny * L.Nz * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        c_scatter<<<c_grid(L, L.Nz), 128, 0, c->stream>>>(L, c->rho_s, c->rqv, 0);
        c->launches++;
        c->moist = 1;
    }
    for (int f = 0; f < 5; ++f) {                       // staged through the perturbation fields (scratch between steps): no sync per field
        if (!src[f]) continue;
        const int nz = (f == CF_RW) ? L.Nz + 1 : L.Nz;
        CC_TRY(c, cudaMemcpyAsync(c->P[f], src[f], (size_t)L.nx * L.Ny * nz * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        c_scatter<<<c_grid(L, nz), 128, 0, c->stream>>>(L, c->P[f], c->U[f], f == CF_RW);
        c->launches++;
    }
    CC_TRY(c, cudaGetLastError());
    CC_TRY(c, cudaStreamSynchronize(c->stream));        // the caller's buffers are free again
    int rc = c_update_state_host(c);
    if (rc) return rc;
    if ((rc = c_store_initial_state(c))) return rc;
    if (c->moist) {   // maybe_prepare_first_time_step!: seed ⟨𝐮⟩ with the velocities, then the first moisture tendency
        const double* vel[3] = {c->u, c->v, c->w};
        for (int a = 0; a < 3; ++a) CC_TRY(c, cudaMemcpyAsync(c->avg[a], vel[a], c->fsize * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        rc = c_moisture_tendency_host(c);
    }
    return rc;
}

// the same CUDA-graph replay as bz_time_step (common.cuh: StepGraphCache); the WS-RK3 step is 43 launches with a fixed substep schedule per dt
int bzc_time_step(bzc_ctx* c, double dt) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.base.device);
    if (!(c->graphs.on((long long)c->L.nx * c->L.Ny * c->L.Nz) && !c->prof_on)) return c_time_step(c, dt);
    StepGraphEntry* g = c->graphs.find(0, dt);
    if (g && g->exec) {
        CC_TRY(c, cudaGraphLaunch(g->exec, c->stream));
        c->launches += g->launches; c->time += dt; c->iteration += 1;
        return BZ_OK;
    }
    if (g && g->seen) {
        const long long l0 = c->launches;
        const double t0 = c->time; const int64_t it0 = c->iteration;   // BZ_KEEP_F64
        cudaGraph_t graph = nullptr;
        CC_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
        int rc = c_time_step(c, dt);
        cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess || !graph) { bzc_set_error(c, "graph capture of the time step failed: %s", cudaGetErrorString(e)); return BZ_ERR_CUDA; }
        e = cudaGraphInstantiate(&g->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { g->exec = nullptr; bzc_set_error(c, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); return BZ_ERR_CUDA; }
        g->launches = c->launches - l0;
        c->time = t0 + dt; c->iteration = it0 + 1;          // c_time_step advanced the clock while recording
        CC_TRY(c, cudaGraphLaunch(g->exec, c->stream));
        return BZ_OK;
    }
    if (!g) g = c->graphs.add(0, dt);
    g->seen = 1;
    return c_time_step(c, dt);
}

int bzc_time_steps(bzc_ctx* c, double dt, int n) {
    for (int s = 0; s < n; ++s) { int rc = bzc_time_step(c, dt); if (rc) return rc; }
    return BZ_OK;
}

int bzc_compute_slow_tendencies(bzc_ctx* c) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.base.device);
    return c_stage_tendencies(c);
}

int bzc_stage_substep_count_and_size(bzc_ctx* c, double dt, double beta, int32_t* n_tau, double* d_tau) {
    if (!c) return BZ_ERR_INVALID;
    int n; double d;
    c_stage_substeps(c, beta, dt, &n, &d);
    if (n_tau) *n_tau = n;
    if (d_tau) *d_tau = d;
    return BZ_OK;
}

int bzc_acoustic_substep_loop(bzc_ctx* c, double dt, double beta) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.base.device);
    int rc = c_substep_loop(c, dt, beta);
    if (rc) return rc;
    // fill_halo_regions! + compute_velocities! at the end of the loop (:1584-1587), then update_state! as in time_step!
    if ((rc = c_update_state_host(c))) return rc;
    return c_moisture_tendency_host(c);
}

int bzc_get_field(bzc_ctx* c, int f, double* out) {
    if (!c || !out) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.base.device);
    const Layout& L = c->L;
    const double* src = nullptr; int zf = 0; double constant = 0.0; bool is_const = false;
    switch (f) {
        case BZC_RHO: src = c->U[CF_RHO]; break;
        case BZC_RHO_U: src = c->U[CF_RU]; break;
        case BZC_RHO_V: src = c->U[CF_RV]; break;
        case BZC_RHO_W: src = c->U[CF_RW]; zf = 1; break;
        case BZC_RHO_THETA: src = c->U[CF_RTH]; break;
        case BZC_U: src = c->u; break;
        case BZC_V: src = c->v; break;
        case BZC_W: src = c->w; zf = 1; break;
        case BZC_THETA: src = c->theta; break;
        case BZC_T: src = c->T; break;
        case BZC_P: src = c->p; break;
        case BZC_G_RHO: src = c->G[CF_RHO]; break;
        case BZC_G_RHO_U: src = c->G[CF_RU]; break;
        case BZC_G_RHO_V: src = c->G[CF_RV]; break;
        case BZC_G_RHO_W: src = c->G[CF_RW]; break;
        case BZC_G_RHO_THETA: src = c->G[CF_RTH]; break;
        case BZC_SLOW_RHO_W: src = c->Gs_rw; zf = 1; break;
        case BZC_EXNER_L: src = c->PiL; break;
        case BZC_THETA_L: src = c->thL; break;
        case BZC_GAMMA_R_L: is_const = true; constant = c->eos.cpd * c->eos.Rd / (c->eos.cpd - c->eos.Rd); break;   // (moist: Cᴸ/Πᴸ, below)
        case BZC_RHO_PERT: src = c->P[CF_RHO]; break;
        case BZC_RHO_THETA_PERT: src = c->P[CF_RTH]; break;
        case BZC_RHO_U_PERT: src = c->P[CF_RU]; break;
        case BZC_RHO_V_PERT: src = c->P[CF_RV]; break;
        case BZC_RHO_W_PERT: src = c->P[CF_RW]; zf = 1; break;
        case BZC_AVG_U: src = c->avg[0]; break;
        case BZC_AVG_V: src = c->avg[1]; break;
        case BZC_AVG_W: src = c->avg[2]; zf = 1; break;
        case BZC_RHO_QV: src = c->rqv; break;
        case BZC_QV: src = c->qv; break;
        case BZC_TOTAL_RHO: src = c->moist ? c->rho_tot : c->U[CF_RHO]; break;
        case BZC_G_RHO_QV: src = c->Grqv; break;
        default: bzc_set_error(c, "unknown field %d", f); return BZ_ERR_INVALID;
    }
    const int nz = zf ? L.Nz + 1 : L.Nz;
    const size_t count = (size_t)L.nx * L.Ny * nz;
    if (is_const && !c->moist) { for (size_t e = 0; e < count; ++e) out[e] = constant; return BZ_OK; }
    if (is_const) {                                     // moist: γᵐRᵐᴸ = Cᴸ / Πᴸ (Cᴸ is what the kernels store)
        std::vector<double> pi(count);
        int rc = bzc_get_field(c, BZC_EXNER_L, pi.data());
        if (rc) return rc;
        c_extract<<<c_grid(L, nz), 128, 0, c->stream>>>(L, c->CL, c->dense);
        c->launches++;
        CC_TRY(c, cudaMemcpyAsync(out, c->dense, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CC_TRY(c, cudaStreamSynchronize(c->stream));
        for (size_t e = 0; e < count; ++e) out[e] /= pi[e];
        return BZ_OK;
    }
    c_extract<<<c_grid(L, nz), 128, 0, c->stream>>>(L, src, c->dense);
    c->launches++;
    CC_TRY(c, cudaGetLastError());
    CC_TRY(c, cudaMemcpyAsync(out, c->dense, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CC_TRY(c, cudaStreamSynchronize(c->stream));
    return BZ_OK;
}

int bzc_get_state(bzc_ctx* c, double* rho, double* ru, double* rv, double* rw, double* rth, double* rqv) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.base.device);
    const Layout& L = c->L;
    double* dst[5] = {rho, ru, rv, rw, rth};
    // the perturbation fields are scratch between steps: stage the five dense copies there and download them back to back
    for (int f = 0; f < 5; ++f) {
        if (!dst[f]) continue;
        const int nz = (f == CF_RW) ? L.Nz + 1 : L.Nz;
        c_extract<<<c_grid(L, nz), 128, 0, c->stream>>>(L, c->U[f], c->P[f]);
        c->launches++;
        CC_TRY(c, cudaMemcpyAsync(dst[f], c->P[f], (size_t)L.nx * L.Ny * nz * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    if (rqv) {
        c_extract<<<c_grid(L, L.Nz), 128, 0, c->stream>>>(L, c->rqv, c->rho_s);
        c->launches++;
        CC_TRY(c, cudaMemcpyAsync(rqv, c->rho_s, (size_t)L.nx * L.Ny * L.Nz * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    CC_TRY(c, cudaGetLastError());
    CC_TRY(c, cudaStreamSynchronize(c->stream));
    return BZ_OK;
}

int bzc_get_clock(bzc_ctx* c, double* time, int64_t* iteration) {   // BZ_KEEP_F64
    if (!c) return BZ_ERR_INVALID;
    if (time) *time = c->time;
    if (iteration) *iteration = c->iteration;
    return BZ_OK;
}

int bzc_synchronize(bzc_ctx* c) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.base.device);
    CC_TRY(c, cudaStreamSynchronize(c->stream));
    return BZ_OK;
}

int bzc_profile_enable(bzc_ctx* c, int on) { if (!c) return BZ_ERR_INVALID; c->prof_on = on; return BZ_OK; }

int bzc_profile_read(bzc_ctx* c, double* ms, int64_t* n) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.base.device);
    CC_TRY(c, cudaStreamSynchronize(c->stream));
    for (size_t e = 0; e < c->prof_fam.size(); ++e) {
        float t = 0.f;
        cudaEventElapsedTime(&t, c->prof_ev[2 * e], c->prof_ev[2 * e + 1]);
        c->prof_ms[c->prof_fam[e]] += t; c->prof_n[c->prof_fam[e]] += 1;
        c->prof_pool.push_back(c->prof_ev[2 * e]); c->prof_pool.push_back(c->prof_ev[2 * e + 1]);
    }
    c->prof_ev.clear(); c->prof_fam.clear();
    for (int f = 0; f < C_NFAM; ++f) { if (ms) ms[f] = c->prof_ms[f]; if (n) n[f] = c->prof_n[f]; c->prof_ms[f] = 0; c->prof_n[f] = 0; }
    return BZ_OK;
}

int64_t bzc_kernel_launch_count(const bzc_ctx* c) { return c ? c->launches : 0; }
void* bzc_stream(bzc_ctx* c) { return c ? (void*)c->stream : nullptr; }
int64_t bzc_device_bytes(const bzc_ctx* c) { return c ? c->bytes : 0; }

}  // extern "C"
