// compressible.cuh — kernels of the compressible split-explicit path (include/breeze_b200_compressible.h):
// Wicker–Skamarock RK3 with linearized acoustic substepping, dry air.
//
// Reference kernels replaced (paths relative to the reference repository):
//   per stage   _compute_linearization_exner_and_theta!, _compute_linearization_mixture_eos!    acoustic_substepping.jl:372-415
//               (folded into the update_state kernel that produces the stage-entry state)
//               compute_x/y/z_momentum_tendency! (SlowTendencyMode), _compute_density_tendency!,
//               compute_potential_temperature_tendency!                                         acoustic_substep_helpers.jl:55-149
//               _assemble_slow_vertical_momentum_tendency!                                      acoustic_substepping.jl:724-748
//               _zero_stage_workspaces!, _initialize_stage_perturbations!, …_with_rewind!        :805-833
//               (folded into the first substep's two kernels: U′ = U⁰ - U_stage on the fly)
//   per substep _explicit_horizontal_step! (A), _build_predictors! + _build_vertical_rhs! (B),
//               solve!(BatchedTridiagonalSolver) (C), _post_solve_recovery! (D),
//               _thermal_divergence_damping! (E) and 7 halo fills                                :859-1144,1442-1552
//   stage end   _finalize_time_averaged_velocity!, _recover_full_state!, _compute_velocities!,
//               _compute_auxiliary_thermodynamic_variables!, _compute_temperature_and_pressure!   :1230-1293, compressible_time_stepping.jl:191-235
//
// B200 design. The substep loop is HBM-bound (1- to 7-point stencils + a column recurrence), so it is cut to TWO launches per
// substep with no halo passes at all (periodic neighbours are addressed by wrapped indices):
//   acoustic_horizontal   E of the previous substep fused with A of this one (both only touch (ρu)′, (ρv)′ pointwise)
//   acoustic_column       one thread per (i, j) column, x fastest across the warp (coalesced 256-byte rows): the upward march
//                         builds the predictors ρ′★, (ρθ)′★, the face right-hand side and runs the forward elimination of
//                         the tridiagonal system with coefficients formed on the fly from Cᴸ = γRᵐᴸ Πᴸ and θᴸ; the downward
//                         march back-substitutes (ρw)′ and recovers ρ′, (ρθ)′ and the ⟨ρ𝐮′⟩ accumulators. Face values of the
//                         level below / above are carried in registers, so every field is read once per march.
// Fields use the layout of common.cuh (4 ghost cells in x / y, none in z) with Nz + 1 levels allocated; only the WENO5 slow
// tendency kernel reads ghost cells (filled once per stage). z-face level 0 is the bottom wall, level Nz the top wall.
#pragma once
#include "common.cuh"
#include "weno.cuh"

struct CEos { double Rd, cpd, pst, g, Rv, cpv; };

// ---- periodic neighbours by wrapped index (interior addressing, no ghost cells needed) -----------------------------
__device__ __forceinline__ long long cxm(const Layout& L, long long n, int i) { return i > 0 ? n - 1 : n + (L.nx - 1); }
__device__ __forceinline__ long long cxp(const Layout& L, long long n, int i) { return i < L.nx - 1 ? n + 1 : n - (L.nx - 1); }
__device__ __forceinline__ long long cym(const Layout& L, long long n, int j) { return j > 0 ? n - L.PX : n + (long long)(L.Ny - 1) * L.PX; }
__device__ __forceinline__ long long cyp(const Layout& L, long long n, int j) { return j < L.Ny - 1 ? n + L.PX : n - (long long)(L.Ny - 1) * L.PX; }

// ---- update_state!: velocities, θ, and the joint (T, p) diagnosis ---------------------------------------------------
// temperature(::LiquidIceDensityState) with NewtonSolver(reltol=0, abstol=1e-4, maxiter=8) (dynamic_states.jl:201-232); p = ρ Rᵈ T
__device__ __forceinline__ void c_temperature_pressure(const CEos& e, double rho, double theta, double qv, double& T, double& p) {
    // mixture constants of MoistureMassFractions(qᵛ, 0, 0); qᵛ = 0 reproduces Rᵈ, cᵖᵈ bit for bit
    const double qd = 1.0 - qv;
    const double Rm = qd * e.Rd + qv * e.Rv, cpm = qd * e.cpd + qv * e.cpv;
    const double kap = Rm / cpm, gam = cpm / (cpm - Rm);
    T = pow(theta, gam) * pow(rho * Rm / e.pst, gam - 1.0) + 0.0;
    double dT = T; int iter = 0;
    while (fabs(dT) > 1e-4 && iter < 8) {
        double Phi = pow(rho * Rm * T / e.pst, kap) * theta;
        dT = -(T - Phi - 0.0) / (1.0 - kap * Phi / T);
        T += dT;
        ++iter;
    }
    p = rho * Rm * T;
}

// grid (x blocks, Ny, Nz + 1); ρᵈ needs valid x / y ghost cells. rqv == nullptr: dry air (ρ = ρᵈ, qᵛ = 0; rho_tot / qv are not written).
__global__ void c_update_state(Layout L, CEos e, const double* __restrict__ rho, const double* __restrict__ ru, const double* __restrict__ rv,
                               const double* __restrict__ rw, const double* __restrict__ rth, const double* __restrict__ rqv,
                               double* __restrict__ u, double* __restrict__ v, double* __restrict__ w, double* __restrict__ theta,
                               double* __restrict__ T, double* __restrict__ p, double* __restrict__ rho_tot, double* __restrict__ qv,
                               double* __restrict__ PiL, double* __restrict__ CL) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= L.nx) return;
    long long n = lidx(L, i, j, k);
    if (k < L.Nz) {
        double r = rho[n];
        double rx = L.flat_x ? r : (r + rho[n - 1]) / 2;
        double ry = L.flat_y ? r : (r + rho[n - L.PX]) / 2;
        u[n] = ru[n] / rx;
        v[n] = rv[n] / ry;
        double th = rth[n] / r;                       // θ = ρθ / ρᵈ (coupling density)
        theta[n] = th;
        double rt = r, q = 0.0;
        if (rqv) {                                     // compute_total_density!: ρ = ρᵈ + ρqᵛ; qᵛ is a mass fraction of the total
            const double m = rqv[n];
            rt = r + m; q = m / rt;
            rho_tot[n] = rt; qv[n] = q;
        }
        double Tn, pn;
        c_temperature_pressure(e, rt, th, q, Tn, pn);
        T[n] = Tn; p[n] = pn;
        // refresh_linearization_basic_state! of the NEXT stage reads exactly this state (nothing touches it in between), so the
        // stage-entry cache is written here: Πᴸ = (p/pˢᵗ)^(Rᵈ/cᵖᵈ), Cᴸ = γᵐRᵐᴸ Πᴸ; θᴸ = ρθ/ρᵈ is the θ field itself
        {
            const double Pi = pow(pn / e.pst, e.Rd / e.cpd);
            const double qd = 1.0 - q;
            const double Rm = qd * e.Rd + q * e.Rv, cpm = qd * e.cpd + q * e.cpv;
            PiL[n] = Pi;
            CL[n] = (cpm * Rm / (cpm - Rm)) * Pi;
        }
    }
    w[n] = (k == 0 || k == L.Nz) ? 0.0 : rw[n] / ((rho[n] + rho[n - L.plane]) / 2);
}

// ---- slow tendencies (WENO5, 3-D coupling density) ------------------------------------------------------------------
struct CSlowArgs {
    const double *rho, *ru, *rv, *rw, *u, *v, *w, *theta, *p;
    const double *rho_tot;               // total density ρᵈ + ρqᵛ for the buoyancy (= rho for dry air)
    const double *p_r, *rho_r;           // Nz each or nullptr (reference_state = nothing)
    double *Grho, *Gru, *Grv, *Grw, *Grth, *Gs_rw;
};

// biased reconstruction at the "face" between f[n - s] and f[n]; in x / y the ghost cells make R = 3 always valid,
// in z the buffer R shrinks next to the walls (weno.cuh red_face / red_center) and only in-range levels are read
// BUF: the scheme's buffer (3: WENO5, the path of record; 4 / 5: WENO7 / WENO9); R <= BUF: buffer in use at this point
template <int BUF = 3>
__device__ __forceinline__ double c_biased(const double* __restrict__ f, long long n, long long s, int R, bool left) {
    if constexpr (BUF >= 5) { if (R >= 5) return weno_hi_mem<5>(f, n, s, left); }
    if constexpr (BUF >= 4) { if (R == 4) return weno_hi_mem<4>(f, n, s, left); }
    if (R >= 3) return biased6c<3>(f[n - 3 * s], f[n - 2 * s], f[n - s], f[n], f[n + s], f[n + 2 * s], left);
    if (R == 2) return left ? weno3z(f[n - 2 * s], f[n - s], f[n]) : weno3z(f[n + s], f[n], f[n - s]);
    return left ? f[n - s] : f[n];
}
template <int BUF = 3>
__device__ __forceinline__ double c_sym(const double* __restrict__ a, long long n, long long s, int R) {
    if constexpr (BUF >= 5) { if (R >= 4) return centered_hi_mem<4>(a, n, s); }
    if constexpr (BUF >= 4) { if (R == 3) return centered_hi_mem<3>(a, n, s); }
    if (R >= 2) return ((7.0 / 12.0) * (a[n - s] + a[n])) - ((1.0 / 12.0) * (a[n - 2 * s] + a[n + s]));
    return 0.5 * (a[n - s] + a[n]);
}

// Every face flux is evaluated ONCE: a block owns 31 x TY columns and marches up a chunk of levels. Each thread evaluates the
// fluxes through the LOW x face, the LOW y face and the TOP z face of its cell; the high x face comes from lane + 1 by warp
// shuffle (tiles overlap by one column: lane 31 only supplies fluxes), the high y face from the row above through shared memory
// (the last row evaluates its own), the bottom z face is carried in registers from the level below.
#ifndef CS_TY
#define CS_TY 8
#endif
#define CS_TX 31
#ifndef CS_MINB
#define CS_MINB 4             // 64 registers per thread: 4 CTAs of 256 threads per SM (A/B: 2 → 697 us, 3 → 570 us, 4 → 494 us at 256x256x64)
#endif
template <int BUF = 3>
__global__ void __launch_bounds__(32 * CS_TY, BUF == 3 ? CS_MINB : 1) c_slow_tendencies(Layout L, CSlowArgs A, double g, int k_chunk) {
    constexpr int BS = BUF - 1;                           // buffer of the symmetric (advecting) interpolation: Centered(order - 1)
    __shared__ double sfy[2][4][CS_TY + 1][32];          // double-buffered by level parity: one CTA barrier per level
    const int lane = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * CS_TX + lane, j = blockIdx.y * CS_TY + ty;
    const int Nz = L.Nz;
    const int kb = blockIdx.z * k_chunk, ke = min(Nz, kb + k_chunk);
    const long long SX = 1, SY = L.PX, SZ = L.plane;
    const bool fx_ = L.flat_x, fy_ = L.flat_y;
    const bool col_ok = (i <= L.nx) && (j < L.Ny) && (fx_ ? i < L.nx : true);       // i = nx: only the low-x-face fluxes are needed
    const bool own = (lane < CS_TX) && (i < L.nx) && (j < L.Ny);
    const bool yrow_ok = (lane < CS_TX) && (i < L.nx) && (j <= L.Ny);             // j = Ny: the ghost row supplies the high face of row Ny - 1
    const int j_top = blockIdx.y * CS_TY + CS_TY;                                 // the row of y faces just above a full tile: one flux kind per warp
    const bool top_ok = (ty < 4) && (j_top <= L.Ny) && (lane < CS_TX) && (i < L.nx);
    const double Ax = L.dy * L.dz, Ay = L.dx * L.dz, Az = L.dx * L.dy, Vinv = 1.0 / (L.dx * L.dy * L.dz);
    // advecting mass fluxes: centred-4 interpolation of the area-weighted momentum; advected velocity: WENO5-Z
    auto symx = [&](const double* a, long long m) { return fx_ ? a[m] : c_sym<BUF>(a, m, SX, BS); };
    auto symy = [&](const double* a, long long m) { return fy_ ? a[m] : c_sym<BUF>(a, m, SY, BS); };
    auto Fuu = [&](long long m1) { double t = Ax * c_sym<BUF>(A.ru, m1, SX, BS); return t * c_biased<BUF>(A.u, m1, SX, BUF, t > 0); };          // centre i (m1 = i + 1)
    auto Fvu = [&](long long m) { double t = Ay * symx(A.rv, m); return t * c_biased<BUF>(A.u, m, SY, BUF, t > 0); };                      // (face i, face j)
    auto Fwu = [&](long long m, int kk) { if (kk == 0 || kk == Nz) return 0.0; double t = Az * symx(A.rw, m); return t * c_biased<BUF>(A.u, m, SZ, red_face(kk, Nz, BUF), t > 0); };
    auto Fuv = [&](long long m) { double t = Ax * symy(A.ru, m); return t * c_biased<BUF>(A.v, m, SX, BUF, t > 0); };
    auto Fvv = [&](long long m1) { double t = Ay * c_sym<BUF>(A.rv, m1, SY, BS); return t * c_biased<BUF>(A.v, m1, SY, BUF, t > 0); };
    auto Fwv = [&](long long m, int kk) { if (kk == 0 || kk == Nz) return 0.0; double t = Az * symy(A.rw, m); return t * c_biased<BUF>(A.v, m, SZ, red_face(kk, Nz, BUF), t > 0); };
    auto Fuw = [&](long long m, int kk) { if (kk == 0) return 0.0; double t = Ax * c_sym<BUF>(A.ru, m, SZ, red_face(kk, Nz, BS)); return t * c_biased<BUF>(A.w, m, SX, BUF, t > 0); };
    auto Fvw = [&](long long m, int kk) { if (kk == 0) return 0.0; double t = Ay * c_sym<BUF>(A.rv, m, SZ, red_face(kk, Nz, BS)); return t * c_biased<BUF>(A.w, m, SY, BUF, t > 0); };
    auto Fww = [&](long long m1, int kc) { double t = Az * c_sym<BUF>(A.rw, m1, SZ, red_center(kc, Nz, BS)); return t * c_biased<BUF>(A.w, m1, SZ, red_center(kc, Nz, BUF), t > 0); };
    auto Tx = [&](long long m) { double t = A.u[m]; return ((A.rho[m] + A.rho[m - SX]) / 2) * (Ax * t * c_biased<BUF>(A.theta, m, SX, BUF, t > 0)); };
    auto Ty = [&](long long m) { double t = A.v[m]; return ((A.rho[m] + A.rho[m - SY]) / 2) * (Ay * t * c_biased<BUF>(A.theta, m, SY, BUF, t > 0)); };
    auto Tz = [&](long long m, int kk) { if (kk == 0 || kk == Nz) return 0.0; double t = A.w[m];
                                         return ((A.rho[m] + A.rho[m - SZ]) / 2) * (Az * t * c_biased<BUF>(A.theta, m, SZ, red_face(kk, Nz, BUF), t > 0)); };
    const long long n0 = lidx(L, min(i, L.nx), min(j, L.Ny), 0);
    const long long n0_top = lidx(L, min(i, L.nx), min(j_top, L.Ny), 0);
    // z-type fluxes through the bottom face of the chunk's first level (Fww: at centre kb - 1)
    double zb_u = 0.0, zb_v = 0.0, zb_w = 0.0, zb_t = 0.0;
    if (own && kb > 0) {
        const long long n = n0 + (long long)kb * SZ;
        zb_u = Fwu(n, kb); zb_v = Fwv(n, kb); zb_w = Fww(n, kb - 1); zb_t = Tz(n, kb);
    }
    for (int k = kb; k < ke; ++k) {
        const long long n = n0 + (long long)k * SZ;
        // low-x-face fluxes of this column (every lane of a valid row, including the overlap lane)
        double xu = 0.0, xv = 0.0, xw = 0.0, xt = 0.0;
        if (col_ok && !fx_) { xu = Fuu(n); xv = Fuv(n); xw = Fuw(n, k); xt = Tx(n); }
        const double xu_e = __shfl_down_sync(0xffffffffu, xu, 1), xv_e = __shfl_down_sync(0xffffffffu, xv, 1);
        const double xw_e = __shfl_down_sync(0xffffffffu, xw, 1), xt_e = __shfl_down_sync(0xffffffffu, xt, 1);
        // low-y-face fluxes → shared memory (this level's buffer); the extra row above a full tile is spread one kind per warp
        double yu = 0.0, yv = 0.0, yw = 0.0, yt = 0.0;
        if (yrow_ok && !fy_) { yu = Fvu(n); yv = Fvv(n); yw = Fvw(n, k); yt = Ty(n); }
        auto& S = sfy[k & 1];
        S[0][ty][lane] = yu; S[1][ty][lane] = yv; S[2][ty][lane] = yw; S[3][ty][lane] = yt;
        if (top_ok && !fy_) {
            const long long nt = n0_top + (long long)k * SZ;
            S[ty][CS_TY][lane] = (ty == 0) ? Fvu(nt) : (ty == 1) ? Fvv(nt) : (ty == 2) ? Fvw(nt, k) : Ty(nt);
        }
        __syncthreads();                                   // the only barrier of the level (the other buffer is written next level)
        if (!own) continue;
        double yu_n = 0.0, yv_n = 0.0, yw_n = 0.0, yt_n = 0.0;
        if (!fy_) { yu_n = S[0][ty + 1][lane]; yv_n = S[1][ty + 1][lane]; yw_n = S[2][ty + 1][lane]; yt_n = S[3][ty + 1][lane]; }
        // top-z-face fluxes (Fww: at centre k)
        const double zt_u = Fwu(n + SZ, k + 1), zt_v = Fwv(n + SZ, k + 1), zt_w = Fww(n + SZ, k), zt_t = Tz(n + SZ, k + 1);
        // Fuu / Fvv live at centres: this thread's "low" value is centre i-1 (j-1), the neighbour's is centre i (j)
        A.Gru[n] = -(Vinv * ((fx_ ? 0.0 : xu_e - xu) + (fy_ ? 0.0 : yu_n - yu) + (zt_u - zb_u)));
        A.Grv[n] = -(Vinv * ((fx_ ? 0.0 : xv_e - xv) + (fy_ ? 0.0 : yv_n - yv) + (zt_v - zb_v)));
        const double Gw = (k >= 1) ? -(Vinv * ((fx_ ? 0.0 : xw_e - xw) + (fy_ ? 0.0 : yw_n - yw) + (zt_w - zb_w))) : 0.0;
        A.Grw[n] = Gw;
        {
            double dxu = fx_ ? 0.0 : Ax * A.ru[n + SX] - Ax * A.ru[n];
            double dyv = fy_ ? 0.0 : Ay * A.rv[n + SY] - Ay * A.rv[n];
            double dzw = Az * A.rw[n + SZ] - Az * A.rw[n];
            A.Grho[n] = -(Vinv * (dxu + dyv + dzw));
        }
        A.Grth[n] = -(Vinv * ((fx_ ? 0.0 : xt_e - xt) + (fy_ ? 0.0 : yt_n - yt) + (zt_t - zb_t)));
        // _assemble_slow_vertical_momentum_tendency!: Gˢρw = (Gρw - ∂z(pᴸ - pᵣ) - g ℑz(ρᴸ - ρᵣ)) (k > 1)
        double Gs = 0.0;
        if (k >= 1) {
            if (A.p_r) {
                double dpk = A.p[n] - A.p_r[k], dpm = A.p[n - SZ] - A.p_r[k - 1];
                double drk = A.rho_tot[n] - A.rho_r[k], drm = A.rho_tot[n - SZ] - A.rho_r[k - 1];
                Gs = Gw - (dpk - dpm) * L.rdz - g * ((drk + drm) / 2);
            } else {
                Gs = Gw - (A.p[n] - A.p[n - SZ]) * L.rdz - g * ((A.rho_tot[n] + A.rho_tot[n - SZ]) / 2);
            }
        }
        A.Gs_rw[n] = Gs;
        zb_u = zt_u; zb_v = zt_v; zb_w = zt_w; zb_t = zt_t;
    }
}

struct CFields5 { double* f[5]; };      // ρ, ρu, ρv, ρw, ρθ
struct CConst5 { const double* f[5]; };

// ---- substep, horizontal part: E (damping of the previous substep) fused with A (explicit step of this substep) ------
struct CHorizArgs {
    double *ru_p, *rv_p;
    const double *rth_p, *rth_old, *thL, *CL, *p, *Gru, *Grv;
    double kx, ky;            // κˣ, κʸ of the damping (0: none)
    double dtau, factor;      // A: Δτ and the perturbation-PGF gate (1 / 0)
    int do_damp, do_step;
    // first substep of a stage (initialize_stage_perturbations!, acoustic_substepping.jl:765-842): the perturbations are the rewind
    // U⁰ - U_stage, formed on the fly instead of by a separate pass; null otherwise
    const double *ru0, *ru, *rv0, *rv, *rth0, *rth;
};

template <bool FIRST>
__global__ void c_acoustic_horizontal(Layout L, CHorizArgs A) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= L.nx) return;
    const long long n = lidx(L, i, j, k);
    const long long nxm = cxm(L, n, i), nym = cym(L, n, j);
    constexpr bool first = FIRST;
    auto rth_at = [&](long long m) { return first ? A.rth0[m] - A.rth[m] : A.rth_p[m]; };
    double ru = first ? A.ru0[n] - A.ru[n] : A.ru_p[n], rv = first ? A.rv0[n] - A.rv[n] : A.rv_p[n];
    const double rt = rth_at(n);
    if (A.do_damp) {          // _thermal_divergence_damping!
        double d0 = rt - A.rth_old[n];
        double th = A.thL[n];
        if (!L.flat_x) {
            double dxd = (d0 - (A.rth_p[nxm] - A.rth_old[nxm])) * L.rdx;
            ru -= A.kx * dxd / ((th + A.thL[nxm]) / 2);
        }
        if (!L.flat_y) {
            double dyd = (d0 - (A.rth_p[nym] - A.rth_old[nym])) * L.rdy;
            rv -= A.ky * dyd / ((th + A.thL[nym]) / 2);
        }
    }
    if (A.do_step) {          // _explicit_horizontal_step!
        double pn = A.p[n], cp = A.CL[n] * rt;
        double dxp = 0.0, dyp = 0.0;
        if (!L.flat_x) dxp = (pn - A.p[nxm]) * L.rdx + A.factor * ((cp - A.CL[nxm] * rth_at(nxm)) * L.rdx);
        if (!L.flat_y) dyp = (pn - A.p[nym]) * L.rdy + A.factor * ((cp - A.CL[nym] * rth_at(nym)) * L.rdy);
        ru += A.dtau * (A.Gru[n] - dxp);
        rv += A.dtau * (A.Grv[n] - dyp);
    }
    A.ru_p[n] = ru; A.rv_p[n] = rv;
}

// ---- substep, vertical part: B + C + D in one column kernel ------------------------------------------------------------
struct CColumnArgs {
    double *rho_p, *rth_p, *rw_p;              // perturbation prognostics (in / out)
    const double *ru_p, *rv_p;                 // after step A
    double *rho_s, *rth_s, *rth_old, *tfac;    // predictors, stashed (ρθ)′, Thomas factors
    double *avg_u, *avg_v, *avg_w;
    const double *Grho, *Grth, *Gs_rw, *thL, *CL;
    const double* sponge;                      // UpperSponge: rate · ramp at the Nz + 1 z-faces, or nullptr (sponge = nothing)
    double dtau, dtm, dts, dm, ds, g, fth, fw;
    // first substep of a stage: ρ′, (ρθ)′, (ρw)′ are the rewind U⁰ - U_stage formed on the fly and the ⟨ρ𝐮′⟩ accumulators start from
    // zero (_zero_stage_workspaces!, _initialize_stage_perturbations!, _initialize_perturbation_with_rewind!); null otherwise
    const double *rho0, *rho, *rth0, *rth, *rw0, *rw;
};

// Everything one level of the upward march reads from HBM. The column recurrence is latency-bound unless many loads are in
// flight per thread, so the march is software-pipelined: the loads of level k + D - 1 are issued (into a register ring of D
// LevelIn) before level k is computed and stored — the DRAM round trips of D - 1 levels overlap with the recurrence.
#ifndef BZ_COL_DEPTH
#define BZ_COL_DEPTH 6            // ring of prefetched levels per thread: loads run BZ_COL_DEPTH - 1 levels ahead of the recurrence
#endif
struct CLevelIn { double rp, tp, ru0, rue, rv0, rvn, th_e, th_w, th_n, th_s, th_up, Grho, Grth, C, w_up, Gs; };
struct CLevelDown { double w, t_up, th_dn, rs, ts, ru, rv, au, av, aw; };

// one thread per column; blockDim.x columns along x per block, blockIdx.y = j
template <bool FIRST>
__global__ void __launch_bounds__(128) c_acoustic_column(Layout L, CColumnArgs A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= L.nx) return;
    const int Nz = L.Nz;
    const long long SZ = L.plane;
    const long long n0 = lidx(L, i, j, 0);
    const long long oxm = cxm(L, n0, i) - n0, oxp = cxp(L, n0, i) - n0, oym = cym(L, n0, j) - n0, oyp = cyp(L, n0, j) - n0;
    const double Ax = L.dy * L.dz, Ay = L.dx * L.dz, Vinv = 1.0 / (L.dx * L.dy * L.dz);
    const double rdz = L.rdz, rdzc = L.rdz, rdzf = L.rdz;
    const double dtm2 = A.dtm * A.dtm;
    const double tiny = 10.0 * 2.220446049250313e-16;
    const bool fx_ = L.flat_x, fy_ = L.flat_y;
    const double* __restrict__ ru_p = A.ru_p; const double* __restrict__ rv_p = A.rv_p;
    const double* __restrict__ thL = A.thL; const double* __restrict__ CL = A.CL;
    const double* __restrict__ Grho = A.Grho; const double* __restrict__ Grth = A.Grth; const double* __restrict__ Gs_rw = A.Gs_rw;
    constexpr bool first = FIRST;

    auto load_up = [&](int k, CLevelIn& q) {
        const long long n = n0 + (long long)k * SZ;
        const bool top = (k + 1 == Nz);
        q.rp = first ? A.rho0[n] - A.rho[n] : A.rho_p[n]; q.tp = first ? A.rth0[n] - A.rth[n] : A.rth_p[n];
        q.ru0 = __ldg(ru_p + n); q.rv0 = __ldg(rv_p + n);
        q.rue = fx_ ? 0.0 : __ldg(ru_p + n + oxp); q.rvn = fy_ ? 0.0 : __ldg(rv_p + n + oyp);
        q.th_e = fx_ ? 0.0 : __ldg(thL + n + oxp); q.th_w = fx_ ? 0.0 : __ldg(thL + n + oxm);
        q.th_n = fy_ ? 0.0 : __ldg(thL + n + oyp); q.th_s = fy_ ? 0.0 : __ldg(thL + n + oym);
        q.th_up = top ? 0.0 : __ldg(thL + n + SZ);
        q.Grho = __ldg(Grho + n); q.Grth = __ldg(Grth + n); q.C = __ldg(CL + n);
        q.w_up = top ? 0.0 : (first ? A.rw0[n + SZ] - A.rw[n + SZ] : A.rw_p[n + SZ]);   // old (ρw)′ at face k+1 (row k+1 comes later)
        q.Gs = __ldg(Gs_rw + n);
    };

    // ---- upward march: predictors (B), face right-hand side, forward elimination (C) --------------------------------
    // carried from the level below (face k lies between cells k-1 and k)
    double C_m = 0.0, rs_m = 0.0, ts_m = 0.0, rp_m = 0.0, tp_m = 0.0;   // Cᴸ, ρ′★, (ρθ)′★, old ρ′, old (ρθ)′ at cell k-1
    double w_m = 0.0, w_0 = 0.0;                                         // old (ρw)′ at faces k-1, k (face 0 is the wall)
    double th_0 = __ldg(thL + n0);
    double thf_m = th_0, thf_0 = th_0;                                   // ℑbz θᴸ at faces k-1, k (one-sided on the walls)
    double cu_prev = 0.0, beta = 1.0, phi_prev = 0.0;
    auto process_up = [&](int k, const CLevelIn& cur) {
        const long long n = n0 + (long long)k * SZ;
        const bool top = (k + 1 == Nz);
        const double w_p = cur.w_up;
        const double th_p = top ? th_0 : cur.th_up;
        const double thf_p = top ? th_0 : (th_p + th_0) / 2;             // face k+1
        const double C_0 = cur.C;
        const double rp_0 = cur.rp, tp_0 = cur.tp;
        // Step B — _build_predictors!
        double dxM = 0.0, dxT = 0.0, dyM = 0.0, dyT = 0.0;
        if (!fx_) {
            dxM = Ax * cur.rue - Ax * cur.ru0;
            dxT = Ax * ((cur.th_e + th_0) / 2) * cur.rue - Ax * ((th_0 + cur.th_w) / 2) * cur.ru0;
        }
        if (!fy_) {
            dyM = Ay * cur.rvn - Ay * cur.rv0;
            dyT = Ay * ((cur.th_n + th_0) / 2) * cur.rvn - Ay * ((th_0 + cur.th_s) / 2) * cur.rv0;
        }
        const double divM = Vinv * (dxM + dyM), divT = Vinv * (dxT + dyT);
        const double dzw = (w_p - w_0) * rdz;
        const double dzT = (thf_p * w_p - thf_0 * w_0) * rdz;
        const double rs_0 = rp_0 + A.dtau * (cur.Grho - divM) - A.dts * dzw;
        const double ts_0 = tp_0 + A.dtau * (A.fth * cur.Grth - divT) - A.dts * dzT;
        // _build_vertical_rhs! at face k, then row k of the forward elimination. Row 0 is the wall: b = 1, c = 0, rhs = 0.
        double phi, cu, t = 0.0;
        if (k == 0) {
            beta = 1.0;
            phi = 0.0;
            cu = 0.0;
        } else {
            double dp_s = (C_0 * ts_0 - C_m * ts_m) * rdz;
            double dp_o = (C_0 * tp_0 - C_m * tp_m) * rdz;
            double Gp = A.dts * dp_o + A.dtm * dp_s;
            double Gb = A.g * (A.dts * ((rp_0 + rp_m) / 2) + A.dtm * ((rs_0 + rs_m) / 2));
            double d2 = ((w_p - w_0) * rdz - (w_0 - w_m) * rdz) * rdz;
            double Gd = -A.ds * d2;
            const double sp = A.sponge ? A.sponge[k] : 0.0;        // level-uniform: one broadcast load
            double rhs = w_0 + A.dtau * A.fw * cur.Gs - Gp - Gb - Gd - fabs(A.dts) * sp * w_0;
            // get_coefficient(::AcousticTridiagLower / Diagonal / Upper) for row k
            double al = -dtm2 * C_m * thf_m * rdzc * rdzf + dtm2 * A.g * rdzc / 2 - A.dm * rdzc * rdzf;
            double b = 1.0 + (dtm2 * thf_0 * (C_0 * rdzc + C_m * rdzc) * rdzf + dtm2 * A.g * (rdzc - rdzc) / 2 + A.dm * (rdzc + rdzc) * rdzf
                              + fabs(A.dtm) * sp);
            cu = -dtm2 * C_0 * thf_p * rdzc * rdzf - dtm2 * A.g * rdzc / 2 - A.dm * rdzc * rdzf;
            t = cu_prev / beta;
            beta = b - al * t;
            phi = (fabs(beta) > tiny) ? (rhs - al * phi_prev) / beta : w_0;
        }
        A.rth_old[n] = tp_0;
        A.rho_s[n] = rs_0; A.rth_s[n] = ts_0;
        A.tfac[n] = t;
        A.rw_p[n] = phi;
        // roll
        cu_prev = cu; phi_prev = phi;
        C_m = C_0; rs_m = rs_0; ts_m = ts_0; rp_m = rp_0; tp_m = tp_0;
        w_m = w_0; w_0 = w_p;
        thf_m = thf_0; thf_0 = thf_p; th_0 = th_p;
    };
    {
        CLevelIn ring[BZ_COL_DEPTH];
#pragma unroll
        for (int d = 0; d < BZ_COL_DEPTH - 1; ++d) if (d < Nz) load_up(d, ring[d]);
        for (int k0 = 0; k0 < Nz; k0 += BZ_COL_DEPTH) {
#pragma unroll
            for (int d = 0; d < BZ_COL_DEPTH; ++d) {
                const int k = k0 + d;
                if (k < Nz) {
                    if (k + BZ_COL_DEPTH - 1 < Nz) load_up(k + BZ_COL_DEPTH - 1, ring[(d + BZ_COL_DEPTH - 1) % BZ_COL_DEPTH]);   // in flight while levels k … are computed
                    process_up(k, ring[d]);
                }
            }
        }
    }

    // ---- downward march: back substitution (C), recovery of ρ′, (ρθ)′ and the ⟨ρ𝐮′⟩ accumulators (D) -----------------
    auto load_down = [&](int k, CLevelDown& q) {
        const long long n = n0 + (long long)k * SZ;
        q.w = A.rw_p[n];
        q.t_up = (k < Nz - 1) ? A.tfac[n + SZ] : 0.0;
        q.th_dn = (k > 0) ? __ldg(thL + n - SZ) : 0.0;
        q.rs = A.rho_s[n]; q.ts = A.rth_s[n];
        q.ru = __ldg(ru_p + n); q.rv = __ldg(rv_p + n);
        q.au = first ? 0.0 : A.avg_u[n]; q.av = first ? 0.0 : A.avg_v[n]; q.aw = first ? 0.0 : A.avg_w[n];
    };
    double w_top = 0.0;                                                  // final (ρw)′ at face k+1 (top wall: 0)
    double th_up = 0.0;                                                  // θᴸ at cell k+1
    double th_k = __ldg(thL + n0 + (long long)(Nz - 1) * SZ);
    auto process_down = [&](int k, const CLevelDown& dc) {
        const long long n = n0 + (long long)k * SZ;
        double w_k = dc.w;
        if (k < Nz - 1) w_k -= dc.t_up * w_top;
        const double th_dn = (k > 0) ? dc.th_dn : th_k;
        const double thf_p = (k + 1 < Nz) ? (th_up + th_k) / 2 : th_k;
        const double thf_0 = (k > 0) ? (th_k + th_dn) / 2 : th_k;
        const double dzw = (w_top - w_k) * rdz;
        const double dzT = (thf_p * w_top - thf_0 * w_k) * rdz;
        A.rw_p[n] = w_k;
        A.rho_p[n] = dc.rs - A.dtm * dzw;
        A.rth_p[n] = dc.ts - A.dtm * dzT;
        A.avg_u[n] = dc.au + dc.ru;
        A.avg_v[n] = dc.av + dc.rv;
        A.avg_w[n] = dc.aw + w_k;
        w_top = w_k; th_up = th_k; th_k = th_dn;
    };
    {
        CLevelDown ring[BZ_COL_DEPTH];
#pragma unroll
        for (int d = 0; d < BZ_COL_DEPTH - 1; ++d) if (Nz - 1 - d >= 0) load_down(Nz - 1 - d, ring[d]);
        for (int k0 = Nz - 1; k0 >= 0; k0 -= BZ_COL_DEPTH) {
#pragma unroll
            for (int d = 0; d < BZ_COL_DEPTH; ++d) {
                const int k = k0 - d;
                if (k >= 0) {
                    if (k - (BZ_COL_DEPTH - 1) >= 0) load_down(k - (BZ_COL_DEPTH - 1), ring[(d + BZ_COL_DEPTH - 1) % BZ_COL_DEPTH]);
                    process_down(k, ring[d]);
                }
            }
        }
    }
}

// ---- stage end ------------------------------------------------------------------------------------------------------
// _finalize_time_averaged_velocity! (ρᴸ needs valid x / y ghost cells; still the stage-entry density here)
__global__ void c_finalize_average(Layout L, const double* __restrict__ rho, const double* __restrict__ ru, const double* __restrict__ rv,
                                   const double* __restrict__ rw, double* __restrict__ avg_u, double* __restrict__ avg_v,
                                   double* __restrict__ avg_w, double inv_n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= L.nx) return;
    long long n = lidx(L, i, j, k);
    double r = rho[n];
    double rx = L.flat_x ? r : (r + rho[n - 1]) / 2;
    double ry = L.flat_y ? r : (r + rho[n - L.PX]) / 2;
    double rz = (k > 0) ? (r + rho[n - L.plane]) / 2 : r;
    rx = (rx == 0.0) ? 1.0 : rx; ry = (ry == 0.0) ? 1.0 : ry; rz = (rz == 0.0) ? 1.0 : rz;
    avg_u[n] = (ru[n] + avg_u[n] * inv_n) / rx;
    avg_v[n] = (rv[n] + avg_v[n] * inv_n) / ry;
    avg_w[n] = (k > 0) ? (rw[n] + avg_w[n] * inv_n) / rz : 0.0;
}

// _recover_full_state!: U ← Uᴸ + U′ in place; with moisture also scalar_rk3_substep!: ρqᵛ ← ρqᵛ⁰ + βΔt Gⁿ.ρqᵛ
__global__ void c_recover(Layout L, CFields5 U, CConst5 P, double* __restrict__ rqv, const double* __restrict__ rqv0,
                          const double* __restrict__ Grqv, double dt_stage) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= L.nx) return;
    long long n = lidx(L, i, j, k);
#pragma unroll
    for (int f = 0; f < 5; ++f) U.f[f][n] = U.f[f][n] + P.f[f][n];
    if (rqv) rqv[n] = rqv0[n] + dt_stage * Grqv[n];
}

// compute_scalar_tendency! for the moisture density (update_atmosphere_model_state.jl:343, dynamics_kernel_functions.jl:132-159):
// Gⁿ.ρqᵛ = -div_ρUc(ρ_total, ⟨𝐮⟩, qᵛ) with the acoustic-mean transport velocities. ρ_total, qᵛ, ⟨𝐮⟩ need valid ghost cells.
template <int BUF = 3>
__global__ void __launch_bounds__(128) c_moisture_tendency(Layout L, const double* __restrict__ rho, const double* __restrict__ q,
                                                           const double* __restrict__ au, const double* __restrict__ av,
                                                           const double* __restrict__ aw, double* __restrict__ G) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= L.nx) return;
    const long long n = lidx(L, i, j, k), SX = 1, SY = L.PX, SZ = L.plane;
    const int Nz = L.Nz;
    const double Ax = L.dy * L.dz, Ay = L.dx * L.dz, Az = L.dx * L.dy, Vinv = 1.0 / (L.dx * L.dy * L.dz);
    auto Fx = [&](long long m) { double t = au[m]; return ((rho[m] + rho[m - SX]) / 2) * (Ax * t * c_biased<BUF>(q, m, SX, BUF, t > 0)); };
    auto Fy = [&](long long m) { double t = av[m]; return ((rho[m] + rho[m - SY]) / 2) * (Ay * t * c_biased<BUF>(q, m, SY, BUF, t > 0)); };
    auto Fz = [&](long long m, int kk) { if (kk == 0 || kk == Nz) return 0.0; double t = aw[m];
                                         return ((rho[m] + rho[m - SZ]) / 2) * (Az * t * c_biased<BUF>(q, m, SZ, red_face(kk, Nz, BUF), t > 0)); };
    double fx = L.flat_x ? 0.0 : Fx(n + SX) - Fx(n);
    double fy = L.flat_y ? 0.0 : Fy(n + SY) - Fy(n);
    double fz = Fz(n + SZ, k + 1) - Fz(n, k);
    G[n] = -(Vinv * (fx + fy + fz));
}

// seed_time_averaged_velocities! / store_initial_state!: plain copies of n doubles
__global__ void c_copy(const double* __restrict__ src, double* __restrict__ dst, long long n) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) dst[e] = src[e];
}

// dense interior (x fastest, nz_out levels) <-> padded field with Nz + 1 levels allocated
__global__ void c_extract(Layout L, const double* __restrict__ src, double* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= L.nx) return;
    dst[((size_t)k * L.Ny + j) * L.nx + i] = src[lidx(L, i, j, k)];
}
__global__ void c_scatter(Layout L, const double* __restrict__ src, double* __restrict__ dst, int zface) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= L.nx) return;
    double v = src[((size_t)k * L.Ny + j) * L.nx + i];
    if (zface && (k == 0 || k == L.Nz)) v = 0.0;           // impenetrable walls
    dst[lidx(L, i, j, k)] = v;
}
__global__ void c_fill(Layout L, double* __restrict__ dst, const double* __restrict__ column, double value, int nz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= L.nx || k >= nz) return;
    dst[lidx(L, i, j, k)] = column ? column[k] : value;
}
