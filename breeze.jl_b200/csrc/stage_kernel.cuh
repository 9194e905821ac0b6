// stage_kernel.cuh — ONE fused kernel per SSP-RK3 stage for the anelastic path.
//
// Replaces, per stage, the reference's 14 per-field launches + ~20 halo passes:
//   _compute_velocities!, _compute_auxiliary_thermodynamic_variables!          (update_atmosphere_model_state.jl:248-292)
//   compute_x/y/z_momentum_tendency!, compute_potential_temperature_tendency!,
//   compute_scalar_tendency!                                                    (:390-411, dynamics_kernel_functions.jl:64-159,
//                                                                                potential_temperature_tendency.jl:66-106, Advection.jl:20-35)
//   buoyancy_forceᶜᶜᶜ                                                           (anelastic_buoyancy.jl:36-72)
//   _ssp_rk3_substep! × 5                                                       (ssp_runge_kutta_3.jl:167-173)
// It reads the five projected prognostics (+ U⁰ in stages 2, 3) once and writes the five predictor fields once:
// 88 / 128 algorithmic bytes per cell.
//
// Decomposition: a CTA owns a column of (TX-1) × TY cells and marches up z. An 8-slot ring of z-planes
// (TX+8) × (TY+6) of the five fields lives in shared memory as VELOCITIES / SPECIFIC values (u, v, w, θ, q —
// converted once per loaded element); every x-, y- and z-stencil of the 15 WENO5 reconstructions per cell is
// read from it. Each thread computes only the fluxes through the LOW x/y faces and the TOP z face of its cell;
// the high-side x/y fluxes come from the neighbouring thread through shared memory (the CTA's last x-column only
// produces fluxes: tiles overlap by one cell in x; the extra y-row of fluxes is spread over five warps), the
// bottom z flux is carried in registers from the previous level. So every face flux is evaluated once.
// Planes are staged either by TMA (cp.async.bulk.tensor, one 3-D box per field and level, mbarrier
// completion, prefetched one level ahead) or by plain coalesced loads (selected at run time; bit-identical).
#pragma once
#include "common.cuh"
#include "weno.cuh"

#define RING 8

struct StageParams {
    CUtensorMap tmap[NPROG];          // 64-byte aligned; only used when use_tma
    Layout L;
    Columns col;
    Thermo th;
    const double* U[NPROG];
    const double* U0[NPROG];
    double* out[NPROG];
    double dt, alpha;
    int mode;                         // 0: RK update → out; 1: tendency G → out
    int nx_u;                         // ρu (and G_ρu) is produced for i < nx_u (nx, or nx + 1 on a multi-GPU slab)
    int k_chunk;                      // levels per z chunk (blockIdx.z selects the chunk)
    int use_tma;
};

// ---- thermodynamics on the fly ---------------------------------------------------------------------------------
__device__ __forceinline__ double sat_vapor_pressure_liquid(const Thermo& th, double T) {
    // clausius_clapeyron.jl:59-68 over a planar liquid surface
    double dcl = th.cpv - th.cl;
    double L0 = th.Ll - dcl * th.Tr_energy;
    return th.ptr * pow(T / th.Ttr, dcl / th.Rv) * exp((1.0 / th.Ttr - 1.0 / T) * L0 / th.Rv);
}

__device__ __forceinline__ double lipt_temperature(const Thermo& th, double theta, double logp, double qv, double ql) {
    // dynamic_states.jl:31-58: T = Π θ + ℒˡ qˡ / cᵖᵐ,  Π = (pᵣ/pˢᵗ)^(Rᵐ/cᵖᵐ)
    double qd = 1.0 - (qv + ql);
    double Rm = qd * th.Rd + qv * th.Rv;
    double cpm = qd * th.cpd + qv * th.cpv + ql * th.cl;
    return exp((Rm / cpm) * logp) * theta + th.Ll * ql / cpm;
}

// Warm-phase saturation adjustment (saturation_adjustment.jl:182-231, Solvers.jl:243-262); returns T, sets qv, ql.
__device__ double saturation_adjust(const Thermo& th, double theta, double pr, double logp, double qt, double& qv, double& ql) {
    qv = qt; ql = 0.0;
    if (theta == 0.0) return 0.0;
    double T1 = lipt_temperature(th, theta, logp, qt, 0.0);
    double Rm1 = (1.0 - qt) * th.Rd + qt * th.Rv;
    double rho1 = pr / (Rm1 * T1);
    double qvs1 = sat_vapor_pressure_liquid(th, T1) / (rho1 * th.Rv * T1);
    if (qt <= qvs1) return T1;
    const double eps = th.Rd / th.Rv;
    auto qsat_adj = [&](double T) { double pvs = sat_vapor_pressure_liquid(th, T); return eps * (1.0 - qt) * pvs / (pr - pvs); };
    auto residual = [&](double T) { double qs = qsat_adj(T); double l = fmax(0.0, qt - qs); return T - lipt_temperature(th, theta, logp, qt - l, l); };
    double l1 = fmax(0.0, qt - qsat_adj(T1));
    double v1 = qt - l1;
    double cpm = (1.0 - qt) * th.cpd + v1 * th.cpv + l1 * th.cl;
    double T2 = T1 + fmax(0.01, 0.5 * (th.Ll * l1 / cpm));
    double x1 = T1, x2 = T2, r1 = residual(x1), r2 = residual(x2);
    int iter = 0;
    while (fabs(r2) > 1e-4 && iter < 20) {
        double slope = (x2 - x1) / (r2 - r1);
        bool valid = isfinite(slope);
        if (!valid) slope = 0.0;
        x1 = x2; r1 = r2;
        x2 -= r2 * slope;
        r2 = residual(x2);
        if (!valid) r2 = 0.0;
        ++iter;
    }
    double qs = qsat_adj(x2);
    ql = fmax(0.0, qt - qs);
    qv = qt - ql;
    return lipt_temperature(th, theta, logp, qv, ql);
}

// buoyancy_forceᶜᶜᶜ: -g ρᵣ (Rᵐᵣ Tᵣ / (Rᵐ T) - 1), reference moisture = 0 (anelastic_buoyancy.jl:36-72)
template <int MICRO>
__device__ __forceinline__ double buoyancy_center(const Thermo& th, const Columns& col, int k, double theta, double q) {
    double T, Rm;
    if (MICRO == BZ_MICROPHYSICS_NONE) {
        if (q == 0.0) { T = col.exner_dry[k] * theta; Rm = th.Rd; }
        else { T = lipt_temperature(th, theta, col.log_p_pst[k], q, 0.0); Rm = (1.0 - q) * th.Rd + q * th.Rv; }
    } else {
        double qv, ql;
        T = saturation_adjust(th, theta, col.p[k], col.log_p_pst[k], q, qv, ql);
        Rm = (1.0 - (qv + ql)) * th.Rd + qv * th.Rv;
    }
    double rho_p = col.rho[k] * (th.Rd * col.T[k] / (Rm * T) - 1.0);
    return -th.g * rho_p;
}

// ---- TMA / mbarrier primitives ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// ---- the kernel ------------------------------------------------------------------------------------------------
template <int TX, int TY, bool HAS_Y>
struct StageShared {
    static constexpr int SW = TX + 8;                 // x: tile origin i0 - 4
    static constexpr int SH = HAS_Y ? TY + 6 : 1;     // y: tile origin j0 - 3
    static constexpr int YO = HAS_Y ? 3 : 0;
    static constexpr int PLANE = (SW * SH + 15) & ~15;   // every field slice stays 128-byte aligned (TMA destination)
    static_assert(!HAS_Y || (TY >= NPROG && TX == 32), "the extra y-face row is spread one flux kind per warp");
    double ring[RING][NPROG][PLANE];
    double fx[NPROG][TY][TX];
    double fy[NPROG][HAS_Y ? TY + 1 : 1][TX];
    uint64_t bar[RING];
};

template <int TX, int TY, bool HAS_Y, int MICRO>
__global__ void __launch_bounds__(2 * TX * TY, 1) stage_kernel(const __grid_constant__ StageParams P) {
    using SM = StageShared<TX, TY, HAS_Y>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    constexpr int SW = SM::SW, SH = SM::SH, YO = SM::YO, NCELL = TX * TY, NT = 2 * NCELL;

    // Two threads per cell column ("roles", warp-uniform): the FP64 pipe needs ~4 warps per scheduler to stay busy and the
    // plane ring leaves room for one CTA per SM, so the 15 flux kinds of a cell are split between two warps sets that share
    // the ring:   role 0: ρu, ρv (all directions) + the x/y fluxes of θ;   role 1: ρw, the z flux of θ, ρq, buoyancy.
    const Layout& L = P.L;
    const int tid = threadIdx.x;
    const int role = tid / NCELL, ctid = tid % NCELL;
    const int tx = ctid % TX, ty = ctid / TX;
    const int i0 = blockIdx.x * (TX - 1), j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int Nz = L.Nz;
    const int kb = blockIdx.z * P.k_chunk;
    if (kb >= Nz) return;
    const int ke = min(Nz, kb + P.k_chunk);
    // The box origin must be 16-byte aligned in global memory for TMA (even x index): tiles with an odd i0 start one
    // column further left (SW has the slack) and shift their threads by one column inside the plane.
    const int xs = i0 & 1;
    const int sx = tx + 4 + xs, sy = ty + YO;         // this thread's cell inside a plane
    const bool flat_x = L.flat_x;

    auto ld = [&](int f, int kk, int x, int y) -> double { return S.ring[kk & (RING - 1)][f][y * SW + x]; };

    // -- plane staging --------------------------------------------------------------------------------------------
    auto scale_of = [&](int f, int kk) -> double {
        if (kk < 0 || kk >= Nz) return 0.0;
        return f == 2 ? P.col.rho_f_inv[kk] : P.col.rho_inv[kk];
    };
    auto load_plane_direct = [&](int kk) {       // plain coalesced loads + conversion to velocities / specific values
        const bool inside = (kk >= 0 && kk < Nz);
        for (int f = 0; f < NPROG; ++f) {
            const double sc = scale_of(f, kk);
            double* dst = S.ring[kk & (RING - 1)][f];
            const double* src = P.U[f] + (long long)(inside ? kk : 0) * L.plane;
            for (int e = tid; e < SW * SH; e += NT) {
                int x = e % SW, y = e / SW;
                int gx = i0 - xs + x;                      // padded x index (i0 - xs - 4 + x + HX)
                int gy = HAS_Y ? (j0 + y + L.HY - 3) : 0;  // padded y index
                double v = 0.0;
                if (inside && gx < L.PX && gy < L.PY && (flat_x ? (x == 4) : true)) {
                    int px = flat_x ? 0 : gx;
                    v = src[(long long)gy * L.PX + px] * sc;
                }
                dst[e] = v;
            }
        }
    };
    auto issue_plane_tma = [&](int kk) {         // one elected thread; OOB levels / columns are zero-filled by the TMA unit
        uint64_t* bar = &S.bar[kk & (RING - 1)];
        mbar_expect_tx(bar, (uint32_t)(NPROG * SW * SH * sizeof(double)));
        for (int f = 0; f < NPROG; ++f)
            tma_load_3d(S.ring[kk & (RING - 1)][f], &P.tmap[f], bar, i0 - xs, HAS_Y ? (j0 + L.HY - 3) : 0, kk);
    };
    auto convert_plane = [&](int kk) {           // raw prognostics → velocities / specific values, in place
        for (int f = 0; f < NPROG; ++f) {
            const double sc = scale_of(f, kk);
            double* dst = S.ring[kk & (RING - 1)][f];
            for (int e = tid; e < SW * SH; e += NT) dst[e] *= sc;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes before a later TMA refill of the slot
    };

    uint32_t phase_bits = 0;                     // one parity bit per ring slot
    if (P.use_tma) {
        if (tid == 0) {
            for (int s = 0; s < RING; ++s) mbar_init(&S.bar[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    auto wait_plane_tma = [&](int kk) {
        int s = kk & (RING - 1);
        mbar_wait(&S.bar[s], (phase_bits >> s) & 1u);
        phase_bits ^= (1u << s);
    };

    const int kstart = (kb > 0) ? kb - 1 : 0;    // a chunk that starts above the ground first rebuilds the carried z fluxes
    // prologue: planes kstart-2 .. kstart+2 (the loop brings in kstart+3)
    if (P.use_tma) {
        if (tid == 0) for (int kk = kstart - 2; kk <= kstart + 3; ++kk) issue_plane_tma(kk);
        for (int kk = kstart - 2; kk <= kstart + 2; ++kk) { wait_plane_tma(kk); convert_plane(kk); }
    } else {
        for (int kk = kstart - 2; kk <= kstart + 2; ++kk) load_plane_direct(kk);
    }

    // carried from the level below: z-type fluxes through the bottom face (role 0: ρu, ρv; role 1: ρw, θ, q), buoyancy below
    double zb0 = 0.0, zb1 = 0.0, zb2 = 0.0, b_below = 0.0;

    const double rdx = L.rdx, rdy = L.rdy, rdz = L.rdz;
    const bool own_cell = (tx < TX - 1) && (j < L.Ny);
    const bool in_x = i < L.nx;
    const int f_first = role == 0 ? 0 : 2, f_count = role == 0 ? 2 : 3;   // fields this thread assembles and stores

    for (int k = kstart; k < ke; ++k) {
        // ---- stage plane k+3 ------------------------------------------------------------------------------------
        if (P.use_tma) {
            wait_plane_tma(k + 3);
            convert_plane(k + 3);
            __syncthreads();                                           // plane k+3 converted; slot of k-4 (== k+4) is free
            if (tid == 0 && k + 1 < ke) issue_plane_tma(k + 4);        // prefetch for the next level
        } else {
            load_plane_direct(k + 3);
            __syncthreads();
        }

        // own-point values for the RK update: issued now, consumed after the flux phase
        const long long n = lidx(L, i, j, k);
        const bool do_store = (k >= kb) && own_cell;
        double Uc[3] = {0.0, 0.0, 0.0}, U0c[3] = {0.0, 0.0, 0.0};
        if (do_store && P.mode == 0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                int f = f_first + a;
                if (a < f_count && (in_x || (f == 0 && i < P.nx_u))) {
                    Uc[a] = P.U[f][n];
                    if (P.alpha != 1.0) U0c[a] = P.U0[f][n];
                }
            }
        }

        const double rho_k = P.col.rho[k];
        const int kf = k + 1;                                          // top face of this cell
        const double rho_ft = P.col.rho_f[kf];                         // ℑz ρ at the top face (kf <= Nz)
        const int Rf_top = red_face(kf, Nz, 3);                        // biased z reconstruction at the top face
        const int Rf_k2 = red_face(k, Nz, 2);                          // symmetric z interpolation to the face k
        const int Rc3 = red_center(k, Nz, 3), Rc2 = red_center(k, Nz, 2);

        // ---- X-type fluxes: through x-face i (or at centre i-1 for ρu); kinds 0 ρu, 1 ρv, 2 ρw, 3 θ, 4 q ---------------
        auto x_flux = [&](int kind) -> double {
            const double u_i = ld(0, k, sx, sy);
            switch (kind) {
                case 0: {   // FUu at centre i-1
                    double ut = rho_k * sym4(ld(0, k, sx - 2, sy), ld(0, k, sx - 1, sy), u_i, ld(0, k, sx + 1, sy), 2);
                    double uh = biased6(ld(0, k, sx - 3, sy), ld(0, k, sx - 2, sy), ld(0, k, sx - 1, sy), u_i, ld(0, k, sx + 1, sy), ld(0, k, sx + 2, sy), 3, ut > 0.0);
                    return ut * uh;
                }
                case 1: {   // FUv at (face i, face j)
                    double ut;
                    if (HAS_Y) ut = rho_k * sym4(ld(0, k, sx, sy - 2), ld(0, k, sx, sy - 1), u_i, ld(0, k, sx, sy + 1), 2);
                    else ut = rho_k * u_i;
                    double vh = biased6(ld(1, k, sx - 3, sy), ld(1, k, sx - 2, sy), ld(1, k, sx - 1, sy), ld(1, k, sx, sy), ld(1, k, sx + 1, sy), ld(1, k, sx + 2, sy), 3, ut > 0.0);
                    return ut * vh;
                }
                case 2: {   // FUw at (face i, z-face k); the wall face k = 0 carries no w tendency
                    if (k < 1) return 0.0;
                    double a0 = (k >= 2) ? P.col.rho[k - 2] * ld(0, k - 2, sx, sy) : 0.0;
                    double a1 = P.col.rho[k - 1] * ld(0, k - 1, sx, sy);
                    double a2 = rho_k * u_i;
                    double a3 = (k + 1 < Nz) ? P.col.rho[k + 1] * ld(0, k + 1, sx, sy) : 0.0;
                    double ut = sym4(a0, a1, a2, a3, Rf_k2);
                    double wh = biased6(ld(2, k, sx - 3, sy), ld(2, k, sx - 2, sy), ld(2, k, sx - 1, sy), ld(2, k, sx, sy), ld(2, k, sx + 1, sy), ld(2, k, sx + 2, sy), 3, ut > 0.0);
                    return ut * wh;
                }
                case 3: {   // tracer mass flux ρ u θ̂
                    double th = biased6(ld(3, k, sx - 3, sy), ld(3, k, sx - 2, sy), ld(3, k, sx - 1, sy), ld(3, k, sx, sy), ld(3, k, sx + 1, sy), ld(3, k, sx + 2, sy), 3, u_i > 0.0);
                    return rho_k * u_i * th;
                }
                default: {
                    double qh = biased6(ld(4, k, sx - 3, sy), ld(4, k, sx - 2, sy), ld(4, k, sx - 1, sy), ld(4, k, sx, sy), ld(4, k, sx + 1, sy), ld(4, k, sx + 2, sy), 3, u_i > 0.0);
                    return rho_k * u_i * qh;
                }
            }
        };

        // ---- Y-type fluxes: through y-face j (or at centre j-1 for ρv); row `yy` of the plane --------------------
        auto y_flux = [&](int kind, int yy) -> double {
            const double v_j = ld(1, k, sx, yy);
            switch (kind) {
                case 0: {   // FVu at (face i, face j)
                    double vt = flat_x ? rho_k * v_j : rho_k * sym4(ld(1, k, sx - 2, yy), ld(1, k, sx - 1, yy), v_j, ld(1, k, sx + 1, yy), 2);
                    double uh = biased6(ld(0, k, sx, yy - 3), ld(0, k, sx, yy - 2), ld(0, k, sx, yy - 1), ld(0, k, sx, yy), ld(0, k, sx, yy + 1), ld(0, k, sx, yy + 2), 3, vt > 0.0);
                    return vt * uh;
                }
                case 1: {   // FVv at centre j-1
                    double vt = rho_k * sym4(ld(1, k, sx, yy - 2), ld(1, k, sx, yy - 1), v_j, ld(1, k, sx, yy + 1), 2);
                    double vh = biased6(ld(1, k, sx, yy - 3), ld(1, k, sx, yy - 2), ld(1, k, sx, yy - 1), v_j, ld(1, k, sx, yy + 1), ld(1, k, sx, yy + 2), 3, vt > 0.0);
                    return vt * vh;
                }
                case 2: {   // FVw at (face j, z-face k)
                    if (k < 1) return 0.0;
                    double a0 = (k >= 2) ? P.col.rho[k - 2] * ld(1, k - 2, sx, yy) : 0.0;
                    double a1 = P.col.rho[k - 1] * ld(1, k - 1, sx, yy);
                    double a2 = rho_k * v_j;
                    double a3 = (k + 1 < Nz) ? P.col.rho[k + 1] * ld(1, k + 1, sx, yy) : 0.0;
                    double vt = sym4(a0, a1, a2, a3, Rf_k2);
                    double wh = biased6(ld(2, k, sx, yy - 3), ld(2, k, sx, yy - 2), ld(2, k, sx, yy - 1), ld(2, k, sx, yy), ld(2, k, sx, yy + 1), ld(2, k, sx, yy + 2), 3, vt > 0.0);
                    return vt * wh;
                }
                case 3: {
                    double th = biased6(ld(3, k, sx, yy - 3), ld(3, k, sx, yy - 2), ld(3, k, sx, yy - 1), ld(3, k, sx, yy), ld(3, k, sx, yy + 1), ld(3, k, sx, yy + 2), 3, v_j > 0.0);
                    return rho_k * v_j * th;
                }
                default: {
                    double qh = biased6(ld(4, k, sx, yy - 3), ld(4, k, sx, yy - 2), ld(4, k, sx, yy - 1), ld(4, k, sx, yy), ld(4, k, sx, yy + 1), ld(4, k, sx, yy + 2), 3, v_j > 0.0);
                    return rho_k * v_j * qh;
                }
            }
        };

        // ---- Z-type fluxes through the top face kf (or at centre k for ρw) ---------------------------------------
        auto z_flux = [&](int kind) -> double {
            const double w_top = ld(2, kf, sx, sy);                    // 0 on the top wall (zero plane)
            switch (kind) {
                case 0: {   // FWu at (face i, z-face kf)
                    double wt = flat_x ? rho_ft * w_top : rho_ft * sym4(ld(2, kf, sx - 2, sy), ld(2, kf, sx - 1, sy), w_top, ld(2, kf, sx + 1, sy), 2);
                    double uh = biased6(ld(0, kf - 3, sx, sy), ld(0, kf - 2, sx, sy), ld(0, kf - 1, sx, sy), ld(0, kf, sx, sy), ld(0, kf + 1, sx, sy), ld(0, kf + 2, sx, sy), Rf_top, wt > 0.0);
                    return wt * uh;
                }
                case 1: {   // FWv at (face j, z-face kf)
                    double wt = HAS_Y ? rho_ft * sym4(ld(2, kf, sx, sy - 2), ld(2, kf, sx, sy - 1), w_top, ld(2, kf, sx, sy + 1), 2) : rho_ft * w_top;
                    double vh = biased6(ld(1, kf - 3, sx, sy), ld(1, kf - 2, sx, sy), ld(1, kf - 1, sx, sy), ld(1, kf, sx, sy), ld(1, kf + 1, sx, sy), ld(1, kf + 2, sx, sy), Rf_top, wt > 0.0);
                    return wt * vh;
                }
                case 2: {   // FWw at centre k: faces k-1 .. k+2 (advecting), k-2 .. k+3 (advected)
                    double a0 = (k >= 1) ? P.col.rho_f[k - 1] * ld(2, k - 1, sx, sy) : 0.0;
                    double a1 = P.col.rho_f[k] * ld(2, k, sx, sy);
                    double a2 = rho_ft * w_top;
                    double a3 = (k + 2 <= Nz) ? P.col.rho_f[k + 2] * ld(2, k + 2, sx, sy) : 0.0;
                    double wt = sym4(a0, a1, a2, a3, Rc2);
                    double wh = biased6(ld(2, k - 2, sx, sy), ld(2, k - 1, sx, sy), ld(2, k, sx, sy), w_top, ld(2, k + 2, sx, sy), ld(2, k + 3, sx, sy), Rc3, wt > 0.0);
                    return wt * wh;
                }
                case 3: {   // tracer mass flux ℑz(ρ) w θ̂
                    double th = biased6(ld(3, kf - 3, sx, sy), ld(3, kf - 2, sx, sy), ld(3, kf - 1, sx, sy), ld(3, kf, sx, sy), ld(3, kf + 1, sx, sy), ld(3, kf + 2, sx, sy), Rf_top, w_top > 0.0);
                    return rho_ft * w_top * th;
                }
                default: {
                    double qh = biased6(ld(4, kf - 3, sx, sy), ld(4, kf - 2, sx, sy), ld(4, kf - 1, sx, sy), ld(4, kf, sx, sy), ld(4, kf + 1, sx, sy), ld(4, kf + 2, sx, sy), Rf_top, w_top > 0.0);
                    return rho_ft * w_top * qh;
                }
            }
        };

        double zt0, zt1, zt2 = 0.0, b_here = 0.0;
        if (role == 0) {
            if (!flat_x) { S.fx[0][ty][tx] = x_flux(0); S.fx[1][ty][tx] = x_flux(1); S.fx[3][ty][tx] = x_flux(3); }
            if (HAS_Y) {
                S.fy[0][ty][tx] = y_flux(0, sy); S.fy[1][ty][tx] = y_flux(1, sy); S.fy[3][ty][tx] = y_flux(3, sy);
                // the extra row of y-faces above the tile: one flux kind per warp (rows 0..2 of this role)
                if (ty == 0) S.fy[0][TY][tx] = y_flux(0, TY + YO);
                else if (ty == 1) S.fy[1][TY][tx] = y_flux(1, TY + YO);
                else if (ty == 2) S.fy[3][TY][tx] = y_flux(3, TY + YO);
            }
            zt0 = z_flux(0); zt1 = z_flux(1);
        } else {
            if (!flat_x) { S.fx[2][ty][tx] = x_flux(2); S.fx[4][ty][tx] = x_flux(4); }
            if (HAS_Y) {
                S.fy[2][ty][tx] = y_flux(2, sy); S.fy[4][ty][tx] = y_flux(4, sy);
                if (ty == 0) S.fy[2][TY][tx] = y_flux(2, TY + YO);
                else if (ty == 1) S.fy[4][TY][tx] = y_flux(4, TY + YO);
            }
            zt0 = z_flux(2); zt1 = z_flux(3); zt2 = z_flux(4);
            b_here = buoyancy_center<MICRO>(P.th, P.col, k, ld(3, k, sx, sy), ld(4, k, sx, sy));
        }

        __syncthreads();                                               // fx / fy complete

        // ---- tendencies, RK update, store -----------------------------------------------------------------------
        if (do_store) {
            double zt[3] = {zt0, zt1, zt2}, zb[3] = {zb0, zb1, zb2};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int f = f_first + a;
                if (a >= f_count) break;
                if (!(in_x || (f == 0 && i < P.nx_u))) continue;
                double g = 0.0;
                if (!flat_x) g += (S.fx[f][ty][tx + 1] - S.fx[f][ty][tx]) * rdx;
                if (HAS_Y) g += (S.fy[f][ty + 1][tx] - S.fy[f][ty][tx]) * rdy;
                g = -(g + (zt[a] - zb[a]) * rdz);
                if (f == 2) g = (k >= 1) ? g + 0.5 * (b_here + b_below) : 0.0;
                double r;
                if (P.mode == 1) r = g;
                else {
                    double un = Uc[a] + P.dt * g;
                    r = (P.alpha == 1.0) ? un : (1.0 - P.alpha) * U0c[a] + P.alpha * un;
                    if (f == 2 && k == 0) r = 0.0;                     // impenetrable bottom wall
                }
                P.out[f][n] = r;
            }
        }
        zb0 = zt0; zb1 = zt1; zb2 = zt2; b_below = b_here;
        // the next level's staging barrier also protects fx / fy
    }
}
