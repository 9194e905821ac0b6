// stage_hi.cuh — the fused stage kernel for WENO(order = 7 / 9) on the anelastic path (the scheme every shipped
// example runs, examples/dry_thermal_bubble.jl:24, examples/bomex.jl:204).
//
// Same contract as stage_kernel.cuh: one launch per SSP-RK3 stage evaluates every tendency of
//   compute_x/y/z_momentum_tendency!, compute_potential_temperature_tendency! / static_energy_tendency, compute_scalar_tendency!
//   (dynamics_kernel_functions.jl:64-159, potential_temperature_tendency.jl:66-106, static_energy_tendency.jl:39-72, Advection.jl:20-35),
//   buoyancy_forceᶜᶜᶜ (anelastic_buoyancy.jl:36-72), the BOMEX forcings / flux BCs, and _ssp_rk3_substep! (ssp_runge_kutta_3.jl:167-173)
// and writes the five predictor fields once; every face flux is evaluated once.
//
// Why a second kernel and not another instantiation of stage_kernel: an order-9 reconstruction is ≈ 170 FP64 instructions against 47 for
// order 5, so the stage is FP64-bound at ≈ 3 x the WENO5 kernel whatever the staging, and the 10-point stencils would need a 12-plane
// ring of (32 + 12) x (8 + 10) tiles = 380 KB. The stencils are therefore read through L1 / L2 (each value is re-read ≈ 25 x from L1,
// ≈ 7.6 ms of L1 time at 512^3 against ≥ 20 ms of FP64 time), from five fields of VELOCITIES / SPECIFIC values (u, v, w, θ | e, q) that
// `specific_fields_kernel` forms once per stage — the reference's _compute_velocities! + the θ / q diagnostics
// (update_atmosphere_model_state.jl:248-292), 80 B per cell.
// Decomposition: a CTA owns 31 x TY columns and marches up a chunk of levels; a thread evaluates the fluxes through the LOW x face, the
// LOW y face and the TOP z face of its cell; the high x face comes from lane + 1 by warp shuffle (tiles overlap by one column), the high
// y face from the row above through shared memory (the row of y faces just above the tile is spread one flux kind per warp), the bottom
// z face is carried in registers.
// Fields carry BUF + 1 ghost cells in x / y (Layout::HX, HY) and ONE extra zero plane on top (level Nz: the top wall of ρw / w), so the
// z stencils of the wall-adjacent reduced-order reconstructions never leave the allocation.
#pragma once
#include "common.cuh"
#include "weno.cuh"
#include "stage_kernel.cuh"      // StageParams, buoyancy_center, thermodynamic helpers

#define HI_TX 31

// biased reconstruction at the "face" between f[n - s] and f[n] with the buffer in use R <= BUF (R < BUF only next to the z walls)
template <int BUF>
__device__ __forceinline__ double hi_biased(const double* __restrict__ f, long long n, long long s, int R, bool left) {
    if (R >= BUF) return weno_hi_mem<BUF>(f, n, s, left);
    if constexpr (BUF >= 5) { if (R == 4) return weno_hi_mem<4>(f, n, s, left); }
    if (R == 3) return biased6c<3>(f[n - 3 * s], f[n - 2 * s], f[n - s], f[n], f[n + s], f[n + 2 * s], left);
    if (R == 2) return left ? weno3z(f[n - 2 * s], f[n - s], f[n]) : weno3z(f[n + s], f[n], f[n - s]);
    return left ? f[n - s] : f[n];
}
// Centered(order = 2 R) interpolation to the face between a[n - s] and a[n]; R <= BUF - 1
template <int BUF>
__device__ __forceinline__ double hi_sym(const double* __restrict__ a, long long n, long long s, int R) {
    if constexpr (BUF >= 5) { if (R >= 4) return centered_hi_mem<4>(a, n, s); }
    if (R == 3) return centered_hi_mem<3>(a, n, s);
    if (R == 2) return ((7.0 / 12.0) * (a[n - s] + a[n])) - ((1.0 / 12.0) * (a[n - 2 * s] + a[n + s]));
    return 0.5 * (a[n - s] + a[n]);
}

struct SpecificFields { const double* U[NPROG]; double* V[NPROG]; };

// _compute_velocities! and the specific thermodynamic variables over the whole padded planes (the ghost cells of U are valid on entry):
// u = ρu / ρᵣ, v = ρv / ρᵣ, w = ρw / ℑzρᵣ, θ | e = ρθ / ρᵣ, q = ρq / ρᵣ — as multiplications by the reciprocals, like stage_kernel does.
__global__ void specific_fields_kernel(Layout L, Columns col, SpecificFields F) {
    const int k = blockIdx.y;
    const double sc = col.rho_inv[k], sf = col.rho_f_inv[k];
    const long long base = (long long)k * L.plane;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < L.plane; e += (long long)gridDim.x * blockDim.x) {
        const long long n = base + e;
#pragma unroll
        for (int f = 0; f < NPROG; ++f) F.V[f][n] = F.U[f][n] * (f == BZ_RHO_W ? sf : sc);
    }
}

#ifndef BZ_HI_MINB
#define BZ_HI_MINB 2     // 128 registers, 2 CTAs of 256 threads per SM: 9.45 -> 6.36 ms per launch at 256^3, order 9 (profiles/r2c_hi_order_variants.txt)
#endif
template <int BUF, int TY, int MICRO, bool FORCED>
__global__ void __launch_bounds__(32 * TY, TY >= 8 ? BZ_HI_MINB : 1) stage_hi_kernel(const __grid_constant__ StageParams P, SpecificFields F) {
    constexpr int BS = BUF - 1;                          // WENO(order) advects with Centered(order - 1)
    constexpr int NK = NPROG;
    __shared__ double sfy[2][NK][TY + 1][32];            // y-face fluxes, double-buffered by level parity: one CTA barrier per level
    const Layout& L = P.L;
    const int lane = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * HI_TX + lane, j = blockIdx.y * TY + ty;
    const int Nz = L.Nz;
    const int kb = blockIdx.z * P.k_chunk, ke = min(Nz, kb + P.k_chunk);
    if (kb >= Nz) return;
    const long long SX = 1, SY = L.PX, SZ = L.plane;
    const bool fx_ = L.flat_x, fy_ = L.flat_y;
    const bool col_ok = (i <= L.nx) && (j < L.Ny) && (fx_ ? i < L.nx : true);      // i = nx: only the low-x-face fluxes are needed
    const bool own = (lane < HI_TX) && (i < L.nx) && (j < L.Ny);
    const bool yrow_ok = (lane < HI_TX) && (i < L.nx) && (j <= L.Ny);
    const int j_top = blockIdx.y * TY + TY;                                        // the row of y faces just above a full tile
    const bool top_ok = (TY >= NK) && (ty < NK) && (j_top <= L.Ny) && (lane < HI_TX) && (i < L.nx);
    const double* __restrict__ ru = P.U[0]; const double* __restrict__ rv = P.U[1]; const double* __restrict__ rw = P.U[2];
    const double* __restrict__ u = F.V[0]; const double* __restrict__ v = F.V[1]; const double* __restrict__ w = F.V[2];
    const double rdx = L.rdx, rdy = L.rdy, rdz = L.rdz;

    // advecting momenta: Centered(order - 1) of ρ𝐮 (identity along a Flat dimension); advected: WENO of the velocity / specific field
    auto symx = [&](const double* a, long long m) { return fx_ ? a[m] : hi_sym<BUF>(a, m, SX, BS); };
    auto symy = [&](const double* a, long long m) { return fy_ ? a[m] : hi_sym<BUF>(a, m, SY, BS); };
    // kinds: 0 ρu, 1 ρv, 2 ρw, 3 θ | e, 4 q.  x fluxes through x-face i (kind 0: at centre i - 1) of the cell with index m at level kk
    auto x_flux = [&](int kind, long long m, int kk, double rho_k) -> double {
        if (kind == 0) { double t = hi_sym<BUF>(ru, m, SX, BS); return t * hi_biased<BUF>(u, m, SX, BUF, positive(t)); }
        if (kind == 1) { double t = symy(ru, m); return t * hi_biased<BUF>(v, m, SX, BUF, positive(t)); }
        if (kind == 2) { if (kk == 0) return 0.0; double t = hi_sym<BUF>(ru, m, SZ, red_face(kk, Nz, BS)); return t * hi_biased<BUF>(w, m, SX, BUF, positive(t)); }
        const double t = u[m];
        return rho_k * t * hi_biased<BUF>(F.V[kind], m, SX, BUF, positive(t));
    };
    auto y_flux = [&](int kind, long long m, int kk, double rho_k) -> double {
        if (kind == 0) { double t = symx(rv, m); return t * hi_biased<BUF>(u, m, SY, BUF, positive(t)); }
        if (kind == 1) { double t = hi_sym<BUF>(rv, m, SY, BS); return t * hi_biased<BUF>(v, m, SY, BUF, positive(t)); }
        if (kind == 2) { if (kk == 0) return 0.0; double t = hi_sym<BUF>(rv, m, SZ, red_face(kk, Nz, BS)); return t * hi_biased<BUF>(w, m, SY, BUF, positive(t)); }
        const double t = v[m];
        return rho_k * t * hi_biased<BUF>(F.V[kind], m, SY, BUF, positive(t));
    };
    // z fluxes through z-face kf (kind 2: at centre kf - 1) above / below the cell; m = index of (i, j, kf)
    auto z_flux = [&](int kind, long long m, int kf) -> double {
        if (kind == 2) {
            const int kc = kf - 1;
            double t = hi_sym<BUF>(rw, m, SZ, red_center(kc, Nz, BS));
            return t * hi_biased<BUF>(w, m, SZ, red_center(kc, Nz, BUF), positive(t));
        }
        if (kf == 0 || kf == Nz) return 0.0;             // w = 0 exactly on the walls
        if (kind == 0) { double t = symx(rw, m); return t * hi_biased<BUF>(u, m, SZ, red_face(kf, Nz, BUF), positive(t)); }
        if (kind == 1) { double t = symy(rw, m); return t * hi_biased<BUF>(v, m, SZ, red_face(kf, Nz, BUF), positive(t)); }
        const double t = w[m];
        return P.col.rho_f[kf] * t * hi_biased<BUF>(F.V[kind], m, SZ, red_face(kf, Nz, BUF), positive(t));
    };
    auto buoy = [&](int kk, long long m, double* cpm_pi) -> double {
        return buoyancy_center<MICRO>(P.th, P.col, kk, P.col.rho[kk], P.col.exner_dry[kk], P.col.T[kk], F.V[3][m], F.V[4][m], cpm_pi);
    };

    const long long n0 = lidx(L, min(i, L.nx), min(j, L.Ny), 0);
    const long long n0_top = lidx(L, min(i, L.nx), min(j_top, L.Ny), 0);
    // carried from the level below: z-type fluxes through the bottom face, buoyancy of the level below
    double zb[NK] = {0.0, 0.0, 0.0, 0.0, 0.0}, b_below = 0.0, b_carry = 0.0;
    if (own && kb > 0) {
        const long long n = n0 + (long long)kb * SZ;
#pragma unroll
        for (int f = 0; f < NK; ++f) zb[f] = z_flux(f, n, kb);
        b_below = buoy(kb - 1, n - SZ, nullptr);
    }
    for (int k = kb; k < ke; ++k) {
        const long long n = n0 + (long long)k * SZ;
        const double rho_k = P.col.rho[k];
        // low-x-face fluxes of this column (every lane of a valid row, including the overlap lane) and the neighbour's by shuffle
        double xf[NK], xe[NK];
#pragma unroll
        for (int f = 0; f < NK; ++f) {
            xf[f] = (col_ok && !fx_) ? x_flux(f, n, k, rho_k) : 0.0;
            xe[f] = __shfl_down_sync(0xffffffffu, xf[f], 1);
        }
        // low-y-face fluxes → shared memory (this level's buffer)
        auto& S = sfy[k & 1];
        double yf[NK];
#pragma unroll
        for (int f = 0; f < NK; ++f) {
            yf[f] = (yrow_ok && !fy_) ? y_flux(f, n, k, rho_k) : 0.0;
            S[f][ty][lane] = yf[f];
        }
        if (top_ok && !fy_) S[ty][TY][lane] = y_flux(ty, n0_top + (long long)k * SZ, k, rho_k);
        __syncthreads();                                   // the only barrier of the level (the other buffer is written next level)
        if (own) {
            double G[NK];
            double zt[NK];
#pragma unroll
            for (int f = 0; f < NK; ++f) {
                zt[f] = z_flux(f, n + SZ, k + 1);
                double g = 0.0;
                if (!fx_) g += (xe[f] - xf[f]) * rdx;
                if (!fy_) g += (((TY >= NK) ? S[f][ty + 1][lane] : 0.0) - yf[f]) * rdy;
                G[f] = -(g + (zt[f] - zb[f]) * rdz);
            }
            double cpm_pi = 1.0, b_here, b_above = 0.0;
            if (MICRO == BZ_THERMO_STATIC_ENERGY) {
                // the ρe tendency needs the buoyancy at k-1, k, k+1 (static_energy_tendency.jl:60-63): evaluated one level ahead and carried
                b_here = (k == kb) ? buoy(k, n, nullptr) : b_carry;
                b_above = (k + 1 < Nz) ? buoy(k + 1, n + SZ, nullptr) : 0.0;
                G[3] -= 0.5 * (w[n] * (0.5 * (b_here + b_below)) + w[n + SZ] * (0.5 * (b_above + b_here)));     // w = 0 on both walls
            } else {
                b_here = buoy(k, n, (FORCED && P.e_tend) ? &cpm_pi : nullptr);
            }
            G[2] = (k >= 1) ? G[2] + 0.5 * (b_here + b_below) : 0.0;
            if (FORCED) {
                // FPlane Coriolis, horizontally uniform forcings, prescribed energy tendency, bottom flux BCs (bz_forcing)
                const double rv_fc = 0.25 * ((rv[n - SX] + rv[n]) + (fy_ ? rv[n - SX] + rv[n] : rv[n + SY - SX] + rv[n + SY]));
                const double ru_cf = 0.25 * ((ru[n] + ru[n + SX]) + (fy_ ? ru[n] + ru[n + SX] : ru[n - SY] + ru[n - SY + SX]));
                G[0] += P.coriolis_f * rv_fc + P.fcol[0][k];
                G[1] += -P.coriolis_f * ru_cf + P.fcol[1][k];
                G[3] += P.fcol[2][k];
                if (P.e_tend) G[3] += rho_k * P.e_tend[k] / cpm_pi;
                G[4] += P.fcol[3][k];
                if (k == 0) {
                    G[3] += P.theta_flux_dz; G[4] += P.q_flux_dz;
                    if (P.drag_dz != 0.0) {
                        G[0] -= P.drag_dz * ru[n] / sqrt(ru[n] * ru[n] + rv_fc * rv_fc);
                        G[1] -= P.drag_dz * rv[n] / sqrt(ru_cf * ru_cf + rv[n] * rv[n]);
                    }
                }
            }
#pragma unroll
            for (int f = 0; f < NK; ++f) {
                double r;
                if (P.mode == 1) r = G[f];
                else {
                    const double un = P.U[f][n] + P.dt * G[f];
                    r = (P.alpha == 1.0) ? un : (1.0 - P.alpha) * P.U0[f][n] + P.alpha * un;
                    if (f == 2 && k == 0) r = 0.0;                     // impenetrable bottom wall
                }
                P.out[f][n] = r;
                zb[f] = zt[f];
            }
            b_below = b_here; b_carry = b_above;
        }
    }
}
