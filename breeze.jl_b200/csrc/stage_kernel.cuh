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
// Decomposition: a CTA owns a column of TX × TY cells and marches up z. An 8-slot ring of z-planes
// (TX+8) × (TY+6) of the five fields lives in shared memory as VELOCITIES / SPECIFIC values (u, v, w, θ, q —
// converted once per loaded element); every x-, y- and z-stencil of the 15 WENO5 reconstructions per cell is
// read from it. Each thread computes only the fluxes through the LOW x/y faces and the TOP z face of its cell;
// the high-side x/y fluxes come from the neighbouring thread through shared memory (the extra column of x-faces and
// the extra row of y-faces at the tile's high edges are spread one flux kind per warp), the bottom z flux is carried
// in registers from the previous level. So every face flux is evaluated once.
// Planes are staged either by TMA (cp.async.bulk.tensor, one 3-D box per field and level, mbarrier
// completion) or by plain coalesced loads (selected at run time; bit-identical).
// Level schedule (default build): planes k-2 .. k+3 are in use, plane k+4 is converted and plane k+5 is in flight — the 8 ring
// slots. The level's CTA barrier is split: a thread ARRIVES after its x/y fluxes and its share of plane k+4, evaluates the
// z fluxes / buoyancy, and WAITS only before the tendency assembly reads the other threads' x/y fluxes. The per-level column
// values (reference density, Exner, conversion scales) come as one 128-byte record relayed through shared memory.
// -DBZ_PLAIN_BARRIER / -DBZ_ROLE1_STORES_Q / -DBZ_EDGE_UNBALANCED rebuild the earlier variants (profiles/r1j_stage_variants.txt).
#pragma once
#include "common.cuh"
#include "weno.cuh"
#include <type_traits>

#define RING 8
#ifdef BZ_F32
// Float32 build: the plane ring is half the size, so TWO CTAs fit an SM once role 1 stores ρq itself (no hand-over buffer: 114.6 KB per
// CTA) and the kernel is compiled for 64 registers (no spills: a Float32 value is one register). Measured at 512^3
// (profiles/r2r_f32_stage_variants.txt): 7.96 -> 7.20 ms per launch; the role-1 variant alone, at one CTA per SM, 8.26 ms.
#ifndef BZ_STAGE_MINB
#define BZ_STAGE_MINB 2
#endif
#ifndef BZ_ROLE1_STORES_Q
#define BZ_ROLE1_STORES_Q 1
#endif
#endif
#ifndef BZ_PLAIN_BARRIER
#define BZ_SPLIT_BARRIER 1     // default: split level barrier (arrive after the x/y fluxes, wait before the tendency assembly)
#endif
#if defined(BZ_SPLIT_BARRIER) && !defined(BZ_ROLE1_STORES_Q)
#define BZ_BALANCED_STORES 1   // default: role 0 stores ρq (role 1 hands over its z-flux difference), evening out the two warp roles
#endif
#if defined(BZ_BALANCED_STORES) && !defined(BZ_SPLIT_BARRIER)
#error "BZ_BALANCED_STORES evaluates a z flux before the level barrier: it needs the split barrier's plane schedule"
#endif

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
    // forcing / Coriolis / bottom flux BCs (bz_forcing); only read by the FORCED instantiation
    const double* fcol[4];            // ρ × horizontally uniform specific forcing of u, v, θ, q per level (device, Nz each)
    const double* e_tend;             // specific energy tendency per level or nullptr
    double coriolis_f, theta_flux_dz, q_flux_dz, drag_dz;   // fluxes already divided by Δz
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
// cpm_pi (optional): cᵖᵐ Π of the cell, for the Fρe / (cᵖᵐ Π) term of the ρθ tendency (potential_temperature_tendency.jl:93-104)
template <int MICRO>
__device__ __forceinline__ double buoyancy_center(const Thermo& th, const Columns& col, int k, double rho_k, double exner_dry_k, double Tr_k,
                                                  double theta, double q, double* cpm_pi = nullptr) {
    double T, Rm, qv = q, ql = 0.0;
    if (MICRO == BZ_THERMO_STATIC_ENERGY) {
        // `theta` carries the specific static energy e; temperature(::StaticEnergyState) = (e - g z)/cᵖᵐ (dynamic_states.jl:283-298)
        T = (theta - th.g * (th.z0 + (k + 0.5) * th.dz)) / ((1.0 - q) * th.cpd + q * th.cpv);
        Rm = (1.0 - q) * th.Rd + q * th.Rv;
    } else if (MICRO == BZ_MICROPHYSICS_NONE) {
        if (q == 0.0) { T = exner_dry_k * theta; Rm = th.Rd; }
        else { T = lipt_temperature(th, theta, col.log_p_pst[k], q, 0.0); Rm = (1.0 - q) * th.Rd + q * th.Rv; }
    } else {
        T = saturation_adjust(th, theta, col.p[k], col.log_p_pst[k], q, qv, ql);
        Rm = (1.0 - (qv + ql)) * th.Rd + qv * th.Rv;
    }
    if (cpm_pi) {
        double cpm = (1.0 - (qv + ql)) * th.cpd + qv * th.cpv + ql * th.cl;
        *cpm_pi = cpm * exp((Rm / cpm) * col.log_p_pst[k]);
    }
    double rho_p = rho_k * (th.Rd * Tr_k / (Rm * T) - 1.0);
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
// The level barrier's wait. A bare try_wait loop re-issues SYNCS + BRA + YIELD every few cycles while the warp waits for the slowest
// warp of the CTA (measured: 38 spins per wait, 11 % of all issued instructions, competing for issue slots with the warps that still
// have flux work): BZ_WAIT_HINT gives try_wait a suspend-time hint (ns) so the hardware parks the warp; BZ_WAIT_SLEEP backs off with nanosleep.
__device__ __forceinline__ void mbar_wait_level(uint64_t* bar, uint32_t parity) {
#if defined(BZ_WAIT_HINT)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LWAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra LWAIT_DONE;\n"
        "bra LWAIT_LOOP;\n"
        "LWAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)BZ_WAIT_HINT) : "memory");
#elif defined(BZ_WAIT_SLEEP)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SWAIT_DONE;\n"
        "SWAIT_LOOP:\n"
        "nanosleep.u32 %2;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SWAIT_DONE;\n"
        "bra SWAIT_LOOP;\n"
        "SWAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)BZ_WAIT_SLEEP) : "memory");
#else
    mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
    static constexpr int ALN = 128 / (int)sizeof(double);
    static constexpr int PLANE = (SW * SH + ALN - 1) & ~(ALN - 1);   // every field slice stays 128-byte aligned (TMA destination)
    static_assert(!HAS_Y || (TY >= NPROG && TX == 32), "the extra y-face row is spread one flux kind per warp");
    double ring[RING][NPROG][PLANE];
    double fx[2][NPROG][TY][TX + 1];                  // double-buffered by level parity: one CTA barrier per level
    double fy[2][NPROG][HAS_Y ? TY + 1 : 1][TX];
    alignas(16) double lev[4][LEV_REC];               // per-level column records (common.cuh), relayed one level ahead
    uint64_t bar[RING];
    uint64_t lbar;                                    // BZ_SPLIT_BARRIER: the level barrier as an mbarrier (arrive early, wait late)
#ifdef BZ_BALANCED_STORES
    double fz[2][TY][TX];                             // z-flux difference of ρq, handed from role 1 to role 0 (which stores ρq)
#endif
};

// Biased reconstruction from six consecutive values with the buffer R known at compile time on the fast path.
template <int R>
__device__ __forceinline__ double biased6c(double v0, double v1, double v2, double v3, double v4, double v5, bool left) {
    if (R >= 3) {
        double a = left ? v0 : v5, b = left ? v1 : v4, c = left ? v2 : v3, d = left ? v3 : v2, e = left ? v4 : v1;
        return weno5z(a, b, c, d, e);
    } else if (R == 2) {
        double a = left ? v1 : v4, b = left ? v2 : v3, c = left ? v3 : v2;
        return weno3z(a, b, c);
    }
    return left ? v2 : v3;
}

#ifndef BZ_STAGE_MINB
#define BZ_STAGE_MINB 1
#endif
template <int TX, int TY, bool HAS_Y, bool FLAT_X, int MICRO, bool FORCED>
__global__ void __launch_bounds__(2 * TX * TY, BZ_STAGE_MINB) stage_kernel(const __grid_constant__ StageParams P) {
    using SM = StageShared<TX, TY, HAS_Y>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SM& S = *reinterpret_cast<SM*>(smem_raw);
    constexpr int SW = SM::SW, SH = SM::SH, YO = SM::YO, NCELL = TX * TY, NT = 2 * NCELL;
    constexpr int PL = SM::PLANE;                 // doubles per field slice
    constexpr int SLOT = NPROG * PL;              // doubles per ring slot

    // Two threads per cell column ("roles", warp-uniform): the FP64 pipe needs ~4 warps per scheduler to stay busy and the
    // plane ring leaves room for one CTA per SM, so the 15 flux kinds of a cell are split between two warp sets that share
    // the ring:   role 0: ρu, ρv (all directions) + the x/y fluxes of θ;   role 1: ρw, the z flux of θ, ρq, buoyancy.
    // Tendency assembly and stores: role 0 → ρu, ρv, ρq;  role 1 → ρw, ρθ (BZ_BALANCED_STORES).
    const Layout& L = P.L;
    const int tid = threadIdx.x;
    const int role = tid / NCELL, ctid = tid % NCELL;
    const int tx = ctid % TX, ty = ctid / TX;
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int Nz = L.Nz;
    const int kb = blockIdx.z * P.k_chunk;
    if (kb >= Nz) return;
    const int ke = min(Nz, kb + P.k_chunk);
    // The box origin must be 16-byte aligned in global memory for TMA (even x index): tiles with an odd i0 start one
    // column further left (SW has the slack) and shift their threads by one column inside the plane.
    const int xs = i0 & 1;
    const int sx = tx + 4 + xs, sy = ty + YO;         // this thread's cell inside a plane
    double* const ring = &S.ring[0][0][0];
    const int toff = sy * SW + sx;                    // this thread's element inside a field slice

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
                if (inside && gx < L.PX && gy < L.PY && (FLAT_X ? (x == 4) : true)) {
                    int px = FLAT_X ? 0 : gx;
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
    auto convert_plane = [&](int kk, double sc_c, double sc_f) {   // raw prognostics → velocities / specific values, in place
        // the five (padded, 128-byte aligned) field slices of a slot are contiguous: 16-byte accesses, a fixed trip count, and the
        // ρw slice (pairs PL .. 3 PL / 2) picked out by comparisons that fold away in the iterations that cannot contain it
        double2* dst = reinterpret_cast<double2*>(S.ring[kk & (RING - 1)][0]);
        constexpr int NPAIR = NPROG * PL / 2;
#pragma unroll
        for (int it = 0; it < (NPAIR + NT - 1) / NT; ++it) {
            const int p = tid + it * NT;
            if ((it + 1) * NT <= NPAIR || p < NPAIR) {
                const bool maybe_f = (it * NT < 3 * PL / 2) && ((it + 1) * NT > PL);
                const double sc = (maybe_f && p >= PL && p < 3 * PL / 2) ? sc_f : sc_c;
                double2 v = dst[p];
                v.x *= sc; v.y *= sc;
                dst[p] = v;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes before a later TMA refill of the slot
    };

    uint32_t phase_bits = 0;                     // one parity bit per ring slot
#ifdef BZ_SPLIT_BARRIER
    // Split level barrier: a thread ARRIVES once its x/y fluxes of level k are in shared memory and its share of plane k+4 is
    // converted, evaluates the z fluxes (which read planes that were complete one level earlier), and only then WAITS for the
    // other warps before it reads their x/y fluxes — the arrival skew between warps is hidden behind the z-flux work.
    // Planes therefore run one level further ahead than with the plain barrier: used k-2 .. k+3, converted k+4, in flight k+5.
    constexpr int AHEAD = 4;
    uint32_t lpar = 0;
    {
        if (tid == 0) {
            for (int s = 0; s < RING; ++s) mbar_init(&S.bar[s], 1);
            mbar_init(&S.lbar, NT);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
#else
    constexpr int AHEAD = 3;
    if (P.use_tma) {
        if (tid == 0) {
            for (int s = 0; s < RING; ++s) mbar_init(&S.bar[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
#endif
    auto wait_plane_tma = [&](int kk) {
        int s = kk & (RING - 1);
        mbar_wait(&S.bar[s], (phase_bits >> s) & 1u);
        phase_bits ^= (1u << s);
    };

    const int kstart = (kb > 0) ? kb - 1 : 0;    // a chunk that starts above the ground first rebuilds the carried z fluxes
    // prologue: planes kstart-2 .. kstart+AHEAD-1 (the loop brings in kstart+AHEAD; every plane issued is also waited for)
    if (P.use_tma) {
        if (tid == 0) for (int kk = kstart - 2; kk <= kstart + AHEAD; ++kk) if (kk < kstart + AHEAD || AHEAD == 3 || kstart + 1 < ke) issue_plane_tma(kk);
        for (int kk = kstart - 2; kk < kstart + AHEAD; ++kk) { wait_plane_tma(kk); convert_plane(kk, scale_of(0, kk), scale_of(2, kk)); }
    } else {
        for (int kk = kstart - 2; kk < kstart + AHEAD; ++kk) load_plane_direct(kk);
    }
    if (tid < LEV_REC) S.lev[kstart & 3][tid] = P.col.lev[(long long)kstart * LEV_REC + tid];
    __syncthreads();                             // prologue planes are converted and visible, and the first level's record

    // carried from the level below: z-type fluxes through the bottom face (role 0: ρu, ρv; role 1: ρw, θ, q), buoyancy below
    double zb0 = 0.0, zb1 = 0.0, zb2 = 0.0, b_below = 0.0;
    double b_carry = 0.0;                        // StaticEnergyFormulation: buoyancy of the level above, evaluated one level ahead

    const double rdx = L.rdx, rdy = L.rdy, rdz = L.rdz;
    const bool own_cell = (j < L.Ny);
    const bool in_x = i < L.nx;
#ifdef BZ_BALANCED_STORES
    // role 0 assembles and stores ρu, ρv and ρq (whose z-flux difference role 1 hands over through shared memory), role 1 ρw and ρθ:
    // 8 reconstructions + 3 stores against 8 reconstructions (one of them a tile-edge flux) + buoyancy + 2 stores per level
    const int f_count = role == 0 ? 3 : 2;
    auto field_of = [&](int a) -> int { return role == 0 ? (a < 2 ? a : 4) : 2 + a; };
#else
    const int f_first = role == 0 ? 0 : 2, f_count = role == 0 ? 2 : 3;   // fields this thread assembles and stores
    auto field_of = [&](int a) -> int { return f_first + a; };
#endif

    // Column values of a level (reference density at centres k-2..k+1 and z-faces k-1..k+2, Exner / Tᵣ, conversion scales of plane
    // k+3; 0 outside the column: those entries only ever multiply zero planes or are ignored by the reduced-order interpolation)
    // come as one record per level (Columns::lev). The last half-warp of the least loaded warp fetches the next level's record
    // at the top of a level and stores it to shared memory just before the level's barrier; four buffers, so the values of
    // level k stay readable through both flux phases while the record of level k+1 lands.
    long long n = lidx(L, i, j, kstart);         // own point, advanced by one plane per level
    const bool relay_lane = tid >= NT - LEV_REC;

    for (int k = kstart; k < ke; ++k) {
        const double2* const rec = reinterpret_cast<const double2*>(S.lev[k & 3]);
        const bool relay = relay_lane && (k + 1 < ke);
        double rec_next = 0.0;
        if (relay) rec_next = P.col.lev[(long long)(k + 1) * LEV_REC + (tid - (NT - LEV_REC))];
        const double2 rc01 = rec[0], rc23 = rec[1];
        const double r_m2 = rc01.x, r_m1 = rc01.y, rho_k = rc23.x, r_p1 = rc23.y;   // ρ at centres k-2, k-1, k (this level), k+1
        // filled in after the level's barrier (only the z-flux phase reads them)
        double f_m1 = 0.0, f_0 = 0.0, rho_ft = 0.0, f_p2 = 0.0;                      // ℑz ρ at faces k-1, k, k+1 (the top face), k+2
#ifdef BZ_BALANCED_STORES
        rho_ft = rec[3].x;                                                            // the ρq z flux is evaluated with the x/y fluxes
#endif
        double ex_k = 0.0, Tr_k = 0.0, nx_ex = 0.0, nx_Tr = 0.0;                      // buoyancy inputs of level k and of level k+1

        // own-point values for the RK update: issued now, consumed after the flux phase
        const bool do_store = (k >= kb) && own_cell;
        double Uc[3] = {0.0, 0.0, 0.0}, U0c[3] = {0.0, 0.0, 0.0};
        if (do_store && P.mode == 0) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                int f = field_of(a);
                if (a < f_count && (in_x || (f == 0 && i < P.nx_u))) {
                    Uc[a] = P.U[f][n];
                    if (P.alpha != 1.0) U0c[a] = P.U0[f][n];
                }
            }
        }

        // this thread's element in the planes k-2 .. k+3 (index m+2)
        const double* Lp[6];
#pragma unroll
        for (int m = 0; m < 6; ++m) Lp[m] = ring + ((k + m - 2) & (RING - 1)) * SLOT + toff;
        const double* const Lk = Lp[2];
        const double* const Lt = Lp[3];                                // level of the top face / of cell k+1

        double zt0 = 0.0, zt1 = 0.0, zt2 = 0.0, b_here = 0.0, cpm_pi = 1.0, b_above = 0.0;

        // One level of flux work. FULL: every z stencil is at full order (2 <= k <= Nz-4): all orders are compile-time.
        auto level = [&](auto full_tag, auto phase_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            constexpr int PHASE = decltype(phase_tag)::value;          // 0: x/y fluxes → shared memory; 1: z fluxes, buoyancy
            const int kf = k + 1;
            const int Rf_top = FULL ? 3 : red_face(kf, Nz, 3);
            const int Rf_k2 = FULL ? 2 : red_face(k, Nz, 2);
            const int Rc3 = FULL ? 3 : red_center(k, Nz, 3), Rc2 = FULL ? 2 : red_center(k, Nz, 2);
            auto bz = [&](double v0, double v1, double v2, double v3, double v4, double v5, int R, bool left) -> double {
                return FULL ? biased6c<3>(v0, v1, v2, v3, v4, v5, left) : biased6(v0, v1, v2, v3, v4, v5, R, left);
            };
            auto sz = [&](double v0, double v1, double v2, double v3, int R) -> double { return FULL ? sym4(v0, v1, v2, v3, 2) : sym4(v0, v1, v2, v3, R); };
#define XS(f, dx) cell[(f) * PL + (dx)]
#define YS(f, dy) row[(f) * PL + (dy) * SW]
#define ZS(f, m) Lp[(m) + 2][(f) * PL]
            // X-type fluxes: through x-face i (or at centre i-1 for ρu); kinds 0 ρu, 1 ρv, 2 ρw, 3 θ, 4 q
            // `cell` is the plane element of the cell whose low x-face is meant (own cell, or the column just right of the tile)
            auto x_flux = [&](auto kind_tag, const double* cell, int zoff) -> double {
                constexpr int kind = decltype(kind_tag)::value;
                const double u_i = XS(0, 0);
                if (kind == 0) {          // FUu at centre i-1
                    double ut = rho_k * sym4(XS(0, -2), XS(0, -1), u_i, XS(0, 1), 2);
                    return ut * biased6c<3>(XS(0, -3), XS(0, -2), XS(0, -1), u_i, XS(0, 1), XS(0, 2), positive(ut));
                } else if (kind == 1) {   // FUv at (face i, face j)
                    double ut = HAS_Y ? rho_k * sym4(cell[-2 * SW], cell[-SW], u_i, cell[SW], 2) : rho_k * u_i;
                    return ut * biased6c<3>(XS(1, -3), XS(1, -2), XS(1, -1), XS(1, 0), XS(1, 1), XS(1, 2), positive(ut));
                } else if (kind == 2) {   // FUw at (face i, z-face k); the wall face k = 0 carries no w tendency
                    if (!FULL && k < 1) return 0.0;
                    double ut = sz(r_m2 * Lp[0][zoff], r_m1 * Lp[1][zoff], rho_k * u_i, r_p1 * Lp[3][zoff], Rf_k2);
                    return ut * biased6c<3>(XS(2, -3), XS(2, -2), XS(2, -1), XS(2, 0), XS(2, 1), XS(2, 2), positive(ut));
                } else {                  // tracer mass flux ρ u ĉ
                    return rho_k * u_i * biased6c<3>(XS(kind, -3), XS(kind, -2), XS(kind, -1), XS(kind, 0), XS(kind, 1), XS(kind, 2), positive(u_i));
                }
            };
            // Y-type fluxes: through y-face j (or at centre j-1 for ρv) of the cell whose plane element is `row`
            auto y_flux = [&](auto kind_tag, const double* row, int zoff) -> double {
                constexpr int kind = decltype(kind_tag)::value;
                const double v_j = YS(1, 0);
                if (kind == 0) {          // FVu at (face i, face j)
                    double vt = FLAT_X ? rho_k * v_j : rho_k * sym4(row[PL - 2], row[PL - 1], v_j, row[PL + 1], 2);
                    return vt * biased6c<3>(YS(0, -3), YS(0, -2), YS(0, -1), YS(0, 0), YS(0, 1), YS(0, 2), positive(vt));
                } else if (kind == 1) {   // FVv at centre j-1
                    double vt = rho_k * sym4(YS(1, -2), YS(1, -1), v_j, YS(1, 1), 2);
                    return vt * biased6c<3>(YS(1, -3), YS(1, -2), YS(1, -1), v_j, YS(1, 1), YS(1, 2), positive(vt));
                } else if (kind == 2) {   // FVw at (face j, z-face k)
                    if (!FULL && k < 1) return 0.0;
                    double vt = sz(r_m2 * Lp[0][PL + zoff], r_m1 * Lp[1][PL + zoff], rho_k * v_j, r_p1 * Lp[3][PL + zoff], Rf_k2);
                    return vt * biased6c<3>(YS(2, -3), YS(2, -2), YS(2, -1), YS(2, 0), YS(2, 1), YS(2, 2), positive(vt));
                } else {
                    return rho_k * v_j * biased6c<3>(YS(kind, -3), YS(kind, -2), YS(kind, -1), YS(kind, 0), YS(kind, 1), YS(kind, 2), positive(v_j));
                }
            };
            // Z-type fluxes through the top face k+1 (or at centre k for ρw)
            auto z_flux = [&](auto kind_tag) -> double {
                constexpr int kind = decltype(kind_tag)::value;
                const double w_top = Lt[2 * PL];                           // 0 on the top wall (zero plane)
                if (kind == 0) {          // FWu at (face i, z-face k+1)
                    double wt = FLAT_X ? rho_ft * w_top : rho_ft * sym4(Lt[2 * PL - 2], Lt[2 * PL - 1], w_top, Lt[2 * PL + 1], 2);
                    return wt * bz(ZS(0, -2), ZS(0, -1), ZS(0, 0), ZS(0, 1), ZS(0, 2), ZS(0, 3), Rf_top, positive(wt));
                } else if (kind == 1) {   // FWv at (face j, z-face k+1)
                    double wt = HAS_Y ? rho_ft * sym4(Lt[2 * PL - 2 * SW], Lt[2 * PL - SW], w_top, Lt[2 * PL + SW], 2) : rho_ft * w_top;
                    return wt * bz(ZS(1, -2), ZS(1, -1), ZS(1, 0), ZS(1, 1), ZS(1, 2), ZS(1, 3), Rf_top, positive(wt));
                } else if (kind == 2) {   // FWw at centre k: faces k-1 .. k+2 (advecting), k-2 .. k+3 (advected)
                    double wt = sz(f_m1 * ZS(2, -1), f_0 * ZS(2, 0), rho_ft * w_top, f_p2 * ZS(2, 2), Rc2);
                    return wt * bz(ZS(2, -2), ZS(2, -1), ZS(2, 0), w_top, ZS(2, 2), ZS(2, 3), Rc3, positive(wt));
                } else {                  // tracer mass flux ℑz(ρ) w ĉ
                    return rho_ft * w_top * bz(ZS(kind, -2), ZS(kind, -1), ZS(kind, 0), ZS(kind, 1), ZS(kind, 2), ZS(kind, 3), Rf_top, positive(w_top));
                }
            };
#ifdef BZ_F32_PACKED
            // Packed Float32 pairs (weno.cuh weno5z_x2): the x / y fluxes of a role are reconstructed two at a time. x_rec / y_rec gather what
            // x_flux / y_flux feed into their reconstruction — the advecting factor and the five values selected by its sign.
            struct Rec6 { float adv, a, b, c, d, e; };
            auto mk = [](float adv, float v0, float v1, float v2, float v3, float v4, float v5, bool left) -> Rec6 {
                return Rec6{adv, left ? v0 : v5, left ? v1 : v4, left ? v2 : v3, left ? v3 : v2, left ? v4 : v1};
            };
            auto x_rec = [&](auto kind_tag, const float* cell, int zoff) -> Rec6 {
                constexpr int kind = decltype(kind_tag)::value;
                const float u_i = XS(0, 0);
                float ut;
                if (kind == 0) ut = rho_k * sym4(XS(0, -2), XS(0, -1), u_i, XS(0, 1), 2);
                else if (kind == 1) ut = HAS_Y ? rho_k * sym4(cell[-2 * SW], cell[-SW], u_i, cell[SW], 2) : rho_k * u_i;
                else if (kind == 2) ut = (!FULL && k < 1) ? 0.0f : sz(r_m2 * Lp[0][zoff], r_m1 * Lp[1][zoff], rho_k * u_i, r_p1 * Lp[3][zoff], Rf_k2);
                else ut = rho_k * u_i;
                return mk(ut, XS(kind, -3), XS(kind, -2), XS(kind, -1), XS(kind, 0), XS(kind, 1), XS(kind, 2), positive(kind >= 3 ? u_i : ut));
            };
            auto y_rec = [&](auto kind_tag, const float* row, int zoff) -> Rec6 {
                constexpr int kind = decltype(kind_tag)::value;
                const float v_j = YS(1, 0);
                float vt;
                if (kind == 0) vt = FLAT_X ? rho_k * v_j : rho_k * sym4(row[PL - 2], row[PL - 1], v_j, row[PL + 1], 2);
                else if (kind == 1) vt = rho_k * sym4(YS(1, -2), YS(1, -1), v_j, YS(1, 1), 2);
                else if (kind == 2) vt = (!FULL && k < 1) ? 0.0f : sz(r_m2 * Lp[0][PL + zoff], r_m1 * Lp[1][PL + zoff], rho_k * v_j, r_p1 * Lp[3][PL + zoff], Rf_k2);
                else vt = rho_k * v_j;
                return mk(vt, YS(kind, -3), YS(kind, -2), YS(kind, -1), YS(kind, 0), YS(kind, 1), YS(kind, 2), positive(kind >= 3 ? v_j : vt));
            };
            // z pairs only in the interior-level instantiation (FULL): there every z reconstruction is the full-order one
            auto z_rec = [&](auto kind_tag) -> Rec6 {
                constexpr int kind = decltype(kind_tag)::value;
                const float w_top = Lt[2 * PL];
                float wt;
                if (kind == 0) wt = rho_ft * sym4(Lt[2 * PL - 2], Lt[2 * PL - 1], w_top, Lt[2 * PL + 1], 2);
                else if (kind == 1) wt = rho_ft * sym4(Lt[2 * PL - 2 * SW], Lt[2 * PL - SW], w_top, Lt[2 * PL + SW], 2);
                else wt = rho_ft * w_top;
                return mk(wt, ZS(kind, -2), ZS(kind, -1), ZS(kind, 0), ZS(kind, 1), ZS(kind, 2), ZS(kind, 3), positive(kind >= 3 ? w_top : wt));
            };
            auto pair_flux = [](const Rec6& A, const Rec6& B, float& fa, float& fb) {
                float ra, rb;
                upk2(weno5z_x2(pk2(A.a, B.a), pk2(A.b, B.b), pk2(A.c, B.c), pk2(A.d, B.d), pk2(A.e, B.e)), ra, rb);
                fa = A.adv * ra; fb = B.adv * rb;
            };
            constexpr bool PACKED = HAS_Y && !FLAT_X;
#else
            constexpr bool PACKED = false;
#endif
            using K0 = std::integral_constant<int, 0>; using K1 = std::integral_constant<int, 1>; using K2 = std::integral_constant<int, 2>;
            using K3 = std::integral_constant<int, 3>; using K4 = std::integral_constant<int, 4>;
            const double* const edge = Lk + (TY - ty) * SW;               // the row of y-faces just above the tile
            const int ezoff = (TY - ty) * SW;
            auto& FX = S.fx[k & 1];
            auto& FY = S.fy[k & 1];
            // the column of x-faces just right of the tile: one flux kind per warp, one lane per row
            const int wrp = ctid >> 5, lane = ctid & 31;
            const int xoff = ((lane + YO) * SW + (TX + 4 + xs)) - toff;
            const double* const xedge = Lk + xoff;
            constexpr int XE0 = HAS_Y ? 3 : 1, XE1 = HAS_Y ? 2 : 1;
            // Same x flux with the kind chosen at run time (only the advecting factor differs between kinds; the reconstruction is one
            // instruction stream), so that ONE warp evaluates the x-edge column of four kinds at once: lane → (kind, row).
            auto x_flux_rt = [&](int kind, const double* cell, int zoff) -> double {
                const double u_i = cell[0];
                double ut;
                if (kind == 0) ut = rho_k * sym4(cell[-2], cell[-1], u_i, cell[1], 2);
                else if (kind == 1) ut = HAS_Y ? rho_k * sym4(cell[-2 * SW], cell[-SW], u_i, cell[SW], 2) : rho_k * u_i;
                else if (kind == 2) ut = (!FULL && k < 1) ? 0.0 : sz(r_m2 * Lp[0][zoff], r_m1 * Lp[1][zoff], rho_k * u_i, r_p1 * Lp[3][zoff], Rf_k2);
                else ut = rho_k * u_i;
                const double* f = cell + kind * PL;
                return ut * biased6c<3>(f[-3], f[-2], f[-1], f[0], f[1], f[2], positive(kind >= 3 ? u_i : ut));
            };
#ifndef BZ_EDGE_UNBALANCED
            constexpr bool BALANCED = HAS_Y && TY == 8 && !FLAT_X;   // all tile-edge fluxes on the lighter role (role 1: 7.3 vs 8 flux units per level)
#else
            constexpr bool BALANCED = false;
#endif
            if (PHASE == 0) {
                if (role == 0) {
#ifdef BZ_F32_PACKED
                    if (PACKED) {       // six reconstructions as three packed pairs
                        float fa, fb;
                        pair_flux(x_rec(K0{}, Lk, 0), x_rec(K1{}, Lk, 0), fa, fb); FX[0][ty][tx] = fa; FX[1][ty][tx] = fb;
                        pair_flux(x_rec(K3{}, Lk, 0), y_rec(K0{}, Lk, 0), fa, fb); FX[3][ty][tx] = fa; FY[0][ty][tx] = fb;
                        pair_flux(y_rec(K1{}, Lk, 0), y_rec(K3{}, Lk, 0), fa, fb); FY[1][ty][tx] = fa; FY[3][ty][tx] = fb;
                    }
#endif
                    if (!FLAT_X) {
                        if (!PACKED) { FX[0][ty][tx] = x_flux(K0{}, Lk, 0); FX[1][ty][tx] = x_flux(K1{}, Lk, 0); FX[3][ty][tx] = x_flux(K3{}, Lk, 0); }
                        if (!BALANCED && lane < TY) {
                            if (wrp == XE0) FX[0][lane][TX] = x_flux(K0{}, xedge, xoff);
                            else if (wrp == XE0 + 1) FX[1][lane][TX] = x_flux(K1{}, xedge, xoff);
                            else if (wrp == XE0 + 2) FX[3][lane][TX] = x_flux(K3{}, xedge, xoff);
                        }
                    }
                    if (HAS_Y) {
                        if (!PACKED) { FY[0][ty][tx] = y_flux(K0{}, Lk, 0); FY[1][ty][tx] = y_flux(K1{}, Lk, 0); FY[3][ty][tx] = y_flux(K3{}, Lk, 0); }
                        // the extra row of y-faces above the tile: one flux kind per warp (rows 0..2 of this role)
                        if (!BALANCED) {
                            if (ty == 0) FY[0][TY][tx] = y_flux(K0{}, edge, ezoff);
                            else if (ty == 1) FY[1][TY][tx] = y_flux(K1{}, edge, ezoff);
                            else if (ty == 2) FY[3][TY][tx] = y_flux(K3{}, edge, ezoff);
                        }
                    }
                } else {
#ifdef BZ_BALANCED_STORES
                    zt2 = z_flux(K4{}); S.fz[k & 1][ty][tx] = zt2 - zb2;      // ρq is stored by role 0
#endif
#ifdef BZ_F32_PACKED
                    if (PACKED) {       // four reconstructions as two packed pairs
                        float fa, fb;
                        pair_flux(x_rec(K2{}, Lk, 0), x_rec(K4{}, Lk, 0), fa, fb); FX[2][ty][tx] = fa; FX[4][ty][tx] = fb;
                        pair_flux(y_rec(K2{}, Lk, 0), y_rec(K4{}, Lk, 0), fa, fb); FY[2][ty][tx] = fa; FY[4][ty][tx] = fb;
                    }
#endif
                    if (!FLAT_X) {
                        if (!PACKED) { FX[2][ty][tx] = x_flux(K2{}, Lk, 0); FX[4][ty][tx] = x_flux(K4{}, Lk, 0); }
                        if (BALANCED) {
                            // x-edge column: warp 0 takes kinds 0..3 (lane = kind * TY + row), warp 1 takes kind 4
                            if (wrp == 0 || (wrp == 1 && lane < TY)) {
                                const int kind = (wrp == 0) ? (lane >> 3) : 4, row = lane & (TY - 1);
                                const int xo = ((row + YO) * SW + (TX + 4 + xs)) - toff;
                                FX[kind][row][TX] = x_flux_rt(kind, Lk + xo, xo);
                            }
                        } else if (lane < TY) {
                            if (wrp == XE1) FX[2][lane][TX] = x_flux(K2{}, xedge, xoff);
                            else if (wrp == XE1 + 1) FX[4][lane][TX] = x_flux(K4{}, xedge, xoff);
                        }
                    }
                    if (HAS_Y) {
                        if (!PACKED) { FY[2][ty][tx] = y_flux(K2{}, Lk, 0); FY[4][ty][tx] = y_flux(K4{}, Lk, 0); }
                        if (BALANCED) {                      // y-edge row: one kind per warp, warps 2..6
                            if (ty == 2) FY[0][TY][tx] = y_flux(K0{}, edge, ezoff);
                            else if (ty == 3) FY[1][TY][tx] = y_flux(K1{}, edge, ezoff);
                            else if (ty == 4) FY[2][TY][tx] = y_flux(K2{}, edge, ezoff);
                            else if (ty == 5) FY[3][TY][tx] = y_flux(K3{}, edge, ezoff);
                            else if (ty == 6) FY[4][TY][tx] = y_flux(K4{}, edge, ezoff);
                        } else {
                            if (ty == 0) FY[2][TY][tx] = y_flux(K2{}, edge, ezoff);
                            else if (ty == 1) FY[4][TY][tx] = y_flux(K4{}, edge, ezoff);
                        }
                    }
                }
            } else {
#ifdef BZ_F32_PACKED
                constexpr bool PACKED_Z = PACKED && FULL;
#else
                constexpr bool PACKED_Z = false;
#endif
                if (role == 0) {
#ifdef BZ_F32_PACKED
                    if (PACKED_Z) pair_flux(z_rec(K0{}), z_rec(K1{}), zt0, zt1);
#endif
                    if (!PACKED_Z) { zt0 = z_flux(K0{}); zt1 = z_flux(K1{}); }
                } else {
#ifdef BZ_BALANCED_STORES
                    zt0 = z_flux(K2{}); zt1 = z_flux(K3{});
#else
                    zt0 = z_flux(K2{});
#ifdef BZ_F32_PACKED
                    if (PACKED_Z) pair_flux(z_rec(K3{}), z_rec(K4{}), zt1, zt2);
#endif
                    if (!PACKED_Z) { zt1 = z_flux(K3{}); zt2 = z_flux(K4{}); }
#endif
                    if (MICRO == BZ_THERMO_STATIC_ENERGY) {
                        // the ρe tendency needs the buoyancy at k-1, k, k+1 (static_energy_tendency.jl:60-63): evaluate it one level ahead
                        b_here = (k == kstart) ? buoyancy_center<MICRO>(P.th, P.col, k, rho_k, ex_k, Tr_k, Lk[3 * PL], Lk[4 * PL]) : b_carry;
                        b_above = (k + 1 < Nz) ? buoyancy_center<MICRO>(P.th, P.col, k + 1, r_p1, nx_ex, nx_Tr, Lt[3 * PL], Lt[4 * PL]) : 0.0;
                    } else
                    b_here = buoyancy_center<MICRO>(P.th, P.col, k, rho_k, ex_k, Tr_k, Lk[3 * PL], Lk[4 * PL], (FORCED && P.e_tend) ? &cpm_pi : nullptr);
                }
            }
#undef XS
#undef YS
#undef ZS
        };
        const bool full = (k >= 2 && k <= Nz - 4);
        using PH0 = std::integral_constant<int, 0>; using PH1 = std::integral_constant<int, 1>;
        // x/y fluxes of level k need the planes k-2 .. k+1 only …
        if (full) level(std::true_type{}, PH0{}); else level(std::false_type{}, PH0{});
        if (relay) S.lev[(k + 1) & 3][tid - (NT - LEV_REC)] = rec_next;
#ifdef BZ_SPLIT_BARRIER
        // plane k+4 (first read by the z stencils of level k+1): its TMA was issued a whole level ago
        if (k + 1 < ke) {
            if (P.use_tma) { wait_plane_tma(k + 4); const double2 sc = rec[7]; convert_plane(k + 4, sc.x, sc.y); }
            else load_plane_direct(k + 4);
        }
        mbar_arrive(&S.lbar);                                          // fx / fy of this level, plane k+4 and the next record are written
        { const double2 a = rec[2], b = rec[3], c = rec[4], d = rec[6];
          f_m1 = a.x; f_0 = a.y; rho_ft = b.x; f_p2 = b.y; ex_k = c.x; Tr_k = c.y; nx_ex = d.x; nx_Tr = d.y; }
        if (full) level(std::true_type{}, PH1{}); else level(std::false_type{}, PH1{});
        mbar_wait_level(&S.lbar, lpar); lpar ^= 1u;                          // every warp has arrived: their fx / fy are visible, plane k-3 is dead
        if (P.use_tma && tid == NT - 32 && k + 2 < ke) issue_plane_tma(k + 5);   // slot of plane k-3; the least loaded warp issues
#else
        // … so plane k+3 (needed by the z stencils) is staged behind them: its TMA had a whole level to land
        if (P.use_tma) { wait_plane_tma(k + 3); const double2 sc = rec[5]; convert_plane(k + 3, sc.x, sc.y); }
        else load_plane_direct(k + 3);
        __syncthreads();                                               // the level's only CTA barrier: plane k+3, fx / fy and the next record are complete
        { const double2 a = rec[2], b = rec[3], c = rec[4], d = rec[6];
          f_m1 = a.x; f_0 = a.y; rho_ft = b.x; f_p2 = b.y; ex_k = c.x; Tr_k = c.y; nx_ex = d.x; nx_Tr = d.y; }
        if (P.use_tma && tid == NT - 32 && k + 1 < ke) issue_plane_tma(k + 4);   // slot of plane k-4; the least loaded warp issues
        if (full) level(std::true_type{}, PH1{}); else level(std::false_type{}, PH1{});
#endif

        // ---- tendencies, RK update, store -----------------------------------------------------------------------
        if (do_store) {
            double zt[3] = {zt0, zt1, zt2}, zb[3] = {zb0, zb1, zb2};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int f = field_of(a);
                if (a >= f_count) break;
                if (!(in_x || (f == 0 && i < P.nx_u))) continue;
                double g = 0.0;
                if (!FLAT_X) g += (S.fx[k & 1][f][ty][tx + 1] - S.fx[k & 1][f][ty][tx]) * rdx;
                if (HAS_Y) g += (S.fy[k & 1][f][ty + 1][tx] - S.fy[k & 1][f][ty][tx]) * rdy;
#ifdef BZ_BALANCED_STORES
                const double dzf = (role == 0 && a == 2) ? S.fz[k & 1][ty][tx] : zt[a] - zb[a];
                g = -(g + dzf * rdz);
#else
                g = -(g + (zt[a] - zb[a]) * rdz);
#endif
                if (f == 2) g = (k >= 1) ? g + 0.5 * (b_here + b_below) : 0.0;
                if (MICRO == BZ_THERMO_STATIC_ENERGY && f == 3)      // - ℑzᵃᵃᶜ(w ℑzᵃᵃᶠ(ρb)); w = 0 on both walls (zero plane above the top)
                    g -= 0.5 * (Lk[2 * PL] * (0.5 * (b_here + b_below)) + Lt[2 * PL] * (0.5 * (b_above + b_here)));
                if (FORCED) {
                    // FPlane Coriolis, horizontally uniform forcings, prescribed energy tendency, bottom flux BCs (bz_forcing)
                    if (f == 0) {
                        double rv_fc = 0.25 * rho_k * ((Lk[PL - 1] + Lk[PL]) + (HAS_Y ? Lk[PL + SW - 1] + Lk[PL + SW] : Lk[PL - 1] + Lk[PL]));
                        g += P.coriolis_f * rv_fc + P.fcol[0][k];
                        if (k == 0 && P.drag_dz != 0.0) { double ru = rho_k * Lk[0]; g -= P.drag_dz * ru / sqrt(ru * ru + rv_fc * rv_fc); }
                    } else if (f == 1) {
                        double ru_cf = 0.25 * rho_k * ((Lk[0] + Lk[1]) + (HAS_Y ? Lk[-SW] + Lk[-SW + 1] : Lk[0] + Lk[1]));
                        g += -P.coriolis_f * ru_cf + P.fcol[1][k];
                        if (k == 0 && P.drag_dz != 0.0) { double rv = rho_k * Lk[PL]; g -= P.drag_dz * rv / sqrt(ru_cf * ru_cf + rv * rv); }
                    } else if (f == 3) {
                        g += P.fcol[2][k];
                        if (P.e_tend) g += rho_k * P.e_tend[k] / cpm_pi;
                        if (k == 0) g += P.theta_flux_dz;
                    } else if (f == 4) {
                        g += P.fcol[3][k];
                        if (k == 0) g += P.q_flux_dz;
                    }
                }
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
        zb0 = zt0; zb1 = zt1; zb2 = zt2; b_below = b_here; b_carry = b_above;
        n += L.plane;
        // fx / fy are double-buffered by level parity: the next level writes the other buffer, and the barrier after that
        // orders it against this level's readers
    }
}
