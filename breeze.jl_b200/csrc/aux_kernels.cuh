// aux_kernels.cuh — halo fills, momentum projection, host <-> device marshalling, diagnostics and reductions.
#pragma once
#include "common.cuh"
#include "stage_kernel.cuh"   // thermodynamic helpers shared with the diagnostics

struct FieldSet { double* f[NPROG + 1]; int n; };

// fill_halo_regions! for Periodic x / y on one GPU: every ghost point copies its periodic image from the interior
// (one launch for up to six fields; corners included). mode bit 0: x ghosts, bit 1: y ghosts (full padded width,
// used after an x-face exchange across GPUs so that the corners pick up the neighbour's data).
__global__ void halo_fill_periodic(Layout L, FieldSet F, int mode) {
    const int ny_rows = (mode & 2) ? 2 * L.HY : 0;                 // ghost rows, full width PX
    const int nx_cols = (mode & 1) ? 2 * L.HX : 0;                 // ghost columns of the interior rows
    const int per_level = ny_rows * L.PX + nx_cols * L.Ny;
    const long long total = (long long)per_level * L.Nz;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int k = (int)(e / per_level), r = (int)(e % per_level);
        int px, py;
        if (r < ny_rows * L.PX) {
            int row = r / L.PX; px = r % L.PX;
            py = row < L.HY ? row : L.Ny + row;                    // rows 0..HY-1 and Ny+HY..Ny+2HY-1
        } else {
            r -= ny_rows * L.PX;
            int col = r % nx_cols; py = L.HY + r / nx_cols;
            px = col < L.HX ? col : L.nx + col;
        }
        int si = px - L.HX, sj = py - L.HY;
        if (mode & 1) { if (si < 0) si += L.nx; else if (si >= L.nx) si -= L.nx; }
        if (sj < 0) sj += L.Ny; else if (sj >= L.Ny) sj -= L.Ny;
        long long dst = ((long long)k * L.PY + py) * L.PX + px;
        long long src = ((long long)k * L.PY + (sj + L.HY)) * L.PX + (si + L.HX);
        for (int f = 0; f < F.n; ++f) F.f[f][dst] = F.f[f][src];
    }
}

// _pressure_correct_momentum! (src/AnelasticEquations/anelastic_time_stepping.jl:45-54), in place.
// cells per thread of project_momentum: 4 in the Float32 build (3.51 -> 3.08 ms per step at 512^3), 1 in Float64 (4 measured 12 % slower:
// 1.37 -> 1.53 ms per launch, profiles/r2t_launches_bench_512.txt against r2l)
#ifdef BZ_F32
#define PROJ_ILP 4
#else
#define PROJ_ILP 1
#endif
// wrap_x / the y wrap: the periodic images of φ are addressed directly, so φ needs no ghost fill on one GPU (across slabs, wrap_x = 0,
// its first ghost column on the left is pulled from the neighbour).
__global__ void project_momentum(Layout L, Columns col, double* __restrict__ ru, double* __restrict__ rv, double* __restrict__ rw,
                                 const double* __restrict__ phi, double dt, int wrap_x) {
    // PROJ_ILP cells per thread, blockDim.x apart: the loads of all of them are in flight together (a 4-byte real moves half the bytes
    // per access, so one cell per thread left the Float32 build at half of the HBM rate)
    const int j = blockIdx.y, k = blockIdx.z;
    const int i0 = blockIdx.x * (blockDim.x * PROJ_ILP) + threadIdx.x;
    double p[PROJ_ILP], px[PROJ_ILP], py[PROJ_ILP], pz[PROJ_ILP], u[PROJ_ILP], v[PROJ_ILP], w[PROJ_ILP];
    const double rc = col.rho[k], rf = col.rho_f[k];
#pragma unroll
    for (int q = 0; q < PROJ_ILP; ++q) {
        const int i = i0 + q * blockDim.x;
        if (i < L.nx) {
            const long long n = lidx(L, i, j, k);
            p[q] = phi[n];
            px[q] = L.flat_x ? p[q] : phi[(wrap_x && i == 0) ? n - 1 + L.nx : n - 1];
            py[q] = L.flat_y ? p[q] : phi[j == 0 ? n - L.PX + (long long)L.Ny * L.PX : n - L.PX];
            pz[q] = k >= 1 ? phi[n - L.plane] : p[q];
            u[q] = ru[n]; v[q] = rv[n]; w[q] = rw[n];
        }
    }
#pragma unroll
    for (int q = 0; q < PROJ_ILP; ++q) {
        const int i = i0 + q * blockDim.x;
        if (i < L.nx) {
            const long long n = lidx(L, i, j, k);
            if (!L.flat_x) ru[n] = u[q] - rc * dt * ((p[q] - px[q]) * L.rdx);
            if (!L.flat_y) rv[n] = v[q] - rc * dt * ((p[q] - py[q]) * L.rdy);
            if (k >= 1) rw[n] = w[q] - rf * dt * ((p[q] - pz[q]) * L.rdz);
        }
    }
}

// The same projection restricted to the x columns [ia, ia + na) and [ib, ib + nb): on a slab the edge columns are corrected first so
// that their exchange with the neighbours overlaps the projection of the interior columns (second stream, api.cu pressure_correct).
__global__ void project_momentum_columns(Layout L, Columns col, double* __restrict__ ru, double* __restrict__ rv, double* __restrict__ rw,
                                         const double* __restrict__ phi, double dt, int ia, int na, int ib, int nb) {
    const int nc = na + nb, k = blockIdx.y;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nc * L.Ny; e += gridDim.x * blockDim.x) {
        const int c = e % nc, j = e / nc;
        const int i = c < na ? ia + c : ib + (c - na);
        long long n = lidx(L, i, j, k);
        double p = phi[n];
        double rc = col.rho[k];
        if (!L.flat_x) ru[n] -= rc * dt * ((p - phi[n - 1]) * L.rdx);
        if (!L.flat_y) rv[n] -= rc * dt * ((p - phi[j == 0 ? n - L.PX + (long long)L.Ny * L.PX : n - L.PX]) * L.rdy);
        if (k >= 1) rw[n] -= col.rho_f[k] * dt * ((p - phi[n - L.plane]) * L.rdz);
    }
}

// dense interior (x fastest, nz_out levels) <- padded field; levels >= Nz are the top wall (0)
__global__ void extract_interior(Layout L, const double* __restrict__ src, double* __restrict__ dst, int nz_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= L.nx) return;
    dst[((size_t)k * L.Ny + j) * L.nx + i] = (k < L.Nz) ? src[lidx(L, i, j, k)] : 0.0;
}
__global__ void scatter_interior(Layout L, const double* __restrict__ src, double* __restrict__ dst, int zero_level0) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= L.nx) return;
    double v = src[((size_t)k * L.Ny + j) * L.nx + i];
    if (zero_level0 && k == 0) v = 0.0;
    dst[lidx(L, i, j, k)] = v;
}

// Diagnostic fields on demand (the stage kernel never materialises them):
// _compute_velocities! / _compute_auxiliary_thermodynamic_variables! (update_atmosphere_model_state.jl:248-292)
template <int MICRO>
__global__ void diagnose_field(Layout L, Columns col, Thermo th, FieldSet U, int which, double* __restrict__ dst, int nz_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;
    if (i >= L.nx) return;
    double v = 0.0;
    if (k < L.Nz) {
        long long n = lidx(L, i, j, k);
        switch (which) {
            case BZ_U: v = U.f[0][n] / col.rho[k]; break;
            case BZ_V: v = U.f[1][n] / col.rho[k]; break;
            case BZ_W: v = U.f[2][n] / col.rho_f[k]; break;
            case BZ_THETA: v = U.f[3][n] / col.rho[k]; break;
            default: {
                double theta = U.f[3][n] / col.rho[k], q = U.f[4][n] / col.rho[k];
                double qv = q, ql = 0.0, T;
                if (MICRO == BZ_THERMO_STATIC_ENERGY) T = (theta - th.g * (th.z0 + (k + 0.5) * th.dz)) / ((1.0 - q) * th.cpd + q * th.cpv);
                else if (MICRO == BZ_MICROPHYSICS_NONE) T = (q == 0.0) ? col.exner_dry[k] * theta : lipt_temperature(th, theta, col.log_p_pst[k], q, 0.0);
                else T = saturation_adjust(th, theta, col.p[k], col.log_p_pst[k], q, qv, ql);
                v = which == BZ_T ? T : (which == BZ_QV ? qv : ql);
            }
        }
    }
    dst[((size_t)k * L.Ny + j) * L.nx + i] = v;
}

#ifdef BZ_F32
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) { atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v)); }
__device__ __forceinline__ bool is_nonfinite(float v) { return (__float_as_int(v) & 0x7f800000) == 0x7f800000; }
#else
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}
// exponent bits all ones <=> NaN or Inf (integer test: no FP64 pipe)
__device__ __forceinline__ bool is_nonfinite(double v) { return (__double2hiint(v) & 0x7ff00000) == 0x7ff00000; }
#endif

// which = 0: max |div(ρu)|; which = 1: max (|u|/Δx + |v|/Δy + |w|/Δz)  (cell_advection_timescale.jl:46-65)
__global__ void reduce_max_kernel(Layout L, Columns col, const double* __restrict__ ru, const double* __restrict__ rv,
                                  const double* __restrict__ rw, int which, double* __restrict__ out) {
    double m = 0.0;
    const long long total = (long long)L.nx * L.Ny * L.Nz;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int i = (int)(e % L.nx), j = (int)((e / L.nx) % L.Ny), k = (int)(e / ((long long)L.nx * L.Ny));
        long long n = lidx(L, i, j, k);
        double v;
        if (which == 0) {
            double d = 0.0;
            if (!L.flat_x) d += (ru[n + 1] - ru[n]) * L.rdx;
            if (!L.flat_y) d += (rv[n + L.PX] - rv[n]) * L.rdy;
            double wt = (k + 1 < L.Nz) ? rw[n + L.plane] : 0.0;
            d += (wt - rw[n]) * L.rdz;
            v = fabs(d);
        } else {
            v = fabs(rw[n] / col.rho_f[k]) * L.rdz;
            if (!L.flat_x) v += fabs(ru[n] / col.rho[k]) * L.rdx;
            if (!L.flat_y) v += fabs(rv[n] / col.rho[k]) * L.rdy;
        }
        m = fmax(m, v);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomic_max_nonneg(out, m);
}

// NaNChecker: flag[0] = 1 if any interior value of the F.n fields is NaN or Inf
__global__ void nonfinite_kernel(Layout L, FieldSet F, int* __restrict__ flag) {
    const long long total = (long long)L.nx * L.Ny * L.Nz;
    bool bad = false;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int i = (int)(e % L.nx), j = (int)((e / L.nx) % L.Ny), k = (int)(e / ((long long)L.nx * L.Ny));
        long long n = lidx(L, i, j, k);
        for (int f = 0; f < F.n; ++f) {
            bad |= is_nonfinite(F.f[f][n]);
        }
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// dense 2-D slice <- dense 3-D interior array (nx x Ny x nz_out, x fastest); axis 0: x = index, 1: y = index, 2: z = index
__global__ void slice_kernel(int nx, int Ny, int nz_out, const double* __restrict__ src, double* __restrict__ dst, int axis, int index) {
    const int n0 = axis == 0 ? Ny : nx, n1 = axis == 2 ? Ny : nz_out;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n0 * n1; e += gridDim.x * blockDim.x) {
        const int a = e % n0, b = e / n0;
        const int i = axis == 0 ? index : a, j = axis == 0 ? a : (axis == 1 ? index : b), k = axis == 2 ? index : b;
        dst[e] = src[((size_t)k * Ny + j) * nx + i];
    }
}

// x-face packing for the slab halo exchange: `w` columns starting at interior index i_src of nf fields -> contiguous buffer
__global__ void pack_x_faces(Layout L, FieldSet F, int i_src, int w, double* __restrict__ buf) {
    const long long per_field = (long long)w * L.Ny * L.Nz;
    const long long total = per_field * F.n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int f = (int)(e / per_field); long long r = e % per_field;
        int c = (int)(r % w), j = (int)((r / w) % L.Ny), k = (int)(r / ((long long)w * L.Ny));
        buf[e] = F.f[f][lidx(L, i_src + c, j, k)];
    }
}
__global__ void unpack_x_faces(Layout L, FieldSet F, int i_dst, int w, const double* __restrict__ buf) {
    const long long per_field = (long long)w * L.Ny * L.Nz;
    const long long total = per_field * F.n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int f = (int)(e / per_field); long long r = e % per_field;
        int c = (int)(r % w), j = (int)((r / w) % L.Ny), k = (int)(r / ((long long)w * L.Ny));
        F.f[f][lidx(L, i_dst + c, j, k)] = buf[e];
    }
}

// Horizontal sums of the four forced prognostics per level (compute_forcing!(::SubsidenceForcing): the averages are
// recomputed every update_state!). grid = Nz blocks; sums[f * Nz + k].
__global__ void column_sums(Layout L, FieldSet U, double* __restrict__ sums) {
    __shared__ double sh[4][256];
    const int k = blockIdx.x;
    const int map[4] = {0, 1, 3, 4};             // ρu, ρv, ρθ, ρq
    double a[4] = {0.0, 0.0, 0.0, 0.0};
    const int n = L.nx * L.Ny;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        long long idx = lidx(L, e % L.nx, e / L.nx, k);
#pragma unroll
        for (int f = 0; f < 4; ++f) a[f] += U.f[map[f]][idx];
    }
#pragma unroll
    for (int f = 0; f < 4; ++f) sh[f][threadIdx.x] = a[f];
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) for (int f = 0; f < 4; ++f) sh[f][threadIdx.x] += sh[f][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 4) sums[threadIdx.x * L.Nz + k] = sh[threadIdx.x][0];
}

// ρ × (subsidence + geostrophic + prescribed) specific forcing per level (src/Forcings/subsidence_forcing.jl:84-100,
// geostrophic_forcings.jl:74-84, specific_forcing.jl:70-74). One block; fcol[f * Nz + k].
__global__ void column_forcing_kernel(int Nz, double dz, double inv_n, const double* __restrict__ rho, const double* __restrict__ sums,
                                      const double* __restrict__ ws, int mask, const double* __restrict__ ug, const double* __restrict__ vg,
                                      const double* __restrict__ q_tend, double coriolis_f, double* __restrict__ fcol) {
    for (int e = threadIdx.x; e < 4 * Nz; e += blockDim.x) {
        const int f = e / Nz, k = e % Nz;
        double F = 0.0;
        if (ws && ((mask >> f) & 1)) {
            auto mean = [&](int kk) { return sums[f * Nz + kk] * inv_n / rho[kk]; };
            double up = (k + 1 < Nz) ? ws[k + 1] * ((mean(k + 1) - mean(k)) / dz) : 0.0;
            double lo = (k > 0) ? ws[k] * ((mean(k) - mean(k - 1)) / dz) : 0.0;
            F -= (k == Nz - 1) ? lo : ((k == 0) ? up : 0.5 * (up + lo));
        }
        if (f == 0 && vg) F += -coriolis_f * vg[k];
        if (f == 1 && ug) F += coriolis_f * ug[k];
        if (f == 3 && q_tend) F += q_tend[k];
        fcol[e] = rho[k] * F;
    }
}
