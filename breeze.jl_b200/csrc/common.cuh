// common.cuh — context, layout and small device helpers shared by the kernels of libbreeze_b200.so.
//
// Device layout of a 3-D field (one x-slab per GPU):
//   padded in x and y by HX / HY = 4 ghost cells (0 in a Flat dimension), NOT padded in z:
//     index(i, j, k) = (k * PY + (j + HY)) * PX + (i + HX),   PX = nx + 2 HX, PY = Ny + 2 HY
//   x fastest (the reference's memory order), so a warp reads 256 contiguous bytes and a z-plane
//   tile (x-range × y-range) is one TMA box. Ghost cells are kept valid by halo_fill (periodic copy on one
//   GPU, NCCL exchange of x-faces across GPUs); Bounded z needs no ghosts because the reconstructions
//   drop their order next to the walls (weno.cuh).
//   rho_w lives on z-faces: level k is the BOTTOM face of cell k, k = 0..Nz-1; face 0 is the bottom wall
//   (stored, always 0) and face Nz is the top wall (not stored, always 0).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/breeze_b200.h"

#define BZ_HALO 4
// machine epsilon of the library's real type (the Thomas pivot rule |β| < 10 eps of the reference's solver)
#ifdef BZ_F32
#define BZ_REAL_EPS 1.1920929e-7
#else
#define BZ_REAL_EPS 2.220446049250313e-16
#endif
// Float32 library: the stage kernel reconstructs its WENO5-Z fluxes two at a time in f32x2 registers (weno.cuh weno5z_x2; 7.19 -> 6.60 ms
// per 512^3 launch, profiles/r2l_f32_packed.txt). -DBZ_F32_SCALAR rebuilds the one-at-a-time form.
#if defined(BZ_F32) && !defined(BZ_F32_SCALAR) && !defined(BZ_F32_PACKED)
#define BZ_F32_PACKED 1
#endif
#define NPROG 5
#define NFAM 8

struct Layout {
    int nx, Ny, Nz;          // local interior size
    int HX, HY;              // ghost widths
    int PX, PY;              // padded row / plane sizes
    long long plane;         // PX * PY
    long long n;             // PX * PY * Nz
    int flat_x, flat_y;
    double dx, dy, dz;
    double rdx, rdy, rdz;    // reciprocals (0 in a Flat dimension: the derivative vanishes)
};

__host__ __device__ __forceinline__ long long lidx(const Layout& L, int i, int j, int k) {
    return ((long long)k * L.PY + (j + L.HY)) * L.PX + (i + L.HX);
}

// per-level column data of the reference state, all length Nz (+1 where noted), device pointers
struct Columns {
    const double* rho;       // ρᵣ at centres
    const double* rho_inv;   // 1/ρᵣ
    const double* rho_f;     // ℑz ρᵣ at z-faces k = 0..Nz (wall faces hold the one-sided value; only ever × 0)
    const double* rho_f_inv; // 1/rho_f
    const double* p;         // pᵣ
    const double* T;         // Tᵣ
    const double* exner_dry; // (pᵣ/pˢᵗ)^(Rᵈ/cᵖᵈ)
    const double* log_p_pst; // log(pᵣ/pˢᵗ)
    const double* lev;       // per-level records of the stage kernel: LEV_REC doubles for each k (layout: upload_columns in api.cu)
};

// One record per level k with every column value the stage kernel's level k consumes, so that a level is one 128-byte
// relay through shared memory instead of fourteen bounds-checked loads rolled through registers:
//   0-3  ρᵣ at centres k-2 .. k+1        4-7  ℑzρᵣ at z-faces k-1 .. k+2      (0 outside the column)
//   8,9  Exner (dry), Tᵣ at k            10,11  1/ρᵣ, 1/ℑzρᵣ of plane k+3 (0 outside)      12,13  Exner, Tᵣ at min(k+1, Nz-1)
//   14,15  1/ρᵣ, 1/ℑzρᵣ of plane k+4 (the plane staged at level k under BZ_SPLIT_BARRIER)
#define LEV_REC 16

struct Thermo {
    double Rd, Rv, cpd, cpv, cl, ci, g, Ll, Li, pst;
    double Tr_energy, Ttr, ptr;
    double z0, dz;           // heights of the cell centres z_k = z0 + (k + 1/2) dz (StaticEnergyFormulation: T = (e - g z)/cᵖᵐ)
    int microphysics;
};

// third value of the stage kernel's MICRO template parameter: StaticEnergyFormulation without microphysics
#define BZ_THERMO_STATIC_ENERGY 2

#define CUDA_TRY(ctx, call)                                                                       \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            bz_set_error((ctx), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return BZ_ERR_CUDA;                                                                   \
        }                                                                                         \
    } while (0)

void bz_set_error(bz_ctx* ctx, const char* fmt, ...);

// ---- CUDA-graph replay of a whole time step ---------------------------------------------------------------------------------------
// A step is 27 (anelastic) to 43 (compressible) launches of pure stream work. On small grids (the shipped 2-D 128 x 128 bubble, BOMEX,
// the README quick-start) the launch overhead on the host is the step time, so a step is captured once per (state-buffer rotation, dt)
// and replayed as ONE graph launch. The first step with a key runs eagerly (function attributes, lazy allocations), the second is
// captured, later ones replay. Anything that changes what the kernels' arguments mean (forcings, reference state, profiling) clears
// the cache. Used for grids of at most 2^21 cells (measured: BOMEX 128 x 128 x 75 1.73 -> 1.61 ms per step; at 256 x 256 x 64 the eager
// launches already keep the GPU busy and the replay measured 1 % slower); BZ_GRAPHS=0 disables it, BZ_GRAPHS=2 forces it for any size.
#include <vector>
struct StepGraphEntry { int key; double dt; int seen; cudaGraphExec_t exec; int cur_after; long long launches; };
struct StepGraphCache {
    std::vector<StepGraphEntry> e;
    int enabled = -1;
    bool on(long long cells) {
        if (enabled < 0) { const char* v = getenv("BZ_GRAPHS"); enabled = v ? atoi(v) : 1; }
        return enabled == 2 || (enabled == 1 && cells <= (1ll << 21));
    }
    StepGraphEntry* find(int key, double dt) { for (auto& x : e) if (x.key == key && x.dt == dt) return &x; return nullptr; }
    StepGraphEntry* add(int key, double dt) {
        if (e.size() >= 8) { if (e.front().exec) cudaGraphExecDestroy(e.front().exec); e.erase(e.begin()); }
        e.push_back(StepGraphEntry{key, dt, 0, nullptr, 0, 0});
        return &e.back();
    }
    void clear() { for (auto& x : e) if (x.exec) cudaGraphExecDestroy(x.exec); e.clear(); }
};
