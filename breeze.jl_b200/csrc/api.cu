// api.cu — the C ABI of include/breeze_b200.h: context, orchestration of one SSP-RK3 step, marshalling.
//
// Per stage (reference: src/TimeSteppers/ssp_runge_kutta_3.jl:225-270):
//   stage_kernel      tendencies of the projected state + RK update          cur -> nxt   (1 launch)
//   halo_fill         ρu, ρv ghosts for the divergence
//   poisson_*         source term + FFT(y) → FFT(x) → Thomas(z) → FFT⁻¹(x) → FFT⁻¹(y) → φ
//   halo_fill φ, project_momentum (in place), halo_fill of the five prognostics
// U⁰ is never copied: three buffer sets rotate (the step's input set IS U⁰ while stages ping-pong on the other two).
#include <cuda_runtime.h>
#include <cuda.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "common.cuh"
// launch bounds of the transforms (CTAs per SM): the register budget follows the word size of the library's real type.
// Float32 sweep at 512^3 (profiles/r2s_f32_fft_launch_bounds.txt): (y, x) = (4, 3) 3.75 + 2.64, (5, 4) 3.54 + 2.65, (6, 5) 3.33 + 2.62,
// (8, 6) 3.22 + 2.49 ms per step (forward + inverse).
#ifdef BZ_F32
#ifndef BZ_FFT_Y_MINB
#define BZ_FFT_Y_MINB 8
#endif
#ifndef BZ_FFT_X_MINB
#define BZ_FFT_X_MINB 6
#endif
#endif
#ifndef BZ_FFT_Y_MINB
#define BZ_FFT_Y_MINB 4
#endif
#ifndef BZ_FFT_X_MINB
#define BZ_FFT_X_MINB 3
#endif
#include "stage_kernel.cuh"
#include "stage_hi.cuh"
#include "poisson.cuh"
#include "aux_kernels.cuh"
#include "comm.cuh"

struct bz_ctx {
    bz_config cfg;
    Layout L;
    Thermo th;
    Columns col;                         // device pointers
    std::vector<double> h_rho, h_p, h_T; // host copies (Nz)
    double* col_store = nullptr;         // one allocation behind `col`
    double* set[3][NPROG] = {};          // three rotating sets of prognostic fields (padded)
    CUtensorMap tmap[3][NPROG];
    int cur = 0;                         // index of the set holding the current state
    double* phi = nullptr;
    // One device allocation ("arena") holds the 15 prognostic buffers, φ and the spectral arrays, so that a single CUDA IPC
    // handle exposes them to the neighbouring ranks (peer loads over NVLink replace NCCL transfers, comm.cuh).
    double* arena = nullptr;
    size_t arena_bytes = 0, off_W = 0, off_W2 = 0;      // byte offsets of W / W2 inside the arena (fields: (s*5+f)*L.n, φ: 15*L.n doubles)
    double* G[NPROG] = {};               // tendencies, allocated on first bz_compute_tendencies
    StepGraphCache graphs;               // captured time steps (common.cuh)
    int buf = 3;                         // buffer of the advection scheme: (order + 1) / 2
    double* V[NPROG] = {};               // WENO(order = 7 / 9): velocities / specific values of the stage's input state (stage_hi.cuh)
    double* dense = nullptr;             // nx*Ny*(Nz+1) staging buffer for host transfers
    double* scalar = nullptr;            // device scalars for reductions
    double* slice_buf = nullptr;         // bz_get_slice staging (grown on demand)
    size_t slice_cap = 0;
    // Poisson solver
    PoissonGeom PG;
    double2* W = nullptr;
    double2* W2 = nullptr;               // transposed layout (multi-GPU only)
    double2 *tw_x = nullptr, *tw_y = nullptr;
    double *lam_x = nullptr, *lam_y = nullptr, *inv_beta = nullptr, *tfac = nullptr;
    long long* ky_base = nullptr;
    int* ky_kstride = nullptr;
    int* ky_owner = nullptr;
    long long* ky_base2 = nullptr;
    // forcing (bz_forcing)
    int forced = 0, subs_mask = 0;
    double coriolis_f = 0, theta_flux = 0, q_flux = 0, drag = 0;
    double* fstore = nullptr;            // ws[Nz+1] | ug | vg | q_tend | e_tend | sums[4Nz] | fcol[4Nz]
    double *d_ws = nullptr, *d_ug = nullptr, *d_vg = nullptr, *d_qt = nullptr, *d_et = nullptr, *d_sums = nullptr, *d_fcol = nullptr;
    int lines_x = 1, lines_y = 1;
    cudaStream_t stream = nullptr;
    // host <-> device marshalling: strided 3-D copies straight between the caller's dense arrays and the padded fields, in z chunks,
    // on two copy streams (one per PCIe direction) so that a download and the upload that follows it run full duplex
    cudaStream_t s2 = nullptr;           // second compute stream: x-halo exchanges overlapped with the Poisson solve / the interior projection (slabs)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_chunk[8] = {};
    int dma_transpose = 0;               // slabs: BZ_DMA_TRANSPOSE=1 runs the transposes as copy-engine transfers pipelined over z chunks; default: peer loads
                                         // inside the transforms. Measured at 2 slabs of 512^3 (profiles/r2v_transpose_variants.txt): peer loads 29.86 ms per step;
                                         // DMA with 1 / 2 / 4 / 8 chunks 31.57 / 30.45 / 30.00 / 30.52 — the copies do overlap the transforms, but they add a pass over
                                         // W2 that the fused peer loads do not have
    int fft_z_chunks = 1;                // z chunks of the pipelined distributed transform (BZ_FFT_Z_CHUNKS). Measured at 2 slabs of 512^3
                                         // (profiles/r2j_fft_z_chunks.txt): 1 -> 29.73, 2 -> 29.77, 4 -> 30.25, 8 -> 30.79 ms per step — the pulls
                                         // already run at NVLink rate and do not overlap the next chunk's transform, so the default stays 1
    int overlap = 1;                     // BZ_NO_OVERLAP=1: every exchange serialised on the main stream (A/B and debugging)
    int scalars_in_flight = 0;           // the θ / q ghosts of set[cur] are being pulled on s2
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_main = nullptr, ev_in = nullptr;
    std::vector<cudaEvent_t> ev_out;     // one per (field, chunk) of the last bz_get_state_async
    const double* out_ptr[NPROG] = {};   // host destinations of the last bz_get_state_async (an upload from the same buffer waits chunk-wise)
    Comm comm;
    int use_tma = 0, z_chunks = 1;
    double time = 0.0;                   // BZ_KEEP_F64 (the clock stays double in the Float32 build)
    int64_t iteration = 0;
    int64_t launches = 0;
    int64_t bytes = 0;
    // profiling
    int prof_on = 0;
    std::vector<cudaEvent_t> prof_ev;    // pairs
    std::vector<cudaEvent_t> prof_pool;  // recycled events: no cudaEventCreate inside the step once the pool is warm
    std::vector<int> prof_fam;
    double prof_ms[NFAM] = {};
    int64_t prof_n[NFAM] = {};
    char err[512] = {};
};

static char g_err[512];

void bz_set_error(bz_ctx* ctx, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx ? ctx->err : g_err, 512, fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
static int dev_alloc(bz_ctx* c, T** p, size_t count) {
    CUDA_TRY(c, cudaMalloc((void**)p, count * sizeof(T)));
    CUDA_TRY(c, cudaMemsetAsync(*p, 0, count * sizeof(T), c->stream));
    c->bytes += (int64_t)(count * sizeof(T));
    return BZ_OK;
}

// Per-kernel-family timing. Events come from a pool that bz_profile_read refills, so a profiled step creates events only
// the first time; at most PROF_MAX_SCOPES scopes are kept between two reads (later ones are not recorded).
#define PROF_MAX_SCOPES 16384
static cudaEvent_t prof_event(bz_ctx* c) {
    cudaEvent_t e = nullptr;
    if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}
struct ProfScope {
    bz_ctx* c; int fam; cudaEvent_t a = nullptr, b = nullptr; bool on;
    ProfScope(bz_ctx* c_, int fam_) : c(c_), fam(fam_), on(c_->prof_on && c_->prof_fam.size() < PROF_MAX_SCOPES) {
        if (!on) return;
        a = prof_event(c); b = prof_event(c);
        cudaEventRecord(a, c->stream);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(b, c->stream);
        c->prof_ev.push_back(a); c->prof_ev.push_back(b); c->prof_fam.push_back(fam);
    }
};

static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
// line lengths of the in-house FFT: 2^m in [8, 2048] or 3 * 2^m in [24, 1536]
static int fft_length_ok(int n) {
    if (is_pow2(n)) return n >= 8 && n <= 2048;
    for (int r = 3; r <= 7; r += 2) if (n % r == 0 && is_pow2(n / r)) return n / r >= 8 && n <= 2048;     // 3 · 2^m .. 1536, 5 · 2^m .. 1280, 7 · 2^m .. 1792
    return 0;
}
static int fft_threads(int N, int lines) { return lines * N / fft_pt(N); }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_tensor_maps(bz_ctx* c, int box_w, int box_h) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_TRY(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { bz_set_error(c, "cuTensorMapEncodeTiled unavailable"); return BZ_ERR_CUDA; }
    PFN_encodeTiled encode = (PFN_encodeTiled)fn;
    const Layout& L = c->L;
    cuuint64_t dims[3] = {(cuuint64_t)L.PX, (cuuint64_t)L.PY, (cuuint64_t)L.Nz};
    cuuint64_t strides[2] = {(cuuint64_t)L.PX * sizeof(double), (cuuint64_t)L.plane * sizeof(double)};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    for (int s = 0; s < 3; ++s)
        for (int f = 0; f < NPROG; ++f) {
            CUresult r = encode(&c->tmap[s][f], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, c->set[s][f], dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { bz_set_error(c, "cuTensorMapEncodeTiled failed (%d)", (int)r); return BZ_ERR_CUDA; }
        }
    return BZ_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// reference state columns (src/Thermodynamics/reference_states.jl:88-123, 326-330)
// ---------------------------------------------------------------------------------------------------------------
static int upload_columns(bz_ctx* c) {
    const int Nz = c->L.Nz;
    const bz_config& g = c->cfg;
    std::vector<double> h((size_t)8 * (Nz + 1) + (size_t)LEV_REC * Nz, 0.0);
    double* rho = &h[0]; double* rho_inv = rho + (Nz + 1); double* rho_f = rho_inv + (Nz + 1); double* rho_f_inv = rho_f + (Nz + 1);
    double* p = rho_f_inv + (Nz + 1); double* T = p + (Nz + 1); double* ex = T + (Nz + 1); double* lg = ex + (Nz + 1);
    for (int k = 0; k < Nz; ++k) {
        rho[k] = c->h_rho[k]; rho_inv[k] = 1.0 / rho[k];
        p[k] = c->h_p[k]; T[k] = c->h_T[k];
        ex[k] = pow(p[k] / g.standard_pressure, c->th.Rd / c->th.cpd);
        lg[k] = log(p[k] / g.standard_pressure);
    }
    for (int k = 1; k < Nz; ++k) rho_f[k] = 0.5 * (rho[k] + rho[k - 1]);
    rho_f[0] = rho[0]; rho_f[Nz] = rho[Nz - 1];          // wall faces: only ever multiply w = 0
    for (int k = 0; k <= Nz; ++k) rho_f_inv[k] = 1.0 / rho_f[k];
    double* lev = lg + (Nz + 1);                           // per-level records of the stage kernel (common.cuh: LEV_REC)
    for (int k = 0; k < Nz; ++k) {
        double* r = lev + (size_t)LEV_REC * k;
        for (int m = 0; m < 4; ++m) {
            int kc = k - 2 + m, kf = k - 1 + m;
            r[m] = (kc >= 0 && kc < Nz) ? rho[kc] : 0.0;
            r[4 + m] = (kf >= 0 && kf <= Nz) ? rho_f[kf] : 0.0;
        }
        r[8] = ex[k]; r[9] = T[k];
        r[10] = (k + 3 < Nz) ? rho_inv[k + 3] : 0.0; r[11] = (k + 3 < Nz) ? rho_f_inv[k + 3] : 0.0;
        int kn = k + 1 < Nz ? k + 1 : Nz - 1;
        r[12] = ex[kn]; r[13] = T[kn];
        r[14] = (k + 4 < Nz) ? rho_inv[k + 4] : 0.0; r[15] = (k + 4 < Nz) ? rho_f_inv[k + 4] : 0.0;
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->col_store, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    double* d = c->col_store;
    c->col.rho = d; c->col.rho_inv = d + (Nz + 1); c->col.rho_f = d + 2 * (Nz + 1); c->col.rho_f_inv = d + 3 * (Nz + 1);
    c->col.p = d + 4 * (Nz + 1); c->col.T = d + 5 * (Nz + 1); c->col.exner_dry = d + 6 * (Nz + 1); c->col.log_p_pst = d + 7 * (Nz + 1);
    c->col.lev = d + 8 * (Nz + 1);
    return BZ_OK;
}

static void default_reference_state(bz_ctx* c) {
    const bz_config& g = c->cfg;
    const int Nz = g.Nz;
    const double Rd = c->th.Rd, cpd = c->th.cpd, grav = c->th.g;
    const double p0 = g.surface_pressure, th0 = g.potential_temperature, pst = g.standard_pressure;
    const double T0 = th0 * pow(p0 / pst, Rd / cpd);
    const double rho0 = p0 / (Rd * (pow(p0 / pst, Rd / cpd) * th0));
    c->h_rho.resize(Nz); c->h_p.resize(Nz); c->h_T.resize(Nz);
    const double dz = (g.z1 - g.z0) / Nz;
    for (int k = 0; k < Nz; ++k) {
        double z = g.z0 + (k + 0.5) * dz;
        double pr = p0 * pow(1 - grav * z / (cpd * T0), cpd / Rd);
        c->h_p[k] = pr;
        c->h_rho[k] = rho0 * pow(pr / p0, 1 - Rd / cpd);
        c->h_T[k] = th0 * pow(pr / pst, Rd / cpd);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Poisson solver setup / solve
// ---------------------------------------------------------------------------------------------------------------
static int setup_thomas(bz_ctx* c) {
    const PoissonGeom& G = c->PG;
    if (G.nky_loc == 0) return BZ_OK;                  // a rank that owns no ky modes (Flat y on rank > 0, or Ny/2 + 1 < n_ranks)
    dim3 grid((G.Nx + 127) / 128, G.nky_loc);
    thomas_setup<<<grid, 128, 0, c->stream>>>(G, c->col.rho, c->col.rho_f, c->L.dz, c->lam_x, c->lam_y, c->inv_beta, c->tfac);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return BZ_OK;
}

static int setup_poisson(bz_ctx* c) {
    const Layout& L = c->L;
    const bz_config& g = c->cfg;
    PoissonGeom& G = c->PG;
    // twiddles and eigenvalues (Oceananigans poisson_eigenvalues: λ = (2 sin(π i / N) / Δ)², Flat: 0)
    std::vector<double2> twx(g.Nx), twy(g.Ny);
    std::vector<double> lx(g.Nx), ly(G.nky);
    for (int i = 0; i < g.Nx; ++i) {
        long double a = -2.0L * M_PIl * i / g.Nx;
        twx[i] = make_double2((double)cosl(a), (double)sinl(a));
        double s = 2 * sin(M_PI * i / g.Nx) / L.dx;
        lx[i] = L.flat_x ? 0.0 : s * s;
    }
    for (int j = 0; j < g.Ny; ++j) {
        long double a = -2.0L * M_PIl * j / g.Ny;
        twy[j] = make_double2((double)cosl(a), (double)sinl(a));
    }
    for (int j = 0; j < G.nky; ++j) {
        double s = 2 * sin(M_PI * j / g.Ny) / L.dy;
        ly[j] = L.flat_y ? 0.0 : s * s;
    }
    int rc;
    if ((rc = dev_alloc(c, &c->tw_x, g.Nx))) return rc;
    if ((rc = dev_alloc(c, &c->tw_y, g.Ny))) return rc;
    if ((rc = dev_alloc(c, &c->lam_x, g.Nx))) return rc;
    if ((rc = dev_alloc(c, &c->lam_y, G.nky))) return rc;
    if ((rc = dev_alloc(c, &c->ky_base, G.nky))) return rc;
    if ((rc = dev_alloc(c, &c->ky_kstride, G.nky))) return rc;
    if ((rc = dev_alloc(c, &c->ky_owner, G.nky))) return rc;
    if ((rc = dev_alloc(c, &c->ky_base2, G.nky))) return rc;
    {
        std::vector<long long> kb(G.nky), kb2(G.nky); std::vector<int> ks(G.nky), ko(G.nky);
        for (int p = 0; p < G.P; ++p) {
            int st, cnt; ky_block(G.nky, G.P, p, &st, &cnt);
            for (int ky = st; ky < st + cnt; ++ky) {
                kb[ky] = (long long)st * L.nx * G.Nz + (long long)(ky - st) * L.nx; ks[ky] = cnt * L.nx;
                ko[ky] = p; kb2[ky] = ((long long)c->comm.rank * G.Nz * cnt + (ky - st)) * L.nx;
            }
        }
        CUDA_TRY(c, cudaMemcpyAsync(c->ky_owner, ko.data(), sizeof(int) * G.nky, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(c->ky_base2, kb2.data(), sizeof(long long) * G.nky, cudaMemcpyHostToDevice, c->stream));
        G.ky_owner = c->ky_owner; G.ky_base2 = c->ky_base2;
        G.off_W = (long long)c->off_W; G.off_W2 = (long long)c->off_W2; G.rank = c->comm.rank;
        CUDA_TRY(c, cudaMemcpyAsync(c->ky_base, kb.data(), sizeof(long long) * G.nky, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(c->ky_kstride, ks.data(), sizeof(int) * G.nky, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        G.ky_base = c->ky_base; G.ky_kstride = c->ky_kstride;
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->tw_x, twx.data(), sizeof(double2) * g.Nx, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->tw_y, twy.data(), sizeof(double2) * g.Ny, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->lam_x, lx.data(), sizeof(double) * g.Nx, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->lam_y, ly.data(), sizeof(double) * G.nky, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    size_t nW2 = (size_t)G.Nx * G.nky_loc * G.Nz;             // transposed layout (W, W2 live in the arena)
    if ((rc = dev_alloc(c, &c->inv_beta, nW2 > 0 ? nW2 : 1))) return rc;
    if ((rc = dev_alloc(c, &c->tfac, nW2 > 0 ? nW2 : 1))) return rc;
    // launch shapes: each thread owns 8 points of a line
    // launch shapes: each thread owns fft_pt(N) points of a line. The y transforms run as the 64-register build (256, 4) — 4 CTAs per SM
    // hide the source term's load latency (round-2 sweep at 512^3, profiles/r2a_fft_sweep.txt: forward 6.84 -> 6.01 ms, inverse 4.41 -> 4.24 ms
    // per step; 8-line / 512-thread tiles and a 64-register x transform measured no gain and were dropped).
    if (!L.flat_y) {
        int lines = 256 * fft_pt(g.Ny) / g.Ny; if (lines < 1) lines = 1;      // 256 threads per CTA
        int half = (L.nx + 1) / 2; if (lines > half) lines = half;
        while (lines & (lines - 1)) lines &= lines - 1;       // power of two (the kernels shift instead of dividing)
        c->lines_y = lines;
        size_t sm = fft_smem_bytes(g.Ny, lines);
        FFT_DISPATCH(g.Ny, {
            CUDA_TRY(c, cudaFuncSetAttribute(poisson_forward_y<FN, 256, BZ_FFT_Y_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            CUDA_TRY(c, cudaFuncSetAttribute(poisson_inverse_y<FN, 256, BZ_FFT_Y_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        })
    }
    if (!L.flat_x) {
        int lines = 128 * fft_pt(g.Nx) / g.Nx; if (lines < 1) lines = 1;      // 128 threads per CTA
        long long nl = (long long)G.Nz * G.nky_loc; if (nl < 1) nl = 1;
        if (lines > nl) lines = (int)nl;
        while (lines & (lines - 1)) lines &= lines - 1;
        c->lines_x = lines;
        size_t sm = fft_smem_bytes(g.Nx, lines);
        FFT_DISPATCH(g.Nx, { CUDA_TRY(c, cudaFuncSetAttribute(fft_x_kernel<FN, 256, BZ_FFT_X_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); })
    }
    return setup_thomas(c);
}

// compute_pressure_correction! (anelastic_time_stepping.jl:26-39): momentum ghosts must be valid on entry.
// Slabs with peer memory: the two transposes of the distributed transform are peer LOADS inside the consuming transforms — fft_x
// (forward) and inverse_y (backward) read the other ranks' spectra straight from their arenas, so a transpose costs no memory pass of
// its own. Two alternatives are built in and measured slower (bz_ctx::dma_transpose, ::fft_z_chunks): pipelining those pulls over z
// chunks across the two streams, and BZ_DMA_TRANSPOSE=1 — per chunk, the second stream waits for every rank (flag barrier), moves the
// chunk with the copy engines (transpose_chunk_dma) and runs the consuming transform while the main stream transforms the next chunk.
#define FFT_Z_CHUNKS_MAX 8

// One z chunk [k0, k1) of a transpose of the distributed transform as copy-engine transfers between the peer-blocked layouts (poisson.cuh):
//   forward : my W2 block p  <-  rank p's W  block `me` (my ky modes of p's columns)
//   backward: my W  block p  <-  rank p's W2 block `me` (p's ky modes of my columns)
// Every block's chunk is one contiguous run, so a transpose is n_ranks cudaMemcpyAsync calls; the DMA engines move them over NVLink while
// the SMs transform the next chunk.
static int transpose_chunk_dma(bz_ctx* c, bool forward, int k0, int k1, cudaStream_t s) {
    const PoissonGeom& G = c->PG;
    const int P = c->comm.n_ranks, me = c->comm.rank, nx = c->L.nx;
    int st_me, cnt_me; ky_block(G.nky, P, me, &st_me, &cnt_me);
    for (int q = 0; q < P; ++q) {
        const int p = (me + q) % P;                                   // start with the local block, then spread the peers round-robin
        int st_p, cnt_p; ky_block(G.nky, P, p, &st_p, &cnt_p);
        const double2* src; double2* dst; size_t count;
        if (forward) {
            count = (size_t)(k1 - k0) * cnt_me * nx;
            src = reinterpret_cast<const double2*>(c->comm.peer_base[p] + c->off_W) + (size_t)st_me * nx * G.Nz + (size_t)k0 * cnt_me * nx;
            dst = c->W2 + (size_t)p * G.Nz * cnt_me * nx + (size_t)k0 * cnt_me * nx;
        } else {
            count = (size_t)(k1 - k0) * cnt_p * nx;
            src = reinterpret_cast<const double2*>(c->comm.peer_base[p] + c->off_W2) + (size_t)me * G.Nz * cnt_p * nx + (size_t)k0 * cnt_p * nx;
            dst = c->W + (size_t)st_p * nx * G.Nz + (size_t)k0 * cnt_p * nx;
        }
        if (count) CUDA_TRY(c, cudaMemcpyAsync(dst, src, count * sizeof(double2), cudaMemcpyDeviceToDevice, s));
    }
    return BZ_OK;
}

static int poisson_solve(bz_ctx* c, double dt) {
    const Layout& L = c->L;
    const PoissonGeom& G = c->PG;
    double** U = c->set[c->cur];
    const double dz_over_dt = L.dz / dt;
    const bool pull = c->comm.p2p && !L.flat_y;        // peer-memory path: the transposes are peer loads inside fft_x / inverse_y
    PeerBases peers;
    for (int p = 0; p < 8; ++p) peers.base[p] = c->comm.peer_base[p];
    const long long n_lines = (long long)G.Nz * G.nky_loc;
    const bool do_x = !L.flat_x;
    int nch = 1;
    if (pull && c->overlap && do_x) { nch = c->fft_z_chunks; if (nch > L.Nz / 8) nch = L.Nz / 8; if (nch < 1) nch = 1; }
    const bool chunked = nch > 1 || (pull && c->overlap && do_x && c->dma_transpose);
    const int kper = (L.Nz + nch - 1) / nch;
    auto launch_fft_x = [&](int inverse, int do_pull, long long line0, long long line1, cudaStream_t s) {
        if (!do_x || line1 <= line0) return;
        int lines = c->lines_x;
        size_t sm = fft_smem_bytes(G.Nx, lines);
        FFT_DISPATCH(G.Nx, (fft_x_kernel<FN, 256, BZ_FFT_X_MINB><<<(unsigned)((line1 - line0 + lines - 1) / lines), fft_threads(G.Nx, lines), sm, s>>>(G, c->W2, line1, c->tw_x, lines, inverse, peers, do_pull, (int)line0)));
        c->launches++;
    };
    // ---- forward: source term + y transform (x-slab layout), transpose, x transform
    if (!chunked) {
        {
            ProfScope ps(c, 1);
            if (!L.flat_y) {
                int lines = c->lines_y;
                dim3 grid((L.nx + 2 * lines - 1) / (2 * lines), L.Nz);
                size_t sm = fft_smem_bytes(G.Ny, lines);
                FFT_DISPATCH(G.Ny, (poisson_forward_y<FN, 256, BZ_FFT_Y_MINB><<<grid, fft_threads(G.Ny, lines), sm, c->stream>>>(L, G, U[0], U[1], U[2], dz_over_dt, c->W, c->tw_y, lines, 0)));
            } else {
                dim3 grid((L.nx + 127) / 128, L.Nz);
                poisson_pack_flat_y<<<grid, 128, 0, c->stream>>>(L, G, U[0], U[1], U[2], dz_over_dt, c->W);
            }
            c->launches++;
        }
        if (c->comm.n_ranks > 1) {
            ProfScope ps(c, 5);
            int rc = pull ? comm_barrier(c->comm, c->stream) : comm_transpose_forward(c->comm, c->W, c->W2, L.nx, G, c->stream, &c->launches);
            if (rc) { bz_set_error(c, "transpose: %s", c->comm.err); return rc; }
        }
        if (n_lines > 0) { ProfScope ps(c, 1); launch_fft_x(0, pull ? 1 : 0, 0, n_lines, c->stream); }
    } else {
        ProfScope ps(c, 1);
        int lines = c->lines_y;
        size_t sm = fft_smem_bytes(G.Ny, lines);
        for (int ch = 0; ch < nch; ++ch) {
            const int k0 = ch * kper, k1 = (k0 + kper < L.Nz) ? k0 + kper : L.Nz;
            if (k1 <= k0) break;
            dim3 grid((L.nx + 2 * lines - 1) / (2 * lines), k1 - k0);
            FFT_DISPATCH(G.Ny, (poisson_forward_y<FN, 256, BZ_FFT_Y_MINB><<<grid, fft_threads(G.Ny, lines), sm, c->stream>>>(L, G, U[0], U[1], U[2], dz_over_dt, c->W, c->tw_y, lines, k0)));
            c->launches++;
            CUDA_TRY(c, cudaEventRecord(c->ev_chunk[ch], c->stream));
            CUDA_TRY(c, cudaStreamWaitEvent(c->s2, c->ev_chunk[ch], 0));
            int rc = comm_barrier(c->comm, c->s2, 1);                   // every rank has transformed chunk ch
            if (rc) { bz_set_error(c, "transpose: %s", c->comm.err); return rc; }
            if (c->dma_transpose && (rc = transpose_chunk_dma(c, true, k0, k1, c->s2))) return rc;
            launch_fft_x(0, c->dma_transpose ? 0 : 1, (long long)k0 * G.nky_loc, (long long)k1 * G.nky_loc, c->s2);
        }
        CUDA_TRY(c, cudaEventRecord(c->ev_join, c->s2));
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    }
    // ---- z: batched Thomas in the transposed layout
    if (G.nky_loc > 0) {
        ProfScope ps(c, 2);
        // one thread per (kx, ky) column: on a rank that keeps few ky modes (8 slabs of 512^3: 4 x 33 blocks of 128) the blocks shrink
        // until the grid covers the machine a few times over
        int tb = 128;
        while (tb > 32 && (long long)((G.Nx + tb - 1) / tb) * G.nky_loc < 4 * 148) tb >>= 1;
        dim3 grid((G.Nx + tb - 1) / tb, G.nky_loc);
        thomas_z<<<grid, tb, 0, c->stream>>>(G, c->W2, c->col.rho_f, L.dz, c->inv_beta, c->tfac);
        c->launches++;
        if (G.ky0 == 0) { remove_mean_mode<<<1, 256, 0, c->stream>>>(G, c->W2); c->launches++; }
    }
    // ---- inverse: x transform, transpose back, y transform → φ
    const double scale = 1.0 / ((L.flat_x ? 1.0 : (double)G.Nx) * (L.flat_y ? 1.0 : (double)G.Ny));
    if (!chunked) {
        if (n_lines > 0) { ProfScope ps(c, 3); launch_fft_x(1, 0, 0, n_lines, c->stream); }
        if (c->comm.n_ranks > 1) {
            ProfScope ps(c, 5);
            int rc = pull ? comm_barrier(c->comm, c->stream) : comm_transpose_backward(c->comm, c->W2, c->W, L.nx, G, c->stream, &c->launches);
            if (rc) { bz_set_error(c, "transpose: %s", c->comm.err); return rc; }
        }
        ProfScope ps(c, 3);
        if (!L.flat_y) {
            int lines = c->lines_y;
            dim3 grid((L.nx + 2 * lines - 1) / (2 * lines), L.Nz);
            size_t sm = fft_smem_bytes(G.Ny, lines);
            FFT_DISPATCH(G.Ny, (poisson_inverse_y<FN, 256, BZ_FFT_Y_MINB><<<grid, fft_threads(G.Ny, lines), sm, c->stream>>>(L, G, c->W, c->phi, c->tw_y, lines, scale, peers, pull ? 1 : 0, 0)));
        } else {
            dim3 grid((L.nx + 127) / 128, L.Nz);
            poisson_unpack_flat_y<<<grid, 128, 0, c->stream>>>(L, G, c->W, c->phi, scale);
        }
        c->launches++;
    } else {
        ProfScope ps(c, 3);
        int lines = c->lines_y;
        size_t sm = fft_smem_bytes(G.Ny, lines);
        for (int ch = 0; ch < nch; ++ch) {
            const int k0 = ch * kper, k1 = (k0 + kper < L.Nz) ? k0 + kper : L.Nz;
            if (k1 <= k0) break;
            launch_fft_x(1, 0, (long long)k0 * G.nky_loc, (long long)k1 * G.nky_loc, c->stream);
            CUDA_TRY(c, cudaEventRecord(c->ev_chunk[ch], c->stream));
            CUDA_TRY(c, cudaStreamWaitEvent(c->s2, c->ev_chunk[ch], 0));
            int rc = comm_barrier(c->comm, c->s2, 1);                   // every rank has inverted its ky modes of chunk ch
            if (rc) { bz_set_error(c, "transpose: %s", c->comm.err); return rc; }
            if (c->dma_transpose && (rc = transpose_chunk_dma(c, false, k0, k1, c->s2))) return rc;
            dim3 grid((L.nx + 2 * lines - 1) / (2 * lines), k1 - k0);
            FFT_DISPATCH(G.Ny, (poisson_inverse_y<FN, 256, BZ_FFT_Y_MINB><<<grid, fft_threads(G.Ny, lines), sm, c->s2>>>(L, G, c->W, c->phi, c->tw_y, lines, scale, peers, c->dma_transpose ? 0 : 1, k0)));
            c->launches++;
        }
        CUDA_TRY(c, cudaEventRecord(c->ev_join, c->s2));
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    }
    CUDA_TRY(c, cudaGetLastError());
    return BZ_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// halo fills
// ---------------------------------------------------------------------------------------------------------------
static int fill_halos(bz_ctx* c, double* const* fields, int nf, int fam, bool exchange_x = true) {
    const Layout& L = c->L;
    if (L.HX == 0 && L.HY == 0) return BZ_OK;
    ProfScope ps(c, fam);
    FieldSet F; F.n = nf;
    for (int f = 0; f < nf; ++f) F.f[f] = fields[f];
    int mode = 3;
    if (c->comm.n_ranks > 1) {
        if (exchange_x) {
            int rc = c->comm.p2p ? comm_pull_x_halos(c->comm, L, F, 0, c->stream, &c->launches, 0, nf == 1 ? 1 : 3)
                                 : comm_exchange_x_halos(c->comm, L, F, c->stream, &c->launches);
            if (rc) { bz_set_error(c, "halo exchange: %s", c->comm.err); return rc; }
        }
        mode = 2;
    }
    if (mode == 2 && L.HY == 0) return BZ_OK;
    long long per_level = (long long)((mode & 2) ? 2 * L.HY * L.PX : 0) + (long long)((mode & 1) ? 2 * L.HX * L.Ny : 0);
    long long total = per_level * L.Nz;
    if (total == 0) return BZ_OK;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    halo_fill_periodic<<<blocks, 256, 0, c->stream>>>(L, F, mode);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return BZ_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// forcing: horizontal means → per-level forcing columns (compute_forcings!, update_atmosphere_model_state.jl:81-86)
// ---------------------------------------------------------------------------------------------------------------
static int update_column_forcing(bz_ctx* c, int set_index) {
    const Layout& L = c->L;
    ProfScope ps(c, 4);
    if (c->d_ws && c->subs_mask) {
        FieldSet U; U.n = NPROG;
        for (int f = 0; f < NPROG; ++f) U.f[f] = c->set[set_index][f];
        column_sums<<<L.Nz, 256, 0, c->stream>>>(L, U, c->d_sums);
        c->launches++;
        if (c->comm.n_ranks > 1) {
            int rc = comm_allreduce_sum_device(c->comm, c->d_sums, (size_t)4 * L.Nz, c->stream);
            if (rc) { bz_set_error(c, "allreduce: %s", c->comm.err); return rc; }
        }
    }
    const double inv_n = 1.0 / ((double)c->cfg.Nx * (double)c->cfg.Ny);
    column_forcing_kernel<<<1, 256, 0, c->stream>>>(L.Nz, L.dz, inv_n, c->col.rho, c->d_sums, c->d_ws, c->subs_mask, c->d_ug, c->d_vg,
                                                     c->d_qt, c->coriolis_f, c->d_fcol);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return BZ_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// stage kernel launch
// ---------------------------------------------------------------------------------------------------------------
template <int TX, int TY, bool HAS_Y, bool FLAT_X, int MICRO, bool FORCED>
static int launch_stage_t(bz_ctx* c, const StageParams& P, int nz_chunks) {
    using SM = StageShared<TX, TY, HAS_Y>;
    static bool configured[64] = {false};            // function attributes are per device: one flag per device ordinal
    auto kern = stage_kernel<TX, TY, HAS_Y, FLAT_X, MICRO, FORCED>;
    const int dev = c->cfg.device;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SM)));
        CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    dim3 grid((P.nx_u + TX - 1) / TX, (c->L.Ny + TY - 1) / TY, nz_chunks);
    kern<<<grid, 2 * TX * TY, sizeof(SM), c->stream>>>(P);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return BZ_OK;
}

template <int BUF, int TY, int MICRO, bool FORCED>
static int launch_stage_hi_t(bz_ctx* c, const StageParams& P, const SpecificFields& F, int chunks) {
    const Layout& L = c->L;
    const dim3 grid((L.nx + HI_TX - 1) / HI_TX, (L.Ny + TY - 1) / TY, chunks), block(32, TY);
    stage_hi_kernel<BUF, TY, MICRO, FORCED><<<grid, block, 0, c->stream>>>(P, F);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return BZ_OK;
}

// WENO(order = 7 / 9): specific fields of the input state, then the fused high-order stage kernel (stage_hi.cuh)
static int launch_stage_hi(bz_ctx* c, StageParams& P, int in) {
    const Layout& L = c->L;
    SpecificFields F;
    for (int f = 0; f < NPROG; ++f) { F.U[f] = c->set[in][f]; F.V[f] = c->V[f]; }
    {
        int bx = (int)((L.plane + 255) / 256); if (bx > 64) bx = 64;
        specific_fields_kernel<<<dim3(bx, L.Nz), 256, 0, c->stream>>>(L, c->col, F);
        c->launches++;
        CUDA_TRY(c, cudaGetLastError());
    }
    // z chunks: enough CTAs for ≈ 3 waves of the machine, at least 8 levels per chunk (a chunk start re-evaluates five z fluxes and one buoyancy)
    const int ty = L.flat_y ? 1 : 8;
    const long long tiles = (long long)((L.nx + HI_TX - 1) / HI_TX) * ((L.Ny + ty - 1) / ty);
    int chunks = c->cfg.z_chunks > 0 ? c->cfg.z_chunks : (int)((3 * 148 * (L.flat_y ? 8 : 1) + tiles - 1) / tiles);
    const int max_chunks = (L.Nz + 7) / 8;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    P.k_chunk = (L.Nz + chunks - 1) / chunks;
    chunks = (L.Nz + P.k_chunk - 1) / P.k_chunk;
    const bool moist = c->cfg.microphysics != BZ_MICROPHYSICS_NONE;
    const bool se = c->cfg.formulation == BZ_FORMULATION_STATIC_ENERGY;
#define LAUNCH_HI(BUF, TY)                                                                                                         \
    (se ? launch_stage_hi_t<BUF, TY, BZ_THERMO_STATIC_ENERGY, false>(c, P, F, chunks)                                              \
        : c->forced ? (moist ? launch_stage_hi_t<BUF, TY, 1, true>(c, P, F, chunks) : launch_stage_hi_t<BUF, TY, 0, true>(c, P, F, chunks)) \
                    : (moist ? launch_stage_hi_t<BUF, TY, 1, false>(c, P, F, chunks) : launch_stage_hi_t<BUF, TY, 0, false>(c, P, F, chunks)))
    if (c->buf == 5) return L.flat_y ? LAUNCH_HI(5, 1) : LAUNCH_HI(5, 8);
    return L.flat_y ? LAUNCH_HI(4, 1) : LAUNCH_HI(4, 8);
#undef LAUNCH_HI
}

// mode 0: out = RK update of set[in]; mode 1: out = tendencies
static int launch_stage(bz_ctx* c, int in, double* const* out, int u0, double dt, double alpha, int mode) {
    ProfScope ps(c, 0);
    StageParams P;
    memset(&P, 0, sizeof(P));
    for (int f = 0; f < NPROG; ++f) {
        P.tmap[f] = c->tmap[in][f];
        P.U[f] = c->set[in][f]; P.U0[f] = c->set[u0][f]; P.out[f] = out[f];
    }
    P.L = c->L; P.col = c->col; P.th = c->th;
    P.dt = dt; P.alpha = alpha; P.mode = mode;
    P.nx_u = c->L.nx;
    P.use_tma = c->use_tma;
    int chunks = c->z_chunks;
    P.k_chunk = (c->L.Nz + chunks - 1) / chunks;
    const bool moist = c->cfg.microphysics != BZ_MICROPHYSICS_NONE;
    if (c->forced) {
        int rc = update_column_forcing(c, in);
        if (rc) return rc;
        for (int f = 0; f < 4; ++f) P.fcol[f] = c->d_fcol + (size_t)f * c->L.Nz;
        P.e_tend = c->d_et;
        P.coriolis_f = c->coriolis_f;
        P.theta_flux_dz = c->theta_flux / c->L.dz; P.q_flux_dz = c->q_flux / c->L.dz; P.drag_dz = c->drag / c->L.dz;
    }
    if (c->buf > 3) return launch_stage_hi(c, P, in);
#define LAUNCH(TX, TY, HY, FX)                                                                                        \
    (c->forced ? (moist ? launch_stage_t<TX, TY, HY, FX, 1, true>(c, P, chunks) : launch_stage_t<TX, TY, HY, FX, 0, true>(c, P, chunks)) \
               : (moist ? launch_stage_t<TX, TY, HY, FX, 1, false>(c, P, chunks) : launch_stage_t<TX, TY, HY, FX, 0, false>(c, P, chunks)))
    if (c->cfg.formulation == BZ_FORMULATION_STATIC_ENERGY) {       // StaticEnergyFormulation: no microphysics, no forcings (bz_create / bz_set_forcing)
        if (!c->L.flat_y) return launch_stage_t<32, 8, true, false, BZ_THERMO_STATIC_ENERGY, false>(c, P, chunks);
        if (c->L.flat_x) return launch_stage_t<32, 1, false, true, BZ_THERMO_STATIC_ENERGY, false>(c, P, chunks);
        return launch_stage_t<128, 1, false, false, BZ_THERMO_STATIC_ENERGY, false>(c, P, chunks);
    }
    if (!c->L.flat_y) return LAUNCH(32, 8, true, false);
    if (c->L.flat_x) return LAUNCH(32, 1, false, true);
    return LAUNCH(128, 1, false, false);
#undef LAUNCH
}

static int y_halo_fill(bz_ctx* c, double* const* fields, int nf, cudaStream_t s) {
    const Layout& L = c->L;
    if (L.HY == 0) return BZ_OK;
    FieldSet F; F.n = nf;
    for (int f = 0; f < nf; ++f) F.f[f] = fields[f];
    long long total = (long long)2 * L.HY * L.PX * L.Nz;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    halo_fill_periodic<<<blocks, 256, 0, s>>>(L, F, 2);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return BZ_OK;
}

// Slabs with peer memory: right after the stage kernel the x ghosts of ρθ and ρq of the new state can travel — the projection does not
// touch them — so their exchange runs on the second stream underneath the whole Poisson solve. Joined in pressure_correct.
static int start_scalar_exchange(bz_ctx* c) {
    if (!(c->comm.n_ranks > 1 && c->comm.p2p && c->overlap)) return BZ_OK;
    double** U = c->set[c->cur];
    CUDA_TRY(c, cudaEventRecord(c->ev_fork, c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent(c->s2, c->ev_fork, 0));
    FieldSet F; F.n = 2; F.f[0] = U[BZ_RHO_THETA]; F.f[1] = U[BZ_RHO_Q];
    int rc = comm_pull_x_halos(c->comm, c->L, F, 0, c->s2, &c->launches, 1, 2);
    if (rc) { bz_set_error(c, "halo exchange: %s", c->comm.err); return rc; }
    c->scalars_in_flight = 1;
    return BZ_OK;
}

// compute_pressure_correction! + make_pressure_correction! on set[cur], then refresh all ghosts
static int pressure_correct(bz_ctx* c, double dt) {
    int rc;
    double** U = c->set[c->cur];
    const Layout& L = c->L;
    // The projection addresses the periodic images of φ directly (project_momentum): φ needs no ghost fill on one GPU, and across slabs
    // only its first ghost column on the left travels. The divergence reads ghosts of ρu, ρv (see source_term for why).
    if (c->comm.n_ranks == 1) { if ((rc = fill_halos(c, U, 2, 4))) return rc; }       // ρu, ρv ghosts for the divergence
    else {                                                                               // slabs: y ghosts of ρv are local; ρu at the first ghost face
        double* uv[1] = {U[1]};                                                          // comes from the right neighbour's first column
        if ((rc = fill_halos(c, uv, 1, 4, false))) return rc;
        ProfScope ps(c, 5);
        if (c->comm.p2p) { FieldSet Fu; Fu.n = 1; Fu.f[0] = U[0]; rc = comm_pull_x_halos(c->comm, c->L, Fu, 1, c->stream, &c->launches, 0, 0); }
        else rc = comm_exchange_u_face(c->comm, c->L, U[0], c->stream, &c->launches);
        if (rc) { bz_set_error(c, "face exchange: %s", c->comm.err); return rc; }
    }
    if ((rc = poisson_solve(c, dt))) return rc;
    if (c->comm.n_ranks > 1) {
        ProfScope ps(c, 5);
        FieldSet F; F.n = 1; F.f[0] = c->phi;
        rc = c->comm.p2p ? comm_pull_x_halos(c->comm, L, F, 2, c->stream, &c->launches, 0, 1)      // φ at i = -1 only
                         : comm_exchange_x_halos(c->comm, L, F, c->stream, &c->launches);
        if (rc) { bz_set_error(c, "halo exchange: %s", c->comm.err); return rc; }
    }
    const int wrap_x = c->comm.n_ranks == 1;
    const bool split = c->comm.n_ranks > 1 && c->comm.p2p && c->overlap && !L.flat_x && L.nx >= 4 * L.HX;
    if (!split) {
        {
            ProfScope ps(c, 4);
            dim3 grid((L.nx + 128 * PROJ_ILP - 1) / (128 * PROJ_ILP), L.Ny, L.Nz);
            project_momentum<<<grid, 128, 0, c->stream>>>(L, c->col, U[0], U[1], U[2], c->phi, dt, wrap_x);
            c->launches++;
            CUDA_TRY(c, cudaGetLastError());
        }
        if (c->scalars_in_flight) {                            // the θ / q exchange of start_scalar_exchange joins here
            CUDA_TRY(c, cudaEventRecord(c->ev_join, c->s2));
            CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
            c->scalars_in_flight = 0;
            if ((rc = fill_halos(c, U, 3, 4))) return rc;      // momentum: x exchange + y ghosts
            return y_halo_fill(c, U + 3, 2, c->stream);        // θ, q: y ghosts (full padded width: after their x ghosts have landed)
        }
        return fill_halos(c, U, NPROG, 4);
    }
    // Slabs, overlapped: project the HX edge columns of either side first; while the second stream waits for the neighbours to have done
    // the same and pulls their edge columns into our momentum ghosts, the main stream projects the interior columns.
    ProfScope ps(c, 4);
    {
        const int nc = 2 * L.HX;
        dim3 grid((nc * L.Ny + 127) / 128, L.Nz);
        project_momentum_columns<<<grid, 128, 0, c->stream>>>(L, c->col, U[0], U[1], U[2], c->phi, dt, 0, L.HX, L.nx - L.HX, L.HX);
        c->launches++;
        CUDA_TRY(c, cudaGetLastError());
    }
    CUDA_TRY(c, cudaEventRecord(c->ev_fork, c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent(c->s2, c->ev_fork, 0));
    {
        FieldSet F; F.n = 3; F.f[0] = U[0]; F.f[1] = U[1]; F.f[2] = U[2];
        rc = comm_pull_x_halos(c->comm, L, F, 0, c->s2, &c->launches, 1, 3);
        if (rc) { bz_set_error(c, "halo exchange: %s", c->comm.err); return rc; }
        if (!c->scalars_in_flight) {                            // bz_set_state / bz_pressure_correct: θ, q travel here as well
            FieldSet S; S.n = 2; S.f[0] = U[3]; S.f[1] = U[4];
            rc = comm_pull_x_halos(c->comm, L, S, 0, c->s2, &c->launches, 1, 2);
            if (rc) { bz_set_error(c, "halo exchange: %s", c->comm.err); return rc; }
        }
        c->scalars_in_flight = 0;
    }
    CUDA_TRY(c, cudaEventRecord(c->ev_join, c->s2));
    {
        const int ni = L.nx - 2 * L.HX;
        dim3 grid((ni * L.Ny + 127) / 128, L.Nz);
        project_momentum_columns<<<grid, 128, 0, c->stream>>>(L, c->col, U[0], U[1], U[2], c->phi, dt, L.HX, ni, 0, 0);
        c->launches++;
        CUDA_TRY(c, cudaGetLastError());
    }
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    return y_halo_fill(c, U, NPROG, c->stream);                 // y ghosts over the full padded width, corners included
}

// ---------------------------------------------------------------------------------------------------------------
// ABI
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

void bz_default_config(bz_config* c) {
    memset(c, 0, sizeof(*c));
    c->abi_version = BZ_ABI_VERSION;
    c->Nx = c->Ny = c->Nz = 8;
    c->topology_x = c->topology_y = BZ_PERIODIC;
    c->x1 = c->y1 = c->z1 = 1.0;
    c->surface_pressure = 101325.0; c->potential_temperature = 288.0; c->standard_pressure = 1e5;
    c->molar_gas_constant = 8.314462618; c->gravitational_acceleration = 9.81;
    c->energy_reference_temperature = 273.15; c->triple_point_temperature = 273.16; c->triple_point_pressure = 611.657;
    c->dry_air_molar_mass = 0.02897; c->dry_air_heat_capacity = 1005.0;
    c->vapor_molar_mass = 0.018015; c->vapor_heat_capacity = 1850.0;
    c->liquid_reference_latent_heat = 2500800.0; c->liquid_heat_capacity = 4181.0;
    c->ice_reference_latent_heat = 2834000.0; c->ice_heat_capacity = 2108.0;
    c->advection_order = 5; c->microphysics = BZ_MICROPHYSICS_NONE;
    c->n_ranks = 1; c->rank = 0; c->device = 0;
}

int bz_abi_version(void) { return BZ_ABI_VERSION; }

const char* bz_last_error(const bz_ctx* c) { return c ? c->err : g_err; }

void bz_destroy(bz_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->s2) cudaStreamSynchronize(c->s2);
    if (c->s_in) cudaStreamSynchronize(c->s_in);
    if (c->s_out) cudaStreamSynchronize(c->s_out);
    c->graphs.clear();
    comm_destroy(c->comm);
    cudaFree(c->arena);
    for (int f = 0; f < NPROG; ++f) { cudaFree(c->G[f]); cudaFree(c->V[f]); }
    cudaFree(c->dense); cudaFree(c->scalar); cudaFree(c->slice_buf); cudaFree(c->col_store);
    cudaFree(c->tw_x); cudaFree(c->tw_y); cudaFree(c->lam_x); cudaFree(c->lam_y);
    cudaFree(c->inv_beta); cudaFree(c->tfac); cudaFree(c->ky_base); cudaFree(c->ky_kstride); cudaFree(c->ky_owner); cudaFree(c->ky_base2); cudaFree(c->fstore);
    for (auto e : c->prof_ev) cudaEventDestroy(e);
    for (auto e : c->prof_pool) cudaEventDestroy(e);
    for (auto e : c->ev_out) cudaEventDestroy(e);
    if (c->ev_main) cudaEventDestroy(c->ev_main);
    if (c->ev_in) cudaEventDestroy(c->ev_in);
    if (c->s2) cudaStreamDestroy(c->s2);
    for (int e = 0; e < 8; ++e) if (c->ev_chunk[e]) cudaEventDestroy(c->ev_chunk[e]);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int bz_create(const bz_config* cfg, bz_ctx** out) {
    if (!cfg || !out) return BZ_ERR_INVALID;
    *out = nullptr;
#define FAIL(code, ...) do { bz_set_error(nullptr, __VA_ARGS__); return (code); } while (0)
    if (cfg->abi_version != BZ_ABI_VERSION) FAIL(BZ_ERR_INVALID, "abi_version %d != %d", cfg->abi_version, BZ_ABI_VERSION);
    if (cfg->Nx < 1 || cfg->Ny < 1 || cfg->Nz < 2) FAIL(BZ_ERR_INVALID, "grid size must be positive (Nz >= 2)");
    if (cfg->advection_order != 5 && cfg->advection_order != 7 && cfg->advection_order != 9)
        FAIL(BZ_ERR_UNSUPPORTED, "WENO(order = 5, 7 or 9) is on the path, got order %d", cfg->advection_order);
    if (cfg->formulation != BZ_FORMULATION_POTENTIAL_TEMPERATURE && cfg->formulation != BZ_FORMULATION_STATIC_ENERGY) FAIL(BZ_ERR_INVALID, "unknown formulation %d", cfg->formulation);
    if (cfg->formulation == BZ_FORMULATION_STATIC_ENERGY && cfg->microphysics != BZ_MICROPHYSICS_NONE)
        FAIL(BZ_ERR_UNSUPPORTED, "StaticEnergyFormulation is on the path without microphysics only");
    const int fx = cfg->topology_x == BZ_FLAT, fy = cfg->topology_y == BZ_FLAT;
    if ((fx && cfg->Nx != 1) || (fy && cfg->Ny != 1)) FAIL(BZ_ERR_INVALID, "a Flat dimension must have size 1");
    if (fx && !fy) FAIL(BZ_ERR_UNSUPPORTED, "(Flat, Periodic, Bounded) is not supported; use (Periodic, Flat, Bounded)");
    if (!fx && !fft_length_ok(cfg->Nx)) FAIL(BZ_ERR_UNSUPPORTED, "Nx must be 2^m, 3 * 2^m, 5 * 2^m or 7 * 2^m with m >= 3 and Nx <= 2048 (in-house FFT), got %d", cfg->Nx);
    if (!fy && !fft_length_ok(cfg->Ny)) FAIL(BZ_ERR_UNSUPPORTED, "Ny must be 2^m, 3 * 2^m, 5 * 2^m or 7 * 2^m with m >= 3 and Ny <= 2048 (in-house FFT), got %d", cfg->Ny);
    const int P = cfg->n_ranks < 1 ? 1 : cfg->n_ranks;
    if (cfg->use_tma == 1 && cfg->advection_order != 5) FAIL(BZ_ERR_UNSUPPORTED, "TMA staging belongs to the WENO(order=5) stage kernel");
    if (P > 1 && (fx || cfg->Nx % P != 0 || (cfg->Nx / P) < 8)) FAIL(BZ_ERR_INVALID, "x-slabs: Nx must be divisible by n_ranks with at least 8 columns per rank");
    if (cfg->rank < 0 || cfg->rank >= P) FAIL(BZ_ERR_INVALID, "rank out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) FAIL(BZ_ERR_CUDA, "no CUDA device: libbreeze_b200 has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) FAIL(BZ_ERR_INVALID, "device ordinal %d out of range (%d devices)", cfg->device, ndev);
    if (cudaSetDevice(cfg->device) != cudaSuccess) FAIL(BZ_ERR_CUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, cfg->device);
    if (prop.major < 10) FAIL(BZ_ERR_UNSUPPORTED, "compute capability %d.%d: this library is built for sm_100a (B200) only", prop.major, prop.minor);
#undef FAIL

    bz_ctx* c = new bz_ctx();
    c->cfg = *cfg;
    c->cfg.n_ranks = P;
    Layout& L = c->L;
    L.nx = cfg->Nx / P; L.Ny = cfg->Ny; L.Nz = cfg->Nz;
    L.flat_x = fx; L.flat_y = fy;
    c->buf = (cfg->advection_order + 1) / 2;
    const int halo = c->buf == 3 ? BZ_HALO : c->buf + 1;              // 4 for WENO5 (the TMA-staged kernel's tile origin), buffer + 1 otherwise
    L.HX = fx ? 0 : halo; L.HY = fy ? 0 : halo;
    L.PX = L.nx + 2 * L.HX; L.PY = L.Ny + 2 * L.HY;
    L.plane = (long long)L.PX * L.PY;
    L.n = L.plane * (L.Nz + (c->buf > 3 ? 1 : 0));                    // orders 7 / 9: one extra zero plane on top (the top wall of ρw, stage_hi.cuh)
    L.dx = fx ? 1.0 : (cfg->x1 - cfg->x0) / cfg->Nx;
    L.dy = fy ? 1.0 : (cfg->y1 - cfg->y0) / cfg->Ny;
    L.dz = (cfg->z1 - cfg->z0) / cfg->Nz;
    L.rdx = fx ? 0.0 : 1.0 / L.dx; L.rdy = fy ? 0.0 : 1.0 / L.dy; L.rdz = 1.0 / L.dz;
    Thermo& th = c->th;
    th.Rd = cfg->molar_gas_constant / cfg->dry_air_molar_mass; th.Rv = cfg->molar_gas_constant / cfg->vapor_molar_mass;
    th.cpd = cfg->dry_air_heat_capacity; th.cpv = cfg->vapor_heat_capacity;
    th.cl = cfg->liquid_heat_capacity; th.ci = cfg->ice_heat_capacity; th.g = cfg->gravitational_acceleration;
    th.Ll = cfg->liquid_reference_latent_heat; th.Li = cfg->ice_reference_latent_heat; th.pst = cfg->standard_pressure;
    th.Tr_energy = cfg->energy_reference_temperature; th.Ttr = cfg->triple_point_temperature; th.ptr = cfg->triple_point_pressure;
    th.microphysics = cfg->microphysics;
    th.z0 = cfg->z0; th.dz = L.dz;

    int rc = BZ_OK;
#define TRY(x) do { rc = (x); if (rc) { strncpy(g_err, c->err, 511); bz_destroy(c); return rc; } } while (0)
#define TRYCUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { bz_set_error(nullptr, "%s: %s", #x, cudaGetErrorString(e_)); bz_destroy(c); return BZ_ERR_CUDA; } } while (0)
    TRYCUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);                 // the exchange stream's CTAs go first whenever an SM has room
        TRYCUDA(cudaStreamCreateWithPriority(&c->s2, cudaStreamNonBlocking, hi));
    }
    for (int e = 0; e < 8; ++e) TRYCUDA(cudaEventCreateWithFlags(&c->ev_chunk[e], cudaEventDisableTiming));
    if (const char* e = getenv("BZ_DMA_TRANSPOSE")) c->dma_transpose = atoi(e) ? 1 : 0;
    c->fft_z_chunks = c->dma_transpose ? 4 : 1;
    if (const char* e = getenv("BZ_FFT_Z_CHUNKS")) { int v = atoi(e); if (v >= 1 && v <= FFT_Z_CHUNKS_MAX) c->fft_z_chunks = v; }
    TRYCUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    TRYCUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    if (const char* e = getenv("BZ_NO_OVERLAP")) c->overlap = atoi(e) ? 0 : 1;
    TRYCUDA(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
    TRYCUDA(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
    TRYCUDA(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    TRYCUDA(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming));
    rc = comm_init(c->comm, cfg, c->stream);
    if (rc) { bz_set_error(nullptr, "comm_init: %s", c->comm.err); bz_destroy(c); return rc; }
    {
        PoissonGeom& G = c->PG;
        G.Nx = cfg->Nx; G.Ny = cfg->Ny; G.Nz = cfg->Nz;
        G.nky = fy ? 1 : cfg->Ny / 2 + 1;
        comm_split_ky(c->comm, G.nky, &G.ky0, &G.nky_loc);
        G.P = c->comm.n_ranks;
        G.nx = L.nx;
        const size_t nW = (size_t)L.nx * G.nky * G.Nz, nW2 = (size_t)G.Nx * G.nky_loc * G.Nz;
        const size_t field_bytes = (size_t)16 * L.n * sizeof(double);
        // one page of barrier flags (comm.cuh) right behind the fields: like off_W / off_W2 its offset must be the SAME on every rank
        // (the peers address it through their own copy of these offsets; only the size of W2 differs between ranks)
        c->comm.off_flags = (long long)((field_bytes + 255) & ~(size_t)255);
        c->comm.off_pack = c->comm.off_flags + 4096;
        {   // packed x faces (comm.cuh): regions for the ρu face (1 column, 1 field), φ (1 field), the scalars (2 fields), all five fields
            const size_t col = (size_t)L.Ny * L.Nz, hx = (size_t)(L.HX > 0 ? L.HX : 1);
            c->comm.pack_start[0] = 0;
            c->comm.pack_start[1] = col;
            c->comm.pack_start[2] = c->comm.pack_start[1] + 2 * hx * col;
            c->comm.pack_start[3] = c->comm.pack_start[2] + 2 * 2 * hx * col;
            c->comm.pack_bytes = (P > 1) ? ((c->comm.pack_start[3] + 2 * NPROG * hx * col) * sizeof(double) + 255) & ~(size_t)255 : 0;
        }
        c->off_W = (size_t)c->comm.off_pack + c->comm.pack_bytes;
        c->off_W2 = (c->off_W + nW * sizeof(double2) + 255) & ~(size_t)255;
        c->arena_bytes = (P > 1) ? c->off_W2 + (nW2 > 0 ? nW2 : 1) * sizeof(double2) : c->off_W2;
        TRYCUDA(cudaMalloc((void**)&c->arena, c->arena_bytes));
        TRYCUDA(cudaMemsetAsync(c->arena, 0, c->arena_bytes, c->stream));
        c->bytes += (int64_t)c->arena_bytes;
        for (int s = 0; s < 3; ++s) for (int f = 0; f < NPROG; ++f) c->set[s][f] = c->arena + (size_t)(s * NPROG + f) * L.n;
        c->phi = c->arena + (size_t)15 * L.n;
        c->W = (double2*)((char*)c->arena + c->off_W);
        c->W2 = (P > 1) ? (double2*)((char*)c->arena + c->off_W2) : c->W;
    }
    TRY(dev_alloc(c, &c->dense, (size_t)L.nx * L.Ny * (L.Nz + 1)));
    TRY(dev_alloc(c, &c->scalar, 8));
    TRY(dev_alloc(c, &c->col_store, (size_t)8 * (L.Nz + 1) + (size_t)LEV_REC * L.Nz));
    default_reference_state(c);
    TRY(upload_columns(c));
    TRY(setup_poisson(c));
    TRY(comm_alloc_buffers(c->comm, L, c->PG, &c->bytes));
    // staging of the stage kernel's operands
    const bool tma_possible = !fx && ((L.PX * sizeof(double)) % 16 == 0) && c->buf == 3;   // TMA: row pitch a multiple of 16 bytes
    if (c->buf > 3) for (int f = 0; f < NPROG; ++f) TRY(dev_alloc(c, &c->V[f], (size_t)L.n));
    c->use_tma = (cfg->use_tma == 2) ? 0 : (tma_possible ? 1 : 0);
    if (cfg->use_tma == 1 && !tma_possible) { bz_set_error(nullptr, "TMA staging needs an even padded row length"); bz_destroy(c); return BZ_ERR_UNSUPPORTED; }
    if (c->use_tma) {
        if (!fy) TRY(make_tensor_maps(c, StageShared<32, 8, true>::SW, StageShared<32, 8, true>::SH));
        else TRY(make_tensor_maps(c, StageShared<128, 1, false>::SW, 1));
    }
    {
        int gx = fy ? (L.nx + 127) / 128 : (L.nx + 31) / 32, gy = fy ? 1 : (L.Ny + 7) / 8;
        // z-chunks (blockIdx.z): one CTA per SM, so the step time is waves x levels per CTA. Model: ceil(tiles * ch / 148) waves of
        // ceil(Nz / ch) + 5 levels (prologue planes + the one replayed level); the smallest chunk count that minimises it wins.
        // Measured optima (scripts/zchunk_sweep.py): 512^3 -> 1 (13.10 ms; 3: 13.18, 6: 13.35), 256^3 -> 4, 128^3 -> 2.
        int best_ch = 1;
        {
            const long long tiles = (long long)gx * gy;
            const int max_ch = L.Nz / 16 > 1 ? L.Nz / 16 : 1;
            double best = -1.0;
            for (int ch = 1; ch <= max_ch; ++ch) {
                double cost = (double)((tiles * ch + 147) / 148) * ((L.Nz + ch - 1) / ch + 5);
                if (best < 0.0 || cost < best) { best = cost; best_ch = ch; }
            }
        }
        c->z_chunks = cfg->z_chunks > 0 ? cfg->z_chunks : best_ch;
        if (c->z_chunks > L.Nz) c->z_chunks = L.Nz;
    }
    // initialize_model_thermodynamics!: θ = θ₀ (anelastic_time_stepping.jl:15-19)
    {
        std::vector<double> h((size_t)L.nx * L.Ny * L.Nz);
        for (int k = 0; k < L.Nz; ++k)
        {
            double th0 = cfg->potential_temperature;
            if (cfg->formulation == BZ_FORMULATION_STATIC_ENERGY)       // set!(model, θ = θ₀) → e = cᵖᵈ Π θ₀ + g z (static_energy_tendency.jl:113-146)
                th0 = c->th.cpd * (pow(c->h_p[k] / cfg->standard_pressure, c->th.Rd / c->th.cpd) * th0) + c->th.g * (cfg->z0 + (k + 0.5) * L.dz);
            for (size_t a = 0; a < (size_t)L.nx * L.Ny; ++a) h[(size_t)k * L.nx * L.Ny + a] = c->h_rho[k] * th0;
        }
        TRYCUDA(cudaMemcpyAsync(c->dense, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        dim3 grid((L.nx + 127) / 128, L.Ny, L.Nz);
        scatter_interior<<<grid, 128, 0, c->stream>>>(L, c->dense, c->set[0][BZ_RHO_THETA], 0);
        c->launches++;
        TRYCUDA(cudaStreamSynchronize(c->stream));
        TRY(fill_halos(c, c->set[0], NPROG, 4));
    }
    TRYCUDA(cudaStreamSynchronize(c->stream));
#undef TRY
#undef TRYCUDA
    *out = c;
    return BZ_OK;
}

int bz_get_reference_state(bz_ctx* c, double* rho, double* p, double* T) {
    if (!c) return BZ_ERR_INVALID;
    for (int k = 0; k < c->L.Nz; ++k) {
        if (rho) rho[k] = c->h_rho[k];
        if (p) p[k] = c->h_p[k];
        if (T) T[k] = c->h_T[k];
    }
    return BZ_OK;
}

int bz_set_reference_state(bz_ctx* c, const double* rho, const double* p, const double* T) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    for (int k = 0; k < c->L.Nz; ++k) {
        if (rho) c->h_rho[k] = rho[k];
        if (p) c->h_p[k] = p[k];
        if (T) c->h_T[k] = T[k];
    }
    c->graphs.clear();
    int rc = upload_columns(c);
    if (rc) return rc;
    return setup_thomas(c);
}

#define COPY_CHUNKS_MAX 32
// z chunks per field: the upload of a chunk can start as soon as the download of the same chunk has landed (BZ_COPY_CHUNKS overrides: sweeps)
static int copy_chunks() {
    static int n = 0;
    if (!n) { const char* e = getenv("BZ_COPY_CHUNKS"); n = e ? atoi(e) : 8; if (n < 1) n = 1; if (n > COPY_CHUNKS_MAX) n = COPY_CHUNKS_MAX; }
    return n;
}
#define COPY_CHUNKS copy_chunks()

static void chunk_range(int nz, int ch, int* k0, int* k1) {
    const int per = (nz + COPY_CHUNKS - 1) / COPY_CHUNKS;
    *k0 = ch * per < nz ? ch * per : nz;
    *k1 = (ch + 1) * per < nz ? (ch + 1) * per : nz;
}

// levels [k0, k1) of a dense host array (x fastest, nx x Ny per level) <-> the interior of a padded device field
static cudaError_t copy_levels(const bz_ctx* c, double* dev_field, double* host, int k0, int k1, bool to_host, cudaStream_t s) {
    const Layout& L = c->L;
    if (k1 <= k0) return cudaSuccess;
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof(p));
    cudaPitchedPtr hp = make_cudaPitchedPtr(host + (size_t)k0 * L.nx * L.Ny, (size_t)L.nx * sizeof(double), (size_t)L.nx * sizeof(double), (size_t)L.Ny);
    cudaPitchedPtr dp = make_cudaPitchedPtr(dev_field + (size_t)k0 * L.plane, (size_t)L.PX * sizeof(double), (size_t)L.PX * sizeof(double), (size_t)L.PY);
    const cudaPos dpos = make_cudaPos((size_t)L.HX * sizeof(double), (size_t)L.HY, 0), hpos = make_cudaPos(0, 0, 0);
    if (to_host) { p.srcPtr = dp; p.srcPos = dpos; p.dstPtr = hp; p.dstPos = hpos; p.kind = cudaMemcpyDeviceToHost; }
    else { p.srcPtr = hp; p.srcPos = hpos; p.dstPtr = dp; p.dstPos = dpos; p.kind = cudaMemcpyHostToDevice; }
    p.extent = make_cudaExtent((size_t)L.nx * sizeof(double), (size_t)L.Ny, (size_t)(k1 - k0));
    return cudaMemcpy3DAsync(&p, s);
}

int bz_set_state_async(bz_ctx* c, const double* ru, const double* rv, const double* rw, const double* rth, const double* rq, int enforce) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    const Layout& L = c->L;
    const double* src[NPROG] = {ru, rv, rw, rth, rq};
    CUDA_TRY(c, cudaEventRecord(c->ev_main, c->stream));          // the fields are free once the work queued so far has completed
    CUDA_TRY(c, cudaStreamWaitEvent(c->s_in, c->ev_main, 0));
    for (int f = 0; f < NPROG; ++f) {
        if (!src[f]) continue;
        double* dst = c->set[c->cur][f];
        // A pending bz_get_state_async gates this upload chunk by chunk: the download of the SAME field still reads the device chunk we
        // are about to overwrite, and a download that targets the host buffer we read from must have landed first.
        int dep = -1;
        for (int g = 0; g < NPROG; ++g) if (c->out_ptr[g] && c->out_ptr[g] == src[f]) dep = g;
        for (int ch = 0; ch < COPY_CHUNKS; ++ch) {
            int k0, k1; chunk_range(L.Nz, ch, &k0, &k1);
            if (c->out_ptr[f]) CUDA_TRY(c, cudaStreamWaitEvent(c->s_in, c->ev_out[f * COPY_CHUNKS + ch], 0));
            if (dep >= 0 && dep != f) CUDA_TRY(c, cudaStreamWaitEvent(c->s_in, c->ev_out[dep * COPY_CHUNKS + ch], 0));
            if (f == BZ_RHO_W && k0 == 0) {                        // the wall face k = 0 is not the caller's to set (always 0)
                CUDA_TRY(c, cudaMemset2DAsync(dst + (size_t)L.HY * L.PX + L.HX, (size_t)L.PX * sizeof(double), 0, (size_t)L.nx * sizeof(double), (size_t)L.Ny, c->s_in));
                k0 = 1;
            }
            CUDA_TRY(c, copy_levels(c, dst, const_cast<double*>(src[f]), k0, k1, false, c->s_in));
        }
    }
    CUDA_TRY(c, cudaEventRecord(c->ev_in, c->s_in));
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_in, 0));
    int rc;
    if ((rc = fill_halos(c, c->set[c->cur], NPROG, 4))) return rc;
    if (enforce && (rc = pressure_correct(c, 1.0))) return rc;
    return BZ_OK;
}

int bz_set_state(bz_ctx* c, const double* ru, const double* rv, const double* rw, const double* rth, const double* rq, int enforce) {
    int rc = bz_set_state_async(c, ru, rv, rw, rth, rq, enforce);
    if (rc) return rc;
    CUDA_TRY(c, cudaStreamSynchronize(c->s_in));                   // the caller's buffers are free again
    return BZ_OK;
}

int bz_set_forcing(bz_ctx* c, const bz_forcing* F) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    c->graphs.clear();                                  // captured steps carry the old forcing arguments
    if (!F) { c->forced = 0; return BZ_OK; }
    if (c->cfg.formulation == BZ_FORMULATION_STATIC_ENERGY) { bz_set_error(c, "forcings are on the path for the potential-temperature formulation only"); return BZ_ERR_UNSUPPORTED; }
    const int Nz = c->L.Nz;
    if (c->L.flat_x && (F->coriolis_f != 0 || F->drag_rho_ustar2 != 0)) { bz_set_error(c, "Coriolis / drag need a non-Flat x"); return BZ_ERR_UNSUPPORTED; }
    const size_t total = (size_t)(Nz + 1) + 4 * (size_t)Nz + 8 * (size_t)Nz;
    if (!c->fstore) { int rc = dev_alloc(c, &c->fstore, total); if (rc) return rc; }
    std::vector<double> h(total, 0.0);
    double* base = c->fstore;
    size_t off = 0;
    auto put = [&](const auto* src, size_t n, auto** dev) {      // the profiles of bz_forcing are double in both precisions
        if (src) { for (size_t q = 0; q < n; ++q) h[off + q] = src[q]; *dev = base + off; } else *dev = nullptr;
        off += n;
    };
    put(F->subsidence_w, Nz + 1, &c->d_ws);
    put(F->geostrophic_u, Nz, &c->d_ug);
    put(F->geostrophic_v, Nz, &c->d_vg);
    put(F->q_tendency, Nz, &c->d_qt);
    put(F->e_tendency, Nz, &c->d_et);
    c->d_sums = base + off; off += (size_t)4 * Nz;
    c->d_fcol = base + off; off += (size_t)4 * Nz;
    CUDA_TRY(c, cudaMemcpyAsync(c->fstore, h.data(), total * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->coriolis_f = F->coriolis_f; c->theta_flux = F->theta_flux; c->q_flux = F->q_flux; c->drag = F->drag_rho_ustar2;
    c->subs_mask = F->subsidence_mask;
    c->forced = 1;
    return BZ_OK;
}

static int time_step_body(bz_ctx* c, double dt) {
    const double alpha[3] = {1.0, 1.0 / 4.0, 2.0 / 3.0};
    const int u0 = c->cur;                             // store_initial_state! without a copy
    int rc;
    for (int s = 0; s < 3; ++s) {
        int in = c->cur;
        int nxt = 0;
        while (nxt == in || nxt == u0) ++nxt;          // the free set
        if ((rc = launch_stage(c, in, c->set[nxt], u0, dt, alpha[s], 0))) return rc;
        c->cur = nxt;
        if ((rc = start_scalar_exchange(c))) return rc;
        if ((rc = pressure_correct(c, alpha[s] * dt))) return rc;
    }
    return BZ_OK;
}

int bz_time_step(bz_ctx* c, double dt) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    int rc;
    // one GPU, no per-kernel events wanted: replay the step as a CUDA graph keyed by (which buffer set holds the state, dt)
    if (c->graphs.on((long long)c->L.nx * c->L.Ny * c->L.Nz) && c->comm.n_ranks == 1 && !c->prof_on) {
        StepGraphEntry* g = c->graphs.find(c->cur, dt);
        if (g && g->exec) {
            CUDA_TRY(c, cudaGraphLaunch(g->exec, c->stream));
            c->cur = g->cur_after; c->launches += g->launches;
        } else if (g && g->seen) {                      // second step with this key: capture it
            const long long l0 = c->launches;
            cudaGraph_t graph = nullptr;
            CUDA_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
            rc = time_step_body(c, dt);
            cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (e != cudaSuccess || !graph) { bz_set_error(c, "graph capture of the time step failed: %s", cudaGetErrorString(e)); return BZ_ERR_CUDA; }
            e = cudaGraphInstantiate(&g->exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { g->exec = nullptr; bz_set_error(c, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); return BZ_ERR_CUDA; }
            g->cur_after = c->cur; g->launches = c->launches - l0;
            CUDA_TRY(c, cudaGraphLaunch(g->exec, c->stream));      // the capture recorded the step without running it
        } else {
            if (!g) g = c->graphs.add(c->cur, dt);
            g->seen = 1;
            if ((rc = time_step_body(c, dt))) return rc;
        }
    } else if ((rc = time_step_body(c, dt))) return rc;
    c->time += dt;
    c->iteration += 1;
    return BZ_OK;
}

int bz_time_steps(bz_ctx* c, double dt, int n) {
    for (int s = 0; s < n; ++s) { int rc = bz_time_step(c, dt); if (rc) return rc; }
    return BZ_OK;
}

int bz_compute_tendencies(bz_ctx* c) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    int rc;
    for (int f = 0; f < NPROG; ++f) if (!c->G[f] && (rc = dev_alloc(c, &c->G[f], (size_t)c->L.n))) return rc;
    return launch_stage(c, c->cur, c->G, c->cur, 0.0, 1.0, 1);
}

static int download_dense(bz_ctx* c, double* host, int nz_out) {
    size_t n = (size_t)c->L.nx * c->L.Ny * nz_out;
    CUDA_TRY(c, cudaMemcpyAsync(host, c->dense, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return BZ_OK;
}

int bz_get_tendency(bz_ctx* c, int f, double* out) {
    if (!c || f < 0 || f >= NPROG || !out) return BZ_ERR_INVALID;
    if (!c->G[f]) { bz_set_error(c, "call bz_compute_tendencies first"); return BZ_ERR_STATE; }
    cudaSetDevice(c->cfg.device);
    const Layout& L = c->L;
    int nz_out = (f == BZ_RHO_W) ? L.Nz + 1 : L.Nz;
    dim3 grid((L.nx + 127) / 128, L.Ny, nz_out);
    extract_interior<<<grid, 128, 0, c->stream>>>(L, c->G[f], c->dense, nz_out);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return download_dense(c, out, nz_out);
}

int bz_pressure_correct(bz_ctx* c, double dt) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    int rc;
    if ((rc = fill_halos(c, c->set[c->cur], NPROG, 4))) return rc;
    return pressure_correct(c, dt);
}

int bz_get_field(bz_ctx* c, int f, double* out) {
    if (!c || !out) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    const Layout& L = c->L;
    int nz_out = (f == BZ_RHO_W || f == BZ_W) ? L.Nz + 1 : L.Nz;
    dim3 grid((L.nx + 127) / 128, L.Ny, nz_out);
    if (f >= 0 && f < NPROG) extract_interior<<<grid, 128, 0, c->stream>>>(L, c->set[c->cur][f], c->dense, nz_out);
    else if (f == BZ_PHI) extract_interior<<<grid, 128, 0, c->stream>>>(L, c->phi, c->dense, nz_out);
    else if (f >= BZ_U && f <= BZ_QL) {
        FieldSet U; U.n = NPROG;
        for (int a = 0; a < NPROG; ++a) U.f[a] = c->set[c->cur][a];
        if (c->cfg.formulation == BZ_FORMULATION_STATIC_ENERGY) diagnose_field<BZ_THERMO_STATIC_ENERGY><<<grid, 128, 0, c->stream>>>(L, c->col, c->th, U, f, c->dense, nz_out);
        else if (c->cfg.microphysics == BZ_MICROPHYSICS_NONE) diagnose_field<0><<<grid, 128, 0, c->stream>>>(L, c->col, c->th, U, f, c->dense, nz_out);
        else diagnose_field<1><<<grid, 128, 0, c->stream>>>(L, c->col, c->th, U, f, c->dense, nz_out);
    } else return BZ_ERR_INVALID;
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    return download_dense(c, out, nz_out);
}

// fills c->dense with interior(field) (nz_out levels) on the stream; no host copy
static int field_to_dense(bz_ctx* c, int f, int* nz_out_p) {
    const Layout& L = c->L;
    int nz_out = (f == BZ_RHO_W || f == BZ_W) ? L.Nz + 1 : L.Nz;
    dim3 grid((L.nx + 127) / 128, L.Ny, nz_out);
    if (f >= 0 && f < NPROG) extract_interior<<<grid, 128, 0, c->stream>>>(L, c->set[c->cur][f], c->dense, nz_out);
    else if (f == BZ_PHI) extract_interior<<<grid, 128, 0, c->stream>>>(L, c->phi, c->dense, nz_out);
    else if (f >= BZ_U && f <= BZ_QL) {
        FieldSet U; U.n = NPROG;
        for (int a = 0; a < NPROG; ++a) U.f[a] = c->set[c->cur][a];
        if (c->cfg.formulation == BZ_FORMULATION_STATIC_ENERGY) diagnose_field<BZ_THERMO_STATIC_ENERGY><<<grid, 128, 0, c->stream>>>(L, c->col, c->th, U, f, c->dense, nz_out);
        else if (c->cfg.microphysics == BZ_MICROPHYSICS_NONE) diagnose_field<0><<<grid, 128, 0, c->stream>>>(L, c->col, c->th, U, f, c->dense, nz_out);
        else diagnose_field<1><<<grid, 128, 0, c->stream>>>(L, c->col, c->th, U, f, c->dense, nz_out);
    } else return BZ_ERR_INVALID;
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    *nz_out_p = nz_out;
    return BZ_OK;
}

int bz_get_slice(bz_ctx* c, int f, int axis, int index, double* out) {
    if (!c || !out || axis < 0 || axis > 2) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    const Layout& L = c->L;
    int nz_out = 0;
    int rc = field_to_dense(c, f, &nz_out);
    if (rc) return rc;
    const int lim = axis == 0 ? L.nx : (axis == 1 ? L.Ny : nz_out);
    if (index < 0 || index >= lim) { bz_set_error(c, "slice index %d out of range [0, %d)", index, lim); return BZ_ERR_INVALID; }
    const size_t n = (size_t)(axis == 0 ? L.Ny : L.nx) * (axis == 2 ? L.Ny : nz_out);
    if (!c->slice_buf || c->slice_cap < n) {
        cudaFree(c->slice_buf); c->slice_buf = nullptr; c->slice_cap = 0;
        CUDA_TRY(c, cudaMalloc((void**)&c->slice_buf, n * sizeof(double)));
        c->slice_cap = n;
    }
    slice_kernel<<<(int)((n + 255) / 256), 256, 0, c->stream>>>(L.nx, L.Ny, nz_out, c->dense, c->slice_buf, axis, index);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(out, c->slice_buf, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return BZ_OK;
}

int bz_state_is_finite(bz_ctx* c, int* finite) {
    if (!c || !finite) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    const Layout& L = c->L;
    int* flag = reinterpret_cast<int*>(c->scalar + 1);
    CUDA_TRY(c, cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
    FieldSet U; U.n = NPROG;
    for (int a = 0; a < NPROG; ++a) U.f[a] = c->set[c->cur][a];
    long long total = (long long)L.nx * L.Ny * L.Nz;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
    nonfinite_kernel<<<blocks, 256, 0, c->stream>>>(L, U, flag);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    int bad = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&bad, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->comm.n_ranks > 1) {
        double v = bad ? 1.0 : 0.0;
        int rc = comm_allreduce_max(c->comm, &v, c->scalar, c->stream);
        if (rc) { bz_set_error(c, "allreduce: %s", c->comm.err); return rc; }
        bad = v > 0.0;
    }
    *finite = bad ? 0 : 1;
    return BZ_OK;
}

int bz_get_state_async(bz_ctx* c, double* ru, double* rv, double* rw, double* rth, double* rq) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    const Layout& L = c->L;
    double* dst[NPROG] = {ru, rv, rw, rth, rq};
    if (c->ev_out.empty()) {
        c->ev_out.resize((size_t)NPROG * COPY_CHUNKS_MAX);
        for (auto& e : c->ev_out) CUDA_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    CUDA_TRY(c, cudaEventRecord(c->ev_main, c->stream));
    CUDA_TRY(c, cudaStreamWaitEvent(c->s_out, c->ev_main, 0));
    for (int f = 0; f < NPROG; ++f) {
        c->out_ptr[f] = dst[f];
        if (!dst[f]) continue;
        for (int ch = 0; ch < COPY_CHUNKS; ++ch) {
            int k0, k1; chunk_range(L.Nz, ch, &k0, &k1);
            CUDA_TRY(c, copy_levels(c, c->set[c->cur][f], dst[f], k0, k1, true, c->s_out));
            CUDA_TRY(c, cudaEventRecord(c->ev_out[f * COPY_CHUNKS + ch], c->s_out));
        }
        if (f == BZ_RHO_W) memset(dst[f] + (size_t)L.Nz * L.nx * L.Ny, 0, (size_t)L.nx * L.Ny * sizeof(double));   // top wall face (not stored on the device)
    }
    return BZ_OK;
}

int bz_get_state(bz_ctx* c, double* ru, double* rv, double* rw, double* rth, double* rq) {
    int rc = bz_get_state_async(c, ru, rv, rw, rth, rq);
    if (rc) return rc;
    CUDA_TRY(c, cudaStreamSynchronize(c->s_out));
    for (int f = 0; f < NPROG; ++f) c->out_ptr[f] = nullptr;       // landed: nothing left to order against
    return BZ_OK;
}

int bz_get_clock(bz_ctx* c, double* time, int64_t* iteration) {   // BZ_KEEP_F64
    if (!c) return BZ_ERR_INVALID;
    if (time) *time = c->time;
    if (iteration) *iteration = c->iteration;
    return BZ_OK;
}

static int reduce_max(bz_ctx* c, int which, double* out) {
    cudaSetDevice(c->cfg.device);
    const Layout& L = c->L;
    double** U = c->set[c->cur];
    CUDA_TRY(c, cudaMemsetAsync(c->scalar, 0, sizeof(double), c->stream));
    long long total = (long long)L.nx * L.Ny * L.Nz;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
    reduce_max_kernel<<<blocks, 256, 0, c->stream>>>(L, c->col, U[0], U[1], U[2], which, c->scalar);
    c->launches++;
    CUDA_TRY(c, cudaGetLastError());
    double v = 0.0;
    CUDA_TRY(c, cudaMemcpyAsync(&v, c->scalar, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->comm.n_ranks > 1) { int rc = comm_allreduce_max(c->comm, &v, c->scalar, c->stream); if (rc) { bz_set_error(c, "allreduce: %s", c->comm.err); return rc; } }
    *out = v;
    return BZ_OK;
}

int bz_cell_advection_timescale(bz_ctx* c, double* tau) {
    if (!c || !tau) return BZ_ERR_INVALID;
    double m; int rc = reduce_max(c, 1, &m);
    if (rc) return rc;
    *tau = 1.0 / m;
    return BZ_OK;
}

int bz_max_abs_divergence(bz_ctx* c, double* out) {
    if (!c || !out) return BZ_ERR_INVALID;
    return reduce_max(c, 0, out);
}

int bz_synchronize(bz_ctx* c) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->s_in));
    CUDA_TRY(c, cudaStreamSynchronize(c->s_out));
    for (int f = 0; f < NPROG; ++f) c->out_ptr[f] = nullptr;
    if (c->comm.p2p) {                                              // a flag barrier that timed out (dead peer) left its mark
        unsigned long long bad = 0;
        CUDA_TRY(c, cudaMemcpy(&bad, c->comm.peer_base[c->comm.rank] + c->comm.off_flags + BZ_FLAG_TIMEOUT_SLOT * 8, 8, cudaMemcpyDeviceToHost));
        if (bad) { bz_set_error(c, "a peer-memory barrier timed out: another rank stopped"); return BZ_ERR_STATE; }
    }
    return BZ_OK;
}

int bz_profile_enable(bz_ctx* c, int on) {
    if (!c) return BZ_ERR_INVALID;
    c->prof_on = on;
    return BZ_OK;
}

int bz_profile_read(bz_ctx* c, double* ms, int64_t* n) {
    if (!c) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (size_t e = 0; e < c->prof_fam.size(); ++e) {
        float t = 0.f;
        cudaEventElapsedTime(&t, c->prof_ev[2 * e], c->prof_ev[2 * e + 1]);
        c->prof_ms[c->prof_fam[e]] += t; c->prof_n[c->prof_fam[e]] += 1;
        c->prof_pool.push_back(c->prof_ev[2 * e]); c->prof_pool.push_back(c->prof_ev[2 * e + 1]);
    }
    c->prof_ev.clear(); c->prof_fam.clear();
    for (int f = 0; f < NFAM; ++f) { if (ms) ms[f] = c->prof_ms[f]; if (n) n[f] = c->prof_n[f]; c->prof_ms[f] = 0; c->prof_n[f] = 0; }
    return BZ_OK;
}

int bz_nccl_unique_id(uint8_t* out128) {
    if (!out128) return BZ_ERR_INVALID;
    return comm_unique_id(out128, g_err);
}

int bz_ipc_export(bz_ctx* c, uint8_t* out64) {
    if (!c || !out64) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    return comm_ipc_export(c->arena, out64, c->err);
}

int bz_ipc_attach(bz_ctx* c, const uint8_t* handles) {
    if (!c || !handles) return BZ_ERR_INVALID;
    cudaSetDevice(c->cfg.device);
    cudaStreamSynchronize(c->stream);
    int rc = comm_ipc_attach(c->comm, c->arena, handles);
    if (rc) bz_set_error(c, "%s", c->comm.err);
    return rc;
}

int64_t bz_kernel_launch_count(const bz_ctx* c) { return c ? c->launches : 0; }
void* bz_stream(bz_ctx* c) { return c ? (void*)c->stream : nullptr; }
int64_t bz_device_bytes(const bz_ctx* c) { return c ? c->bytes : 0; }

}  // extern "C"

// second hot-path family: compressible WS-RK3 with acoustic substepping (include/breeze_b200_compressible.h; bzc_ / bzcf_ entry points)
#include "compressible_api.cuh"
