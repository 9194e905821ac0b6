// poisson.cuh — in-house anelastic pressure solver: slab FFT (y real-to-complex, x complex) + batched Thomas in z.
//
// Replaces compute_anelastic_source_term! + Oceananigans' solve!(::FourierTridiagonalPoissonSolver)
// (src/AnelasticEquations/anelastic_pressure_solver.jl:84-105; oracle: compute_pressure_correction).
// No cuFFT: the transforms are Stockham radix-8/4/2 passes on lines held in shared memory, FP64, with a
// precomputed twiddle table, so that the source term (Δz·div(ρu)/Δt) is formed while the first pass loads its
// lines and the inverse's last pass writes φ straight into the padded field.
//
//   pass 1  poisson_forward_y : rhs on the fly → real-to-complex FFT along y (two real x-columns per complex line)
//                               → W[k][ky][i], ky = 0..Ny/2                                   (x-slab layout)
//   (multi-GPU: all-to-all so that x becomes local and ky distributed → W[k][ky_local][kx_global])
//   pass 2  fft_x (forward)   : complex FFT along x, in place
//   pass 3  thomas_z          : tridiagonal solve in z per (kx, ky); 1/β and the elimination factors are
//                               precomputed once (they depend on the grid and ρᵣ only); the singular (0,0) mode is
//                               pinned and its vertical mean removed (= the reference's φ .-= mean(φ))
//   pass 4  fft_x (inverse), (transpose back), pass 5 poisson_inverse_y : complex-to-real along y → φ
#pragma once
#include "common.cuh"

struct PoissonGeom {
    int Nx;        // GLOBAL x size (length of the x transforms)
    int Ny;        // y size (length of the y transforms), 1 if Flat
    int Nz;
    int nky;       // number of ky modes held after the y transform: Ny/2 + 1 (1 if Flat y)
    int nky_loc;   // ky modes on this rank in the transposed layout
    int ky0;       // first global ky of this rank in the transposed layout
    int P;         // ranks (x-slabs)
    int nx;        // columns per rank, Nx / P
    const long long* ky_base;   // per ky: offset of (k = 0, ky, i = 0) in the peer-blocked W   (device, nky entries)
    const int* ky_kstride;      // per ky: level stride of its block = count_p * nx              (device, nky entries)
    // peer-memory path (CUDA IPC): the all-to-all transposes become peer loads inside the consuming FFT kernels
    const int* ky_owner;        // per ky: rank that owns it in the transposed layout
    const long long* ky_base2;  // per ky: offset of (k = 0, ky, i = 0) of THIS rank's block in the owner's W2
    long long off_W, off_W2;    // byte offsets of W / W2 in every rank's arena
    int rank;
};

struct PeerBases { const char* base[8]; };   // arena base address of every rank, mapped into this process

// Peer-blocked layouts of the two spectral arrays. The all-to-all of the distributed FFT then moves CONTIGUOUS blocks
// straight between them (no pack / unpack passes); with one rank both reduce to the plain [k][ky][x] order.
//   W  (x-slab side,   nx × nky × Nz):      block p = the ky range owned by rank p  → [p][k][ky - start_p][i]
//   W2 (transposed,    Nx × nky_loc × Nz):  block p = the x columns of rank p       → [p][k][ky_loc][i_p]
__host__ __device__ __forceinline__ void ky_block(int nky, int P, int p, int* start, int* count) {
    int base = nky / P, rem = nky % P;
    *count = base + (p < rem ? 1 : 0);
    *start = p * base + (p < rem ? p : rem);
}
__device__ __forceinline__ size_t w_index(const PoissonGeom& G, int k, int ky, int i) {
    if (G.P == 1) return ((size_t)k * G.nky + ky) * G.nx + i;
    return (size_t)__ldg(&G.ky_base[ky]) + (size_t)k * __ldg(&G.ky_kstride[ky]) + i;
}
__device__ __forceinline__ size_t w2_index(const PoissonGeom& G, int k, int kyl, int kx) {
    const int p = kx / G.nx, i = kx - p * G.nx;
    return (((size_t)p * G.Nz + k) * G.nky_loc + kyl) * G.nx + i;
}

// ---- register butterflies (forward: e^{-2πi/R}) ----------------------------------------------------------------
struct cpx { double x, y; };
__device__ __forceinline__ cpx operator+(cpx a, cpx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cpx operator-(cpx a, cpx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cpx mul_mi(cpx a) { return {a.y, -a.x}; }    // × (-i)

__device__ __forceinline__ void dft2(cpx& a, cpx& b) { cpx t = a - b; a = a + b; b = t; }
__device__ __forceinline__ void dft4(cpx& x0, cpx& x1, cpx& x2, cpx& x3) {
    cpx s02 = x0 + x2, d02 = x0 - x2, s13 = x1 + x3, d13 = mul_mi(x1 - x3);
    x0 = s02 + s13; x2 = s02 - s13; x1 = d02 + d13; x3 = d02 - d13;
}
__device__ __forceinline__ void dft3(cpx& x0, cpx& x1, cpx& x2) {          // forward: W3 = e^{-2πi/3} = -1/2 - i √3/2
    const double c = -0.5, sn = 0.86602540378443864676;
    cpx s = x1 + x2, d = x1 - x2;
    cpx m = {x0.x + c * s.x, x0.y + c * s.y};
    cpx r = {sn * d.y, -sn * d.x};                                        // -i sin(2π/3) (x1 - x2)
    x0 = x0 + s; x1 = m + r; x2 = m - r;
}
// forward DFT of odd prime length R = 5 / 7 in the symmetric form: with S_k = x_k + x_{R-k}, D_k = x_k - x_{R-k},
//   X_m = x_0 + Σ_k cos(2π m k / R) S_k - i Σ_k sin(2π m k / R) D_k,   X_{R-m} = the same with + i      (m, k = 1 .. (R-1)/2)
template <int R>
__device__ __forceinline__ void dft_odd(cpx* v) {
    constexpr int H = (R - 1) / 2;
    // cos / sin (2π j / R), j = 1 .. H
    constexpr double C5[2] = {0.30901699437494742410, -0.80901699437494742410}, S5[2] = {0.95105651629515357212, 0.58778525229247312917};
    constexpr double C7[3] = {0.62348980185873353053, -0.22252093395631440429, -0.90096886790241912624},
                     S7[3] = {0.78183148246802980871, 0.97492791218182360702, 0.43388373911755812048};
    cpx S[H], D[H];
#pragma unroll
    for (int k = 0; k < H; ++k) { S[k] = v[k + 1] + v[R - 1 - k]; D[k] = v[k + 1] - v[R - 1 - k]; }
    const cpx x0 = v[0];
    cpx sum = x0;
#pragma unroll
    for (int k = 0; k < H; ++k) sum = sum + S[k];
    v[0] = sum;
#pragma unroll
    for (int m = 1; m <= H; ++m) {
        cpx a = x0, b = {0.0, 0.0};
#pragma unroll
        for (int k = 1; k <= H; ++k) {
            const int j = (m * k) % R;                       // cos(2π j / R) = cos(2π (R - j) / R), sin changes sign
            const int jj = j <= H ? j : R - j;
            const double c = (R == 5) ? C5[jj - 1] : C7[jj - 1];
            const double sn = ((R == 5) ? S5[jj - 1] : S7[jj - 1]) * (j <= H ? 1.0 : -1.0);
            a = {a.x + c * S[k - 1].x, a.y + c * S[k - 1].y};
            b = {b.x + sn * D[k - 1].x, b.y + sn * D[k - 1].y};
        }
        const cpx r = {b.y, -b.x};                          // -i b
        v[m] = a + r; v[R - m] = a - r;
    }
}
__device__ __forceinline__ void dft8(cpx* v) {
    const double h = 0.70710678118654752440;
    cpx a0 = v[0] + v[4], a4 = v[0] - v[4];
    cpx a1 = v[1] + v[5], a5 = v[1] - v[5];
    cpx a2 = v[2] + v[6], a6 = v[2] - v[6];
    cpx a3 = v[3] + v[7], a7 = v[3] - v[7];
    a5 = {h * (a5.x + a5.y), h * (a5.y - a5.x)};          // × W8   = (1 - i)/√2
    a6 = mul_mi(a6);                                       // × W8²  = -i
    a7 = {h * (a7.y - a7.x), -h * (a7.x + a7.y)};          // × W8³  = (-1 - i)/√2
    dft4(a0, a1, a2, a3);
    dft4(a4, a5, a6, a7);
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = a4; v[3] = a5; v[5] = a6; v[7] = a7;
}

// Shared-memory layout of the FFT lines (doubles; 16 bank pairs of 8 bytes):
//   element n of a line sits at pidx(n) = n + n/16  → unit-stride and stride-8 accesses of a half-warp are conflict-free
//   (the 8-contiguous/stride-64 scatter of the second radix-8 pass is 2-way);
//   line pitch LP ≡ 4 (mod 16) and the imaginary array starts ≡ 2 (mod 16) after the real one → the transposing accesses of
//   the y kernels (lanes spanning 4 lines × 2..4 consecutive elements × re/im) are conflict-free as well.
__host__ __device__ __forceinline__ int pidx(int n) { return n + (n >> 4); }
__host__ __device__ __forceinline__ int line_pitch(int N) { return ((N + (N >> 4) + 15) & ~15) + 4; }
__host__ __device__ __forceinline__ int imag_offset(int N, int lines) { int o = lines * line_pitch(N); return o + ((2 - o) & 15); }
__host__ __device__ __forceinline__ size_t fft_smem_bytes(int N, int lines) { return (size_t)(imag_offset(N, lines) + lines * line_pitch(N)) * sizeof(double); }

// Line lengths: N = 2^m (8 .. 2048), 3 · 2^m (24 .. 1536; the reference benchmarks 768 x 768 x 256, Benchmarks.yml:41), 5 · 2^m (40 .. 1280)
// or 7 · 2^m (56 .. 1792; 896^3 is the largest Float32 case of the reference's memory table, benchmarking/README.md:225-233).
// Points per thread: 8 for the powers of two (radix-8 / 4 / 2 passes), 12 for 3 · 2^m (radix-4 / 2 passes and one radix-3 pass last, so
// that every earlier pass keeps a power-of-two Ns). 5 · 2^m and 7 · 2^m also run 8 points per thread through radix-8 / 4 / 2 passes; their
// last pass (radix 5 / 7) has N / R butterflies for N / 8 threads of a line, so its second round is partly idle (guarded).
__host__ __device__ constexpr int fft_odd_radix(int N) { return (N % 3 == 0) ? 3 : (N % 5 == 0) ? 5 : (N % 7 == 0) ? 7 : 1; }
__host__ __device__ constexpr int fft_pt(int N) { return (N % 3 == 0) ? 12 : 8; }

// One Stockham pass of radix R over lines of compile-time length N held as re[l*LP + pidx(n)], im[l*LP + pidx(n)].
// Each thread owns PT/R butterflies (PT complex values in registers): blockDim.x == lines * N / PT. N, R, Ns are
// compile-time, so every index below folds to shifts and immediates (the run-time-N version spent 85 % of its
// instructions on index arithmetic).
// padded offset of a compile-time displacement D from an element whose padded index is already known: when D is a multiple
// of 16 the padding term folds to a constant (pidx(n + D) = pidx(n) + D + D/16)
template <int D> __device__ __forceinline__ int pidx_plus(int n, int pn) { return (D % 16 == 0) ? pn + D + D / 16 : pidx(n + D); }

// IN: the pass takes its inputs from `ld(line, element)` (global memory) instead of shared memory; OUT: it hands its outputs to
// `st(line, element, value)` instead of storing them to shared memory. The x transform uses both at its two ends: element j + r N/R of the
// first pass and element j + r Ns of the last pass are contiguous across the threads of a line for every r, so the accesses coalesce and
// the transform keeps two of its four shared-memory round trips and four of its eight barriers.
struct NoLineIO {
    __device__ __forceinline__ cpx operator()(int, int) const { return cpx{0.0, 0.0}; }
    __device__ __forceinline__ void operator()(int, int, cpx) const {}
};
template <int N, int R, int Ns, bool IN = false, bool OUT = false, typename LD = NoLineIO, typename ST = NoLineIO>
__device__ __forceinline__ void stockham_pass(double* __restrict__ re, double* __restrict__ im, const double2* __restrict__ tw,
                                              const LD& ld = LD(), const ST& st = ST()) {
    constexpr int PT = fft_pt(N), PER_LINE = N / PT, NR = N / R, ITEMS = (NR + PER_LINE - 1) / PER_LINE, LP = ((N + (N >> 4) + 15) & ~15) + 4, STEP = N / (Ns * R);
    constexpr bool GUARD = (NR % PER_LINE) != 0;       // radix 5 / 7 with 8 points per thread: the last round of butterflies is partial
    static_assert(N % (Ns * R) == 0 && (Ns & (Ns - 1)) == 0, "pass shape");
    const int l = threadIdx.x / PER_LINE, t = threadIdx.x % PER_LINE;
    double* lre = re + l * LP;
    double* lim = im + l * LP;
    cpx v[ITEMS][R];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int j = t + it * PER_LINE;            // butterfly index in [0, N/R)
        if (GUARD && j >= NR) continue;
        const int pj = pidx(j);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if constexpr (IN) v[it][r] = ld(l, j + r * NR);
            else {
                const int o = (NR % 16 == 0) ? pj + r * (NR + NR / 16) : pidx(j + r * NR);
                v[it][r] = {lre[o], lim[o]};
            }
        }
    }
    if constexpr (!IN) __syncthreads();             // every thread has read its inputs before anybody overwrites them
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int j = t + it * PER_LINE;
        if (GUARD && j >= NR) continue;
        const int k = j & (Ns - 1);
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                double2 w = __ldg(&tw[r * k * STEP]);
                v[it][r] = cmul(v[it][r], cpx{w.x, w.y});
            }
        }
        if constexpr (R == 8) dft8(v[it]);
        else if constexpr (R == 4) dft4(v[it][0], v[it][1], v[it][2], v[it][3]);
        else if constexpr (R == 3) dft3(v[it][0], v[it][1], v[it][2]);
        else if constexpr (R == 5 || R == 7) dft_odd<R>(v[it]);
        else dft2(v[it][0], v[it][1]);
        const int j0 = (j - k) * R + k;
        const int pj0 = pidx(j0);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if constexpr (OUT) st(l, j0 + r * Ns, v[it][r]);
            else {
                const int o = (Ns % 16 == 0) ? pj0 + r * (Ns + Ns / 16) : pidx(j0 + r * Ns);
                lre[o] = v[it][r].x; lim[o] = v[it][r].y;
            }
        }
    }
    if constexpr (!OUT) __syncthreads();
}

// Forward DFT (e^{-2πi nk/N}) of every line; N = 2^m (8 .. 2048) or 3 · 2^m (24 .. 1536). Callers conjugate for the inverse.
// IO = false: the lines are in shared memory on entry and on exit (y transforms). IO = true: the first pass reads through `ld`, the last
// pass writes through `st`, shared memory only carries the values between passes (x transform).
template <int N, int Ns, int REM, bool IN, typename LD>     // REM = remaining power-of-two factor of a 3 · 2^m length: radix-4 passes, then one radix-2 if odd
__device__ __forceinline__ void pow2_passes_r4(double* __restrict__ re, double* __restrict__ im, const double2* __restrict__ tw, const LD& ld) {
    if constexpr (REM >= 4) { stockham_pass<N, 4, Ns, IN, false, LD>(re, im, tw, ld); pow2_passes_r4<N, Ns * 4, REM / 4, false, LD>(re, im, tw, ld); }
    else if constexpr (REM == 2) stockham_pass<N, 2, Ns, IN, false, LD>(re, im, tw, ld);
}
template <int N, bool IO = false, typename LD = NoLineIO, typename ST = NoLineIO>
__device__ __forceinline__ void fft_lines_smem(double* __restrict__ re, double* __restrict__ im, const double2* __restrict__ tw,
                                               const LD& ld = LD(), const ST& st = ST()) {
    if constexpr (N % 5 == 0 || N % 7 == 0) {
        constexpr int R = fft_odd_radix(N), M = N / R;
        constexpr int m = (M == 8) ? 3 : (M == 16) ? 4 : (M == 32) ? 5 : (M == 64) ? 6 : (M == 128) ? 7 : 8;
        static_assert((1 << m) == M, "N must be 5 * 2^m (40 .. 1280) or 7 * 2^m (56 .. 1792)");
        constexpr int n8 = m / 3, rem = m % 3;      // radix-8 passes over the power-of-two factor, one radix-4 / 2 pass, the odd radix last
        stockham_pass<N, 8, 1, IO, false, LD>(re, im, tw, ld);
        if constexpr (n8 >= 2) stockham_pass<N, 8, 8>(re, im, tw);
        constexpr int Ns = 1 << (n8 * 3);
        if constexpr (rem == 2) stockham_pass<N, 4, Ns>(re, im, tw);
        else if constexpr (rem == 1) stockham_pass<N, 2, Ns>(re, im, tw);
        stockham_pass<N, R, M, false, IO, LD, ST>(re, im, tw, ld, st);
    } else if constexpr (N % 3 == 0) {
        constexpr int M = N / 3;
        static_assert((M & (M - 1)) == 0 && M >= 8 && M <= 512, "N must be 3 * 2^m, 24 <= N <= 1536");
        pow2_passes_r4<N, 1, M, IO, LD>(re, im, tw, ld);
        stockham_pass<N, 3, M, false, IO, LD, ST>(re, im, tw, ld, st);
    } else {
        constexpr int m = (N == 8) ? 3 : (N == 16) ? 4 : (N == 32) ? 5 : (N == 64) ? 6 : (N == 128) ? 7 : (N == 256) ? 8 : (N == 512) ? 9 : (N == 1024) ? 10 : 11;
        static_assert((1 << m) == N, "N must be a power of two in [8, 2048]");
        constexpr int n8 = m / 3, rem = m % 3;      // radix-8 passes, then one radix-4 / radix-2 pass for the remainder
        stockham_pass<N, 8, 1, IO, IO && n8 == 1 && rem == 0, LD, ST>(re, im, tw, ld, st);
        if constexpr (n8 >= 2) stockham_pass<N, 8, 8, false, IO && n8 == 2 && rem == 0, LD, ST>(re, im, tw, ld, st);
        if constexpr (n8 >= 3) stockham_pass<N, 8, 64, false, IO && n8 == 3 && rem == 0, LD, ST>(re, im, tw, ld, st);
        constexpr int Ns = 1 << (n8 * 3);
        if constexpr (rem == 2) stockham_pass<N, 4, Ns, false, IO, LD, ST>(re, im, tw, ld, st);
        else if constexpr (rem == 1) stockham_pass<N, 2, Ns, false, IO, LD, ST>(re, im, tw, ld, st);
    }
}

// ---- source term ------------------------------------------------------------------------------------------------
// _compute_anelastic_source_term!: rhs = Δzᶜ · divᶜᶜᶜ(ρu, ρv, ρw) / Δt  (anelastic_pressure_solver.jl:99-105)
__device__ __forceinline__ double source_term(const Layout& L, const double* __restrict__ ru, const double* __restrict__ rv,
                                              const double* __restrict__ rw, int i, int j, int k, double dz_over_dt) {
    long long n = lidx(L, i, j, k);
    double d = 0.0;
    // reads the first ghost column of ρu and the first ghost row of ρv (filled before the solve). Addressing the periodic images here
    // instead was measured: the extra selects cost the 64-register y transform its loads in flight (1.35 -> 2.10 ms per launch at 512^3).
    if (!L.flat_x) d += (ru[n + 1] - ru[n]) * L.rdx;
    if (!L.flat_y) d += (rv[n + L.PX] - rv[n]) * L.rdy;
    double wt = (k + 1 < L.Nz) ? rw[n + L.plane] : 0.0;
    d += (wt - rw[n]) * L.rdz;
    return d * dz_over_dt;
}

// ---- pass 1: source term + real-to-complex FFT along y ----------------------------------------------------------
// grid (ceil(nx / (2*lines)), Nz); block lines*Ny/8 <= MAXT threads; smem 2*lines*line_pitch(Ny) doubles.
// MAXT / MINB: launch bounds. The default (256, 3) is the tuned configuration; (512, 2) exists for the wide-tile experiment
// (8 lines per CTA = 128-byte rows of the momentum fields; BZ_FFT_LINES_Y).
template <int N, int MAXT = 256, int MINB = 3>
__global__ void __launch_bounds__(MAXT, (fft_odd_radix(N) > 1 && MINB > 2) ? 2 : MINB) poisson_forward_y(Layout L, PoissonGeom G, const double* __restrict__ ru, const double* __restrict__ rv,
                                  const double* __restrict__ rw, double dz_over_dt, double2* __restrict__ W,
                                  const double2* __restrict__ tw_y, int lines, int k_base) {
    extern __shared__ double sm[];
    constexpr int LP = ((N + (N >> 4) + 15) & ~15) + 4;
    double* re = sm;
    double* im = sm + imag_offset(N, lines);
    const int k = blockIdx.y + k_base;               // z chunks: the distributed transform is pipelined over levels (api.cu poisson_solve)
    const int XB = 2 * lines;                        // lines is a power of two
    const int xb_shift = 31 - __clz(XB);
    const int ib = blockIdx.x * XB;
    // blockDim.x == lines * N / PT and XB == 2 * lines: every thread evaluates 2 PT source-term values, PT at a time so that
    // their loads are in flight together
    constexpr int PT = fft_pt(N);
    for (int half = 0; half < 2; ++half) {
        double v[PT];
#pragma unroll
        for (int it = 0; it < PT; ++it) {
            int e = threadIdx.x + (half * PT + it) * blockDim.x;
            int c = e & (XB - 1), y = e >> xb_shift;
            int i = ib + c;
            v[it] = (i < L.nx) ? source_term(L, ru, rv, rw, i, y, k, dz_over_dt) : 0.0;
        }
#pragma unroll
        for (int it = 0; it < PT; ++it) {
            int e = threadIdx.x + (half * PT + it) * blockDim.x;
            int c = e & (XB - 1), y = e >> xb_shift;
            ((c & 1) ? im : re)[(c >> 1) * LP + pidx(y)] = v[it];
        }
    }
    __syncthreads();
    fft_lines_smem<N>(re, im, tw_y);
    // untangle the two real transforms: A = (Z[ky] + conj Z[N-ky])/2, B = (Z[ky] - conj Z[N-ky])/(2i)
    for (int e = threadIdx.x; e < G.nky * XB; e += blockDim.x) {
        int c = e & (XB - 1), ky = e >> xb_shift;
        int i = ib + c;
        if (i >= L.nx) continue;
        int l = c >> 1, km = ky ? N - ky : 0;
        double zr = re[l * LP + pidx(ky)], zi = im[l * LP + pidx(ky)];
        double yr = re[l * LP + pidx(km)], yi = im[l * LP + pidx(km)];
        double2 o = (c & 1) ? make_double2(0.5 * (zi + yi), -0.5 * (zr - yr)) : make_double2(0.5 * (zr + yr), 0.5 * (zi - yi));
        W[w_index(G, k, ky, i)] = o;
    }
}

// Flat y: no y transform; W[k][0][i] = rhs + 0i.
__global__ void poisson_pack_flat_y(Layout L, PoissonGeom G, const double* __restrict__ ru, const double* __restrict__ rv,
                                    const double* __restrict__ rw, double dz_over_dt, double2* __restrict__ W) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (i < L.nx) W[w_index(G, k, 0, i)] = make_double2(source_term(L, ru, rv, rw, i, 0, k, dz_over_dt), 0.0);
}

// ---- pass 5: complex-to-real inverse FFT along y → φ (padded field interior) -------------------------------------
template <int N, int MAXT = 256, int MINB = 3>
__global__ void __launch_bounds__(MAXT, (fft_odd_radix(N) > 1 && MINB > 2) ? 2 : MINB) poisson_inverse_y(Layout L, PoissonGeom G, const double2* __restrict__ W, double* __restrict__ phi,
                                  const double2* __restrict__ tw_y, int lines, double scale, PeerBases peers, int pull, int k_base) {
    extern __shared__ double sm[];
    constexpr int LP = ((N + (N >> 4) + 15) & ~15) + 4;
    double* re = sm;
    double* im = sm + imag_offset(N, lines);
    const int k = blockIdx.y + k_base;
    const int XB = 2 * lines;                        // lines is a power of two
    const int xb_shift = 31 - __clz(XB);
    const int ib = blockIdx.x * XB;
    // rebuild the packed spectrum Z = A + iB (Hermitian halves), conjugated for the inverse-by-forward trick
    // (N/2 + 1) * lines <= (PT/2 + 1) * blockDim.x elements; loads are batched ahead of their use
    {
        constexpr int NIT = fft_pt(N) / 2 + 1;
        const int total = G.nky * lines;
        double2 A[NIT], B[NIT];
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int e = threadIdx.x + it * blockDim.x;
            A[it] = make_double2(0.0, 0.0); B[it] = make_double2(0.0, 0.0);
            if (e < total) {
                int l = e & (lines - 1), ky = e >> (xb_shift - 1);
                int i = ib + 2 * l;
                if (pull) {   // transposed spectrum read straight from its owner's W2 over NVLink (replaces the backward all-to-all)
                    const double2* src = reinterpret_cast<const double2*>(peers.base[__ldg(&G.ky_owner[ky])] + G.off_W2)
                                         + __ldg(&G.ky_base2[ky]) + (long long)k * __ldg(&G.ky_kstride[ky]);
                    if (i < L.nx) A[it] = __ldcv(src + i);
                    if (i + 1 < L.nx) B[it] = __ldcv(src + i + 1);
                } else {
                    if (i < L.nx) A[it] = W[w_index(G, k, ky, i)];
                    if (i + 1 < L.nx) B[it] = W[w_index(G, k, ky, i + 1)];
                }
            }
        }
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            int e = threadIdx.x + it * blockDim.x;
            if (e < total) {
                int l = e & (lines - 1), ky = e >> (xb_shift - 1);
                // Z[ky] = A + iB, Z[N-ky] = conj(A) + i conj(B); store conj(Z)
                re[l * LP + pidx(ky)] = A[it].x - B[it].y;
                im[l * LP + pidx(ky)] = -(A[it].y + B[it].x);
                if (ky > 0 && ky < N / 2) {
                    re[l * LP + pidx(N - ky)] = A[it].x + B[it].y;
                    im[l * LP + pidx(N - ky)] = -(B[it].x - A[it].y);
                }
            }
        }
    }
    __syncthreads();
    fft_lines_smem<N>(re, im, tw_y);
    for (int e = threadIdx.x; e < N * XB; e += blockDim.x) {
        int c = e & (XB - 1), y = e >> xb_shift;
        int i = ib + c;
        if (i >= L.nx) continue;
        int l = c >> 1;
        double v = (c & 1) ? -im[l * LP + pidx(y)] : re[l * LP + pidx(y)];
        phi[lidx(L, i, y, k)] = v * scale;
    }
}

__global__ void poisson_unpack_flat_y(Layout L, PoissonGeom G, const double2* __restrict__ W, double* __restrict__ phi, double scale) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (i < L.nx) phi[lidx(L, i, 0, k)] = W[w_index(G, k, 0, i)].x * scale;
}

// ---- passes 2 / 4: complex FFT along x, in place; one line = Nx contiguous complex numbers ----------------------
// grid ceil(n_lines / lines); block lines*Nx/8 <= 256 threads; smem 2*lines*line_pitch(Nx) doubles.
template <int N, int MAXT = 256, int MINB = 3>
__global__ void __launch_bounds__(MAXT, (fft_odd_radix(N) > 1 && MINB > 2) ? 2 : MINB) fft_x_kernel(PoissonGeom G, double2* __restrict__ W, long long n_lines, const double2* __restrict__ tw_x, int lines, int inverse,
                                                       PeerBases peers, int pull, int line_base) {
    extern __shared__ double sm[];
    double* re = sm;
    double* im = sm + imag_offset(N, lines);
    const int l0 = line_base + blockIdx.x * lines;      // first line of this CTA; line = k * nky_loc + ky_loc (fits an int); lines [line_base, n_lines)
    const double sgn = inverse ? -1.0 : 1.0;
    const int k0 = l0 / G.nky_loc, r0 = l0 - k0 * G.nky_loc;    // one division per CTA; lines advance (k, ky) incrementally
    constexpr int PT = fft_pt(N);
    auto w2_of = [&](int l, int x) -> size_t {          // offset of element x of the CTA's line l in the peer-blocked W2
        int ky = r0 + l, k = k0;
        while (ky >= G.nky_loc) { ky -= G.nky_loc; ++k; }
        const int p = x / G.nx;
        return (((size_t)p * G.Nz + k) * G.nky_loc + ky) * G.nx + (x - p * G.nx);
    };
    // The first pass reads its PT inputs per thread straight from global memory and the last pass writes its outputs straight back
    // (coalesced: for every r the threads of a line touch contiguous elements); lines beyond n_lines read zeros and are not written.
    const bool p1 = (G.P == 1);
    const double2* __restrict__ src1 = W + (size_t)l0 * N;
    double2* __restrict__ dst1 = W + (size_t)l0 * N;
    auto ld = [&](int l, int x) -> cpx {
        if (l0 + l >= n_lines) return cpx{0.0, 0.0};
        double2 v;
        if (p1) v = src1[l * N + x];                    // one rank: the CTA's lines are ONE contiguous chunk of W
        else if (pull) {   // x-slab spectrum read straight from the rank that owns these columns (replaces the forward all-to-all)
            int ky = r0 + l, k = k0;
            while (ky >= G.nky_loc) { ky -= G.nky_loc; ++k; }
            const int p = x / G.nx;
            const double2* src = reinterpret_cast<const double2*>(peers.base[p] + G.off_W)
                                 + ((long long)G.ky0 * G.Nz + (long long)k * G.nky_loc + ky) * G.nx + (x - p * G.nx);
            v = __ldcv(src);
        } else v = W[w2_of(l, x)];
        return cpx{v.x, sgn * v.y};
    };
    auto st = [&](int l, int x, cpx v) {
        if (l0 + l >= n_lines) return;
        const double2 o = make_double2(v.x, sgn * v.y);
        if (p1) dst1[l * N + x] = o; else W[w2_of(l, x)] = o;
    };
    fft_lines_smem<N, true>(re, im, tw_x, ld, st);
}

// ---- pass 3: batched Thomas in z -----------------------------------------------------------------------------------
// Setup (once per grid / ρᵣ): the forward-elimination pivots depend on (kx, ky, k) only.
//   D_k = -(ρ̄ᶠ_{k+1}/Δz + ρ̄ᶠ_k/Δz) - ρ_k Δz (λx + λy)   (one-sided at k = 0, Nz-1; anelastic_pressure_solver.jl:39-62)
//   a_k = ρ̄ᶠ_{k+1}/Δz                                      (:72-78)
//   t_k = a_{k-1}/β_{k-1},  β_k = D_k - a_{k-1} t_k,  and 1/β_k := 0 where |β_k| <= 10 eps (the reference's elision
//   of the update for the singular (0,0) mode; it then keeps a stale value, here 0 — any constant is removed with the mean).
__global__ void thomas_setup(PoissonGeom G, const double* __restrict__ rho, const double* __restrict__ rho_f, double dz,
                             const double* __restrict__ lam_x, const double* __restrict__ lam_y,
                             double* __restrict__ inv_beta, double* __restrict__ tfac) {
    int kx = blockIdx.x * blockDim.x + threadIdx.x, ky = blockIdx.y;
    if (kx >= G.Nx || ky >= G.nky_loc) return;
    const double lam = lam_x[kx] + lam_y[G.ky0 + ky];
    const size_t stride = (size_t)G.nky_loc * G.Nx;
    size_t n = (size_t)ky * G.Nx + kx;
    const int Nz = G.Nz;
    double beta = 0.0;
    for (int k = 0; k < Nz; ++k, n += stride) {
        double up = (k + 1 < Nz) ? rho_f[k + 1] / dz : 0.0;
        double lo = (k > 0) ? rho_f[k] / dz : 0.0;
        double D = -(up + lo) - rho[k] * dz * lam;
        double t = 0.0;
        if (k > 0) { t = lo / beta; beta = D - lo * t; } else beta = D;
        tfac[n] = t;
        inv_beta[n] = (fabs(beta) > 10.0 * BZ_REAL_EPS) ? 1.0 / beta : 0.0;
    }
}

// grid (ceil(Nx/128), nky_loc), block 128: thread = one (kx, ky) column, coalesced over kx. Levels are processed in
// batches of TB whose loads are all issued before the dependent recurrence, so each thread keeps TB loads in flight.
#define TB 8
__global__ void thomas_z(PoissonGeom G, double2* __restrict__ W, const double* __restrict__ rho_f, double dz,
                         const double* __restrict__ inv_beta, const double* __restrict__ tfac) {
    int kx = blockIdx.x * blockDim.x + threadIdx.x, ky = blockIdx.y;
    if (kx >= G.Nx) return;
    const size_t stride = (size_t)G.nky_loc * G.Nx;                    // level stride of the factor arrays (plain layout)
    const size_t n0 = (size_t)ky * G.Nx + kx;
    const size_t wstride = (size_t)G.nky_loc * G.nx;                   // level stride of W2 inside its peer block
    const size_t w0 = w2_index(G, 0, ky, kx);
    const int Nz = G.Nz;
    const double rdz = 1.0 / dz;
    double2 prev = make_double2(0.0, 0.0);
    for (int kb = 0; kb < Nz; kb += TB) {
        double2 f[TB]; double ib[TB];
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            int k = kb + b;
            if (k < Nz) { f[b] = W[w0 + k * wstride]; ib[b] = inv_beta[n0 + k * stride]; }
        }
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            int k = kb + b;
            if (k < Nz) {
                double a = (k > 0) ? rho_f[k] * rdz : 0.0;
                prev = make_double2((f[b].x - a * prev.x) * ib[b], (f[b].y - a * prev.y) * ib[b]);
                W[w0 + k * wstride] = prev;
            }
        }
    }
    // back substitution: φ_k -= t_{k+1} φ_{k+1}; prev holds φ_{Nz-1}
    for (int kt = Nz - 2; kt >= 0; kt -= TB) {
        double2 v[TB]; double t[TB];
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            int k = kt - b;
            if (k >= 0) { v[b] = W[w0 + k * wstride]; t[b] = tfac[n0 + (k + 1) * stride]; }
        }
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            int k = kt - b;
            if (k >= 0) {
                prev = make_double2(v[b].x - t[b] * prev.x, v[b].y - t[b] * prev.y);
                W[w0 + k * wstride] = prev;
            }
        }
    }
}

// φ .-= mean(φ): only the (kx, ky) = (0, 0) column carries the mean. One block.
__global__ void remove_mean_mode(PoissonGeom G, double2* __restrict__ W) {
    __shared__ double sr[256], si[256];
    const size_t stride = (size_t)G.nky_loc * G.nx;                    // (kx, ky) = (0, 0) lives in peer block 0
    double ar = 0.0, ai = 0.0;
    for (int k = threadIdx.x; k < G.Nz; k += blockDim.x) { double2 v = W[k * stride]; ar += v.x; ai += v.y; }
    sr[threadIdx.x] = ar; si[threadIdx.x] = ai;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { sr[threadIdx.x] += sr[threadIdx.x + s]; si[threadIdx.x] += si[threadIdx.x + s]; }
        __syncthreads();
    }
    double mr = sr[0] / G.Nz, mi = si[0] / G.Nz;
    for (int k = threadIdx.x; k < G.Nz; k += blockDim.x) { double2 v = W[k * stride]; W[k * stride] = make_double2(v.x - mr, v.y - mi); }
}

// Host-side dispatch on the line length (2^m, 3 · 2^m, 5 · 2^m, 7 · 2^m).
#define FFT_DISPATCH(n, CALL)                                                                                    \
    switch (n) {                                                                                                  \
        case 8: { constexpr int FN = 8; CALL; break; }       case 16: { constexpr int FN = 16; CALL; break; }     \
        case 32: { constexpr int FN = 32; CALL; break; }     case 64: { constexpr int FN = 64; CALL; break; }     \
        case 128: { constexpr int FN = 128; CALL; break; }   case 256: { constexpr int FN = 256; CALL; break; }   \
        case 512: { constexpr int FN = 512; CALL; break; }   case 1024: { constexpr int FN = 1024; CALL; break; } \
        case 2048: { constexpr int FN = 2048; CALL; break; }                                                      \
        case 24: { constexpr int FN = 24; CALL; break; }     case 48: { constexpr int FN = 48; CALL; break; }     \
        case 96: { constexpr int FN = 96; CALL; break; }     case 192: { constexpr int FN = 192; CALL; break; }   \
        case 384: { constexpr int FN = 384; CALL; break; }   case 768: { constexpr int FN = 768; CALL; break; }   \
        case 1536: { constexpr int FN = 1536; CALL; break; }                                                      \
        case 40: { constexpr int FN = 40; CALL; break; }     case 80: { constexpr int FN = 80; CALL; break; }     \
        case 160: { constexpr int FN = 160; CALL; break; }   case 320: { constexpr int FN = 320; CALL; break; }   \
        case 640: { constexpr int FN = 640; CALL; break; }   case 1280: { constexpr int FN = 1280; CALL; break; } \
        case 56: { constexpr int FN = 56; CALL; break; }     case 112: { constexpr int FN = 112; CALL; break; }   \
        case 224: { constexpr int FN = 224; CALL; break; }   case 448: { constexpr int FN = 448; CALL; break; }   \
        case 896: { constexpr int FN = 896; CALL; break; }   case 1792: { constexpr int FN = 1792; CALL; break; } \
        default: break;                                                                                           \
    }
