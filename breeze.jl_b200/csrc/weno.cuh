// weno.cuh — WENO5-Z / WENO3-Z / upwind biased reconstructions and Centered(4) interpolation, FP64.
//
// Same scheme the reference obtains from Oceananigans' WENO(order = 5) (oracle:
// oracle/breeze_oracle.c weno5_biased / weno3_biased / symmetric_interp / red_face / red_center), written for the
// FP64 pipe of sm_100a, which — not HBM — bounds the fused stage kernel:
//   * smoothness indicators in difference form  β = 13/4 (δ²ψ)² + 3/4 (δ̃ψ)²  (= 3 × Jiang–Shu, the reference's
//     scaling, so ε = 1e-8 means the same thing). Algebraically identical to the reference's quadratic forms but
//     21 instead of 30 FP64 instructions and free of their cancellation error;
//   * the three weight divisions and the normalising division collapse into ONE reciprocal:
//       Σ α_r q_r / Σ α_r,  α_r = C_r (1 + τ²/b_r²),  b_r = β_r + ε
//       = Σ C_r (b_r² + τ²) Π_{s≠r} b_s² q_r / Σ C_r (b_r² + τ²) Π_{s≠r} b_s²
//   * that reciprocal is MUFU.RCP64H + two Newton steps (≤ 2 ulp) instead of the IEEE division sequence.
// Results differ from the oracle's by FP64 round-off only (tests/test_gpu_parity.py states the tolerance).
#pragma once

#define WENO_EPS 1e-8

// (BZ_F32: the Float32 library is compiled from a retyped copy of these sources, breeze.jl_b200/make_f32.py; the few places where the
// two precisions need different instructions — bit tests, the reciprocal, the range of the WENO weight products — sit under #ifdef BZ_F32.)
#ifdef BZ_F32
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#else
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#ifdef BZ_RCP_TWO_NEWTON
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
#else
    // one cubically convergent step: r (1 + e + e²), e = 1 - x r  (MUFU.RCP64H is good to ~2⁻²³ ⇒ ≤ 1 ulp after it)
    double e = fma(-x, r, 1.0);
    e = fma(e, e, e);
    r = fma(r, e, r);
#endif
    return r;
}
#endif

// sign test on the high word (integer pipe instead of DSETP on the FP64 pipe); differs from `x > 0` only for
// x = ±0, where the flux it selects for is multiplied by that zero
#ifdef BZ_F32
__device__ __forceinline__ bool positive(float x) { return __float_as_int(x) >= 0; }
#else
__device__ __forceinline__ bool positive(double x) { return __double2hiint(x) >= 0; }
#endif

// Left-biased value at the face between c and d from the five cells a b c | d e (a = ψ[i-3] … e = ψ[i+1]).
// Written with explicit fma() for the minimal FP64 instruction count (43 + one reciprocal):
//   β' = β / 0.75 = (13/3) s² + t²  (ε scaled alike: the weights only see the ratios τ/(β+ε));
//   Σ ω_r q_r = q₁ + ω₀ (q₀ - q₁) + ω₂ (q₂ - q₁)  with  q₀ - q₁ = (s₁ - s₀)/6,  q₂ - q₁ = (s₂ - s₁)/3
// (s_r are the second differences already formed for β), so only the central candidate polynomial is evaluated.
// (An FP32-weights variant of this function — indicators, τ and α_r on the FMA pipe, polynomials in FP64 — was measured and rejected:
// 17.45 against 11.28 ms per stage-kernel launch at 512^3, because an FP64 <-> FP32 conversion issues at 8.5 cycles per warp instruction on
// this part; profiles/r2_stage_fp32_weights_variant.txt, commit b2d200f carries the source.)
#ifdef BZ_F32
// Float32: the product form below would underflow (b_r ≥ 1.3e-8 ⇒ Π b_r² ≈ 1e-48 < FLT_MIN), so the weights are formed from the
// ratios τ / b_r with the single-instruction FP32 reciprocal, which costs one MUFU each.
__device__ __forceinline__ double weno5z(double a, double b, double c, double d, double e) {
    const double K = 13.0 / 3.0, EPSP = WENO_EPS / 0.75;
    double s0 = fma(-2.0, d, c) + e, t0 = fma(3.0, c, fma(-4.0, d, e));
    double s1 = fma(-2.0, c, b) + d, t1 = b - d;
    double s2 = fma(-2.0, b, a) + c, t2 = fma(3.0, c, fma(-4.0, b, a));
    double b0 = fma(s0 * K, s0, fma(t0, t0, EPSP));
    double b1 = fma(s1 * K, s1, fma(t1, t1, EPSP));
    double b2 = fma(s2 * K, s2, fma(t2, t2, EPSP));
    double tau = b0 - b2;
    double r0 = tau * fast_rcp(b0), r1 = tau * fast_rcp(b1), r2 = tau * fast_rcp(b2);
    double a0 = 0.3 * fma(r0, r0, 1.0), a1 = 0.6 * fma(r1, r1, 1.0), a2 = 0.1 * fma(r2, r2, 1.0);
    double rs = fast_rcp(a0 + a1 + a2);
    double qc = fma(-1.0 / 6.0, b, fma(5.0 / 6.0, c, (1.0 / 3.0) * d));
    return fma((a0 * rs) * (1.0 / 6.0), s1 - s0, fma((a2 * rs) * (1.0 / 3.0), s2 - s1, qc));
}
#else
__device__ __forceinline__ double weno5z(double a, double b, double c, double d, double e) {
    const double K = 13.0 / 3.0, EPSP = WENO_EPS / 0.75;
    double s0 = fma(-2.0, d, c) + e, t0 = fma(3.0, c, fma(-4.0, d, e));   // stencil (c, d, e)
    double s1 = fma(-2.0, c, b) + d, t1 = b - d;                          // stencil (b, c, d)
    double s2 = fma(-2.0, b, a) + c, t2 = fma(3.0, c, fma(-4.0, b, a));   // stencil (a, b, c)
    // b_r = β'_r + ε' in one chain; τ = |b0 - b2| only enters squared
    double b0 = fma(s0 * K, s0, fma(t0, t0, EPSP));
    double b1 = fma(s1 * K, s1, fma(t1, t1, EPSP));
    double b2 = fma(s2 * K, s2, fma(t2, t2, EPSP));
    double tau = b0 - b2;
    double tt = tau * tau;
    double q0 = b0 * b0, q1 = b1 * b1, q2 = b2 * b2;
    // un-normalised weights × Π b_s² (the common factor cancels in the ratio):  (q_r + τ²) Π_{s≠r} q_s = Π q + τ² Π_{s≠r} q_s
    double m12 = q1 * q2, m02 = q0 * q2, m01 = q0 * q1;
    double pq = q0 * m12;
    double w0 = fma(tt, m12, pq);
    double w1 = fma(tt, m02, pq);
    double w2 = fma(tt, m01, pq);
    double qc = fma(-1.0 / 6.0, b, fma(5.0 / 6.0, c, (1.0 / 3.0) * d));   // central candidate (stencil b, c, d)
    double num = fma(0.5 * w0, s1 - s0, ((1.0 / 3.0) * w2) * (s2 - s1));  // 3 w0 (q0 - qc) + w2 (q2 - qc)
    double den = fma(3.0, w0, fma(6.0, w1, w2));
    return fma(num, fast_rcp(den), qc);
}
#endif

#ifdef BZ_F32_PACKED
// Packed Float32 pair: TWO independent left-biased WENO5-Z reconstructions of one thread evaluated in f32x2 registers (Blackwell's
// add / mul / fma.rn.f32x2: two results per lane and instruction at the issue cost of one — scripts/ubench/fp32_rate.cu measured 2.11 cycles
// per FFMA2 warp instruction against 1.38 per FFMA). The Float32 stage kernel is issue-bound, so halving the arithmetic issue slots of a
// pair is the lever; the reciprocals stay scalar (MUFU). Same ratio form as the scalar Float32 weno5z above.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float x, float y) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
__device__ __forceinline__ void upk2(f32x2_t v, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) { f32x2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) { f32x2_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) { f32x2_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2_t rcp2(f32x2_t a) { float x, y; upk2(a, x, y); return pk2(fast_rcp(x), fast_rcp(y)); }
__device__ __forceinline__ f32x2_t cst2(float c) { return pk2(c, c); }
// (a, b, c, d, e) of reconstruction A in the .x halves, of reconstruction B in the .y halves; returns (value A, value B)
__device__ __forceinline__ f32x2_t weno5z_x2(f32x2_t a, f32x2_t b, f32x2_t c, f32x2_t d, f32x2_t e) {
    const f32x2_t M2 = cst2(-2.0f), M4 = cst2(-4.0f), P3 = cst2(3.0f), K = cst2(13.0f / 3.0f), EPSP = cst2((float)(WENO_EPS / 0.75)), ONE = cst2(1.0f);
    f32x2_t s0 = add2(fma2(M2, d, c), e), t0 = fma2(P3, c, fma2(M4, d, e));
    f32x2_t s1 = add2(fma2(M2, c, b), d), t1 = sub2(b, d);
    f32x2_t s2 = add2(fma2(M2, b, a), c), t2 = fma2(P3, c, fma2(M4, b, a));
    f32x2_t b0 = fma2(mul2(s0, K), s0, fma2(t0, t0, EPSP));
    f32x2_t b1 = fma2(mul2(s1, K), s1, fma2(t1, t1, EPSP));
    f32x2_t b2 = fma2(mul2(s2, K), s2, fma2(t2, t2, EPSP));
    f32x2_t tau = sub2(b0, b2);
    f32x2_t r0 = mul2(tau, rcp2(b0)), r1 = mul2(tau, rcp2(b1)), r2 = mul2(tau, rcp2(b2));
    f32x2_t a0 = mul2(cst2(0.3f), fma2(r0, r0, ONE)), a1 = mul2(cst2(0.6f), fma2(r1, r1, ONE)), a2 = mul2(cst2(0.1f), fma2(r2, r2, ONE));
    f32x2_t rs = rcp2(add2(add2(a0, a1), a2));
    f32x2_t qc = fma2(cst2(-1.0f / 6.0f), b, fma2(cst2(5.0f / 6.0f), c, mul2(cst2(1.0f / 3.0f), d)));
    return fma2(mul2(mul2(a0, rs), cst2(1.0f / 6.0f)), sub2(s1, s0), fma2(mul2(mul2(a2, rs), cst2(1.0f / 3.0f)), sub2(s2, s1), qc));
}
#endif

// WENO3-Z: left-biased value at the face between b and c from a b | c.
#ifdef BZ_F32
__device__ __forceinline__ double weno3z(double a, double b, double c) {      // ratio form (range, as for weno5z)
    double d0 = c - b, d1 = b - a;
    double b0 = d0 * d0, b1 = d1 * d1;
    double tau = fabs(b0 - b1);
    double r0 = tau * fast_rcp(b0 + WENO_EPS), r1 = tau * fast_rcp(b1 + WENO_EPS);
    double a0 = (2.0 / 3.0) * fma(r0, r0, 1.0), a1 = (1.0 / 3.0) * fma(r1, r1, 1.0);
    return (a0 * (0.5 * (b + c)) + a1 * (1.5 * b - 0.5 * a)) * fast_rcp(a0 + a1);
}
#else
__device__ __forceinline__ double weno3z(double a, double b, double c) {
    double d0 = c - b, d1 = b - a;
    double b0 = d0 * d0, b1 = d1 * d1;
    double tau = fabs(b0 - b1);
    double tt = tau * tau;
    b0 += WENO_EPS; b1 += WENO_EPS;
    double q0 = b0 * b0, q1 = b1 * b1;
    double w0 = 2.0 * ((q0 + tt) * q1);      // C = (2, 1)/3
    double w1 = (q1 + tt) * q0;
    double p0 = b + c;                       // × 2
    double p1 = 3.0 * b - a;
    return (w0 * p0 + w1 * p1) * fast_rcp(2.0 * (w0 + w1));
}
#endif

// Six consecutive values v0..v5 = ψ[i-3..i+2] around the "face" between v2 and v3; R = buffer (3, 2 or 1).
__device__ __forceinline__ double biased6(double v0, double v1, double v2, double v3, double v4, double v5, int R, bool left) {
    if (R >= 3) {
        double a = left ? v0 : v5, b = left ? v1 : v4, c = left ? v2 : v3, d = left ? v3 : v2, e = left ? v4 : v1;
        return weno5z(a, b, c, d, e);
    } else if (R == 2) {
        double a = left ? v1 : v4, b = left ? v2 : v3, c = left ? v3 : v2;
        return weno3z(a, b, c);
    }
    return left ? v2 : v3;
}

// Centered(order = 4) interpolation to the face between v1 and v2 from v0 v1 | v2 v3; R = 2 (4th) or 1 (2nd order).
__device__ __forceinline__ double sym4(double v0, double v1, double v2, double v3, int R) {
    if (R >= 2) return ((7.0 / 12.0) * (v1 + v2)) - ((1.0 / 12.0) * (v0 + v3));
    return 0.5 * (v1 + v2);
}

// ---- WENO(order = 7 / 9)-Z and Centered(6 / 8) -----------------------------------------------------------
// Table-driven restatement of oracle/oracle_weno.h (weno_hi_window / centered_hi) for the kernels that read their stencils through
// L1 / L2 (stage_hi.cuh, compressible.cuh): ≈ 170 FP64 per order-9 reconstruction. Smoothness indicators as quadratic forms in the
// FIRST DIFFERENCES of the stencil (no |ψ|² cancellation; the oracle's beta form 1); every division is MUFU.RCP64H + one cubic step.
#include "weno_tables.cuh"
// w[0 .. 2R-2]: window of the biased reconstruction, upwind cell at w[R-1] (mirror the window for the right bias).
template <int R>
__device__ __forceinline__ double weno_hi(const double (&w)[2 * R - 1]) {
    using T = WenoTab<R>;
    double p[R], beta[R], tau = 0.0;
    double dw[2 * R - 2];                                  // first differences of the window, shared by the R stencils
#pragma unroll
    for (int j = 0; j < 2 * R - 2; ++j) dw[j] = w[j + 1] - w[j];
#pragma unroll
    for (int st = 0; st < R; ++st) {
        double q = 0.0, b = 0.0;
#pragma unroll
        for (int a = 0; a < R; ++a) q = fma(T::C(st, a), w[R - 1 - st + a], q);
#pragma unroll
        for (int a = 0; a < R - 1; ++a) {
            double row = 0.0;
#pragma unroll
            for (int c = a; c < R - 1; ++c) row = fma(T::M(st, a, c), dw[R - 1 - st + c], row);
            b = fma(dw[R - 1 - st + a], row, b);
        }
        p[st] = q;
        beta[st] = T::BS() * b;
        tau = fma(T::G(st), beta[st], tau);
    }
    tau = fabs(tau);
    double num = 0.0, den = 0.0;
#pragma unroll
    for (int st = 0; st < R; ++st) {
        const double rr = tau * fast_rcp(beta[st] + WENO_EPS);
        const double al = T::D(st) * fma(rr, rr, 1.0);
        num = fma(al, p[st], num);
        den += al;
    }
    return num * fast_rcp(den);
}
// biased value at the face between f[n - s] and f[n] from a field in memory (stride s), buffer R = 4 or 5. Kept out of line: one copy of
// the ≈ 250-instruction body per order instead of one per call site (25 sites in stage_hi_kernel); the call costs ≈ 1 % of the body.
#ifdef BZ_WENO_HI_INLINE
#define BZ_HI_LINKAGE __forceinline__
#else
#define BZ_HI_LINKAGE __noinline__
#endif
template <int R>
__device__ BZ_HI_LINKAGE double weno_hi_mem(const double* __restrict__ f, long long n, long long s, bool left) {
    // window w[j] = left ? f[n + (j - R) s] : f[n + (R - 1 - j) s]: ONE selected start pointer and a signed step, 2R - 1 loads — instead of
    // 2R loads at both windows' addresses (a 64-bit multiply-add each) and 4R - 2 selecting moves
    const double* p = f + n + (left ? -(long long)R * s : (long long)(R - 1) * s);
    const long long step = left ? s : -s;
    double w[2 * R - 1];
#pragma unroll
    for (int j = 0; j < 2 * R - 1; ++j) { w[j] = *p; p += step; }
    return weno_hi<R>(w);
}
// Centered(order = 2R) value at the face between a[n - s] and a[n], R = 3 or 4
template <int R>
__device__ __forceinline__ double centered_hi_mem(const double* __restrict__ a, long long n, long long s) {
    const double* lo = a + n - s;
    const double* hi = a + n;
    double v = 0.0;
#pragma unroll
    for (int j = 0; j < R; ++j) { v = fma(R == 4 ? CENTERED8_C[j] : CENTERED6_C[j], *lo + *hi, v); lo -= s; hi += s; }
    return v;
}

// Order reduction next to the Bounded z walls: z-face k (0..Nz) from centres, and centre k (0..Nz-1) from z-faces.
__device__ __forceinline__ int red_face(int k, int Nz, int B) { return max(1, min(B, min(k, Nz - k))); }
__device__ __forceinline__ int red_center(int k, int Nz, int B) { return max(1, min(B, min(k + 1, Nz - k))); }
