/* oracle_weno.h — WENO5-Z / WENO3-Z / upwind biased reconstructions and Centered(4) interpolation of the CPU oracle.
 *
 * TEST INFRASTRUCTURE (see breeze_oracle.c): shared by the anelastic oracle (breeze_oracle.c) and the compressible
 * oracle (breeze_oracle_compressible.c). Restates Oceananigans.Advection for WENO(order = 5). */
#ifndef ORACLE_WENO_H
#define ORACLE_WENO_H
#include <math.h>
#include <stddef.h>

/* ------------------------------------------------------------------------------------------------ */
/* WENO / centred reconstructions (Oceananigans.Advection; PARITY UNPINNED)  */
/* ------------------------------------------------------------------------------------------------ */
#define WENO_EPS 1e-8

/* Option 0 (default): smoothness indicators as the reference evaluates them (quadratic forms in the stencil values).
 * Option 1: the algebraically identical difference form β = 13/4 (δ²ψ)² + 3/4 (δ̃ψ)², free of the quadratic forms'
 * cancellation error (|ψ|² · eps instead of |δψ|² · eps). Used by the tests to separate "different arithmetic for the
 * same scheme" from genuine disagreement: the CUDA kernels use the difference form. */
extern int g_beta_form;     /* defined in breeze_oracle.c (orc_set_beta_form) */

/* Left-biased reconstruction at the face whose upwind cell is s[2] and downwind cell is s[3]:
 * s[0..4] = psi[i-3], psi[i-2], psi[i-1], psi[i], psi[i+1] for face i. Right bias = mirrored arguments. */
static inline double weno5_biased(double m3, double m2, double m1, double p0, double p1) {
    /* candidate polynomials (uniform grid) */
    double q0 = (2 * m1 + 5 * p0 - p1) / 6;
    double q1 = (-m2 + 5 * m1 + 2 * p0) / 6;
    double q2 = (2 * m3 - 7 * m2 + 11 * m1) / 6;
    /* smoothness indicators as quadratic forms, stencils (m1,p0,p1), (m2,m1,p0), (m3,m2,m1) */
    double b0, b1, b2;
    if (g_beta_form == 0) {
        b0 = m1 * (10 * m1 - 31 * p0 + 11 * p1) + p0 * (25 * p0 - 19 * p1) + 4 * p1 * p1;
        b1 = m2 * (4 * m2 - 13 * m1 + 5 * p0) + m1 * (13 * m1 - 13 * p0) + 4 * p0 * p0;
        b2 = m3 * (4 * m3 - 19 * m2 + 11 * m1) + m2 * (25 * m2 - 31 * m1) + 10 * m1 * m1;
    } else {
        double s0 = (m1 - 2 * p0) + p1, t0 = (3 * m1 - 4 * p0) + p1;
        double s1 = (m2 - 2 * m1) + p0, t1 = m2 - p0;
        double s2 = (m3 - 2 * m2) + m1, t2 = (m3 - 4 * m2) + 3 * m1;
        b0 = 3.25 * s0 * s0 + 0.75 * t0 * t0;
        b1 = 3.25 * s1 * s1 + 0.75 * t1 * t1;
        b2 = 3.25 * s2 * s2 + 0.75 * t2 * t2;
    }
    /* WENO-Z weights */
    double tau = fabs(b0 - b2);
    double r0 = tau / (b0 + WENO_EPS), r1 = tau / (b1 + WENO_EPS), r2 = tau / (b2 + WENO_EPS);
    double a0 = 0.3 * (1 + r0 * r0), a1 = 0.6 * (1 + r1 * r1), a2 = 0.1 * (1 + r2 * r2);
    return (a0 * q0 + a1 * q1 + a2 * q2) / (a0 + a1 + a2);
}

/* WENO3 (buffer 2): s = psi[i-2], psi[i-1], psi[i] for a left-biased face i. */
static inline double weno3_biased(double m2, double m1, double p0) {
    double q0 = 0.5 * m1 + 0.5 * p0;
    double q1 = -0.5 * m2 + 1.5 * m1;
    double d0 = p0 - m1, d1 = m1 - m2;
    double b0 = d0 * d0, b1 = d1 * d1;
    double tau = fabs(b0 - b1);
    double r0 = tau / (b0 + WENO_EPS), r1 = tau / (b1 + WENO_EPS);
    double a0 = (2.0 / 3.0) * (1 + r0 * r0), a1 = (1.0 / 3.0) * (1 + r1 * r1);
    return (a0 * q0 + a1 * q1) / (a0 + a1);
}

/* ---- WENO(order = 7 / 9) (buffers 4, 5): SURVEY §8f rank 4 (the reference's shipped examples use WENO(order = 9)) ------------------
 * Candidates, optimal weights and smoothness forms are DERIVED from their definitions in exact arithmetic
 * (scripts/derive_weno_coefficients.py → oracle_weno_tables.h; the same derivation reproduces the order-5 constants above).
 * Recalled from upstream Oceananigans, not verifiable here (PARITY UNPINNED like the rest of this header):
 *   - the WENO-Z global indicator  tau_7 = |b0 + 3 b1 - 3 b2 - b3|,  tau_9 = |b0 + 2 b1 - 6 b2 + 2 b3 + b4|  (Castro et al. 2011);
 *   - the scaling of the stored forms: 240 x JS / 1000 (order 7) and 5040 x JS / 100000 (order 9) — it only matters relative to eps;
 *   - exponent 2 and eps = 1e-8 as for order 5. */
#include "oracle_weno_tables.h"
#define WENO7_BSCALE 0.24
#define WENO9_BSCALE 0.0504

/* w[0 .. 2R-2]: the window of the biased reconstruction, upwind cell at w[R-1], downwind side at w[R .. 2R-2]. */
static inline double weno_hi_window(const double* w, int R) {
    const double* C = (R == 4) ? &WENO7_C[0][0] : &WENO9_C[0][0];
    const double* D = (R == 4) ? WENO7_D : WENO9_D;
    const double* B = (R == 4) ? &WENO7_B[0][0][0] : &WENO9_B[0][0][0];
    const double* M = (R == 4) ? &WENO7_M[0][0][0] : &WENO9_M[0][0][0];   /* the same forms in the stencil's first differences */
    const double bscale = (R == 4) ? WENO7_BSCALE : WENO9_BSCALE;
    static const double G4[4] = {1, 3, -3, -1}, G5[5] = {1, 2, -6, 2, 1};
    const double* G = (R == 4) ? G4 : G5;
    double p[5], beta[5], tau = 0;
    for (int st = 0; st < R; ++st) {
        const double* v = w + (R - 1 - st);            /* stencil st covers w[R-1-st .. 2R-2-st] */
        double q = 0, b = 0;
        for (int a = 0; a < R; ++a) q += C[st * R + a] * v[a];
        if (g_beta_form == 0) {                        /* quadratic forms in the stencil VALUES (how the reference stores them) */
            for (int a = 0; a < R; ++a) {
                double row = 0;
                for (int c = a; c < R; ++c) row += B[(st * R + a) * R + c] * v[c];
                b += v[a] * row;
            }
        } else {                                       /* algebraically identical, in the first differences: no |psi|^2 cancellation; the CUDA form */
            double d[4];
            for (int a = 0; a < R - 1; ++a) d[a] = v[a + 1] - v[a];
            for (int a = 0; a < R - 1; ++a) {
                double row = 0;
                for (int c = a; c < R - 1; ++c) row += M[(st * (R - 1) + a) * (R - 1) + c] * d[c];
                b += d[a] * row;
            }
        }
        p[st] = q; beta[st] = bscale * b;
        tau += G[st] * beta[st];
    }
    tau = fabs(tau);
    double num = 0, den = 0;
    for (int st = 0; st < R; ++st) {
        double rr = tau / (beta[st] + WENO_EPS);
        double al = D[st] * (1 + rr * rr);
        num += al * p[st]; den += al;
    }
    return num / den;
}

/* kept out of line so that the order-5 call sites (the CPU baseline that bench.py times) stay as small as they were */
static __attribute__((noinline)) double weno_hi_biased(const double* psi, ptrdiff_t s, int R, int left) {
    double w[9];
    for (int j = 0; j < 2 * R - 1; ++j) w[j] = left ? psi[(j - R) * s] : psi[(R - 1 - j) * s];
    return weno_hi_window(w, R);
}

/* Biased interpolation of psi (stride s) to "face" i, i.e. between psi[i-1] and psi[i]; R = buffer in use
 * (5: WENO9, 4: WENO7, 3: WENO5, 2: WENO3, 1: first-order upwind); left != 0 selects the left (upwind = i-1) bias. */
static inline double biased_interp(const double* psi, ptrdiff_t s, int R, int left) {
    if (__builtin_expect(R >= 4, 0)) return weno_hi_biased(psi, s, R > 5 ? 5 : R, left);
    if (left) {
        if (R >= 3) return weno5_biased(psi[-3 * s], psi[-2 * s], psi[-s], psi[0], psi[s]);
        if (R == 2) return weno3_biased(psi[-2 * s], psi[-s], psi[0]);
        return psi[-s];
    } else {
        if (R >= 3) return weno5_biased(psi[2 * s], psi[s], psi[0], psi[-s], psi[-2 * s]);
        if (R == 2) return weno3_biased(psi[s], psi[0], psi[-s]);
        return psi[0];
    }
}

static __attribute__((noinline)) double centered_hi(const double* a, ptrdiff_t s, int R) {
    double v = 0;
    if (R >= 4) { for (int j = 0; j < 4; ++j) v += CENTERED8_C[3 - j] * (a[(-1 - j) * s] + a[j * s]); return v; }
    for (int j = 0; j < 3; ++j) v += CENTERED6_C[2 - j] * (a[(-1 - j) * s] + a[j * s]);
    return v;
}

/* Centered(order = 2R) symmetric interpolation to "face" i (between a[i-1], a[i]); R = 4: 8th, 3: 6th, 2: 4th order, 1: 2nd.
 * (WENO(order = n) advects with Centered(order = n - 1), recalled from upstream: buffer R_sym = R_weno - 1.) */
static inline double symmetric_interp(const double* a, ptrdiff_t s, int R) {
    if (__builtin_expect(R >= 3, 0)) return centered_hi(a, s, R);
    if (R >= 2) return (7 * (a[-s] + a[0]) - (a[-2 * s] + a[s])) / 12;
    return 0.5 * (a[-s] + a[0]);
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* Order reduction next to the Bounded z walls (Appendix A.3): the largest buffer whose left- and
 * right-biased stencils both stay inside the domain. Face k (0..Nz) from centres 0..Nz-1: */
static inline int red_face(int k, int Nz, int B) { return imax(1, imin(B, imin(k, Nz - k))); }
/* centre k (0..Nz-1) from faces 0..Nz: */
static inline int red_center(int k, int Nz, int B) { return imax(1, imin(B, imin(k + 1, Nz - k))); }

#endif /* ORACLE_WENO_H */
