"""Exact derivation of the finite-volume WENO coefficients of buffer r = 2 .. 6 (order 2r - 1) on a uniform grid, from their
definitions (Jiang & Shu 1996; Balsara & Shu 2000), in rational arithmetic:

  * candidate reconstructions  p_s(x_{i+1/2}) = sum_j c[s][j] * psi[i - s + j],  s = 0 .. r-1  (s = left shift of the stencil;
    left-biased: the face is the RIGHT face of cell i; stencil s covers cells i-s .. i-s+r-1);
  * optimal (linear) weights d[s] with  sum_s d[s] p_s = the (2r-1)-th order reconstruction on the union stencil;
  * smoothness indicators  beta_s = sum_{l=1}^{r-1} dx^(2l-1) * int_{cell i} (d^l p_s / dx^l)^2 dx  as quadratic forms
    beta_s = sum_{a<=b} B[s][a][b] psi_a psi_b  (upper-triangular storage, psi_a = psi[i - s + a]).

    python scripts/derive_weno_coefficients.py            # prints the tables and checks r = 3 against the oracle's WENO5 constants
    python scripts/derive_weno_coefficients.py --header   # writes oracle/oracle_weno_tables.h (r = 4, 5: WENO7, WENO9)

Test infrastructure (feeds the CPU oracle); not part of the product path."""
import sys
from fractions import Fraction as F


def poly_mul(a, b):
    out = [F(0)] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] += x * y
    return out


def poly_der(a):
    return [a[k] * k for k in range(1, len(a))] or [F(0)]


def poly_int(a, lo, hi):
    return sum(c * (F(hi) ** (k + 1) - F(lo) ** (k + 1)) / (k + 1) for k, c in enumerate(a))


def solve(A, b):
    """Gaussian elimination in exact arithmetic; A is n x n (list of rows), b a list of right-hand sides (n x m)."""
    n = len(A)
    M = [list(A[i]) + list(b[i]) for i in range(n)]
    for c in range(n):
        p = next(r for r in range(c, n) if M[r][c] != 0)
        M[c], M[p] = M[p], M[c]
        piv = M[c][c]
        M[c] = [x / piv for x in M[c]]
        for r in range(n):
            if r != c and M[r][c] != 0:
                f = M[r][c]
                M[r] = [x - f * y for x, y in zip(M[r], M[c])]
    return [row[n:] for row in M]


def stencil_polynomials(cells):
    """For cells (integer offsets j: cell j = [j - 1/2, j + 1/2], dx = 1) return, for each cell a, the coefficients of the
    polynomial of degree len(cells) - 1 whose cell averages are delta_{a,b} — i.e. p(x) = sum_a psi_a * basis[a](x)."""
    n = len(cells)
    # cell average of x^k over cell j
    A = [[(F(2 * j + 1, 2) ** (k + 1) - F(2 * j - 1, 2) ** (k + 1)) / (k + 1) for k in range(n)] for j in cells]
    I = [[F(1) if a == b else F(0) for b in range(n)] for a in range(n)]
    X = solve(A, I)                      # X[k][a] = coefficient of x^k in basis[a]
    return [[X[k][a] for k in range(n)] for a in range(n)]


def derive(r):
    c, B = [], []
    for s in range(r):
        cells = [j - s for j in range(r)]                 # offsets relative to cell i (= 0)
        basis = stencil_polynomials(cells)
        # value at the right face x = 1/2
        c.append([sum(co * F(1, 2) ** k for k, co in enumerate(basis[a])) for a in range(r)])
        # smoothness indicator as a quadratic form
        Bs = [[F(0)] * r for _ in range(r)]
        for a in range(r):
            for b in range(a, r):
                tot = F(0)
                da, db = basis[a], basis[b]
                for l in range(1, r):
                    da, db = poly_der(da), poly_der(db)
                    tot += poly_int(poly_mul(da, db), F(-1, 2), F(1, 2))     # dx = 1
                Bs[a][b] = tot * (1 if a == b else 2)
        B.append(Bs)
    # optimal weights: match the (2r-1)-point reconstruction on cells -(r-1) .. r-1
    big = stencil_polynomials([j - (r - 1) for j in range(2 * r - 1)])
    full = [sum(co * F(1, 2) ** k for k, co in enumerate(big[a])) for a in range(2 * r - 1)]   # coefficient of psi[i - (r-1) + a]
    d = [None] * r
    # stencil s covers cells -s .. -s+r-1; the cell -(r-1) only appears in stencil r-1, then peel off
    resid = list(full)
    for s in range(r - 1, -1, -1):
        lead = (r - 1) - s                                 # index in `full` of this stencil's leftmost cell
        d[s] = resid[lead] / c[s][0]
        for a in range(r):
            resid[lead + a] -= d[s] * c[s][a]
    assert all(x == 0 for x in resid), resid
    assert sum(d) == 1
    return c, d, B


def difference_forms(B):
    """The same smoothness indicators as quadratic forms in the FIRST DIFFERENCES d_a = psi_{a+1} - psi_a of the stencil
    (beta of a constant vanishes, so beta depends on the differences only): beta = sum_{a<=b} M[a][b] d_a d_b.
    Free of the cancellation of the value forms (|psi|^2 eps -> |d psi|^2 eps): what the CUDA kernels evaluate."""
    out = []
    for Bs in B:
        r = len(Bs)
        sym = [[(Bs[min(a, b)][max(a, b)] / (1 if a == b else 2)) for b in range(r)] for a in range(r)]
        assert all(sum(row) == 0 for row in sym)                      # B * 1 = 0
        # psi_j = psi_0 + sum_{a<j} d_a  ->  L[j][a] = 1 if a < j
        M = [[sum(sym[i][j] * (1 if a < i else 0) * (1 if b < j else 0) for i in range(r) for j in range(r)) for b in range(r - 1)] for a in range(r - 1)]
        out.append([[M[a][b] * (1 if a == b else 2) if b >= a else F(0) for b in range(r - 1)] for a in range(r - 1)])
    return out


def centered(m):
    """Centered(order = 2m) finite-volume reconstruction at the face between cells -1 and 0 from cells -m .. m-1."""
    cells = list(range(-m, m))
    basis = stencil_polynomials(cells)
    return [sum(co * F(-1, 2) ** k for k, co in enumerate(basis[a])) for a in range(2 * m)]      # face at x = -1/2


def common_scale(B):
    """Smallest positive rational that makes every entry of the quadratic forms an integer (the reference's r = 3 tables are the
    Jiang-Shu forms times 12 / 4 = 3; see check below)."""
    from math import gcd
    dens = [x.denominator for Bs in B for row in Bs for x in row if x != 0]
    l = 1
    for q in dens:
        l = l * q // gcd(l, q)
    return l


def fmt(x):
    return f"{x.numerator}.0 / {x.denominator}.0" if x.denominator != 1 else f"{x.numerator}.0"


def main():
    if "--header" not in sys.argv:
        for r in range(2, 6):
            c, d, B = derive(r)
            print(f"--- buffer r = {r} (order {2 * r - 1}) ---")
            print(" candidates:", [[str(x) for x in row] for row in c])
            print(" optimal weights:", [str(x) for x in d])
            L = common_scale(B)
            print(f" smoothness forms x {L}:", [[[str(x * L) for x in row] for row in Bs] for Bs in B])
        # r = 3 against the constants of oracle/oracle_weno.h: candidates (2,5,-1)/6, (-1,5,2)/6, (2,-7,11)/6,
        # weights (3/10, 3/5, 1/10), forms = 3 x Jiang-Shu: (10,-31,11,25,-19,4), (4,-13,5,13,-13,4), (4,-19,11,25,-31,10)
        c, d, B = derive(3)
        assert c == [[F(1, 3), F(5, 6), F(-1, 6)], [F(-1, 6), F(5, 6), F(1, 3)], [F(1, 3), F(-7, 6), F(11, 6)]], c
        assert d == [F(3, 10), F(3, 5), F(1, 10)], d
        flat = [[3 * Bs[0][0], 3 * Bs[0][1], 3 * Bs[0][2], 3 * Bs[1][1], 3 * Bs[1][2], 3 * Bs[2][2]] for Bs in B]
        assert flat == [[10, -31, 11, 25, -19, 4], [4, -13, 5, 13, -13, 4], [4, -19, 11, 25, -31, 10]], flat
        print("r = 3 reproduces the oracle's WENO5 constants (forms = 3 x Jiang-Shu)")
        import random
        random.seed(1)
        for r in (3, 4, 5):
            _, _, B = derive(r)
            M = difference_forms(B)
            v = [F(random.randint(-50, 50), random.randint(1, 9)) for _ in range(r)]
            d1 = [v[a + 1] - v[a] for a in range(r - 1)]
            for st in range(r):
                bv = sum(B[st][a][b] * v[a] * v[b] for a in range(r) for b in range(a, r))
                bd = sum(M[st][a][b] * d1[a] * d1[b] for a in range(r - 1) for b in range(a, r - 1))
                assert bv == bd, (r, st)
        print("difference forms == value forms (exact, random rational data)")
        for m in range(1, 5):
            print(f" Centered({2 * m}):", [str(x) for x in centered(m)])
        assert centered(2) == [F(-1, 12), F(7, 12), F(7, 12), F(-1, 12)]
        return
    if "--cuda" in sys.argv:
        return write_cuda_tables()
    out = ["/* oracle_weno_tables.h — GENERATED by scripts/derive_weno_coefficients.py --header (exact rational derivation from the",
           " * definitions; do not edit). Finite-volume WENO of buffer r = 4 (order 7) and r = 5 (order 9), uniform grid, left-biased at the",
           " * face between psi[-1] and psi[0]: stencil s covers psi[-1 - s + a], a = 0 .. r-1. TEST INFRASTRUCTURE (CPU oracle). */",
           "#ifndef ORACLE_WENO_TABLES_H", "#define ORACLE_WENO_TABLES_H"]
    for r in (4, 5):
        c, d, B = derive(r)
        out.append(f"static const double WENO{2 * r - 1}_C[{r}][{r}] = {{")
        for row in c:
            out.append("    {" + ", ".join(fmt(x) for x in row) + "},")
        out.append("};")
        out.append(f"static const double WENO{2 * r - 1}_D[{r}] = {{" + ", ".join(fmt(x) for x in d) + "};")
        out.append(f"/* beta_s = sum_(a<=b) B[s][a][b] psi_a psi_b (unscaled Jiang-Shu / Balsara-Shu definition) */")
        out.append(f"static const double WENO{2 * r - 1}_B[{r}][{r}][{r}] = {{")
        for Bs in B:
            out.append("    {" + ", ".join("{" + ", ".join(fmt(x) for x in row) + "}" for row in Bs) + "},")
        out.append("};")
        M = difference_forms(B)
        out.append(f"/* the same forms in the first differences d_a = psi_(a+1) - psi_a of the stencil: beta_s = sum_(a<=b) M[s][a][b] d_a d_b */")
        out.append(f"static const double WENO{2 * r - 1}_M[{r}][{r - 1}][{r - 1}] = {{")
        for Ms in M:
            out.append("    {" + ", ".join("{" + ", ".join(fmt(x) for x in row) + "}" for row in Ms) + "},")
        out.append("};")
    for m in (3, 4):
        out.append(f"/* Centered(order = {2 * m}) reconstruction at the face between a[-1] and a[0] from a[-{m}] .. a[{m - 1}] */")
        out.append(f"static const double CENTERED{2 * m}_C[{2 * m}] = {{" + ", ".join(fmt(x) for x in centered(m)) + "};")
    out.append("#endif")
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "oracle_weno_tables.h")
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", path)


def write_cuda_tables():
    """breeze.jl_b200/csrc/weno_tables.cuh: the same tables as device constants for the high-order reconstructions of weno.cuh."""
    import os
    out = ["// weno_tables.cuh — GENERATED by scripts/derive_weno_coefficients.py --header --cuda (exact rational derivation; do not edit).",
           "// Finite-volume WENO of buffer R = 4 (order 7) and R = 5 (order 9): candidates C, optimal weights D, smoothness forms M",
           "// (upper-triangular, in the first differences of the stencil; unscaled Jiang-Shu / Balsara-Shu definition), WENO-Z global-indicator combination G, and the scale BS the",
           "// stored forms carry upstream (oracle/oracle_weno.h documents what is derived and what is recalled).",
           "// C and M live in __constant__ memory: an FP64 instruction takes c[bank][offset] as an operand directly, whereas a folded 64-bit immediate",
           "// costs two UMOV per use (measured in stage_hi_kernel: 27 % of all issued instructions were UMOV before this change).",
           "#pragma once"]
    G = {4: [1, 3, -3, -1], 5: [1, 2, -6, 2, 1]}
    BS = {4: "0.24", 5: "0.0504"}
    for r in (4, 5):
        n = f"WENO{2 * r - 1}"
        c, d, B = derive(r)
        out.append(f"static __device__ const double {n}_G[{r}] = {{" + ", ".join(f"{g}.0" for g in G[r]) + "};")
        out.append(f"static __constant__ double {n}_D[{r}] = {{" + ", ".join(fmt(x) for x in d) + "};")
        out.append(f"static __constant__ double {n}_C[{r}][{r}] = {{")
        for row in c:
            out.append("    {" + ", ".join(fmt(x) for x in row) + "},")
        out.append("};")
        out.append(f"// smoothness forms in the first differences of the stencil (difference_forms() of the script)")
        out.append(f"static __constant__ double {n}_M[{r}][{r - 1}][{r - 1}] = {{")
        for Ms in difference_forms(B):
            out.append("    {" + ", ".join("{" + ", ".join(fmt(x) for x in row) + "}" for row in Ms) + "},")
        out.append("};")
    out.append("template <int R> struct WenoTab;")
    for r in (4, 5):
        n = f"WENO{2 * r - 1}"
        out.append(f"template <> struct WenoTab<{r}> {{")
        out.append(f"    static __device__ __forceinline__ double BS() {{ return {BS[r]}; }}")
        out.append(f"    static __device__ __forceinline__ double G(int s) {{ return {n}_G[s]; }}")
        out.append(f"    static __device__ __forceinline__ double D(int s) {{ return {n}_D[s]; }}")
        out.append(f"    static __device__ __forceinline__ double C(int s, int a) {{ return {n}_C[s][a]; }}")
        out.append(f"    static __device__ __forceinline__ double M(int s, int a, int c) {{ return {n}_M[s][a][c]; }}")
        out.append("};")
    for m in (3, 4):
        cc = centered(m)
        out.append(f"// Centered(order = {2 * m}) at the face between a[-1] and a[0]: coefficient of (a[-1-j] + a[j]), j = 0 .. {m - 1}")
        out.append(f"static __constant__ double CENTERED{2 * m}_C[{m}] = {{" + ", ".join(fmt(cc[m + j]) for j in range(m)) + "};")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "breeze.jl_b200", "csrc", "weno_tables.cuh")
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
