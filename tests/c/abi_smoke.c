/* abi_smoke.c — drives libbreeze_b200.so from plain C through include/breeze_b200.h and include/breeze_b200_compressible.h:
 * no Python, no torch, plain pointers only. Built and run by tests/test_c_abi.py.
 *   abi_smoke anelastic|compressible NX NY NZ STEPS   → prints "checksum <sum of |field|>" lines, exit 0
 * Without a CUDA device bz_create / bzc_create must fail with the "no CPU fallback" message (exit 3). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "breeze_b200.h"
#include "breeze_b200_compressible.h"

static double bubble(double x, double y, double z, double zc) {
    double r = sqrt(x * x + y * y + (z - zc) * (z - zc)) / 2000.0;
    if (r > 1) r = 1;
    double c = cos(M_PI / 2 * r);
    return 300.0 + 2.0 * c * c;
}

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s anelastic|compressible NX NY NZ STEPS\n", argv[0]); return 2; }
    const int compressible = strcmp(argv[1], "compressible") == 0;
    const int Nx = atoi(argv[2]), Ny = atoi(argv[3]), Nz = atoi(argv[4]), steps = atoi(argv[5]);
    const size_t nc = (size_t)Nx * Ny * Nz, nw = (size_t)Nx * Ny * (Nz + 1);
    double* th = malloc(nc * sizeof(double)); double* out = malloc(nw * sizeof(double));
    if (!compressible) {
        bz_config cfg; bz_default_config(&cfg);
        cfg.Nx = Nx; cfg.Ny = Ny; cfg.Nz = Nz;
        cfg.x0 = -10e3; cfg.x1 = 10e3; cfg.y0 = -10e3; cfg.y1 = 10e3; cfg.z0 = 0; cfg.z1 = 10e3;
        cfg.potential_temperature = 300.0;
        bz_ctx* ctx = NULL;
        if (bz_create(&cfg, &ctx) != BZ_OK) { fprintf(stderr, "bz_create: %s\n", bz_last_error(NULL)); return 3; }
        double* rho = malloc(Nz * sizeof(double));
        bz_get_reference_state(ctx, rho, NULL, NULL);
        for (int k = 0; k < Nz; ++k) for (int j = 0; j < Ny; ++j) for (int i = 0; i < Nx; ++i) {
            double x = cfg.x0 + (i + 0.5) * (cfg.x1 - cfg.x0) / Nx, y = cfg.y0 + (j + 0.5) * (cfg.y1 - cfg.y0) / Ny, z = (k + 0.5) * cfg.z1 / Nz;
            th[i + (size_t)Nx * (j + (size_t)Ny * k)] = rho[k] * bubble(x, y, z, 2000.0);
        }
        if (bz_set_state(ctx, NULL, NULL, NULL, th, NULL, 1) != BZ_OK) { fprintf(stderr, "bz_set_state: %s\n", bz_last_error(ctx)); return 4; }
        if (bz_time_steps(ctx, 2.0, steps) != BZ_OK) { fprintf(stderr, "bz_time_steps: %s\n", bz_last_error(ctx)); return 5; }
        const int ids[3] = {BZ_RHO_W, BZ_RHO_THETA, BZ_PHI};
        for (int f = 0; f < 3; ++f) {
            if (bz_get_field(ctx, ids[f], out) != BZ_OK) return 6;
            double s = 0; size_t n = ids[f] == BZ_RHO_W ? nw : nc;
            for (size_t e = 0; e < n; ++e) s += fabs(out[e]);
            printf("checksum %d %.17g\n", ids[f], s);
        }
        double t; int64_t it; bz_get_clock(ctx, &t, &it);
        printf("clock %.17g %lld launches %lld\n", t, (long long)it, (long long)bz_kernel_launch_count(ctx));
        bz_destroy(ctx); free(rho);
    } else {
        bzc_config cfg; bzc_default_config(&cfg);
        cfg.base.Nx = Nx; cfg.base.Ny = Ny; cfg.base.Nz = Nz;
        cfg.base.x0 = -5e3; cfg.base.x1 = 5e3; cfg.base.y0 = -5e3; cfg.base.y1 = 5e3; cfg.base.z0 = 0; cfg.base.z1 = 10e3;
        cfg.base.potential_temperature = 300.0; cfg.substeps = 6;
        bzc_ctx* ctx = NULL;
        if (bzc_create(&cfg, &ctx) != BZ_OK) { fprintf(stderr, "bzc_create: %s\n", bzc_last_error(NULL)); return 3; }
        double* rho_r = malloc(Nz * sizeof(double)); double* rho = malloc(nc * sizeof(double));
        bzc_get_reference_state(ctx, NULL, rho_r, NULL);
        for (int k = 0; k < Nz; ++k) for (int j = 0; j < Ny; ++j) for (int i = 0; i < Nx; ++i) {
            double x = -5e3 + (i + 0.5) * 10e3 / Nx, y = -5e3 + (j + 0.5) * 10e3 / Ny, z = (k + 0.5) * 10e3 / Nz;
            size_t e = i + (size_t)Nx * (j + (size_t)Ny * k);
            rho[e] = rho_r[k]; th[e] = rho_r[k] * bubble(x, y, z, 3000.0);
        }
        if (bzc_set_state(ctx, rho, NULL, NULL, NULL, th, NULL) != BZ_OK) { fprintf(stderr, "bzc_set_state: %s\n", bzc_last_error(ctx)); return 4; }
        if (bzc_time_steps(ctx, 2.0, steps) != BZ_OK) { fprintf(stderr, "bzc_time_steps: %s\n", bzc_last_error(ctx)); return 5; }
        const int ids[3] = {BZC_RHO_W, BZC_RHO_THETA, BZC_P};
        for (int f = 0; f < 3; ++f) {
            if (bzc_get_field(ctx, ids[f], out) != BZ_OK) return 6;
            double s = 0; size_t n = ids[f] == BZC_RHO_W ? nw : nc;
            for (size_t e = 0; e < n; ++e) s += fabs(out[e]);
            printf("checksum %d %.17g\n", ids[f], s);
        }
        double t; int64_t it; bzc_get_clock(ctx, &t, &it);
        printf("clock %.17g %lld launches %lld\n", t, (long long)it, (long long)bzc_kernel_launch_count(ctx));
        bzc_destroy(ctx); free(rho_r); free(rho);
    }
    free(th); free(out);
    return 0;
}
