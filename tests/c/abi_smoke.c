/* Plain C99 consumer of the C ABI: proves include/tvl1_b200.h is C, links against the shared object,
 * and (on a GPU box) solves one small pair.  Exit code 0 = ok, 3 = no device (the library has no CPU
 * fallback), anything else = failure. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "tvl1_b200.h"

int main(void)
{
    const int nx = 96, ny = 64;
    tvl1_ctx *ctx = NULL;
    int rc = tvl1_create(0, &ctx);
    if (rc == TVL1_ERR_NODEVICE) { printf("no device: %s\n", tvl1_last_error(NULL)); return 3; }
    if (rc != TVL1_OK) { printf("create failed: %s\n", tvl1_last_error(NULL)); return 1; }
    float *I0 = malloc(sizeof(float) * nx * ny), *I1 = malloc(sizeof(float) * nx * ny);
    float *u1 = malloc(sizeof(float) * nx * ny), *u2 = malloc(sizeof(float) * nx * ny);
    for (int i = 0; i < ny; i++)
        for (int j = 0; j < nx; j++) {
            I0[i * nx + j] = 128.f + 60.f * sinf(0.21f * j + 0.13f * i);
            I1[i * nx + j] = 128.f + 60.f * sinf(0.21f * (j - 0.6f) + 0.13f * (i - 0.3f));
        }
    tvl1_params p;
    tvl1_default_params(&p);
    p.nscales = 2;
    int iters[2 * 5];
    double errs[2 * 5];
    rc = tvl1_solve_f32(ctx, I0, I1, u1, u2, nx, ny, &p, iters, errs);
    if (rc != TVL1_OK) { printf("solve failed: %s\n", tvl1_last_error(ctx)); return 1; }
    double m1 = 0, m2 = 0;
    for (int i = 16; i < ny - 16; i++)
        for (int j = 16; j < nx - 16; j++) { m1 += u1[i * nx + j]; m2 += u2[i * nx + j]; }
    m1 /= (double) (nx - 32) * (ny - 32);
    m2 /= (double) (nx - 32) * (ny - 32);
    printf("mean flow (%.3f, %.3f), first-warp iterations %d\n", m1, m2, iters[0]);
    p.nscales = 7;                                    /* level 5 is 3 px wide: narrower than the 6-tap zoom window */
    rc = tvl1_solve_f32(ctx, I0, I1, u1, u2, nx, ny, &p, NULL, NULL);
    if (rc != TVL1_ERR_SIGMA) { printf("expected TVL1_ERR_SIGMA, got %d\n", rc); return 1; }
    tvl1_destroy(ctx);
    free(I0); free(I1); free(u1); free(u2);
    return (fabs(m1 - 0.6) < 0.15 && fabs(m2 - 0.3) < 0.15 && iters[0] >= 1) ? 0 : 2;
}
