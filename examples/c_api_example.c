/*
 * Minimal C client of libtsvgp.so (include/tsvgp.h): Gaussian regression, a few natural-gradient steps, ELBO and predictions.
 *   gcc -Iinclude examples/c_api_example.c -o c_api_example -Lt-svgp_b200 -ltsvgp -Wl,-rpath,$PWD/t-svgp_b200 -lm
 * Needs a B200 to run (there is no CPU fallback); it compiles and links anywhere (tests/test_abi_and_host.py checks that).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "tsvgp.h"

#define CHECK(call)                                                                        \
    do {                                                                                   \
        int rc_ = (call);                                                                  \
        if (rc_ != TSVGP_OK) {                                                             \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, tsvgp_last_error(ctx));    \
            return 1;                                                                      \
        }                                                                                  \
    } while (0)

int main(void) {
    enum { N = 2000, M = 40, D = 1 };
    static double X[N * D], Y[N], Z[M * D], Xs[5], mean[5], var[5];
    unsigned s = 12345u;
    for (int i = 0; i < N; ++i) {
        s = s * 1664525u + 1013904223u;
        X[i] = 2.0 * (s >> 8) / 16777216.0 - 1.0;
        s = s * 1664525u + 1013904223u;
        Y[i] = sin(6.0 * X[i]) + 0.2 * ((s >> 8) / 16777216.0 - 0.5);
    }
    for (int i = 0; i < M; ++i) Z[i] = -1.0 + 2.0 * i / (M - 1);
    for (int i = 0; i < 5; ++i) Xs[i] = -0.8 + 0.4 * i;

    tsvgp_ctx* ctx = NULL;
    if (tsvgp_create(&ctx, 0) != TSVGP_OK) {
        fprintf(stderr, "tsvgp_create: %s\n", tsvgp_last_error(NULL));
        return 1;
    }
    const double lengthscale = 0.2;
    CHECK(tsvgp_set_kernel(ctx, TSVGP_KERNEL_SE, 1.0, &lengthscale, 1));
    CHECK(tsvgp_set_likelihood(ctx, TSVGP_LIK_GAUSSIAN, 0.05, 0.0, 0, NULL, NULL));
    CHECK(tsvgp_set_inducing(ctx, Z, M, D, NULL));
    CHECK(tsvgp_set_data(ctx, X, Y, N, D, NULL));
    for (int it = 0; it < 3; ++it) {
        double elbo_before = 0.0;
        CHECK(tsvgp_natgrad_step(ctx, /*lr=*/0.9, /*jitter=*/1e-9, /*scale=*/1.0, &elbo_before));
        printf("step %d: ELBO before the step %.6f\n", it, elbo_before);
    }
    double elbo = 0.0;
    CHECK(tsvgp_elbo(ctx, 1.0, &elbo));
    CHECK(tsvgp_predict_f(ctx, Xs, 5, D, NULL, mean, var));
    printf("ELBO %.6f\n", elbo);
    for (int i = 0; i < 5; ++i) printf("f(% .2f) = % .4f +- %.4f   (sin(6x) = % .4f)\n", Xs[i], mean[i], sqrt(var[i]), sin(6.0 * Xs[i]));
    tsvgp_destroy(ctx);
    return 0;
}
