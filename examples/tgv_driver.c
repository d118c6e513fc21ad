/* tgv_driver.c -- a plain-C driver over the C ABI of libfen_gpu.so (include/fen_gpu.h): the call sequence of
 * FEN's test/small_test/navier_stokes/taylor_green_vortex/taylor_green_vortex.f90 (grid setup, init_solver,
 * set_timestep, initial condition written into host arrays, time loop with advance_solution and the status line),
 * with the two explicit transfer points (push / pull) that replace the reference's direct pokes into module arrays.
 *
 *   gcc -std=c99 -Iinclude examples/tgv_driver.c -Lfen_b200 -lfen_gpu -Wl,-rpath,$PWD/fen_b200 -lm -o tgv_driver
 *   ./tgv_driver 64 20          # N, steps -- needs a CUDA device; without one fen_gpu_create fails (no CPU fallback)
 *
 * tests/test_host_logic.py compiles it (the header must be valid C, not only C++) and runs it without a GPU to check
 * the error path. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fen_gpu.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != FEN_OK) {                                                         \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, fen_gpu_last_error());     \
            return rc_;                                                              \
        }                                                                            \
    } while (0)

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 64;
    const int nstep = argc > 2 ? atoi(argv[2]) : 10;
    const double pi = acos(-1.0);
    const double L = 2.0 * pi;

    fen_grid_desc g;
    memset(&g, 0, sizeof(g));
    g.nx = g.ny = n;
    g.nz = 1;
    g.ndim = 2;
    g.delta = L / (double)(float)n;          /* grid.f90:140: Lx/float(Nx) */
    g.rank = 0;
    g.nranks = 1;
    g.device = -1;                           /* bc[] = 0: periodic */
    fen_ctx* ctx = NULL;
    CHECK(fen_gpu_create(&g, &ctx));

    fen_ns_params prm;
    CHECK(fen_gpu_get_params(ctx, &prm));
    prm.density = 1.0;
    prm.viscosity = 1.0;
    CHECK(fen_gpu_set_params(ctx, &prm));
    CHECK(fen_gpu_init_solver(ctx));
    double dt = 0.0;
    CHECK(fen_gpu_set_timestep(ctx, 2.0, &dt));

    /* host arrays f(0:N+1, 0:N+1, 0:2), x fastest (scalar.f90:79-81); initial condition of
     * taylor_green_vortex.f90:86-113 at the staggered locations */
    const size_t m = (size_t)(n + 2);
    const size_t cnt = m * m * 3;
    double* u = (double*)calloc(cnt, sizeof(double));
    double* v = (double*)calloc(cnt, sizeof(double));
    double* p = (double*)calloc(cnt, sizeof(double));
    for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) {
            const size_t o = (size_t)i + m * ((size_t)j + m * 1);
            const double xc = (i - 0.5) * g.delta, yc = (j - 0.5) * g.delta;
            const double xf = i * g.delta, yf = j * g.delta;
            u[o] = -cos(xf) * sin(yc);
            v[o] = sin(xc) * cos(yf);
            p[o] = -0.25 * (cos(2.0 * xc) + cos(2.0 * yc));
        }
    CHECK(fen_gpu_push(ctx, FEN_VX, u, 1));
    CHECK(fen_gpu_push(ctx, FEN_VY, v, 1));
    CHECK(fen_gpu_push(ctx, FEN_P, p, 1));
    CHECK(fen_gpu_update_ghost_nodes(ctx, FEN_VX, 2));
    CHECK(fen_gpu_update_ghost_nodes(ctx, FEN_P, 1));

    double time = 0.0;
    char line[256];
    for (int step = 1; step <= nstep; ++step) {
        time += dt;
        CHECK(fen_gpu_navier_stokes_solver(ctx, step, &dt));        /* advance_solution */
        CHECK(fen_gpu_status_line(ctx, step, time, dt, line, (int)sizeof(line)));
        if (step == 1 || step == nstep) printf("%s\n", line);
    }
    CHECK(fen_gpu_pull(ctx, FEN_VX, u, 1));
    /* analytic solution: u = -cos x sin y exp(-2 nu t) (postpro.py:48) */
    double emax = 0.0;
    for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) {
            const size_t o = (size_t)i + m * ((size_t)j + m * 1);
            const double e = fabs(u[o] + cos(i * g.delta) * sin((j - 0.5) * g.delta) * exp(-2.0 * time));
            if (e > emax) emax = e;
        }
    printf("max |u - u_exact| after %d steps: %.3e\n", nstep, emax);
    free(u); free(v); free(p);
    CHECK(fen_gpu_destroy_solver(ctx));
    CHECK(fen_gpu_destroy(ctx));
    return 0;
}
