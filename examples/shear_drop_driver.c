/* shear_drop_driver.c -- a plain-C driver of the two-phase path over the C ABI of libfen_gpu.so (include/fen_gpu.h):
 * the call sequence of FEN's test/small_test/multiphase/shear_drop/shear_drop.f90 (a -DMF build): module parameters,
 * init_solver with the distance function of the drop, set_timestep, moving-wall boundary values, the linear shear as
 * initial velocity, the time loop, and the deformation of the vof = 0.5 contour as shear_drop/deformation.py measures
 * it -- to be compared with the Basilisk value the reference ships (Re1Ca02b.csv: D = 0.1204 at t = 1).
 *
 *   gcc -std=c99 -Iinclude examples/shear_drop_driver.c -Lfen_b200 -lfen_gpu -Wl,-rpath,$PWD/fen_b200 -lm -o shear_drop
 *   ./shear_drop 0.2 1.0        # capillary number, end time -- needs a CUDA device (no CPU fallback)
 *
 * tests/test_host_logic.py compiles and links it; tests/test_gpu_zz_rising_bubble.py runs it on a GPU box. */
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

/* shear_drop.f90:113-121: positive inside the drop of radius 0.5 centred in the 2 x 2 box */
static double circle(void* user, double x, double y) {
    (void)user;
    return -(sqrt((x - 1.0) * (x - 1.0) + (y - 1.0) * (y - 1.0)) - 0.5);
}

/* deformation.py:45-69: extreme distances of the vof = 0.5 contour from the box centre, contour points by linear
 * interpolation on the lines joining cell centres */
static double deformation(const double* vof, int n, double delta) {
    const size_t m = (size_t)(n + 2);
    double dmax = 0.0, dmin = 1.0e30;
    for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) {
            const double f0 = vof[i + m * (j + m)] - 0.5;
            const double xc = (i - 0.5) * delta, yc = (j - 0.5) * delta;
            for (int dir = 0; dir < 2; ++dir) {
                if ((dir == 0 && i == n) || (dir == 1 && j == n)) continue;
                const double f1 = vof[(i + (dir == 0)) + m * ((j + (dir == 1)) + m)] - 0.5;
                if (f0 * f1 >= 0.0) continue;
                const double t = f0 / (f0 - f1);
                const double x = xc + (dir == 0 ? t * delta : 0.0), y = yc + (dir == 1 ? t * delta : 0.0);
                const double d = sqrt((x - 1.0) * (x - 1.0) + (y - 1.0) * (y - 1.0));
                if (d > dmax) dmax = d;
                if (d < dmin) dmin = d;
            }
        }
    return (dmax - dmin) / (dmax + dmin);
}

int main(int argc, char** argv) {
    const double Ca = argc > 1 ? atof(argv[1]) : 0.2;
    const double Tmax = argc > 2 ? atof(argv[2]) : 1.0;
    const int n = 64;
    const double U = 1.0, a = 0.5, Re = 1.0, Lx = 2.0;

    fen_grid_desc g;
    memset(&g, 0, sizeof(g));
    g.nx = g.ny = n;
    g.nz = 1;
    g.ndim = 2;
    g.delta = Lx / (double)(float)n;                 /* grid.f90:140 */
    g.bc[0] = g.bc[1] = FEN_BC_PERIODIC;             /* shear_drop.f90:47-50 */
    g.bc[2] = g.bc[3] = FEN_BC_WALL;
    g.nranks = 1;
    g.device = -1;
    fen_ctx* ctx = NULL;
    CHECK(fen_gpu_create(&g, &ctx));

    /* module variables set before init_solver (:54-62) */
    fen_mf_params mp;
    CHECK(fen_gpu_mf_get_params(ctx, &mp));
    mp.beta = 1.0;
    mp.rho_0 = 1.0;
    mp.rho_1 = mp.rho_0;
    mp.mu_0 = mp.rho_0 * U * 2.0 * a / Re;
    mp.mu_1 = 1.0 * mp.mu_0;
    mp.sigma = U * mp.mu_0 / Ca;
    CHECK(fen_gpu_mf_set_params(ctx, &mp));
    CHECK(fen_gpu_init_solver_mf(ctx, circle, NULL, 0.0, 0.0));        /* :64 */
    double dt = 0.0;
    CHECK(fen_gpu_set_timestep(ctx, U, &dt));                          /* :76 */
    const double top = U, bottom = -U;                                 /* :77-78 v%x%bc%top = U, %bottom = -U */
    CHECK(fen_gpu_set_bc_plane(ctx, FEN_VX, FEN_TOP, &top, 1));
    CHECK(fen_gpu_set_bc_plane(ctx, FEN_VX, FEN_BOTTOM, &bottom, 1));

    const size_t m = (size_t)(n + 2), cnt = m * m * 3;
    double* u = (double*)calloc(cnt, sizeof(double));
    double* vof = (double*)calloc(cnt, sizeof(double));
    for (int j = 1; j <= n; ++j)                                       /* :80-88 linear shear */
        for (int i = 1; i <= n; ++i) u[i + m * (j + m)] = -U + 2.0 * U * ((j - 0.5) * g.delta) / Lx;
    CHECK(fen_gpu_push(ctx, FEN_VX, u, 1));
    CHECK(fen_gpu_update_ghost_nodes(ctx, FEN_VX, 2));

    double time = 0.0, D = 0.0;
    char line[256];
    int step = 0;
    while (time <= Tmax) {                                             /* :92-100 */
        ++step;
        time += dt;
        CHECK(fen_gpu_navier_stokes_solver(ctx, step, &dt));
        if (step % 1024 == 0 || time > Tmax) {
            CHECK(fen_gpu_status_line(ctx, step, time, dt, line, (int)sizeof(line)));
            CHECK(fen_gpu_pull(ctx, FEN_VOF, vof, 1));
            D = deformation(vof, n, g.delta);
            printf("%s D: %.6f\n", line, D);
        }
    }
    double i1 = 0.0, i2 = 0.0;
    CHECK(fen_gpu_check_vof_integral(ctx, &i1, &i2));
    printf("steps %d  deformation %.6f  drop volume integral %.12e\n", step, D, i1);
    free(u); free(vof);
    CHECK(fen_gpu_destroy_solver(ctx));
    CHECK(fen_gpu_destroy(ctx));
    return 0;
}
