/* fen_oracle_mf_c.c -- plain C (C99 + OpenMP) restatement of FEN's two-phase fractional step (the -DMF build) for the
 * wave cases of the reference: 2-D, x periodic, walls in y (pn Poisson: FFT in x + Thomas in y), MTHINC volume of
 * fluid, variable-viscosity stress divergence, CSF surface tension, constant-coefficient pressure splitting.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as oracle/fen_oracle.py): nothing under fen_b200/ may link or call this.  It is
 * (1) a second checker of the two-phase path, written from the reference's explicit loops, not from the numpy
 * restatement -- tests/test_oracle_c.py holds the two to 1e-11 of each other -- and (2) the CPU baseline of
 * `bench.py --case wave2d` on all host threads (the reference itself, Fortran + MPI + FFTW3 + 2decomp, cannot be built
 * in this image).  Paths below are relative to /root/reference.  FFTW's r2c / c2r (INSTALL.sh:16-17, FFTW 3.3.10) are
 * restated by a textbook radix-2 complex FFT of the row (nx a power of two).
 *
 * Layout: every field is the reference's 2-D f(0:nx+1, 0:ny+1), x fastest (src/scalar.f90:79-81): plane 1 of the numpy
 * oracle's Fortran-ordered (nx+2, ny+2, 3) arrays.  Scope: constant_CFL off, S = 0, time step fixed by the caller.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cplx;

typedef struct FoMF {
    int nx, ny;
    long sy, n;
    double delta, rho0, rho1, mu0, mu1, sigma, beta, cut, dt_o, g[2], rhomin, irhomin;
    int x_first, quadratic;
    int vof_bc_y;                    /* boundary type of vof on bottom / top: 2 (wall) until the first advect_vof, then
                                        0 -- `vof = vof1` copies vof1's default periodic types (hazard H13) */
    double *p, *phi, *rho, *mu, *u, *v, *vof, *h, *d, *curv, *nrx, *nry, *lx, *ly, *phat, *po, *vof1, *vof2;
    double *dvx, *dvy, *dvox, *dvoy, *rfx, *rfy, *gpx, *gpy, *ghx, *ghy;
    cplx* C;                         /* [nx/2+1][ny], kx fastest */
    int mc;
    double *mwn_x, *ta, *tb, *tc, *c1;
    cplx* tw;
    double maxdiv, maxvel;
    double ubc[2];                   /* uniform Dirichlet values of u on the bottom / top wall */
} FoMF;

#define AT(s, i, j) ((long)(i) + (s)->sy * (long)(j))
static const double SMALL = 1.0e-14;                                  /* global.f90:16 */
#define RP (0.5 * (1.0 + 1.0 / sqrt(3.0)))                            /* volume_of_fluid.f90:33 */
#define RM (0.5 * (1.0 - 1.0 / sqrt(3.0)))                            /* :34 */

static double f32(long n) { return (double)(float)n; }               /* Fortran float(n) (hazard H1) */

int fomf_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void fomf_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- radix-2 FFT, in place, unnormalised; sign = -1 forward, +1 backward ------------------------------------------ */
static void fft(cplx* x, int n, int sign, const cplx* tw) {
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cplx t = x[i]; x[i] = x[j]; x[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1, step = n / len;
        for (int s = 0; s < n; s += len)
            for (int q = 0; q < half; ++q) {
                const cplx w = tw[q * step];
                const double wi = sign < 0 ? w.im : -w.im;
                cplx* a = x + s + q;
                cplx* b = a + half;
                const double tr = b->re * w.re - b->im * wi, ti = b->re * wi + b->im * w.re;
                b->re = a->re - tr; b->im = a->im - ti;
                a->re += tr; a->im += ti;
            }
    }
}

static double* new_field(const FoMF* s) { return (double*)calloc((size_t)s->n, sizeof(double)); }   /* scalar.f90:84 */

FoMF* fomf_create(int nx, int ny, double delta, double rho0, double rho1, double mu0, double mu1, double sigma,
                  double beta) {
    if (nx < 2 || (nx & (nx - 1)) || ny < 2) return NULL;
    FoMF* s = (FoMF*)calloc(1, sizeof(FoMF));
    s->nx = nx; s->ny = ny; s->sy = nx + 2; s->n = (long)(nx + 2) * (ny + 2);
    s->delta = delta; s->rho0 = rho0; s->rho1 = rho1; s->mu0 = mu0; s->mu1 = mu1; s->sigma = sigma; s->beta = beta;
    s->cut = 1.0e-8;                                                  /* volume_of_fluid.f90:37 */
    s->x_first = 1; s->quadratic = 1;                                 /* :30, :27 */
    s->vof_bc_y = 2;
    s->rhomin = rho0 < rho1 ? rho0 : rho1;                            /* solver.f90:92-93 */
    s->irhomin = 1.0 / s->rhomin;
    double** all[] = {&s->p, &s->phi, &s->rho, &s->mu, &s->u, &s->v, &s->vof, &s->h, &s->d, &s->curv, &s->nrx, &s->nry,
                      &s->lx, &s->ly, &s->phat, &s->po, &s->vof1, &s->vof2, &s->dvx, &s->dvy, &s->dvox, &s->dvoy,
                      &s->rfx, &s->rfy, &s->gpx, &s->gpy, &s->ghx, &s->ghy};
    for (size_t q = 0; q < sizeof(all) / sizeof(all[0]); ++q) *all[q] = new_field(s);
    s->mc = nx / 2 + 1;
    s->C = (cplx*)calloc((size_t)s->mc * ny, sizeof(cplx));
    s->c1 = (double*)calloc((size_t)s->mc * ny, sizeof(double));
    s->tw = (cplx*)malloc(sizeof(cplx) * (size_t)nx);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int m = 0; m < nx; ++m) {
        long double a = -two_pi * (long double)m / (long double)nx;
        s->tw[m].re = (double)cosl(a);
        s->tw[m].im = (double)sinl(a);
    }
    /* init_poisson_pn, poisson.f90:176-237 */
    const double pi = acos(-1.0);                                     /* global.f90:14 */
    s->mwn_x = (double*)malloc(sizeof(double) * (size_t)nx);
    for (int i = 1; i <= nx; ++i) s->mwn_x[i - 1] = 2.0 * (cos(2.0 * pi * (i - 1.0) / f32(nx)) - 1.0) / (delta * delta);
    s->ta = (double*)malloc(sizeof(double) * (size_t)ny);
    s->tb = (double*)malloc(sizeof(double) * (size_t)ny);
    s->tc = (double*)malloc(sizeof(double) * (size_t)ny);
    for (int j = 0; j < ny; ++j) { s->ta[j] = 1.0 / (delta * delta); s->tb[j] = -2.0 / (delta * delta); s->tc[j] = 1.0 / (delta * delta); }
    s->tb[0] = s->tb[0] + s->ta[0];
    s->tb[ny - 1] = s->tb[ny - 1] + s->tc[ny - 1];
    s->ta[0] = 0.0;
    s->tc[ny - 1] = 0.0;
    return s;
}

void fomf_destroy(FoMF* s) {
    if (!s) return;
    double* all[] = {s->p, s->phi, s->rho, s->mu, s->u, s->v, s->vof, s->h, s->d, s->curv, s->nrx, s->nry, s->lx, s->ly,
                     s->phat, s->po, s->vof1, s->vof2, s->dvx, s->dvy, s->dvox, s->dvoy, s->rfx, s->rfy, s->gpx, s->gpy,
                     s->ghx, s->ghy, s->mwn_x, s->ta, s->tb, s->tc, s->c1};
    for (size_t q = 0; q < sizeof(all) / sizeof(all[0]); ++q) free(all[q]);
    free(s->C); free(s->tw); free(s);
}

long fomf_field_size(const FoMF* s) { return s->n; }
/* 0 p, 1 phi, 2 rho, 3 mu, 4 u, 5 v, 6 vof, 7 h, 8 d, 9 curv, 10 norm x, 11 norm y, 12 l x, 13 l y, 14 p_hat, 15 p_o,
 * 16 dv_o x, 17 dv_o y */
static double* field(FoMF* s, int id) {
    double* f[] = {s->p, s->phi, s->rho, s->mu, s->u, s->v, s->vof, s->h, s->d, s->curv, s->nrx, s->nry, s->lx, s->ly,
                   s->phat, s->po, s->dvox, s->dvoy};
    return (id >= 0 && id < 18) ? f[id] : NULL;
}
void fomf_set_field(FoMF* s, int id, const double* src) { memcpy(field(s, id), src, sizeof(double) * (size_t)s->n); }
void fomf_get_field(FoMF* s, int id, double* dst) { memcpy(dst, field(s, id), sizeof(double) * (size_t)s->n); }
void fomf_set_params(FoMF* s, double dt_o, double g0, double g1) { s->dt_o = dt_o; s->g[0] = g0; s->g[1] = g1; }
void fomf_set_wall_velocity(FoMF* s, double bottom, double top) { s->ubc[0] = bottom; s->ubc[1] = top; }
int fomf_x_first(const FoMF* s) { return s->x_first; }
void fomf_set_vof_state(FoMF* s, int x_first, int vof_bc_y) { s->x_first = x_first; s->vof_bc_y = vof_bc_y; }
int fomf_vof_bc_y(const FoMF* s) { return s->vof_bc_y; }
double fomf_maxdiv(const FoMF* s) { return s->maxdiv; }
double fomf_maxcfl(const FoMF* s, double dt) { return dt * s->maxvel / s->delta; }      /* navier_stokes.f90:617 */

/* scalar%update_ghost_nodes (src/scalar.f90:255-345), x periodic; bottom / top of type ty: 0 periodic (prow = 1), 1
 * Dirichlet with value 0 -- loc 'y' (the wall-normal staggered component): ghost = bc and the last interior face = bc
 * too; otherwise ghost = 2 bc - f -- and 2 Neumann.  The x faces go first over the whole (j) extent, then the y faces over
 * the whole (i) extent: this order fills the corner ghosts the stencils read (hazard H3).  vb / vt: the (uniform)
 * boundary values bc%bottom / bc%top. */
static void ghosts_v(const FoMF* s, double* f, char loc, int ty, double vb, double vt) {
    const int nx = s->nx, ny = s->ny;
    for (int j = 0; j <= ny + 1; ++j) f[AT(s, 0, j)] = f[AT(s, nx, j)];
    for (int j = 0; j <= ny + 1; ++j) f[AT(s, nx + 1, j)] = f[AT(s, 1, j)];
    for (int i = 0; i <= nx + 1; ++i) {                                /* bottom */
        if (ty == 0) f[AT(s, i, 0)] = f[AT(s, i, ny)];
        else if (ty == 1) f[AT(s, i, 0)] = (loc == 'y') ? vb : 2.0 * vb - f[AT(s, i, 1)];
        else f[AT(s, i, 0)] = f[AT(s, i, 1)];
    }
    for (int i = 0; i <= nx + 1; ++i) {                                /* top */
        if (ty == 0) f[AT(s, i, ny + 1)] = f[AT(s, i, 1)];
        else if (ty == 1) {
            if (loc == 'y') { f[AT(s, i, ny)] = vt; f[AT(s, i, ny + 1)] = vt; }
            else f[AT(s, i, ny + 1)] = 2.0 * vt - f[AT(s, i, ny)];
        } else f[AT(s, i, ny + 1)] = f[AT(s, i, ny)];
    }
}
static void ghosts(const FoMF* s, double* f, char loc, int ty) { ghosts_v(s, f, loc, ty, 0.0, 0.0); }
static void ghosts_velocity(const FoMF* s) {      /* vector%update_ghost_nodes, vector.f90:82-109: walls -> Dirichlet */
    ghosts_v(s, s->u, 'x', 1, s->ubc[0], s->ubc[1]);   /* v%x%bc%bottom / top: moving walls (shear_drop.f90:77-78) */
    ghosts(s, s->v, 'y', 1);
}

/* ---- volume_of_fluid_mod ------------------------------------------------------------------------------------------- */
typedef struct { double cx, cy, a10, a01, a20, a02; } Surf;

/* coefficients of Eq. 12 of Ii et al. from the cell's normal and curvature terms (volume_of_fluid.f90:254-266, 610-634) */
static Surf surface(double nx, double ny, double lx, double ly) {
    Surf q;
    const double ax = fabs(nx), ay = fabs(ny);
    if (ax == (ax > ay ? ax : ay)) { q.cx = 0.0; q.cy = 1.0; }
    else { q.cx = 1.0; q.cy = 0.0; }
    q.a10 = nx - 0.5 * q.cx * lx;
    q.a01 = ny - 0.5 * q.cy * ly;
    q.a20 = 0.5 * q.cx * lx;
    q.a02 = 0.5 * q.cy * ly;
    return q;
}
static double P(const Surf* q, double x, double y) {                  /* :400-411 */
    return q->cx * q->a20 * (x * x) + q->cy * q->a02 * (y * y) + q->a10 * x + q->a01 * y;
}
static double solve_quadratic(double a, double b, double c) {         /* :415-430 */
    const double x1 = (-b + sqrt(b * b - 4.0 * a * c)) / (2.0 * a);
    const double x2 = (-b - sqrt(b * b - 4.0 * a * c)) / (2.0 * a);
    return x1 > x2 ? x1 : x2;
}

static void compute_norm(FoMF* s) {                                   /* :307-396 */
    const int nx = s->nx, ny = s->ny;
    const double delta = s->delta, idelta = 1.0 / delta, idelta2 = 1.0 / (delta * delta);
    const double* F = s->vof;
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; ++j) {
        const int jp = j + 1, jm = j - 1;
        for (int i = 1; i <= nx; ++i) {
            const int ip = i + 1, im = i - 1;
            double mx[4], my[4], nrx[4], nry[4];
            mx[0] = 0.5 * (F[AT(s, i, jm)] + F[AT(s, i, j)] - F[AT(s, im, jm)] - F[AT(s, im, j)]) * idelta;
            mx[1] = 0.5 * (F[AT(s, i, j)] + F[AT(s, i, jp)] - F[AT(s, im, j)] - F[AT(s, im, jp)]) * idelta;
            mx[2] = 0.5 * (F[AT(s, ip, j)] + F[AT(s, ip, jp)] - F[AT(s, i, j)] - F[AT(s, i, jp)]) * idelta;
            mx[3] = 0.5 * (F[AT(s, ip, jm)] + F[AT(s, ip, j)] - F[AT(s, i, jm)] - F[AT(s, i, j)]) * idelta;
            const double mxc = 0.25 * (mx[0] + mx[1] + mx[2] + mx[3]);
            my[0] = 0.5 * (F[AT(s, im, j)] + F[AT(s, i, j)] - F[AT(s, im, jm)] - F[AT(s, i, jm)]) * idelta;
            my[1] = 0.5 * (F[AT(s, im, jp)] + F[AT(s, i, jp)] - F[AT(s, im, j)] - F[AT(s, i, j)]) * idelta;
            my[2] = 0.5 * (F[AT(s, i, jp)] + F[AT(s, ip, jp)] - F[AT(s, i, j)] - F[AT(s, ip, j)]) * idelta;
            my[3] = 0.5 * (F[AT(s, i, j)] + F[AT(s, ip, j)] - F[AT(s, i, jm)] - F[AT(s, ip, jm)]) * idelta;
            const double myc = 0.25 * (my[0] + my[1] + my[2] + my[3]);
            for (int c = 0; c < 4; ++c) {
                nrx[c] = mx[c] / sqrt(mx[c] * mx[c] + my[c] * my[c] + SMALL);
                nry[c] = my[c] / sqrt(mx[c] * mx[c] + my[c] * my[c] + SMALL);
            }
            const long c0 = AT(s, i, j);
            s->nrx[c0] = mxc / sqrt(mxc * mxc + myc * myc + SMALL);
            s->nry[c0] = myc / sqrt(mxc * mxc + myc * myc + SMALL);
            if (s->quadratic) {
                s->lx[c0] = 0.5 * delta * (nrx[3] + nrx[2] - nrx[1] - nrx[0]);
                s->ly[c0] = 0.5 * delta * (nry[1] + nry[2] - nry[0] - nry[3]);
            } else {
                s->lx[c0] = 0.0;
                s->ly[c0] = 0.0;
            }
            s->curv[c0] = -(s->lx[c0] + s->ly[c0]) * idelta2;
        }
    }
    /* curv, norm, l were wired by allocate_vof_fields (:68-198): Wall -> Neumann */
    ghosts(s, s->curv, 'c', 2);
    ghosts(s, s->nrx, 'c', 2);
    ghosts(s, s->nry, 'c', 2);
    ghosts(s, s->lx, 'c', 2);
    ghosts(s, s->ly, 'c', 2);
}

static void get_h_from_vof(FoMF* s) {                                 /* :228-303 */
    compute_norm(s);
    const int nx = s->nx, ny = s->ny;
    const double beta = s->beta, rp = RP, rm = RM;
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            const long c0 = AT(s, i, j);
            const double f = s->vof[c0];
            if (f <= s->cut || f >= (1.0 - s->cut)) {
                s->h[c0] = f;
                s->d[c0] = 0.0;
                continue;
            }
            const Surf q = surface(s->nrx[c0], s->nry[c0], s->lx[c0], s->ly[c0]);
            const double A = (1.0 - q.cx) * exp(2.0 * beta * q.a10) + (1.0 - q.cy) * exp(2.0 * beta * q.a01);
            const double Bp = (1.0 - q.cx) * exp(2.0 * beta * P(&q, 0.0, rp)) + (1.0 - q.cy) * exp(2.0 * beta * P(&q, rp, 0.0));
            const double Bm = (1.0 - q.cx) * exp(2.0 * beta * P(&q, 0.0, rm)) + (1.0 - q.cy) * exp(2.0 * beta * P(&q, rm, 0.0));
            const double Q = (1.0 - q.cx) * exp(2.0 * beta * q.a10 * (2.0 * f - 1.0)) +
                             (1.0 - q.cy) * exp(2.0 * beta * q.a01 * (2.0 * f - 1.0));
            const double aa = A * Bm * Bp * (A - Q);
            const double bb = A * (Bp + Bm) * (1.0 - Q);
            const double cc = 1.0 - A * Q;
            s->d[c0] = log(solve_quadratic(aa, bb, cc)) / (2.0 * beta);
            s->h[c0] = 0.5 * (1.0 + tanh(beta * (P(&q, 0.5, 0.5) + s->d[c0])));
        }
    ghosts(s, s->h, 'c', 2);
    ghosts(s, s->d, 'c', 2);
}

static double an_int(const FoMF* s, const Surf* q, double a, double b, double xa, double xb, double ya, double yb,
                     double c, double d0) {                           /* :660-672 */
    return 0.5 * (b - a + 1.0 / (c * s->beta) * log(cosh(s->beta * (P(q, xb, yb) + d0)) /
                                                     cosh(s->beta * (P(q, xa, ya) + d0))));
}
static double num_int(double a, double b, double qrm, double qrp) { return 0.5 * (qrm + qrp) * (b - a); }   /* :646-656 */

/* compute_flux(dir, i, j, k, u, dt, delta), :558-642: the flux through the + face of cell (i, j) in direction dir */
static double compute_flux(const FoMF* s, int dir, int i, int j, double u, double dt) {
    const double delta = s->delta, rp = RP, rm = RM;
    double xa, xb, ya, yb, sgn;
    int ii = i, jj = j;
    if (dir == 1) {
        if (u >= 0.0) { xa = 1.0 - dt * u / delta; xb = 1.0; sgn = 1.0; }
        else { xa = 0.0; xb = -dt * u / delta; sgn = -1.0; ii = i + 1; }
        ya = 0.0; yb = 1.0;
    } else {
        xa = 0.0; xb = 1.0;
        if (u >= 0.0) { ya = 1.0 - dt * u / delta; yb = 1.0; sgn = 1.0; }
        else { ya = 0.0; yb = -dt * u / delta; sgn = -1.0; jj = j + 1; }
    }
    const long c0 = AT(s, ii, jj);
    const double f = s->vof[c0];
    if (f <= s->cut || f >= (1.0 - s->cut)) return sgn * delta * f * (xb - xa) * (yb - ya);
    const Surf q = surface(s->nrx[c0], s->nry[c0], s->lx[c0], s->ly[c0]);
    const double d0 = s->d[c0];
    if (q.cx == 0.0)      /* |n_x| dominant: analytical integration in x, numerical in y (hazard H15: rm*(ya+yb)) */
        return sgn * delta * num_int(ya, yb, an_int(s, &q, xa, xb, xa, xb, rm * (ya + yb), rm * (yb + ya), q.a10, d0),
                                     an_int(s, &q, xa, xb, xa, xb, rp * (ya + yb), rp * (yb + ya), q.a10, d0));
    return sgn * delta * num_int(xa, xb, an_int(s, &q, ya, yb, rm * (xa + xb), rm * (xa + xb), ya, yb, q.a01, d0),
                                 an_int(s, &q, ya, yb, rp * (xa + xb), rp * (xa + xb), ya, yb, q.a01, d0));
}

static void sweep(FoMF* s, int dir, double dt, const double* src, double* dst) {     /* :457-466, :478-487 */
    const int nx = s->nx, ny = s->ny;
    const double delta = s->delta;
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            double fp, fm, du;
            if (dir == 1) {
                fp = compute_flux(s, 1, i, j, s->u[AT(s, i, j)], dt);
                fm = compute_flux(s, 1, i - 1, j, s->u[AT(s, i - 1, j)], dt);
                du = s->u[AT(s, i, j)] - s->u[AT(s, i - 1, j)];
            } else {
                fp = compute_flux(s, 2, i, j, s->v[AT(s, i, j)], dt);
                fm = compute_flux(s, 2, i, j - 1, s->v[AT(s, i, j - 1)], dt);
                du = s->v[AT(s, i, j)] - s->v[AT(s, i, j - 1)];
            }
            dst[AT(s, i, j)] = (src[AT(s, i, j)] - (fp - fm) / delta) / (1.0 - dt * du / delta);
        }
}

void fomf_advect_vof(FoMF* s, double dt) {                            /* :434-554 */
    const int nx = s->nx, ny = s->ny;
    const double delta = s->delta;
    memset(s->vof1, 0, sizeof(double) * (size_t)s->n);                 /* vof1%allocate, vof2%allocate (:448-449) */
    memset(s->vof2, 0, sizeof(double) * (size_t)s->n);
    get_h_from_vof(s);
    const int d1 = s->x_first ? 1 : 2, d2 = s->x_first ? 2 : 1;
    sweep(s, d1, dt, s->vof, s->vof1);
    memcpy(s->vof, s->vof1, sizeof(double) * (size_t)s->n);            /* vof = vof1 (:470, :510): types too, H13 */
    s->vof_bc_y = 0;
    ghosts(s, s->vof, 'c', s->vof_bc_y);
    get_h_from_vof(s);
    sweep(s, d2, dt, s->vof1, s->vof2);
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            const long c0 = AT(s, i, j);
            const double dux = s->u[c0] - s->u[AT(s, i - 1, j)], dvy = s->v[c0] - s->v[AT(s, i, j - 1)];
            if (s->x_first) s->vof[c0] = s->vof2[c0] - dt * (s->vof1[c0] * dux / delta + s->vof2[c0] * dvy / delta);
            else s->vof[c0] = s->vof2[c0] - dt * (s->vof2[c0] * dux / delta + s->vof1[c0] * dvy / delta);
        }
    s->x_first = !s->x_first;
    ghosts(s, s->vof, 'c', s->vof_bc_y);
}

void fomf_update_material_properties(FoMF* s) {                       /* multiphase.f90:121-137 (whole arrays) */
#pragma omp parallel for schedule(static)
    for (long e = 0; e < s->n; ++e) {
        const double f = s->vof[e];
        s->rho[e] = s->rho1 * f + s->rho0 * (1.0 - f);
        s->mu[e] = s->mu1 * f + s->mu0 * (1.0 - f);
    }
    ghosts(s, s->rho, 'c', 2);
    ghosts(s, s->mu, 'c', 2);
}

/* ---- poisson_solver_pn, src/poisson.f90:306-412 --------------------------------------------------------------------- */
static void poisson_pn(FoMF* s, double* phi) {
    const int nx = s->nx, ny = s->ny, mc = s->mc;
    const double inx = f32(nx);
#pragma omp parallel
    {
        cplx* row = (cplx*)malloc(sizeof(cplx) * (size_t)nx);
#pragma omp for schedule(static)
        for (int j = 1; j <= ny; ++j) {
            for (int i = 0; i < nx; ++i) { row[i].re = phi[AT(s, i + 1, j)]; row[i].im = 0.0; }
            fft(row, nx, -1, s->tw);
            for (int k = 0; k < mc; ++k) {                             /* outc_x = outc_x / float(nx) (:341) */
                s->C[(long)k + (long)mc * (j - 1)].re = row[k].re / inx;
                s->C[(long)k + (long)mc * (j - 1)].im = row[k].im / inx;
            }
        }
        free(row);
    }
    const double *a = s->ta, *b = s->tb, *c = s->tc;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < mc; ++k) {                                     /* one tridiagonal system per kx (:346-385) */
        const double lam = s->mwn_x[k];
        cplx* X = s->C + k;
        double* c1 = s->c1 + k;
        const long st = mc;
        c1[0] = c[0] / (b[0] + lam);
        X[0].re = X[0].re / (b[0] + lam);
        X[0].im = X[0].im / (b[0] + lam);
        for (int j = 1; j < ny - 1; ++j) {
            c1[st * j] = c[j] / (b[j] - a[j] * c1[st * (j - 1)] + lam);
            const double den = b[j] + lam - a[j] * c1[st * (j - 1)];
            X[st * j].re = (X[st * j].re - a[j] * X[st * (j - 1)].re) / den;
            X[st * j].im = (X[st * j].im - a[j] * X[st * (j - 1)].im) / den;
        }
        const int l = ny - 1;
        const double frac = b[l] + lam - a[l] * c1[st * (l - 1)];
        if (frac != 0.0) {
            X[st * l].re = (X[st * l].re - a[l] * X[st * (l - 1)].re) / frac;
            X[st * l].im = (X[st * l].im - a[l] * X[st * (l - 1)].im) / frac;
        } else {
            X[st * l].re = 0.0;
            X[st * l].im = 0.0;
        }
        for (int j = ny - 2; j >= 0; --j) {
            X[st * j].re = X[st * j].re - c1[st * j] * X[st * (j + 1)].re;
            X[st * j].im = X[st * j].im - c1[st * j] * X[st * (j + 1)].im;
        }
    }
#pragma omp parallel
    {
        cplx* row = (cplx*)malloc(sizeof(cplx) * (size_t)nx);
#pragma omp for schedule(static)
        for (int j = 1; j <= ny; ++j) {
            /* FFTW c2r: the half spectrum and its Hermitian image; the imaginary parts of the DC and Nyquist
             * coefficients are ignored */
            for (int k = 0; k < mc; ++k) row[k] = s->C[(long)k + (long)mc * (j - 1)];
            row[0].im = 0.0;
            row[nx / 2].im = 0.0;
            for (int k = mc; k < nx; ++k) { row[k].re = row[nx - k].re; row[k].im = -row[nx - k].im; }
            fft(row, nx, +1, s->tw);
            for (int i = 0; i < nx; ++i) phi[AT(s, i + 1, j)] = row[i].re;
        }
        free(row);
    }
    double mean = 0.0;                                                 /* serial running sum (:398-405) */
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) mean = mean + phi[AT(s, i, j)];
    const double m = mean / f32((long)nx * ny * 1);
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) phi[AT(s, i, j)] = phi[AT(s, i, j)] - m;
}
void fomf_poisson_solve(FoMF* s, double* phi) { poisson_pn(s, phi); }

/* ---- navier_stokes_mod, -DMF ------------------------------------------------------------------------------------------ */
static double sq(double x) { return x * x; }

static void explicit_terms(FoMF* s) {                                 /* compute_explicit_terms, navier_stokes.f90:217-257 */
    const int nx = s->nx, ny = s->ny;
    const double idelta = 1.0 / s->delta;
    const double *u = s->u, *v = s->v, *mu = s->mu, *curv = s->curv, *vof = s->vof;
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; ++j) {
        const int jp = j + 1, jm = j - 1;
        for (int i = 1; i <= nx; ++i) {
            const int ip = i + 1, im = i - 1;
            const long c0 = AT(s, i, j);
            double rx = 0.0, ry = 0.0;
            /* add_advection, :261-353 (2-D) */
            const double uuip = 0.25 * sq(u[AT(s, ip, j)] + u[c0]);
            const double uuim = 0.25 * sq(u[AT(s, im, j)] + u[c0]);
            const double uvjp = (u[AT(s, i, jp)] + u[c0]) * (v[AT(s, ip, j)] + v[c0]) * 0.25;
            const double uvjm = (u[c0] + u[AT(s, i, jm)]) * (v[AT(s, ip, jm)] + v[AT(s, i, jm)]) * 0.25;
            rx = rx - (uuip - uuim) * idelta - (uvjp - uvjm) * idelta;
            const double vuip = (v[AT(s, ip, j)] + v[c0]) * (u[AT(s, i, jp)] + u[c0]) * 0.25;
            const double vuim = (v[c0] + v[AT(s, im, j)]) * (u[AT(s, im, jp)] + u[AT(s, im, j)]) * 0.25;
            const double vvjp = 0.25 * sq(v[AT(s, i, jp)] + v[c0]);
            const double vvjm = 0.25 * sq(v[AT(s, i, jm)] + v[c0]);
            ry = ry - (vuip - vuim) * idelta - (vvjp - vvjm) * idelta;
            /* add_diffusion, variable viscosity, :405-452 */
            const double tauxxip = 2.0 * mu[AT(s, ip, j)] * (u[AT(s, ip, j)] - u[c0]) * idelta;
            const double tauxxim = 2.0 * mu[c0] * (u[c0] - u[AT(s, im, j)]) * idelta;
            const double dtauxxdx = (tauxxip - tauxxim) * idelta;
            const double tauxyjp = 0.25 * (mu[c0] + mu[AT(s, ip, j)] + mu[AT(s, i, jp)] + mu[AT(s, ip, jp)]) *
                                   ((u[AT(s, i, jp)] - u[c0]) * idelta + (v[AT(s, ip, j)] - v[c0]) * idelta);
            const double tauxyjm = 0.25 * (mu[AT(s, i, jm)] + mu[AT(s, ip, jm)] + mu[c0] + mu[AT(s, ip, j)]) *
                                   ((u[c0] - u[AT(s, i, jm)]) * idelta + (v[AT(s, ip, jm)] - v[AT(s, i, jm)]) * idelta);
            const double dtauxydy = (tauxyjp - tauxyjm) * idelta;
            rx = rx + (dtauxxdx + dtauxydy) / s->rfx[c0];
            const double tauyxip = tauxyjp;
            const double tauyxim = 0.25 * (mu[AT(s, im, j)] + mu[c0] + mu[AT(s, im, jp)] + mu[AT(s, i, jp)]) *
                                   ((u[AT(s, im, jp)] - u[AT(s, im, j)]) * idelta + (v[c0] - v[AT(s, im, j)]) * idelta);
            const double dtauyxdx = (tauyxip - tauyxim) * idelta;
            const double tauyyjp = 2.0 * mu[AT(s, i, jp)] * (v[AT(s, i, jp)] - v[c0]) * idelta;
            const double tauyyjm = 2.0 * mu[c0] * (v[c0] - v[AT(s, i, jm)]) * idelta;
            const double dtauyydy = (tauyyjp - tauyyjm) * idelta;
            ry = ry + (dtauyxdx + dtauyydy) / s->rfy[c0];
            /* add_surface_tension, :458-501 */
            rx = rx + s->sigma * 0.5 * (curv[AT(s, ip, j)] + curv[c0]) * (vof[AT(s, ip, j)] - vof[c0]) * idelta / s->rfx[c0];
            ry = ry + s->sigma * 0.5 * (curv[AT(s, i, jp)] + curv[c0]) * (vof[AT(s, i, jp)] - vof[c0]) * idelta / s->rfy[c0];
            /* + S / rhof with S = 0 (:245-255) */
            rx = rx + 0.0 / s->rfx[c0];
            ry = ry + 0.0 / s->rfy[c0];
            s->dvx[c0] = rx;
            s->dvy[c0] = ry;
        }
    }
}

static void predicted_velocity_field(FoMF* s, double dt) {            /* navier_stokes.f90:140-213 */
    const int nx = s->nx, ny = s->ny;
    const double A = 1.0 + 0.5 * dt / s->dt_o, B = -0.5 * dt / s->dt_o, idelta = 1.0 / s->delta;
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; ++j)                                      /* center_to_face(rho, rhof), fields.f90:175-206 */
        for (int i = 1; i <= nx; ++i) {
            const long c0 = AT(s, i, j);
            s->rfx[c0] = 0.5 * (s->rho[AT(s, i + 1, j)] + s->rho[c0]);
            s->rfy[c0] = 0.5 * (s->rho[AT(s, i, j + 1)] + s->rho[c0]);
        }
    explicit_terms(s);
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            const long c0 = AT(s, i, j);
            const double gpx = (s->p[AT(s, i + 1, j)] - s->p[c0]) * idelta, gpy = (s->p[AT(s, i, j + 1)] - s->p[c0]) * idelta;
            const double ghx = (s->phat[AT(s, i + 1, j)] - s->phat[c0]) * idelta;
            const double ghy = (s->phat[AT(s, i, j + 1)] - s->phat[c0]) * idelta;
            double rx = -gpx / s->rfx[c0] + A * s->dvx[c0] + B * s->dvox[c0] + s->g[0];          /* :169 */
            double ry = -gpy / s->rfy[c0] + A * s->dvy[c0] + B * s->dvoy[c0] + s->g[1];          /* :170 */
            rx = rx + gpx / s->rfx[c0] - s->irhomin * gpx - (1.0 / s->rfx[c0] - s->irhomin) * ghx;   /* :176-177 */
            ry = ry + gpy / s->rfy[c0] - s->irhomin * gpy - (1.0 / s->rfy[c0] - s->irhomin) * ghy;   /* :178-179 */
            s->gpx[c0] = rx;
            s->gpy[c0] = ry;
        }
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            const long c0 = AT(s, i, j);
            s->u[c0] = s->u[c0] + dt * s->gpx[c0];                     /* :187-198 */
            s->v[c0] = s->v[c0] + dt * s->gpy[c0];
            s->dvox[c0] = s->dvx[c0];                                  /* :201-205 */
            s->dvoy[c0] = s->dvy[c0];
        }
    ghosts_velocity(s);
}

void fomf_step(FoMF* s, double dt) {                                  /* navier_stokes_solver, :50-136, -DMF */
    const int nx = s->nx, ny = s->ny;
    const double idelta = 1.0 / s->delta;
    fomf_advect_vof(s, dt);                                           /* :82 */
    fomf_update_material_properties(s);                               /* :85 */
#pragma omp parallel for schedule(static)
    for (long e = 0; e < s->n; ++e) s->phat[e] = 2.0 * s->p[e] - s->po[e];   /* :91 (whole arrays) */
    ghosts(s, s->phat, 'c', 2);                                       /* :95 */
    predicted_velocity_field(s, dt);                                  /* :105 */
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; ++j)                                      /* :111-113: phi = div(v) rhomin / dt */
        for (int i = 1; i <= nx; ++i) {
            const long c0 = AT(s, i, j);
            const double dv = (s->u[c0] - s->u[AT(s, i - 1, j)]) * idelta + (s->v[c0] - s->v[AT(s, i, j - 1)]) * idelta;
            s->phi[c0] = dv * s->rhomin / dt;
        }
    poisson_pn(s, s->phi);                                            /* :123 */
    ghosts(s, s->phi, 'c', 2);                                        /* :124 */
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= ny; ++j)                                      /* correct_velocity_field, :505-546 (MF: :527-528) */
        for (int i = 1; i <= nx; ++i) {
            const long c0 = AT(s, i, j);
            const double gx = (s->phi[AT(s, i + 1, j)] - s->phi[c0]) * idelta, gy = (s->phi[AT(s, i, j + 1)] - s->phi[c0]) * idelta;
            s->u[c0] = s->u[c0] - gx * dt * s->irhomin;
            s->v[c0] = s->v[c0] - gy * dt * s->irhomin;
        }
    ghosts_velocity(s);
#pragma omp parallel for schedule(static)
    for (long e = 0; e < s->n; ++e) {                                  /* update_pressure, :550-566: p_o = p; p = p + phi */
        s->po[e] = s->p[e];
        s->p[e] = s->p[e] + s->phi[e];
    }
    ghosts(s, s->p, 'c', 2);
    double maxdiv = -HUGE_VAL, maxvel = 0.0;                           /* checks, :570-619 */
#pragma omp parallel for schedule(static) reduction(max : maxdiv) reduction(max : maxvel)
    for (int j = 1; j <= ny; ++j)
        for (int i = 1; i <= nx; ++i) {
            const long c0 = AT(s, i, j);
            const double dv = (s->u[c0] - s->u[AT(s, i - 1, j)]) * idelta + (s->v[c0] - s->v[AT(s, i, j - 1)]) * idelta;
            const double vel = fabs(s->u[c0]) + fabs(s->v[c0]);
            if (dv > maxdiv) maxdiv = dv;
            if (vel > maxvel) maxvel = vel;
        }
    s->maxdiv = maxdiv;
    s->maxvel = maxvel;
}
