/* fen_oracle_c.c -- plain C (C99 + OpenMP) restatement of FEN's single-phase 3-D fractional step on a fully
 * periodic box (the ppp path of BASELINE configs[1]) and, with foc_set_zwalls, on a channel: x and y periodic,
 * walls in z, ppn Poisson = FFT in x / y + Thomas in z (BASELINE configs[2]).
 *
 * TEST INFRASTRUCTURE ONLY (same rules as oracle/fen_oracle.py): nothing under fen_b200/ may link or call this;
 * it is (1) a second, independent checker -- tests/test_oracle_c.py pins it against the numpy oracle and the
 * committed golden fixture -- and (2) the CPU baseline bench.py times on the GPU box's host cores with all
 * threads (`cpu_baseline.kind = "port"`): the reference itself (Fortran + MPI + FFTW3 + 2decomp) cannot be built
 * in this image, so its loops are restated here one by one, in the reference's operation order.  Paths below are
 * relative to /root/reference.  Third-party arithmetic: FFTW 3.3.10 (INSTALL.sh:16-17) unnormalised forward
 * exp(-i theta) / backward exp(+i theta) transforms are restated by a textbook radix-2 FFT (identical up to
 * round-off, which the 1e-12 parity bound absorbs).
 *
 * Layout: every field is the reference's f(0:nx+1, 0:ny+1, 0:nz+1), x fastest (src/scalar.f90:79-81), i.e. the
 * same memory as the numpy oracle's Fortran-ordered arrays, so Python can hand its arrays in directly.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cplx;

typedef struct FoC {
    int nx, ny, nz;
    long sy, sz, n;                 /* strides and total size of a ghosted field */
    double delta, rho, mu, dt_o;
    double g[3];
    double *p, *phi, *u, *v, *w;    /* gl = 1 */
    double *dvx, *dvy, *dvz, *dvox, *dvoy, *dvoz, *lx, *ly, *lz;   /* same layout, ghosts unused */
    cplx* C;                        /* [nx/2+1][ny][nz], kx fastest */
    int mc;                         /* nx/2 + 1 */
    double *mwn_x, *mwn_y, *mwn_z;  /* modified wavenumbers, poisson.f90:627-629, 645-647, 663-665 */
    cplx *tw_x, *tw_y, *tw_z;       /* exp(-2 pi i m / n) */
    double maxdiv, maxvel;
    int zwall;                      /* bc(5:6) = 'Wall': Neumann p / phi, no-slip velocity, ppn Poisson */
    double *ta, *tb, *tc;           /* tridiagonal coefficients, poisson.f90:744-757 */
} FoC;

#define IDX(s, i, j, k) ((long)(i) + (s)->sy * (long)(j) + (s)->sz * (long)(k))

static double f32(long n) { return (double)(float)n; }      /* Fortran float(n): default real (hazard H1) */

int foc_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void foc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- radix-2 FFT, in place, unnormalised; sign = -1 forward, +1 backward --------------------------------- */
static cplx* make_twiddles(int n) {
    cplx* t = (cplx*)malloc(sizeof(cplx) * (size_t)(n > 0 ? n : 1));
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int m = 0; m < n; ++m) {
        long double a = -two_pi * (long double)m / (long double)n;
        t[m].re = (double)cosl(a);
        t[m].im = (double)sinl(a);
    }
    return t;
}
static void fft(cplx* x, int n, int sign, const cplx* tw) {
    for (int i = 1, j = 0; i < n; ++i) {           /* bit reversal */
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

/* ---- containers --------------------------------------------------------------------------------------- */
static double* new_field(const FoC* s) { return (double*)calloc((size_t)s->n, sizeof(double)); }   /* scalar.f90:84 */

static double* mwn(int n, double delta) {
    const double pi = acos(-1.0);                                  /* global.f90:14 */
    double* m = (double*)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) m[i] = 2.0 * (cos(2.0 * pi * (double)i / f32(n)) - 1.0) / (delta * delta);
    return m;
}

FoC* foc_create(int nx, int ny, int nz, double delta, double rho, double mu) {
    FoC* s = (FoC*)calloc(1, sizeof(FoC));
    s->nx = nx; s->ny = ny; s->nz = nz;
    s->sy = nx + 2; s->sz = (long)(nx + 2) * (ny + 2); s->n = s->sz * (nz + 2);
    s->delta = delta; s->rho = rho; s->mu = mu; s->dt_o = 0.0;
    double** f[] = {&s->p, &s->phi, &s->u, &s->v, &s->w, &s->dvx, &s->dvy, &s->dvz, &s->dvox, &s->dvoy, &s->dvoz,
                    &s->lx, &s->ly, &s->lz};
    for (size_t q = 0; q < sizeof(f) / sizeof(f[0]); ++q) *f[q] = new_field(s);
    s->mc = nx / 2 + 1;
    s->C = (cplx*)malloc(sizeof(cplx) * (size_t)s->mc * ny * nz);
    s->mwn_x = mwn(nx, delta); s->mwn_y = mwn(ny, delta); s->mwn_z = mwn(nz, delta);
    s->tw_x = make_twiddles(nx); s->tw_y = make_twiddles(ny); s->tw_z = make_twiddles(nz);
    return s;
}
void foc_destroy(FoC* s) {
    if (!s) return;
    double* f[] = {s->p, s->phi, s->u, s->v, s->w, s->dvx, s->dvy, s->dvz, s->dvox, s->dvoy, s->dvoz, s->lx, s->ly,
                   s->lz, s->mwn_x, s->mwn_y, s->mwn_z};
    for (size_t q = 0; q < sizeof(f) / sizeof(f[0]); ++q) free(f[q]);
    free(s->C); free(s->tw_x); free(s->tw_y); free(s->tw_z);
    free(s->ta); free(s->tb); free(s->tc);
    free(s);
}

/* grid%setup with bc(5:6) = 'Wall' + init_poisson_ppn's tridiagonal (poisson.f90:744-757): a = c = 1/delta**2,
 * b = -2/delta**2, Neumann folding b(1) += a(1), b(nz) += c(nz), then a(1) = c(nz) = 0 */
void foc_set_zwalls(FoC* s, int on) {
    s->zwall = on ? 1 : 0;
    free(s->ta); free(s->tb); free(s->tc);
    s->ta = s->tb = s->tc = NULL;
    if (!on) return;
    const int n = s->nz;
    const double d2 = s->delta * s->delta;
    s->ta = (double*)malloc(sizeof(double) * (size_t)n);
    s->tb = (double*)malloc(sizeof(double) * (size_t)n);
    s->tc = (double*)malloc(sizeof(double) * (size_t)n);
    for (int k = 0; k < n; ++k) { s->ta[k] = 1.0 / d2; s->tb[k] = -2.0 / d2; s->tc[k] = 1.0 / d2; }
    s->tb[0] = s->tb[0] + s->ta[0];
    s->tb[n - 1] = s->tb[n - 1] + s->tc[n - 1];
    s->ta[0] = 0.0;
    s->tc[n - 1] = 0.0;
}
/* 0 p, 1 phi, 2 u, 3 v, 4 w, 5..7 dv_o */
double* foc_field(FoC* s, int id) {
    double* f[] = {s->p, s->phi, s->u, s->v, s->w, s->dvox, s->dvoy, s->dvoz};
    return (id >= 0 && id < 8) ? f[id] : NULL;
}
long foc_field_size(const FoC* s) { return s->n; }
void foc_set_params(FoC* s, double dt_o, double g0, double g1, double g2) {
    s->dt_o = dt_o; s->g[0] = g0; s->g[1] = g1; s->g[2] = g2;
}

/* scalar%update_ghost_nodes, all faces periodic, single rank (src/scalar.f90:255-388): x-left, x-right over the
 * full (j, k) extent, then y over the full (i, k) extent, then z over full planes -- this order fills the edge and
 * corner ghosts the advection stencil reads (hazard H3). */
void foc_update_ghosts(const FoC* s, double* f) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
#pragma omp parallel for schedule(static)
    for (int k = 0; k <= nz + 1; ++k)
        for (int j = 0; j <= ny + 1; ++j) {
            f[IDX(s, 0, j, k)] = f[IDX(s, nx, j, k)];            /* :257 */
            f[IDX(s, nx + 1, j, k)] = f[IDX(s, 1, j, k)];        /* :276 */
        }
#pragma omp parallel for schedule(static)
    for (int k = 0; k <= nz + 1; ++k)
        for (int i = 0; i <= nx + 1; ++i) {
            f[IDX(s, i, 0, k)] = f[IDX(s, i, ny, k)];            /* :297-300 */
            f[IDX(s, i, ny + 1, k)] = f[IDX(s, i, 1, k)];        /* :322-325 */
        }
#pragma omp parallel for schedule(static)
    for (int j = 0; j <= ny + 1; ++j)
        for (int i = 0; i <= nx + 1; ++i) {
            f[IDX(s, i, j, 0)] = f[IDX(s, i, j, nz)];            /* :348-351 */
            f[IDX(s, i, j, nz + 1)] = f[IDX(s, i, j, 1)];        /* :370-373 */
        }
}

/* update_ghost_nodes with the channel's wiring (navier_stokes.f90:780-1017): x and y periodic as above; at the z
 * walls kind 0 = p, phi: Neumann (scalar.f90:361-362, :384-385); kind 1 = u, v: Dirichlet 0 on a tangential component,
 * ghost = 2 bc - f (:353-354, :375-376); kind 2 = w, the wall-normal component: ghost = bc and, on the back wall, the
 * last interior face too (:355, :377-378; hazard H3).  Periodic box: all kinds are the periodic copy. */
static void update_ghosts_kind(const FoC* s, double* f, int kind) {
    if (!s->zwall) { foc_update_ghosts(s, f); return; }
    const int nx = s->nx, ny = s->ny, nz = s->nz;
#pragma omp parallel for schedule(static)
    for (int k = 0; k <= nz + 1; ++k)
        for (int j = 0; j <= ny + 1; ++j) {
            f[IDX(s, 0, j, k)] = f[IDX(s, nx, j, k)];
            f[IDX(s, nx + 1, j, k)] = f[IDX(s, 1, j, k)];
        }
#pragma omp parallel for schedule(static)
    for (int k = 0; k <= nz + 1; ++k)
        for (int i = 0; i <= nx + 1; ++i) {
            f[IDX(s, i, 0, k)] = f[IDX(s, i, ny, k)];
            f[IDX(s, i, ny + 1, k)] = f[IDX(s, i, 1, k)];
        }
    const double bc = 0.0;
#pragma omp parallel for schedule(static)
    for (int j = 0; j <= ny + 1; ++j)
        for (int i = 0; i <= nx + 1; ++i) {
            if (kind == 0) {
                f[IDX(s, i, j, 0)] = f[IDX(s, i, j, 1)];
                f[IDX(s, i, j, nz + 1)] = f[IDX(s, i, j, nz)];
            } else if (kind == 1) {
                f[IDX(s, i, j, 0)] = 2.0 * bc - f[IDX(s, i, j, 1)];
                f[IDX(s, i, j, nz + 1)] = 2.0 * bc - f[IDX(s, i, j, nz)];
            } else {
                f[IDX(s, i, j, 0)] = bc;
                f[IDX(s, i, j, nz)] = bc;
                f[IDX(s, i, j, nz + 1)] = bc;
            }
        }
}

/* ---- Poisson: poisson_solver_ppn, src/poisson.f90:1038-1173 --------------------------------------------------- */
static void poisson_solve_ppn(FoC* s, double* phi) {
    const int nx = s->nx, ny = s->ny, nz = s->nz, mc = s->mc;
    cplx* C = s->C;
    const double fnx = f32(nx), fny = f32(ny);
#pragma omp parallel
    {
        const int nmax = nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz);
        cplx* line = (cplx*)malloc(sizeof(cplx) * (size_t)nmax);
        cplx* d1 = (cplx*)malloc(sizeof(cplx) * (size_t)nz);
        double* c1 = (double*)malloc(sizeof(double) * (size_t)nz);
        /* r2c along x and outc_x = outc_x/float(nx)  (:1066-1074) */
#pragma omp for collapse(2) schedule(static)
        for (int k = 1; k <= nz; ++k)
            for (int j = 1; j <= ny; ++j) {
                const double* row = phi + IDX(s, 1, j, k);
                for (int i = 0; i < nx; ++i) { line[i].re = row[i]; line[i].im = 0.0; }
                fft(line, nx, -1, s->tw_x);
                cplx* dst = C + (size_t)mc * ((size_t)(j - 1) + (size_t)ny * (k - 1));
                for (int i = 0; i < mc; ++i) { dst[i].re = line[i].re / fnx; dst[i].im = line[i].im / fnx; }
            }
        /* c2c along y and /float(ny)  (:1080-1087) */
#pragma omp for collapse(2) schedule(static)
        for (int k = 0; k < nz; ++k)
            for (int i = 0; i < mc; ++i) {
                cplx* base = C + i + (size_t)mc * ny * k;
                for (int j = 0; j < ny; ++j) line[j] = base[(size_t)mc * j];
                fft(line, ny, -1, s->tw_y);
                for (int j = 0; j < ny; ++j) { base[(size_t)mc * j].re = line[j].re / fny; base[(size_t)mc * j].im = line[j].im / fny; }
            }
        /* Thomas along z, one (i, j) system at a time, the reference's expressions (:1092-1135) */
        const double *a = s->ta, *b = s->tb, *c = s->tc;
#pragma omp for collapse(2) schedule(static)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < mc; ++i) {
                cplx* base = C + i + (size_t)mc * j;
                const size_t st = (size_t)mc * ny;
                const double mx = s->mwn_x[i], my = s->mwn_y[j];
                double factor = 1.0 / (b[0] + mx + my);                                  /* :1096 */
                c1[0] = c[0] * factor;
                d1[0].re = base[0].re * factor; d1[0].im = base[0].im * factor;
                for (int k = 1; k < nz - 1; ++k) {                                       /* :1102-1110 */
                    factor = 1.0 / (b[k] + mx + my - a[k] * c1[k - 1]);
                    c1[k] = c[k] * factor;
                    d1[k].re = (base[st * k].re - a[k] * d1[k - 1].re) * factor;
                    d1[k].im = (base[st * k].im - a[k] * d1[k - 1].im) * factor;
                }
                if (nz > 1) {                                                            /* :1112-1121 */
                    const int k = nz - 1;
                    factor = (b[k] + mx + my - a[k] * c1[k - 1]);
                    if (factor != 0.0) {
                        d1[k].re = (base[st * k].re - a[k] * d1[k - 1].re) / factor;
                        d1[k].im = (base[st * k].im - a[k] * d1[k - 1].im) / factor;
                    } else {
                        d1[k].re = 0.0; d1[k].im = 0.0;                                  /* exact-zero pivot, hazard H5 */
                    }
                }
                base[st * (nz - 1)] = d1[nz - 1];                                        /* :1124-1128 */
                for (int k = nz - 2; k >= 0; --k) {                                      /* :1129-1135 */
                    base[st * k].re = d1[k].re - c1[k] * base[st * (k + 1)].re;
                    base[st * k].im = d1[k].im - c1[k] * base[st * (k + 1)].im;
                }
            }
        /* inverse y (:1141-1145) */
#pragma omp for collapse(2) schedule(static)
        for (int k = 0; k < nz; ++k)
            for (int i = 0; i < mc; ++i) {
                cplx* base = C + i + (size_t)mc * ny * k;
                for (int j = 0; j < ny; ++j) line[j] = base[(size_t)mc * j];
                fft(line, ny, +1, s->tw_y);
                for (int j = 0; j < ny; ++j) base[(size_t)mc * j] = line[j];
            }
        /* phi%f = 0, then c2r along x (:1151-1156) */
#pragma omp for schedule(static)
        for (long q = 0; q < s->n; ++q) phi[q] = 0.0;
#pragma omp for collapse(2) schedule(static)
        for (int k = 1; k <= nz; ++k)
            for (int j = 1; j <= ny; ++j) {
                const cplx* src = C + (size_t)mc * ((size_t)(j - 1) + (size_t)ny * (k - 1));
                for (int i = 0; i < mc; ++i) line[i] = src[i];
                line[0].im = 0.0;
                line[nx / 2].im = 0.0;
                for (int i = 1; i < nx - nx / 2; ++i) { line[nx - i].re = src[i].re; line[nx - i].im = -src[i].im; }
                fft(line, nx, +1, s->tw_x);
                double* row = phi + IDX(s, 1, j, k);
                for (int i = 0; i < nx; ++i) row[i] = line[i].re;
            }
        free(line); free(d1); free(c1);
    }
    /* mean removal: the serial running sum of one rank, then phi%f = phi%f - mean/float(nx*ny*nz) over the WHOLE array
     * (:1159-1171; hazards H4, H8) */
    double mean_phi = 0.0;
    for (int k = 1; k <= nz; ++k)
        for (int j = 1; j <= ny; ++j) {
            const double* row = phi + IDX(s, 1, j, k);
            for (int i = 0; i < nx; ++i) mean_phi = mean_phi + row[i];
        }
    const double shift = mean_phi / f32((long)nx * ny * nz);
#pragma omp parallel for schedule(static)
    for (long q = 0; q < s->n; ++q) phi[q] = phi[q] - shift;
}

/* ---- Poisson: poisson_solver_ppp, src/poisson.f90:941-1034 -------------------------------------------------- */
void foc_poisson_solve(FoC* s, double* phi) {
    if (s->zwall) { poisson_solve_ppn(s, phi); return; }
    const int nx = s->nx, ny = s->ny, nz = s->nz, mc = s->mc;
    cplx* C = s->C;
    /* r2c along x, one line per (j, k)  (:965-969); only nx/2+1 outputs are defined (hazard H7) */
#pragma omp parallel
    {
        cplx* line = (cplx*)malloc(sizeof(cplx) * (size_t)(nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz)));
#pragma omp for collapse(2) schedule(static)
        for (int k = 1; k <= nz; ++k)
            for (int j = 1; j <= ny; ++j) {
                const double* row = phi + IDX(s, 1, j, k);
                for (int i = 0; i < nx; ++i) { line[i].re = row[i]; line[i].im = 0.0; }
                fft(line, nx, -1, s->tw_x);
                memcpy(C + (size_t)mc * ((size_t)(j - 1) + (size_t)ny * (k - 1)), line, sizeof(cplx) * (size_t)mc);
            }
        /* c2c along y (:975-979) */
#pragma omp for collapse(2) schedule(static)
        for (int k = 0; k < nz; ++k)
            for (int i = 0; i < mc; ++i) {
                cplx* base = C + i + (size_t)mc * ny * k;
                for (int j = 0; j < ny; ++j) line[j] = base[(size_t)mc * j];
                fft(line, ny, -1, s->tw_y);
                for (int j = 0; j < ny; ++j) base[(size_t)mc * j] = line[j];
            }
        /* c2c along z, normalise, divide by the modified wavenumbers, inverse z (:985-1012) */
        const double norm = f32((long)nx * ny * nz);                     /* :992 float(nx*ny*nz) */
#pragma omp for collapse(2) schedule(static)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < mc; ++i) {
                cplx* base = C + i + (size_t)mc * j;
                const size_t st = (size_t)mc * ny;
                for (int k = 0; k < nz; ++k) line[k] = base[st * k];
                fft(line, nz, -1, s->tw_z);
                for (int k = 0; k < nz; ++k) {
                    const double lam = s->mwn_x[i] + s->mwn_y[j] + s->mwn_z[k];
                    if (lam == 0.0) { line[k].re = 0.0; line[k].im = 0.0; }      /* :998-999 */
                    else { line[k].re = (line[k].re / norm) / lam; line[k].im = (line[k].im / norm) / lam; }
                }
                fft(line, nz, +1, s->tw_z);
                for (int k = 0; k < nz; ++k) base[st * k] = line[k];
            }
        /* inverse y (:1018-1022) */
#pragma omp for collapse(2) schedule(static)
        for (int k = 0; k < nz; ++k)
            for (int i = 0; i < mc; ++i) {
                cplx* base = C + i + (size_t)mc * ny * k;
                for (int j = 0; j < ny; ++j) line[j] = base[(size_t)mc * j];
                fft(line, ny, +1, s->tw_y);
                for (int j = 0; j < ny; ++j) base[(size_t)mc * j] = line[j];
            }
        /* c2r along x (:1028-1032): Hermitian completion; the imaginary parts of DC / Nyquist are ignored */
#pragma omp for collapse(2) schedule(static)
        for (int k = 1; k <= nz; ++k)
            for (int j = 1; j <= ny; ++j) {
                const cplx* src = C + (size_t)mc * ((size_t)(j - 1) + (size_t)ny * (k - 1));
                for (int i = 0; i < mc; ++i) line[i] = src[i];
                line[0].im = 0.0;
                line[nx / 2].im = 0.0;
                for (int i = 1; i < nx - nx / 2; ++i) { line[nx - i].re = src[i].re; line[nx - i].im = -src[i].im; }
                fft(line, nx, +1, s->tw_x);
                double* row = phi + IDX(s, 1, j, k);
                for (int i = 0; i < nx; ++i) row[i] = line[i].re;
            }
        free(line);
    }
}

/* ---- navier_stokes_mod ------------------------------------------------------------------------------------ */
static double sq(double x) { return x * x; }

/* compute_explicit_terms (navier_stokes.f90:217-257): dv = 0; add_advection (:261-353); add_diffusion const-mu
 * branch (:384-404 -> laplacian_of_vector, fields.f90:298-343); S = 0 */
static void explicit_terms(FoC* s) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    const double id = 1.0 / s->delta, id2 = 1.0 / (s->delta * s->delta);
    const double* u = s->u; const double* v = s->v; const double* w = s->w;
    const long sy = s->sy, sz = s->sz;
    const double rf = 0.5 * (s->rho + s->rho);                       /* center_to_face, fields.f90:197-200 */
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= nz; ++k)
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i) {
                const long c = IDX(s, i, j, k);
                /* x component, :297-310 */
                double uuip = 0.25 * sq(u[c + 1] + u[c]);
                double uuim = 0.25 * sq(u[c - 1] + u[c]);
                double uvjp = (u[c + sy] + u[c]) * (v[c + 1] + v[c]) * 0.25;
                double uvjm = (u[c] + u[c - sy]) * (v[c + 1 - sy] + v[c - sy]) * 0.25;
                double uwkp = (u[c + sz] + u[c]) * (w[c + 1] + w[c]) * 0.25;
                double uwkm = (u[c] + u[c - sz]) * (w[c + 1 - sz] + w[c - sz]) * 0.25;
                double dx = 0.0 - (uuip - uuim) * id - (uvjp - uvjm) * id;
                dx = dx - (uwkp - uwkm) * id;
                /* y component, :315-327 */
                double vuip = (v[c + 1] + v[c]) * (u[c + sy] + u[c]) * 0.25;
                double vuim = (v[c] + v[c - 1]) * (u[c - 1 + sy] + u[c - 1]) * 0.25;
                double vvjp = 0.25 * sq(v[c + sy] + v[c]);
                double vvjm = 0.25 * sq(v[c - sy] + v[c]);
                double vwkp = (v[c + sz] + v[c]) * (w[c + sy] + w[c]) * 0.25;
                double vwkm = (v[c] + v[c - sz]) * (w[c + sy - sz] + w[c - sz]) * 0.25;
                double dy = 0.0 - (vuip - vuim) * id - (vvjp - vvjm) * id;
                dy = dy - (vwkp - vwkm) * id;
                /* z component, :334-347 */
                double wuip = (w[c] + w[c + 1]) * (u[c] + u[c + sz]) * 0.25;
                double wuim = (w[c] + w[c - 1]) * (u[c - 1] + u[c - 1 + sz]) * 0.25;
                double wvjp = (w[c] + w[c + sy]) * (v[c] + v[c + sz]) * 0.25;
                double wvjm = (w[c] + w[c - sy]) * (v[c - sy] + v[c - sy + sz]) * 0.25;
                double wwkp = (w[c] + w[c + sz]) * (w[c] + w[c + sz]) * 0.25;
                double wwkm = (w[c] + w[c - sz]) * (w[c] + w[c - sz]) * 0.25;
                double dz = 0.0 - (wuip - wuim) * id - (wvjp - wvjm) * id - (wwkp - wwkm) * id;
                /* laplacian_of_vector, fields.f90:326-337; dv += mu*lap/rhof, navier_stokes.f90:394-397 */
                double lx = ((u[c + 1] - 2.0 * u[c] + u[c - 1]) + (u[c + sy] - 2.0 * u[c] + u[c - sy])) * id2 +
                            (u[c + sz] - 2.0 * u[c] + u[c - sz]) * id2;
                double ly = ((v[c + 1] - 2.0 * v[c] + v[c - 1]) + (v[c + sy] - 2.0 * v[c] + v[c - sy])) * id2 +
                            (v[c + sz] - 2.0 * v[c] + v[c - sz]) * id2;
                double lz = ((w[c + 1] - 2.0 * w[c] + w[c - 1]) + (w[c + sy] - 2.0 * w[c] + w[c - sy]) +
                             (w[c + sz] - 2.0 * w[c] + w[c - sz])) * id2;
                s->dvx[c] = dx + s->mu * lx / rf;
                s->dvy[c] = dy + s->mu * ly / rf;
                s->dvz[c] = dz + s->mu * lz / rf;
            }
}

/* predicted_velocity_field, navier_stokes.f90:140-213 */
static void predicted_velocity_field(FoC* s, double dt) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    const double A = 1.0 + 0.5 * dt / s->dt_o, B = -0.5 * dt / s->dt_o;        /* :157-158 */
    const double id = 1.0 / s->delta;
    const double rf = 0.5 * (s->rho + s->rho);
    const long sy = s->sy, sz = s->sz;
    explicit_terms(s);                                                          /* :164 */
    const double* p = s->p;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= nz; ++k)
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i) {
                const long c = IDX(s, i, j, k);
                /* gradient(p), fields.f90:55-58; RHS = -grad_p/rhof + A dv + B dv_o + g, :169-172; v += dt RHS */
                const double gx = (p[c + 1] - p[c]) * id, gy = (p[c + sy] - p[c]) * id, gz = (p[c + sz] - p[c]) * id;
                const double rx = -gx / rf + A * s->dvx[c] + B * s->dvox[c] + s->g[0];
                const double ry = -gy / rf + A * s->dvy[c] + B * s->dvoy[c] + s->g[1];
                const double rz = -gz / rf + A * s->dvz[c] + B * s->dvoz[c] + s->g[2];
                s->lx[c] = s->u[c] + dt * rx;           /* into temporaries: neighbours still read the old field */
                s->ly[c] = s->v[c] + dt * ry;
                s->lz[c] = s->w[c] + dt * rz;
                s->dvox[c] = s->dvx[c];                  /* :201-205 */
                s->dvoy[c] = s->dvy[c];
                s->dvoz[c] = s->dvz[c];
            }
    double* t;
    t = s->u; s->u = s->lx; s->lx = t;
    t = s->v; s->v = s->ly; s->ly = t;
    t = s->w; s->w = s->lz; s->lz = t;
    update_ghosts_kind(s, s->u, 1); update_ghosts_kind(s, s->v, 1); update_ghosts_kind(s, s->w, 2);   /* :208 */
}

/* navier_stokes_solver, navier_stokes.f90:50-136 (constant_CFL off) */
void foc_step(FoC* s, double dt) {
    const int nx = s->nx, ny = s->ny, nz = s->nz;
    const double id = 1.0 / s->delta;
    const long sy = s->sy, sz = s->sz;
    const double rf = 0.5 * (s->rho + s->rho);
    predicted_velocity_field(s, dt);                                          /* :105 */
    /* divergence(v, phi); phi = phi*rho/dt  (fields.f90:144-147, navier_stokes.f90:111-121) */
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= nz; ++k)
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i) {
                const long c = IDX(s, i, j, k);
                double d = (s->u[c] - s->u[c - 1]) * id + (s->v[c] - s->v[c - sy]) * id;
                d = d + (s->w[c] - s->w[c - sz]) * id;
                s->phi[c] = d * s->rho / dt;
            }
    foc_poisson_solve(s, s->phi);                                             /* :123 */
    update_ghosts_kind(s, s->phi, 0);                                         /* :124 */
    /* correct_velocity_field (:505-546): v -= grad(phi)*dt/rhof ; update_pressure (:550-566): p += phi */
    const double* f = s->phi;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= nz; ++k)
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i) {
                const long c = IDX(s, i, j, k);
                s->u[c] = s->u[c] - ((f[c + 1] - f[c]) * id) * dt / rf;
                s->v[c] = s->v[c] - ((f[c + sy] - f[c]) * id) * dt / rf;
                s->w[c] = s->w[c] - ((f[c + sz] - f[c]) * id) * dt / rf;
                s->p[c] = s->p[c] + f[c];
            }
    update_ghosts_kind(s, s->u, 1); update_ghosts_kind(s, s->v, 1); update_ghosts_kind(s, s->w, 2);   /* :544 */
    update_ghosts_kind(s, s->p, 0);                                           /* :564 */
    /* checks (:570-619): signed max of the divergence (H6), max |u|+|v|+|w| */
    double md = -1.0e300, mv = 0.0;
#pragma omp parallel for collapse(2) schedule(static) reduction(max : md, mv)
    for (int k = 1; k <= nz; ++k)
        for (int j = 1; j <= ny; ++j)
            for (int i = 1; i <= nx; ++i) {
                const long c = IDX(s, i, j, k);
                double d = (s->u[c] - s->u[c - 1]) * id + (s->v[c] - s->v[c - sy]) * id;
                d = d + (s->w[c] - s->w[c - sz]) * id;
                if (d > md) md = d;
                const double vel = fabs(s->u[c]) + fabs(s->v[c]) + fabs(s->w[c]);
                if (vel > mv) mv = vel;
            }
    s->maxdiv = md;
    s->maxvel = mv;
}
double foc_maxdiv(const FoC* s) { return s->maxdiv; }
double foc_maxcfl(const FoC* s, double dt) { return dt * s->maxvel / s->delta; }     /* :617 */

/* copy a Fortran-ordered ghosted array in / out (the pointers foc_field returns move when fields ping-pong) */
void foc_set_field(FoC* s, int id, const double* src) { memcpy(foc_field(s, id), src, sizeof(double) * (size_t)s->n); }
void foc_get_field(FoC* s, int id, double* dst) { memcpy(dst, foc_field(s, id), sizeof(double) * (size_t)s->n); }
