/* oracle.c -- CPU ORACLE (test infrastructure, not product code): plain-C, single-precision, OpenMP
 * restatement of the reference's element stiffness + Newmark hot path, used
 *   (1) as the timed CPU baseline of bench.py ("cpu_baseline", "--impl reference", kind = "port": the
 *       reference itself cannot be built here -- no Eigen/FFTW/MPI/Boost/NetCDF), and
 *   (2) as an independent cross-check of oracle/axisem_oracle.py (tests/test_c_oracle.py).
 * Pinned by tests/test_golden_reference.py against vectors produced by the reference's own sources (oracle/_ref, built by
 * oracle/Makefile.ref against the Eigen/FFTW stand-ins of oracle/shim); the reference holds no golden vectors itself.
 * Only tests/, __graft_entry__.smoke() and bench.py may load this library.
 *
 * It consumes the flattened per-group arrays of OracleDomain (axisem_oracle.py), so both oracles share
 * inputs.  Parallelism mirrors the reference's "one MPI rank per core, elements independent, additive
 * assembly": OpenMP over elements, atomic adds into the shared point stiffness.
 * FFTW (3.3.4, un-vendored dependency of the reference) is replaced by a mixed-radix complex FFT;
 * two real columns are transformed per complex FFT.  S/ = /root/reference/SOLVER/src/.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float re, im; } cpx;

static inline cpx c_add(cpx a, cpx b) { cpx r = {a.re + b.re, a.im + b.im}; return r; }
static inline cpx c_sub(cpx a, cpx b) { cpx r = {a.re - b.re, a.im - b.im}; return r; }
static inline cpx c_mul(cpx a, cpx b) { cpx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static inline cpx c_scl(cpx a, float s) { cpx r = {a.re * s, a.im * s}; return r; }
static inline cpx c_ia(cpx a, float al) { cpx r = {-al * a.im, al * a.re}; return r; }   /* i*alpha*a */
static const cpx C0 = {0.f, 0.f};

/* ------------------------------------------------------------------ mixed-radix complex FFT */
#define MAXFAC 32
typedef struct {
    int n, nfac, fac[2 * MAXFAC];
    cpx *tw;     /* exp(-2 pi i k / n) */
} fft_plan;

static void plan_init(fft_plan *p, int n) {
    p->n = n;
    p->tw = (cpx *)malloc(sizeof(cpx) * (size_t)n);
    for (int k = 0; k < n; ++k) {
        double a = -2.0 * M_PI * k / n;
        p->tw[k].re = (float)cos(a);
        p->tw[k].im = (float)sin(a);
    }
    int m = n, nf = 0, f = 4;
    while (m > 1) {
        while (m % f) {
            if (f == 4) f = 2;
            else if (f == 2) f = 3;
            else f += 2;
            if (f > 32000 || (long)f * f > m) f = m;
        }
        m /= f;
        p->fac[nf++] = f;
        p->fac[nf++] = m;
    }
    p->nfac = nf / 2;
}
static void plan_free(fft_plan *p) { free(p->tw); }

static void bfly2(cpx *F, size_t fs, const fft_plan *st, int m, int inv) {
    for (int u = 0; u < m; ++u) {
        cpx w = st->tw[fs * u];
        if (inv) w.im = -w.im;
        cpx t = c_mul(F[u + m], w);
        F[u + m] = c_sub(F[u], t);
        F[u] = c_add(F[u], t);
    }
}
static void bfly4(cpx *F, size_t fs, const fft_plan *st, int m, int inv) {
    for (int u = 0; u < m; ++u) {
        cpx w1 = st->tw[fs * u], w2 = st->tw[2 * fs * u], w3 = st->tw[3 * fs * u];
        if (inv) { w1.im = -w1.im; w2.im = -w2.im; w3.im = -w3.im; }
        cpx a0 = F[u], a1 = c_mul(F[u + m], w1), a2 = c_mul(F[u + 2 * m], w2), a3 = c_mul(F[u + 3 * m], w3);
        cpx s02 = c_add(a0, a2), d02 = c_sub(a0, a2), s13 = c_add(a1, a3), d13 = c_sub(a1, a3);
        cpx r;   /* -i * d13 (forward) or +i * d13 (inverse) */
        if (!inv) { r.re = d13.im; r.im = -d13.re; } else { r.re = -d13.im; r.im = d13.re; }
        F[u] = c_add(s02, s13);
        F[u + 2 * m] = c_sub(s02, s13);
        F[u + m] = c_add(d02, r);
        F[u + 3 * m] = c_sub(d02, r);
    }
}
static void bfly_generic(cpx *F, size_t fs, const fft_plan *st, int m, int p, int inv) {
    cpx scratch[64];
    const int n = st->n;
    for (int u = 0; u < m; ++u) {
        for (int q = 0; q < p; ++q) scratch[q] = F[u + q * m];
        int k = u;
        for (int q1 = 0; q1 < p; ++q1) {
            size_t twi = 0;
            cpx acc = scratch[0];
            for (int q = 1; q < p; ++q) {
                twi += fs * (size_t)k;
                if (twi >= (size_t)n) twi %= (size_t)n;
                cpx w = st->tw[twi];
                if (inv) w.im = -w.im;
                acc = c_add(acc, c_mul(scratch[q], w));
            }
            F[k] = acc;
            k += m;
        }
    }
}
static void fft_work(cpx *Fout, const cpx *f, size_t fs, int in_stride, const int *fac, const fft_plan *st, int inv) {
    const int p = fac[0], m = fac[1];
    if (m == 1) {
        for (int i = 0; i < p; ++i) Fout[i] = f[(size_t)i * fs * in_stride];
    } else {
        for (int i = 0; i < p; ++i) fft_work(Fout + (size_t)i * m, f + (size_t)i * fs * in_stride, fs * p, in_stride, fac + 2, st, inv);
    }
    if (p == 2) bfly2(Fout, fs, st, m, inv);
    else if (p == 4) bfly4(Fout, fs, st, m, inv);
    else bfly_generic(Fout, fs, st, m, p, inv);
}
static void fft_exec(const fft_plan *st, const cpx *in, cpx *out, int inv) {
    if (st->n == 1) { out[0] = in[0]; return; }
    fft_work(out, in, 1, 1, st->fac, st, inv);
}

/* ------------------------------------------------------------------ group description (mirrors _Group) */
typedef struct {
    int E, M, Nr, axial, nyq, fluid, law, is3d, tiso, ncoef;
    int att_kind, nsls, do_kappa, P;      /* att_kind: 0 none, 1 full, 2 cg4 */
    const int *pidx;                      /* [E][25] row into the point arrays */
    const float *dsdxii, *dsdeta, *dzdxii, *dzdeta, *inv_s;   /* [E][25] each */
    const double *theta;                  /* [E][25] or NULL */
    const float *coef;                    /* [ncoef][E][rows][25] (fluid: K [E][rows][25]) */
    const float *alpha, *beta, *gamma;    /* [E][nsls] */
    const float *dk3, *dmu, *dmu2;        /* [E][rows][P] */
    float *memvar;                        /* [nsls][E][R][6][P] (x2 floats when 1D) */
    float *stressR;                       /* [E][R][6][P] */
} orc_group;

static const int CG4[4] = {6, 8, 16, 18};

/* Gradient::computeGrad6 (S/core/element/grad/Gradient.cpp:206-265) for one mode */
static void grad6_mode(const float *Gxi, const float *Geta, const float *g0, const float *g1, const float *g2, const float *g3,
                       const float *g4, int axial, int alpha, const cpx *u /*[3][25]*/, cpx *e /*[6][25]*/) {
    const float *dsdxii = g0, *dsdeta = g1, *dzdxii = g2, *dzdeta = g3, *inv_s = g4;
    cpx GU[3][25], UG[3][25];
    const float al = (float)alpha;
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) {
                cpx a = C0, b = C0;
                for (int k = 0; k < 5; ++k) {
                    a = c_add(a, c_scl(u[c * 25 + k * 5 + j], Gxi[k * 5 + i]));
                    b = c_add(b, c_scl(u[c * 25 + i * 5 + k], Geta[k * 5 + j]));
                }
                GU[c][i * 5 + j] = a;
                UG[c][i * 5 + j] = b;
            }
    for (int p = 0; p < 25; ++p) {
        cpx ds[3], dz[3];
        for (int c = 0; c < 3; ++c) {
            ds[c] = c_add(c_scl(GU[c][p], dzdeta[p]), c_scl(UG[c][p], dzdxii[p]));
            dz[c] = c_add(c_scl(GU[c][p], dsdeta[p]), c_scl(UG[c][p], dsdxii[p]));
        }
        cpx v0 = c_add(u[p], c_ia(u[25 + p], al)), v1 = c_sub(c_ia(u[p], al), u[25 + p]), v2 = c_ia(u[50 + p], al);
        e[0 * 25 + p] = ds[0];
        e[1 * 25 + p] = c_scl(v0, inv_s[p]);
        e[2 * 25 + p] = dz[2];
        e[3 * 25 + p] = c_add(dz[1], c_scl(v2, inv_s[p]));
        e[4 * 25 + p] = c_add(dz[0], ds[2]);
        e[5 * 25 + p] = c_add(ds[1], c_scl(v1, inv_s[p]));
    }
    if (axial) {
        for (int j = 0; j < 5; ++j) {
            cpx gv0 = c_add(GU[0][j], c_ia(GU[1][j], al)), gv1 = c_sub(c_ia(GU[0][j], al), GU[1][j]), gv2 = c_ia(GU[2][j], al);
            e[1 * 25 + j] = c_add(e[1 * 25 + j], c_scl(gv0, dzdeta[j]));
            e[5 * 25 + j] = c_add(e[5 * 25 + j], c_scl(gv1, dzdeta[j]));
            e[3 * 25 + j] = c_add(e[3 * 25 + j], c_scl(gv2, dzdeta[j]));
            if (alpha == 1) {
                cpx uv0 = c_add(UG[0][j], c_ia(UG[1][j], al)), uv1 = c_sub(c_ia(UG[0][j], al), UG[1][j]);
                e[1 * 25 + j] = c_add(e[1 * 25 + j], c_scl(uv0, dzdxii[j]));
                e[5 * 25 + j] = c_add(e[5 * 25 + j], c_scl(uv1, dzdxii[j]));
            }
        }
    }
}

/* Gradient::computeQuad6 (Gradient.cpp:267-322) for one mode */
static void quad6_mode(const float *Gxi, const float *Geta, const float *g0, const float *g1, const float *g2, const float *g3,
                       const float *g4, int axial, int beta, const cpx *s /*[6][25]*/, cpx *f /*[3][25]*/) {
    const float *dsdxii = g0, *dsdeta = g1, *dzdxii = g2, *dzdeta = g3, *inv_s = g4;
    const float be = -(float)beta;
    const int pa[3] = {0, 5, 4}, pb[3] = {4, 3, 2};
    cpx g[3][25], X[3][25], Y[3][25];
    for (int p = 0; p < 25; ++p) {
        g[0][p] = c_add(s[1 * 25 + p], c_ia(s[5 * 25 + p], be));
        g[1][p] = c_sub(c_ia(s[1 * 25 + p], be), s[5 * 25 + p]);
        g[2][p] = c_ia(s[3 * 25 + p], be);
        for (int c = 0; c < 3; ++c) {
            X[c][p] = c_add(c_scl(s[pa[c] * 25 + p], dzdeta[p]), c_scl(s[pb[c] * 25 + p], dsdeta[p]));
            Y[c][p] = c_add(c_scl(s[pa[c] * 25 + p], dzdxii[p]), c_scl(s[pb[c] * 25 + p], dsdxii[p]));
        }
    }
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 5; ++i)
            for (int j = 0; j < 5; ++j) {
                cpx a = c_scl(g[c][i * 5 + j], inv_s[i * 5 + j]);
                for (int k = 0; k < 5; ++k) {
                    a = c_add(a, c_scl(X[c][k * 5 + j], Gxi[i * 5 + k]));
                    a = c_add(a, c_scl(Y[c][i * 5 + k], Geta[j * 5 + k]));
                }
                if (axial) {
                    a = c_add(a, c_scl(g[c][j], Gxi[i * 5 + 0] * dzdeta[j]));
                    if (beta == 1 && c < 2 && i == 0)
                        for (int k = 0; k < 5; ++k) a = c_add(a, c_scl(g[c][k], dzdxii[k] * Geta[j * 5 + k]));
                }
                f[c * 25 + i * 5 + j] = a;
            }
}

/* fluid: Gradient::computeGrad / computeQuad (Gradient.cpp:26-82) */
static void grad_fluid_mode(const float *Gxi, const float *Geta, const float *g0, const float *g1, const float *g2, const float *g3,
                            const float *g4, int axial, int alpha, const cpx *u /*[25]*/, cpx *e /*[3][25]*/) {
    const float *dsdxii = g0, *dsdeta = g1, *dzdxii = g2, *dzdeta = g3, *inv_s = g4;
    const float al = (float)alpha;
    cpx GU[25], UG[25];
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) {
            cpx a = C0, b = C0;
            for (int k = 0; k < 5; ++k) {
                a = c_add(a, c_scl(u[k * 5 + j], Gxi[k * 5 + i]));
                b = c_add(b, c_scl(u[i * 5 + k], Geta[k * 5 + j]));
            }
            GU[i * 5 + j] = a;
            UG[i * 5 + j] = b;
        }
    for (int p = 0; p < 25; ++p) {
        e[p] = c_add(c_scl(GU[p], dzdeta[p]), c_scl(UG[p], dzdxii[p]));
        e[25 + p] = c_scl(c_ia(u[p], al), inv_s[p]);
        e[50 + p] = c_add(c_scl(GU[p], dsdeta[p]), c_scl(UG[p], dsdxii[p]));
    }
    if (axial)
        for (int j = 0; j < 5; ++j) e[25 + j] = c_add(e[25 + j], c_scl(c_ia(GU[j], al), dzdeta[j]));
}
static void quad_fluid_mode(const float *Gxi, const float *Geta, const float *g0, const float *g1, const float *g2, const float *g3,
                            const float *g4, int axial, int beta, const cpx *s /*[3][25]*/, cpx *f /*[25]*/) {
    const float *dsdxii = g0, *dsdeta = g1, *dzdxii = g2, *dzdeta = g3, *inv_s = g4;
    const float be = -(float)beta;
    cpx g[25], X[25], Y[25];
    for (int p = 0; p < 25; ++p) {
        g[p] = c_ia(s[25 + p], be);
        X[p] = c_add(c_scl(s[p], dzdeta[p]), c_scl(s[50 + p], dsdeta[p]));
        Y[p] = c_add(c_scl(s[p], dzdxii[p]), c_scl(s[50 + p], dsdxii[p]));
    }
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) {
            cpx a = c_scl(g[i * 5 + j], inv_s[i * 5 + j]);
            for (int k = 0; k < 5; ++k) {
                a = c_add(a, c_scl(X[k * 5 + j], Gxi[i * 5 + k]));
                a = c_add(a, c_scl(Y[i * 5 + k], Geta[j * 5 + k]));
            }
            if (axial) a = c_add(a, c_scl(g[j], Gxi[i * 5 + 0] * dzdeta[j]));
            f[i * 5 + j] = a;
        }
}

/* CrdTransTIsoSolid (S/core/element/crd/CrdTransTIsoSolid.cpp:14-42), real or complex via two passes on floats */
static void rot_fwd(float *u[6], int n, float s1, float c1, float s2, float c2) {
    for (int k = 0; k < n; ++k) {
        float sum = u[0][k] + u[2][k], dif = u[0][k] - u[2][k], u3 = u[3][k];
        u[0][k] = 0.5f * (sum + c2 * dif - s2 * u[4][k]);
        u[2][k] = sum - u[0][k];
        u[4][k] = c2 * u[4][k] + s2 * dif;
        u[3][k] = c1 * u3 + s1 * u[5][k];
        u[5][k] = c1 * u[5][k] - s1 * u3;
    }
}
static void rot_bwd(float *u[6], int n, float s1, float c1, float s2, float c2) {
    for (int k = 0; k < n; ++k) {
        float sum = u[0][k] + u[2][k], dif = (u[0][k] - u[2][k]) * 0.5f, u3 = u[3][k];
        u[0][k] = 0.5f * sum + c2 * dif + s2 * u[4][k];
        u[2][k] = sum - u[0][k];
        u[4][k] = c2 * u[4][k] - s2 * dif;
        u[3][k] = c1 * u3 - s1 * u[5][k];
        u[5][k] = c1 * u[5][k] + s1 * u3;
    }
}

/* constitutive law on `n` scalars per component (real samples, or re/im parts of modes) */
static void stress_n(int law, const float *c /*ncoef values at this point[, phi] with stride cs*/, size_t cs, float *const e[6],
                     float *const s[6], int n, size_t estride) {
    for (int k = 0; k < n; ++k) {
        const size_t o = (size_t)k * estride;
        float E0 = e[0][o], E1 = e[1][o], E2 = e[2][o], E3 = e[3][o], E4 = e[4][o], E5 = e[5][o];
        if (law == 0) {             /* Isotropic1D.cpp:9-26 / Isotropic3D.cpp:10-27 */
            float lam = c[0], mu = c[cs], mu2 = mu + mu, sii = lam * (E0 + E1 + E2);
            s[0][o] = sii + mu2 * E0; s[1][o] = sii + mu2 * E1; s[2][o] = sii + mu2 * E2;
            s[3][o] = mu * E3; s[4][o] = mu * E4; s[5][o] = mu * E5;
        } else if (law == 1) {      /* TransverselyIsotropic1D.cpp:9-26 */
            float A = c[0], C = c[cs], F = c[2 * cs], L = c[3 * cs], N = c[4 * cs], N2 = N + N;
            float e01 = E0 + E1, t = A * e01 + F * E2;
            s[0][o] = t - N2 * E1; s[1][o] = t - N2 * E0; s[2][o] = C * E2 + F * e01;
            s[3][o] = L * E3; s[4][o] = L * E4; s[5][o] = N * E5;
        } else {                    /* Anisotropic1D.cpp:9-54 */
            static const int idx[6][6] = {{0, 1, 2, 3, 4, 5},    {1, 6, 7, 8, 9, 10},   {2, 7, 11, 12, 13, 14},
                                          {3, 8, 12, 15, 16, 17}, {4, 9, 13, 16, 18, 19}, {5, 10, 14, 17, 19, 20}};
            float ev[6] = {E0, E1, E2, E3, E4, E5};
            for (int i = 0; i < 6; ++i) {
                float a = 0.f;
                for (int j = 0; j < 6; ++j) a += c[idx[i][j] * cs] * ev[j];
                s[i][o] = a;
            }
        }
    }
}

/* SLS cell update (Attenuation3D_Full.cpp:16-50 etc.) on one scalar per component */
static inline void att_cell(int nsls, const float *al, const float *be, const float *ga, float dk3, float dmu, float dmu2, int do_kappa,
                            const float e[6], float s[6], float *mem /*[nsls] stride ms, comp stride cst*/, size_t ms, size_t cst,
                            float *R /*comp stride cst*/) {
    const float third = (float)(1.0 / 3.0);
    float Rn[6];
    float e3 = (e[0] + e[1] + e[2]) * third;
    if (do_kappa) {
        float s3 = dk3 * e3;
        Rn[0] = s3 + dmu2 * (e[0] - e3); Rn[1] = s3 + dmu2 * (e[1] - e3); Rn[2] = s3 + dmu2 * (e[2] - e3);
    } else {
        Rn[0] = dmu2 * (e[0] - e3); Rn[1] = dmu2 * (e[1] - e3); Rn[2] = -(Rn[0] + Rn[1]);
    }
    Rn[3] = dmu * e[3]; Rn[4] = dmu * e[4]; Rn[5] = dmu * e[5];
    for (int c = 0; c < 6; ++c) {
        float Ro = R[c * cst];
        for (int k = 0; k < nsls; ++k) {
            float m = mem[k * ms + c * cst];
            s[c] -= m;
            m = al[k] * m + be[k] * Ro;
            m += ga[k] * Rn[c];
            mem[k * ms + c * cst] = m;
        }
        R[c * cst] = Rn[c];
    }
}

static inline void atomic_sub(float *p, float v) {
#pragma omp atomic
    *p -= v;
}

/* Element::computeStiff for every element of a group (SolidElement.cpp:43-65, 404-443; FluidElement.cpp:43-65, 333-355).
 * displ/stiff: point arrays [npoint][ncomp][Mmax] complex; nlive[npoint] = Nu_p - nyq_p + 1. */
void orc_compute_stiff(const orc_group *g, const float *G_GLL, const float *G_GLJ, const float *displ_, float *stiff_,
                       const int *nlive, int Mmax) {
    const int E = g->E, M = g->M, N = g->Nr, ncomp = g->fluid ? 1 : 3, nstr = g->fluid ? 3 : 6;
    const float *Gxi = g->axial ? G_GLJ : G_GLL, *Geta = G_GLL;
    const cpx *displ = (const cpx *)displ_;
    fft_plan plan;
    if (g->is3d) plan_init(&plan, N);
#pragma omp parallel
    {
        cpx *u = (cpx *)malloc(sizeof(cpx) * (size_t)M * ncomp * 25);
        cpx *e = (cpx *)malloc(sizeof(cpx) * (size_t)M * nstr * 25);
        cpx *s = (cpx *)malloc(sizeof(cpx) * (size_t)M * nstr * 25);
        cpx *f = (cpx *)malloc(sizeof(cpx) * (size_t)M * ncomp * 25);
        cpx *zin = NULL, *zph = NULL;
        float *ph = NULL;
        if (g->is3d) {
            zin = (cpx *)malloc(sizeof(cpx) * (size_t)N);
            zph = (cpx *)malloc(sizeof(cpx) * (size_t)N);
            ph = (float *)malloc(sizeof(float) * (size_t)N * 12);   /* e[6][N], s[6][N] of one point */
        }
#pragma omp for schedule(dynamic, 1)
        for (int ie = 0; ie < E; ++ie) {
            const int *pid = g->pidx + (size_t)ie * 25;
            const float *g0 = g->dsdxii + (size_t)ie * 25, *g1 = g->dsdeta + (size_t)ie * 25, *g2 = g->dzdxii + (size_t)ie * 25,
                        *g3 = g->dzdeta + (size_t)ie * 25, *g4 = g->inv_s + (size_t)ie * 25;
            /* gather (SolidPoint.cpp:175-195) */
            for (int a = 0; a < M; ++a)
                for (int c = 0; c < ncomp; ++c)
                    for (int p = 0; p < 25; ++p) {
                        cpx v = C0;
                        if (a < nlive[pid[p]]) v = displ[((size_t)pid[p] * ncomp + c) * Mmax + a];
                        if (a == 0) v.im = 0.f;
                        u[((size_t)a * ncomp + c) * 25 + p] = v;
                    }
            /* gradient */
            for (int a = 0; a < M; ++a) {
                cpx *ea = e + (size_t)a * nstr * 25;
                if (g->nyq && a == M - 1) { memset(ea, 0, sizeof(cpx) * nstr * 25); continue; }
                if (!g->fluid) grad6_mode(Gxi, Geta, g0, g1, g2, g3, g4, g->axial, a, u + (size_t)a * 75, ea);
                else grad_fluid_mode(Gxi, Geta, g0, g1, g2, g3, g4, g->axial, a, u + (size_t)a * 25, ea);
            }
            float trig[4][25];
            if (g->tiso)
                for (int p = 0; p < 25; ++p) {
                    double th = g->theta[(size_t)ie * 25 + p];
                    trig[0][p] = (float)sin(th); trig[1][p] = (float)cos(th);
                    trig[2][p] = (float)sin(2 * th); trig[3][p] = (float)cos(2 * th);
                }
            if (g->tiso)   /* strain -> RTZ, per mode and point, on re and im parts */
                for (int a = 0; a < M; ++a)
                    for (int p = 0; p < 25; ++p) {
                        float *ur[6], *ui[6];
                        for (int c = 0; c < 6; ++c) { ur[c] = &e[((size_t)a * 6 + c) * 25 + p].re; ui[c] = &e[((size_t)a * 6 + c) * 25 + p].im; }
                        rot_fwd(ur, 1, trig[0][p], trig[1][p], trig[2][p], trig[3][p]);
                        rot_fwd(ui, 1, trig[0][p], trig[1][p], trig[2][p], trig[3][p]);
                    }
            /* constitutive law */
            if (!g->is3d) {
                for (int a = 0; a < M; ++a)
                    for (int p = 0; p < 25; ++p) {
                        cpx *ea = e + (size_t)a * nstr * 25 + p, *sa = s + (size_t)a * nstr * 25 + p;
                        if (g->fluid) {
                            const float K = g->coef[(size_t)ie * 25 + p];   /* Acoustic1D.cpp:8-16 */
                            for (int c = 0; c < 3; ++c) sa[c * 25] = c_scl(ea[c * 25], K);
                            continue;
                        }
                        const float *cf = g->coef + (size_t)ie * 25 + p;
                        const size_t cs = (size_t)E * 25;
                        float *er[6], *ei[6], *sr[6], *si[6];
                        for (int c = 0; c < 6; ++c) { er[c] = &ea[c * 25].re; ei[c] = &ea[c * 25].im; sr[c] = &sa[c * 25].re; si[c] = &sa[c * 25].im; }
                        stress_n(g->law, cf, cs, er, sr, 1, 0);
                        stress_n(g->law, cf, cs, ei, si, 1, 0);
                        if (g->att_kind) {
                            int q = p;
                            if (g->att_kind == 2) { q = -1; for (int k = 0; k < 4; ++k) if (CG4[k] == p) q = k; }
                            if (q >= 0) {
                                const int P = g->P;
                                const size_t mo = ((size_t)ie * P + q);   /* rows = 1 */
                                for (int part = 0; part < 2; ++part) {
                                    float ev[6], sv[6];
                                    for (int c = 0; c < 6; ++c) { ev[c] = part ? *ei[c] : *er[c]; sv[c] = part ? *si[c] : *sr[c]; }
                                    /* state [nsls][E][M][6][P] complex -> float index */
                                    float *mem = g->memvar + 2 * ((((size_t)0 * E + ie) * M + a) * 6 * P + q) + part;
                                    float *R = g->stressR + 2 * ((((size_t)ie) * M + a) * 6 * P + q) + part;
                                    att_cell(g->nsls, g->alpha + (size_t)ie * g->nsls, g->beta + (size_t)ie * g->nsls,
                                             g->gamma + (size_t)ie * g->nsls, g->dk3[mo], g->dmu[mo], g->dmu2[mo], g->do_kappa, ev, sv,
                                             mem, 2 * (size_t)E * M * 6 * P, 2 * (size_t)P, R);
                                    for (int c = 0; c < 6; ++c) { if (part) *si[c] = sv[c]; else *sr[c] = sv[c]; }
                                }
                            }
                        }
                    }
            } else {
                const int npair = g->fluid ? 2 : 3;
                const float invN = 1.0f / (float)N;
                for (int p = 0; p < 25; ++p) {
                    /* c2r, two real columns per complex transform (FieldFFT.cpp:19-25; SolverFFTW_N6.cpp:51-53) */
                    for (int pr = 0; pr < npair; ++pr) {
                        const int ca = 2 * pr, cb = 2 * pr + 1, hasb = cb < nstr;
                        for (int k = 0; k < N; ++k) zin[k] = C0;
                        for (int k = 0; k < M; ++k) {
                            cpx A = e[((size_t)k * nstr + ca) * 25 + p], B = hasb ? e[((size_t)k * nstr + cb) * 25 + p] : C0;
                            if (k == 0 || 2 * k == N) { A.im = 0.f; B.im = 0.f; }
                            zin[k].re = A.re - B.im; zin[k].im = A.im + B.re;
                            if (k >= 1 && 2 * k < N) { zin[N - k].re = A.re + B.im; zin[N - k].im = B.re - A.im; }
                        }
                        fft_exec(&plan, zin, zph, 1);
                        for (int k = 0; k < N; ++k) { ph[(size_t)ca * N + k] = zph[k].re; if (hasb) ph[(size_t)cb * N + k] = zph[k].im; }
                    }
                    if (g->fluid) {
                        const float *K = g->coef + ((size_t)ie * N) * 25 + p;   /* [E][rows][25] */
                        for (int c = 0; c < 3; ++c)
                            for (int k = 0; k < N; ++k) ph[(size_t)(6 + c) * N + k] = K[(size_t)k * 25] * ph[(size_t)c * N + k];
                    } else {
                        float *er[6], *sr[6];
                        for (int c = 0; c < 6; ++c) { er[c] = ph + (size_t)c * N; sr[c] = ph + (size_t)(6 + c) * N; }
                        const size_t cs = (size_t)E * N * 25;
                        for (int k = 0; k < N; ++k) {
                            float *ek[6], *sk[6];
                            for (int c = 0; c < 6; ++c) { ek[c] = er[c] + k; sk[c] = sr[c] + k; }
                            stress_n(g->law, g->coef + ((size_t)ie * N + k) * 25 + p, cs, ek, sk, 1, 0);
                        }
                        if (g->att_kind) {
                            int q = p;
                            if (g->att_kind == 2) { q = -1; for (int k = 0; k < 4; ++k) if (CG4[k] == p) q = k; }
                            if (q >= 0) {
                                const int P = g->P;
                                for (int k = 0; k < N; ++k) {
                                    float ev[6], sv[6];
                                    for (int c = 0; c < 6; ++c) { ev[c] = er[c][k]; sv[c] = sr[c][k]; }
                                    const size_t mo = ((size_t)ie * N + k) * P + q;
                                    float *mem = g->memvar + (((size_t)ie * N + k) * 6) * P + q;
                                    float *R = g->stressR + (((size_t)ie * N + k) * 6) * P + q;
                                    att_cell(g->nsls, g->alpha + (size_t)ie * g->nsls, g->beta + (size_t)ie * g->nsls,
                                             g->gamma + (size_t)ie * g->nsls, g->dk3[mo], g->dmu[mo], g->dmu2[mo], g->do_kappa, ev, sv,
                                             mem, (size_t)E * N * 6 * P, (size_t)P, R);
                                    for (int c = 0; c < 6; ++c) sr[c][k] = sv[c];
                                }
                            }
                        }
                    }
                    /* r2c, scaled by 1/Nr (SolverFFTW_N6.cpp:45-49) */
                    for (int pr = 0; pr < npair; ++pr) {
                        const int ca = 2 * pr, cb = 2 * pr + 1, hasb = cb < nstr;
                        for (int k = 0; k < N; ++k) { zin[k].re = ph[(size_t)(6 + ca) * N + k]; zin[k].im = hasb ? ph[(size_t)(6 + cb) * N + k] : 0.f; }
                        fft_exec(&plan, zin, zph, 0);
                        for (int k = 0; k < M; ++k) {
                            cpx A, B, zk = zph[k];
                            if (k >= 1 && 2 * k < N) {
                                cpx w = zph[N - k];
                                A.re = 0.5f * invN * (zk.re + w.re); A.im = 0.5f * invN * (zk.im - w.im);
                                B.re = 0.5f * invN * (zk.im + w.im); B.im = 0.5f * invN * (w.re - zk.re);
                            } else { A.re = invN * zk.re; A.im = 0.f; B.re = invN * zk.im; B.im = 0.f; }
                            s[((size_t)k * nstr + ca) * 25 + p] = A;
                            if (hasb) s[((size_t)k * nstr + cb) * 25 + p] = B;
                        }
                    }
                }
            }
            if (g->tiso)
                for (int a = 0; a < M; ++a)
                    for (int p = 0; p < 25; ++p) {
                        float *ur[6], *ui[6];
                        for (int c = 0; c < 6; ++c) { ur[c] = &s[((size_t)a * 6 + c) * 25 + p].re; ui[c] = &s[((size_t)a * 6 + c) * 25 + p].im; }
                        rot_bwd(ur, 1, trig[0][p], trig[1][p], trig[2][p], trig[3][p]);
                        rot_bwd(ui, 1, trig[0][p], trig[1][p], trig[2][p], trig[3][p]);
                    }
            /* quadrature + scatter (SolidPoint.cpp:197-209) */
            for (int a = 0; a < M; ++a) {
                if (g->nyq && a == M - 1) continue;
                cpx *fa = f + (size_t)a * ncomp * 25;
                if (a == 0) for (int k = 0; k < nstr * 25; ++k) s[k].im = 0.f;
                if (!g->fluid) quad6_mode(Gxi, Geta, g0, g1, g2, g3, g4, g->axial, a, s + (size_t)a * 150, fa);
                else quad_fluid_mode(Gxi, Geta, g0, g1, g2, g3, g4, g->axial, a, s + (size_t)a * 75, fa);
                for (int c = 0; c < ncomp; ++c)
                    for (int p = 0; p < 25; ++p) {
                        if (a >= nlive[pid[p]]) continue;
                        float *dst = stiff_ + 2 * (((size_t)pid[p] * ncomp + c) * Mmax + a);
                        atomic_sub(dst, fa[c * 25 + p].re);
                        if (a > 0) atomic_sub(dst + 1, fa[c * 25 + p].im);
                    }
            }
        }
        free(u); free(e); free(s); free(f);
        if (g->is3d) { free(zin); free(zph); free(ph); }
    }
    if (g->is3d) plan_free(&plan);
}

/* Point::updateNewmark with Mass1D (SolidPoint.cpp:23-38, 216-238; FluidPoint.cpp:23-44, 197-207; Mass1D.cpp:12-18).
 * arrays [npoint][ncomp][Mmax] complex; invmass[npoint]; nu, nr, axial per point. */
void orc_update_newmark(int npoint, int ncomp, int Mmax, const int *nu, const int *nr, const unsigned char *axial,
                        const unsigned char *surf, const float *invmass, float *displ_, float *veloc_, float *accel_, float *stiff_,
                        double dt) {
    cpx *displ = (cpx *)displ_, *veloc = (cpx *)veloc_, *accel = (cpx *)accel_, *stiff = (cpx *)stiff_;
    const float hdt = (float)(0.5 * dt), fdt = (float)dt, hdd = (float)(0.5 * dt * dt);
#pragma omp parallel for schedule(static)
    for (int p = 0; p < npoint; ++p) {
        const int n = nu[p] + 1, nyq = (nr[p] % 2 == 0);
        cpx *st = stiff + (size_t)p * ncomp * Mmax;
        if (surf && surf[p]) {
            for (int k = 0; k < ncomp * Mmax; ++k) {
                size_t i = (size_t)p * ncomp * Mmax + k;
                displ[i] = veloc[i] = accel[i] = stiff[i] = C0;
            }
            continue;
        }
        /* maskField; a scalar mass commutes with it, so one application suffices */
        for (int c = 0; c < ncomp; ++c) st[(size_t)c * Mmax].im = 0.f;
        if (axial[p]) {
            if (ncomp == 3) {
                st[0] = C0; st[Mmax] = C0;
                if (n > 1) {
                    cpx s0 = st[1], s1 = st[Mmax + 1];
                    st[1].re = 0.5f * (s0.re + s1.im); st[1].im = 0.5f * (s0.im - s1.re);
                    st[Mmax + 1].re = 0.5f * (s1.re - s0.im); st[Mmax + 1].im = 0.5f * (s1.im + s0.re);
                    st[2 * Mmax + 1] = C0;
                    for (int c = 0; c < 3; ++c) for (int a = 2; a < n; ++a) st[(size_t)c * Mmax + a] = C0;
                }
            } else {
                for (int a = 1; a < n; ++a) st[a] = C0;
            }
        }
        if (nyq) for (int c = 0; c < ncomp; ++c) st[(size_t)c * Mmax + n - 1] = C0;
        for (int c = 0; c < ncomp; ++c)
            for (int a = 0; a < n; ++a) {
                size_t i = ((size_t)p * ncomp + c) * Mmax + a;
                cpx acc = c_scl(stiff[i], invmass[p]);
                veloc[i].re += hdt * (accel[i].re + acc.re); veloc[i].im += hdt * (accel[i].im + acc.im);
                accel[i] = acc;
                displ[i].re += fdt * veloc[i].re + hdd * acc.re; displ[i].im += fdt * veloc[i].im + hdd * acc.im;
                stiff[i] = C0;
            }
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's CPU arm sets the thread count explicitly: launchers such as torchrun export OMP_NUM_THREADS=1 */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
