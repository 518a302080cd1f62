// oracle/shim/fftw3.h -- TEST INFRASTRUCTURE, not product code.
//
// Stand-in for the four FFTW 3.3 entry points the reference calls on the hot path (SolverFFTW.h:11-25 macros;
// SolverFFTW_{1,3,N3,N6,N9}.cpp:30-33, 46-52): fftw[f]_plan_many_dft_r2c / _c2r, fftw[f]_execute, fftw[f]_destroy_plan.
// FFTW is not in this image.  Semantics are FFTW's documented ones: rank-1 batched transforms with (stride, dist)
// layouts, r2c forward sign -1, c2r backward sign +1 (imaginary parts of the DC and Nyquist inputs ignored), both
// unnormalised.  The sums are evaluated directly in double precision, O(n^2) per column: small test sizes only.
#pragma once
#include <cmath>
#include <vector>

#define FFTW_ESTIMATE (1U << 6)
#define FFTW_MEASURE (0U)
#define FFTW_PATIENT (1U << 5)
#define FFTW_EXHAUSTIVE (1U << 3)

template <class R>
struct ax_fftw_plan_s {
    int n, howmany;
    bool r2c;
    R *rp;
    R (*cp)[2];
    int rstride, rdist, cstride, cdist;
    std::vector<double> cs, sn;
};

template <class R>
inline ax_fftw_plan_s<R> *ax_fftw_make(int rank, const int *n, int howmany, R *rp, int rstride, int rdist, R (*cp)[2], int cstride,
                                       int cdist, bool r2c) {
    (void)rank;
    ax_fftw_plan_s<R> *p = new ax_fftw_plan_s<R>;
    p->n = n[0]; p->howmany = howmany; p->r2c = r2c; p->rp = rp; p->cp = cp;
    p->rstride = rstride; p->rdist = rdist; p->cstride = cstride; p->cdist = cdist;
    p->cs.resize(p->n); p->sn.resize(p->n);
    for (int k = 0; k < p->n; ++k) {
        const double a = 2.0 * 3.14159265358979323846264338327950288 * (double)k / (double)p->n;
        p->cs[k] = std::cos(a); p->sn[k] = std::sin(a);
    }
    return p;
}
template <class R>
inline void ax_fftw_run(const ax_fftw_plan_s<R> *p) {
    const int n = p->n, nc = n / 2 + 1;
    std::vector<double> x(n);
    for (int h = 0; h < p->howmany; ++h) {
        R *r = p->rp + (long)h * p->rdist;
        R (*c)[2] = p->cp + (long)h * p->cdist;
        if (p->r2c) {
            for (int j = 0; j < n; ++j) x[j] = (double)r[(long)j * p->rstride];
            for (int k = 0; k < nc; ++k) {
                double re = 0.0, im = 0.0;
                for (int j = 0; j < n; ++j) {
                    const int t = (int)(((long)j * k) % n);
                    re += x[j] * p->cs[t];
                    im -= x[j] * p->sn[t];
                }
                c[(long)k * p->cstride][0] = (R)re;
                c[(long)k * p->cstride][1] = (R)im;
            }
        } else {
            for (int j = 0; j < n; ++j) {
                double s = (double)c[0][0];
                for (int k = 1; k < nc; ++k) {
                    const int t = (int)(((long)j * k) % n);
                    const double re = (double)c[(long)k * p->cstride][0], im = (double)c[(long)k * p->cstride][1];
                    if (2 * k == n) s += re * p->cs[t];                       // Nyquist: real part only, counted once
                    else s += 2.0 * (re * p->cs[t] - im * p->sn[t]);
                }
                x[j] = s;
            }
            for (int j = 0; j < n; ++j) r[(long)j * p->rstride] = (R)x[j];
        }
    }
}

typedef float fftwf_complex[2];
typedef double fftw_complex[2];
typedef ax_fftw_plan_s<float> *fftwf_plan;
typedef ax_fftw_plan_s<double> *fftw_plan;

inline fftwf_plan fftwf_plan_many_dft_r2c(int rank, const int *n, int howmany, float *in, const int *, int istride, int idist,
                                          fftwf_complex *out, const int *, int ostride, int odist, unsigned) {
    return ax_fftw_make<float>(rank, n, howmany, in, istride, idist, out, ostride, odist, true);
}
inline fftwf_plan fftwf_plan_many_dft_c2r(int rank, const int *n, int howmany, fftwf_complex *in, const int *, int istride, int idist,
                                          float *out, const int *, int ostride, int odist, unsigned) {
    return ax_fftw_make<float>(rank, n, howmany, out, ostride, odist, in, istride, idist, false);
}
inline void fftwf_execute(const fftwf_plan p) { ax_fftw_run<float>(p); }
inline void fftwf_destroy_plan(fftwf_plan p) { delete p; }
inline fftw_plan fftw_plan_many_dft_r2c(int rank, const int *n, int howmany, double *in, const int *, int istride, int idist,
                                        fftw_complex *out, const int *, int ostride, int odist, unsigned) {
    return ax_fftw_make<double>(rank, n, howmany, in, istride, idist, out, ostride, odist, true);
}
inline fftw_plan fftw_plan_many_dft_c2r(int rank, const int *n, int howmany, fftw_complex *in, const int *, int istride, int idist,
                                        double *out, const int *, int ostride, int odist, unsigned) {
    return ax_fftw_make<double>(rank, n, howmany, out, ostride, odist, in, istride, idist, false);
}
inline void fftw_execute(const fftw_plan p) { ax_fftw_run<double>(p); }
inline void fftw_destroy_plan(fftw_plan p) { delete p; }
// wisdom (SolverFFTW.cpp:18-64): the stand-in has no planner, so there is nothing to import or export
inline int fftwf_import_wisdom_from_string(const char *) { return 0; }
inline int fftw_import_wisdom_from_string(const char *) { return 0; }
inline int fftwf_export_wisdom_to_filename(const char *) { return 1; }
inline int fftw_export_wisdom_to_filename(const char *) { return 1; }
